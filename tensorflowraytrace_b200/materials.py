"""
Refractive-index models n(wavelength [nm]) -- the public surface of tfrt/materials.py:25-104.

Every material is a callable (wavelength tensor -> index tensor, evaluated with torch on whatever
device the input lives on, like the reference's TF callables) that also carries a `descriptor`
(kind, coefficients) so the fused CUDA reaction kernel can evaluate the same formula per ray
without leaving the GPU (include/tfrt_b200.h: TfrtMaterial).  Coefficients are the published
refractiveindex.info fits the reference uses (wavelength converted to nm).
"""

import torch

MAT_CONSTANT, MAT_SELLMEIER, MAT_LAURENT, MAT_POLY = 0, 1, 2, 3


class Material:
    """Callable refractive index with a kernel descriptor."""

    def __init__(self, kind, coefficients, name=None):
        self.kind = int(kind)
        self.coefficients = [float(c) for c in coefficients]
        self.name = name or f"material(kind={kind})"

    @property
    def descriptor(self):
        return self.kind, self.coefficients

    def __call__(self, x):
        x = torch.as_tensor(x)
        c = self.coefficients
        if self.kind == MAT_CONSTANT:
            return c[0] * torch.ones_like(x)
        x2 = x * x
        if self.kind == MAT_SELLMEIER:
            s = torch.ones_like(x)
            for k in range(3):
                if c[2 * k] != 0.0:
                    s = s + c[2 * k] * x2 / (x2 - c[2 * k + 1])
            return torch.sqrt(s)
        if self.kind == MAT_LAURENT:
            i2 = 1.0 / x2
            return torch.sqrt(c[0] + c[1] * x2 + c[2] * x2 * x2 + c[3] * i2 + c[4] * i2 ** 2
                              + c[5] * i2 ** 3 + c[6] * i2 ** 4)
        if self.kind == MAT_POLY:
            return c[0] + c[1] * x2 + c[2] / x2
        raise ValueError(f"unknown material kind {self.kind}")

    def __repr__(self):
        return f"<Material {self.name}>"


def build_constant_material(n):
    """Material with a constant refractive index (tfrt/materials.py:25-46)."""
    return Material(MAT_CONSTANT, [n], f"constant({n})")


def sellmeier(b1, c1, b2=0.0, c2=0.0, b3=0.0, c3=0.0, name="sellmeier"):
    """n = sqrt(1 + sum B x^2 / (x^2 - C)), C in nm^2."""
    return Material(MAT_SELLMEIER, [b1, c1, b2, c2, b3, c3, 0.0], name)


vacuum = Material(MAT_CONSTANT, [1.0], "vacuum")                      # materials.py:103-104
reflective = Material(MAT_CONSTANT, [0.0], "reflective")              # materials.py:95-96 (mirror)
acrylic = Material(                                                    # materials.py:52-61
    MAT_LAURENT, [2.1778, 6.1209e-9, -1.5004e-15, 2.3678e4, -4.2137e9, 7.3417e14, -4.5042e19],
    "acrylic")
crown_glass = sellmeier(1.1273555e0, 7.20341707e3, 1.24412303e-1, 2.69835916e4,      # :64-70
                        8.27100531e-1, 1.00384588e8, "crown_glass")
flint_glass = sellmeier(1.34533359e0, 9.97743871e3, 2.09073176e-1, 4.70450767e4,     # :73-79
                        9.37357162e-1, 1.11886764e8, "flint_glass")
fused_silica = sellmeier(6.961663e-1, 4.679148e3, 4.079426e-1, 1.3512063e4,          # :82-88
                         8.974794e-1, 9.7934002538e7, "fused_silica")
polycarbonate = sellmeier(1.4182e0, 2.1304e4, name="polycarbonate")                 # :91-92
soda_lime = Material(MAT_POLY, [1.5130e0, -3.169e-9, 3.962e3], "soda_lime")          # :99-100

BY_NAME = {
    "vacuum": vacuum, "reflective": reflective, "acrylic": acrylic, "crown_glass": crown_glass,
    "flint_glass": flint_glass, "fused_silica": fused_silica, "polycarbonate": polycarbonate,
    "soda_lime": soda_lime,
}


def descriptor_of(material):
    """Kernel descriptor of a materials-list entry ({"n": callable} or a bare callable)."""
    fn = material["n"] if isinstance(material, dict) else material
    if isinstance(fn, Material):
        return fn.descriptor
    raise TypeError(
        "tensorflowraytrace_b200: the fused reaction kernel evaluates refractive indices on the "
        f"GPU and needs materials built from tensorflowraytrace_b200.materials, got {fn!r}.  "
        "Use build_constant_material / sellmeier / the named materials, or "
        "StandardReaction('value') with per-boundary n_in / n_out fields.")
