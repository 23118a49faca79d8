"""
Oracle restatement of tfrt/materials.py:25-104 (float64 torch).  TEST INFRASTRUCTURE ONLY.
Wavelength in nm -> refractive index.
"""

import torch


def build_constant_material(n):
    """materials.py:25-46"""
    return lambda x: n * torch.ones_like(x)


def acrylic(x):
    """materials.py:52-61"""
    return torch.sqrt(
        2.1778
        + 6.1209e-9 * x ** 2
        - 1.5004e-15 * x ** 4
        + 2.3678e4 * x ** -2
        - 4.2137e9 * x ** -4
        + 7.3417e14 * x ** -6
        - 4.5042e19 * x ** -8
    )


def crown_glass(x):
    """materials.py:64-70"""
    return torch.sqrt(
        1
        + 1.1273555e0 * x ** 2 / (x ** 2 - 7.20341707e3)
        + 1.24412303e-1 * x ** 2 / (x ** 2 - 2.69835916e4)
        + 8.27100531e-1 * x ** 2 / (x ** 2 - 1.00384588e8)
    )


def flint_glass(x):
    """materials.py:73-79"""
    return torch.sqrt(
        1
        + 1.34533359e0 * x ** 2 / (x ** 2 - 9.97743871e3)
        + 2.09073176e-1 * x ** 2 / (x ** 2 - 4.70450767e4)
        + 9.37357162e-1 * x ** 2 / (x ** 2 - 1.11886764e8)
    )


def fused_silica(x):
    """materials.py:82-88"""
    return torch.sqrt(
        1
        + 6.961663e-1 * x ** 2 / (x ** 2 - 4.679148e3)
        + 4.079426e-1 * x ** 2 / (x ** 2 - 1.3512063e4)
        + 8.974794e-1 * x ** 2 / (x ** 2 - 9.7934002538e7)
    )


def polycarbonate(x):
    """materials.py:91-92"""
    return torch.sqrt(1 + 1.4182e0 * x ** 2 / (x ** 2 - 2.1304e4))


def reflective(x):
    """materials.py:95-96"""
    return torch.zeros_like(x)


def soda_lime(x):
    """materials.py:99-100"""
    return 1.5130e0 - 3.169e-9 * x ** 2 + 3.962e3 * x ** -2


def vacuum(x):
    """materials.py:103-104"""
    return torch.ones_like(x)


BY_NAME = {
    "vacuum": vacuum,
    "acrylic": acrylic,
    "crown_glass": crown_glass,
    "flint_glass": flint_glass,
    "fused_silica": fused_silica,
    "polycarbonate": polycarbonate,
    "reflective": reflective,
    "soda_lime": soda_lime,
}
