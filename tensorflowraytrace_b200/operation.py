"""
Ray operations: the hook API of tfrt/operation.py (annotate / preprocess / main / postprocess,
signature properties, exclusions) with the same names and meanings.

`StandardReaction` (refraction / reflection / TIR with material look-up, tfrt/operation.py:200-307)
is what the fused CUDA trace loop implements; `OpticalEngine.ray_trace` recognises it by type and
runs the whole iteration loop on the GPU.  Its `main` is also available on its own (used by
`OpticalEngine.single_pass`, the projection-at-a-time path taken when other operations are in
the list): there the reaction is the same formulas as differentiable torch ops on the GPU
(`OpticalEngine._react_projection`, fp64 arithmetic on the fp32 fields in 3D), not the fused
kernel.
"""

import torch


class RayOperation:
    """Base class of modular ray operations (tfrt/operation.py:25-162)."""

    def __init__(self, active=True):
        self.active = active

    @property
    def input_signature(self):
        return set()

    @property
    def output_signature(self):
        return set()

    @property
    def optical_signature(self):
        return set()

    @property
    def stop_signature(self):
        return set()

    @property
    def target_signature(self):
        return set()

    @property
    def material_signature(self):
        return set()

    @property
    def simple_ray_inheritance(self):
        return set()

    def annotate(self, engine):
        pass

    def preprocess(self, engine, proj_result):
        pass

    def main(self, engine, proj_result):
        return {}

    def postprocess(self, engine, proj_result, new_rays):
        pass

    @property
    def exclusions(self):
        return set()


class OldestAncestor(RayOperation):
    """Relates output rays to the source ray they descend from (tfrt/operation.py:166-196).
    `annotate` writes `oldest_ancestor` = running source index into every source; the field is
    inherited by every descendant."""

    @property
    def input_signature(self):
        return {"oldest_ancestor"}

    @property
    def output_signature(self):
        return self.input_signature

    @property
    def simple_ray_inheritance(self):
        return self.input_signature

    def annotate(self, engine):
        start = 0
        for source in engine.optical_system._sources:
            count = source["x_start"].shape[0]
            source["oldest_ancestor"] = torch.arange(
                start, start + count, device=source["x_start"].device)
            start += count


class StandardReaction(RayOperation):
    """Refraction with possible reflection / TIR (tfrt/operation.py:200-307).

    refractive_index_type : "index" -> optical boundaries carry integer `mat_in` / `mat_out`
        indices into `system.materials`, rays carry `wavelength` (nm);
        "value" -> optical boundaries carry float `n_in` / `n_out`.
    """

    def __init__(self, refractive_index_type="index", **kwargs):
        super().__init__(**kwargs)
        if refractive_index_type not in {"index", "value"}:
            raise ValueError(
                f"StandardReaction: received invalid value {refractive_index_type}.  "
                "Must be 'index' or 'value'.")
        self._refractive_index_type = refractive_index_type

    @property
    def refractive_index_type(self):
        return self._refractive_index_type

    @property
    def input_signature(self):
        return {"wavelength"} if self._refractive_index_type == "index" else set()

    @property
    def output_signature(self):
        return self.input_signature

    @property
    def simple_ray_inheritance(self):
        return self.input_signature

    @property
    def optical_signature(self):
        if self._refractive_index_type == "index":
            return {"mat_in", "mat_out"}
        return {"n_in", "n_out"}

    @property
    def material_signature(self):
        return {"n"} if self._refractive_index_type == "index" else set()

    def main(self, engine, proj_result):
        """One reaction step on a projection result (the unfused form of the loop body)."""
        if "active" not in proj_result["rays"]:
            return {}
        return engine._react_projection(self, proj_result)


class GhostThrough(RayOperation):
    """Test operation: rays pass straight through optical surfaces
    (tfrt/operation.py:311-333).  The 3D z_start follows the new start point (the reference's
    `z_start = rays["z_start"]` at :330 is a typo the new rays would not survive)."""

    def main(self, engine, proj_result):
        if "active" not in proj_result["rays"]:
            return {}
        rays = proj_result["rays"]["active"]
        new_rays = {
            "x_start": rays["x_end"],
            "y_start": rays["y_end"],
            "x_end": 2 * rays["x_end"] - rays["x_start"],
            "y_end": 2 * rays["y_end"] - rays["y_start"],
        }
        if engine.dimension == 3:
            new_rays["z_start"] = rays["z_end"]
            new_rays["z_end"] = 2 * rays["z_end"] - rays["z_start"]
        valid = torch.ones_like(rays["x_end"], dtype=torch.bool)
        return {"active": {"rays": new_rays, "valid": valid}}
