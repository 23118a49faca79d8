"""
Oracle restatement of the 3D hot path of tfrt/engine.py + tfrt/operation.py (float64 torch CPU).
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

What is restated (reference file:line):
  amalgamate                         tfrt/engine.py:50-76
  OpticalSystem3D._merge_boundaries  tfrt/engine.py:971-1018
  OpticalSystem3D.intersect          tfrt/engine.py:1020-1078
  OpticalSystem3D._intersection      tfrt/engine.py:1102-1166
  OpticalEngine.process_projection_3D tfrt/engine.py:1988-2191
  OpticalEngine.single_pass          tfrt/engine.py:2193-2302
  OpticalEngine.ray_trace            tfrt/engine.py:2311-2330
  StandardReaction.main              tfrt/operation.py:255-307

The dense (M, N) broadcast of the reference is kept, but evaluated in chunks over the ray axis
so that each temporary stays bounded.  Chunking preserves results: the `2*reduce_max(ray_u)`
fill (engine.py:1144) only has to exceed every *valid* ray_u, which any per-chunk maximum
does; it changes only the garbage returned for rays with no valid hit.
"""

import torch

from . import geometry

OPTICAL = 0
STOP = 1
TARGET = 2

TRIANGLE_GEO_SIG = ("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2", "norm")


def amalgamate(stuff, signature=None):
    """tfrt/engine.py:50-76"""
    processed = [item for item in stuff if bool(item)]
    if not processed:
        return {}
    if not signature:
        signature = None
        for item in processed:
            keys = set(item.keys())
            signature = keys if signature is None else (signature & keys)
    return {
        field: torch.cat([item[field] for item in processed], 0) for field in signature
    }


class OracleSystem3D:
    """Field-dict restatement of OpticalSystem3D (tfrt/engine.py:870-1166)."""

    def __init__(
        self,
        sources,
        optical=None,
        stop=None,
        target=None,
        materials=None,
        intersect_epsilion=1e-10,
        size_epsilion=1e-10,
        ray_start_epsilion=1e-10,
        chunk_bytes=256 << 20,
    ):
        self._amalgamated_sources = amalgamate(sources) if isinstance(sources, list) else sources
        self._amalgamated_optical = amalgamate(optical) if isinstance(optical, list) else (optical or {})
        self._amalgamated_stop = amalgamate(stop) if isinstance(stop, list) else (stop or {})
        self._amalgamated_target = amalgamate(target) if isinstance(target, list) else (target or {})
        self.materials = materials or []
        self.intersect_epsilion = intersect_epsilion
        self.size_epsilion = size_epsilion
        self.ray_start_epsilion = ray_start_epsilion
        self.chunk_bytes = chunk_bytes
        self.tests_done = 0  # ray-triangle tests evaluated (for the CPU baseline metric)
        self._merge_boundaries()

    def _merge_boundaries(self):
        """tfrt/engine.py:971-1018: label, count, concatenate optical -> stop -> target."""
        counts = []
        parts = []
        for cat, d in ((OPTICAL, self._amalgamated_optical), (STOP, self._amalgamated_stop),
                       (TARGET, self._amalgamated_target)):
            if bool(d):
                d["catagory"] = cat * torch.ones_like(d["xp"], dtype=torch.int64)
                counts.append(d["xp"].shape[0])
            else:
                counts.append(0)
            parts.append(d)
        self._optical_count, self._stop_count, self._target_count = counts
        self._merged = amalgamate(parts, set(TRIANGLE_GEO_SIG) | {"catagory"})

    def intersect(self, rays):
        """tfrt/engine.py:1020-1078"""
        result = {}
        if bool(self._merged):
            m = self._merged
            (result["x"], result["y"], result["z"], result["valid"], result["ray_u"],
             result["trig_u"], result["trig_v"], result["gather_ray"],
             result["gather_trig"]) = self._intersection(
                rays["x_start"], rays["y_start"], rays["z_start"],
                rays["x_end"], rays["y_end"], rays["z_end"],
                m["xp"], m["yp"], m["zp"], m["x1"], m["y1"], m["z1"], m["x2"], m["y2"], m["z2"],
                self.intersect_epsilion, self.size_epsilion, self.ray_start_epsilion,
            )
            result["norm"] = m["norm"][result["gather_trig"]]
        return result

    def _intersection(self, rx1, ry1, rz1, rx2, ry2, rz2, xp, yp, zp, x1, y1, z1, x2, y2, z2,
                      intersect_epsilion, size_epsilion, ray_start_epsilion):
        """tfrt/engine.py:1102-1166.

        The dense (M, N) pass decides `valid` and `argmin`; those are not differentiable in the
        reference either.  The six gathered values (x, y, z, ray_u, trig_u, trig_v) are the
        dense tensors' entries at (closest, ray); here they are re-evaluated 1-to-1 with the
        same formula on the same operands (bit-identical values) so that autograd sees exactly
        the reference's differentiable path without keeping (M, N) graphs alive.
        """
        N = rx1.shape[0]
        M = xp.shape[0]
        tri = (xp, yp, zp, x1, y1, z1, x2, y2, z2)
        chunk = max(1, int(self.chunk_bytes // (8 * max(M, 1))))
        closest_parts, valid_parts = [], []
        with torch.no_grad():
            for s in range(0, N, chunk):
                e = min(N, s + chunk)
                ray = (rx1[s:e], ry1[s:e], rz1[s:e], rx2[s:e], ry2[s:e], rz2[s:e])
                _, _, _, valid, ray_u, trig_u, trig_v = geometry.line_triangle_intersect(
                    *ray, *tri, intersect_epsilion)
                # engine.py:1138-1141
                valid = torch.logical_and(valid, trig_u >= -size_epsilion)
                valid = torch.logical_and(valid, trig_v >= -size_epsilion)
                valid = torch.logical_and(valid, trig_u + trig_v <= 1 + size_epsilion)
                valid = torch.logical_and(valid, ray_u >= ray_start_epsilion)
                # engine.py:1144-1148 (torch.argmin returns the first minimal index, like Eigen)
                inf = 2 * torch.max(ray_u) * torch.ones_like(ray_u)
                ray_u = torch.where(valid, ray_u, inf)
                closest_parts.append(torch.argmin(ray_u, dim=0))
                valid_parts.append(torch.any(valid, dim=0))
                self.tests_done += (e - s) * M
        if N == 0:
            closest_trig = torch.zeros((0,), dtype=torch.int64)
            valid = torch.zeros((0,), dtype=torch.bool)
        else:
            closest_trig = torch.cat(closest_parts)
            valid = torch.cat(valid_parts)
        gather_ray = torch.arange(N, dtype=torch.int64)
        gather_triangle = closest_trig

        x, y, z, _, ray_u, trig_u, trig_v = geometry.raw_line_triangle_intersect(
            rx1, ry1, rz1, rx2, ry2, rz2, *[t[gather_triangle] for t in tri], intersect_epsilion)
        return x, y, z, valid, ray_u, trig_u, trig_v, gather_ray, gather_triangle

    def second_best(self, rays, exclude_trig):
        """Test helper (no reference equivalent): for each ray, the smallest valid ray_u over
        all triangles except `exclude_trig[ray]`; +inf where none.  Used to classify near-ties."""
        m = self._merged
        tri = tuple(m[k] for k in ("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2"))
        N = rays["x_start"].shape[0]
        M = tri[0].shape[0]
        chunk = max(1, int(self.chunk_bytes // (8 * max(M, 1))))
        out = []
        with torch.no_grad():
            for s in range(0, N, chunk):
                e = min(N, s + chunk)
                ray = tuple(rays[k][s:e] for k in ("x_start", "y_start", "z_start", "x_end", "y_end", "z_end"))
                _, _, _, valid, ray_u, trig_u, trig_v = geometry.line_triangle_intersect(
                    *ray, *tri, self.intersect_epsilion)
                valid = valid & (trig_u >= -self.size_epsilion) & (trig_v >= -self.size_epsilion) \
                    & (trig_u + trig_v <= 1 + self.size_epsilion) & (ray_u >= self.ray_start_epsilion)
                ray_u = torch.where(valid, ray_u, torch.full_like(ray_u, float("inf")))
                ray_u[exclude_trig[s:e], torch.arange(e - s)] = float("inf")
                out.append(torch.min(ray_u, dim=0).values)
        return torch.cat(out) if out else torch.zeros((0,), dtype=torch.float64)


class OracleEngine3D:
    """Restatement of OpticalEngine (3D) with a single StandardReaction operation."""

    def __init__(
        self,
        optical_system,
        compile_technical_intersections=False,
        compile_stopped_rays=False,
        compile_dead_rays=False,
        compile_finished_rays=True,
        compile_active_rays=True,
        dead_ray_length=None,
        new_ray_length=1.0,
        simple_ray_inheritance=("wavelength",),
        refractive_index_type="index",
    ):
        self.optical_system = optical_system
        self.compile_technical_intersections = compile_technical_intersections
        self.compile_stopped_rays = compile_stopped_rays
        self.compile_dead_rays = compile_dead_rays
        self.compile_finished_rays = compile_finished_rays
        self.compile_active_rays = compile_active_rays
        self.dead_ray_length = dead_ray_length
        self.new_ray_length = new_ray_length
        self.simple_ray_inheritance = set(simple_ray_inheritance)
        self.refractive_index_type = refractive_index_type
        self.clear_ray_history()
        self.last_projection_result = {}
        self.last_intersect_result = []  # test helper: per-iteration intersect() output
        self.last_input_rays = []        # test helper: rays entering each iteration

    # --- history (engine.py:1379-1403, 2304-2309)
    def clear_ray_history(self):
        self._active_rays = []
        self._finished_rays = []
        self._stopped_rays = []
        self._dead_rays = []

    @property
    def active_rays(self):
        return amalgamate(self._active_rays)

    @property
    def finished_rays(self):
        return amalgamate(self._finished_rays)

    @property
    def stopped_rays(self):
        return amalgamate(self._stopped_rays)

    @property
    def dead_rays(self):
        return amalgamate(self._dead_rays)

    @property
    def all_rays(self):
        return amalgamate(self._active_rays + self._finished_rays + self._dead_rays + self._stopped_rays)

    # --- engine.py:1988-2191
    def process_projection_3D(self, input_rays):
        system = self.optical_system
        result = {"rays": {}}
        input_ray_fields = list(input_rays.keys())

        self.last_input_rays.append({k: v.detach().clone() for k, v in input_rays.items()})
        intersect_result = system.intersect(input_rays)
        self.last_intersect_result.append(
            {k: (v.detach() if torch.is_tensor(v) else v) for k, v in intersect_result.items()})
        valid = intersect_result["valid"]

        if self.compile_dead_rays:
            is_dead = torch.logical_not(valid)
            dead_rays = {f: input_rays[f][is_dead] for f in input_ray_fields}
            if bool(self.dead_ray_length):
                for c in "xyz":
                    dead_rays[c + "_end"] = dead_rays[c + "_start"] + self.dead_ray_length * (
                        dead_rays[c + "_end"] - dead_rays[c + "_start"])
            self._dead_rays.append(dead_rays)
            result["rays"]["dead"] = dead_rays

        # engine.py:2052-2060
        for c in "xyz":
            input_rays[c + "_end"] = torch.where(valid, intersect_result[c], input_rays[c + "_end"])

        boundary_type = system._merged["catagory"][intersect_result["gather_trig"]]

        is_active = torch.logical_and(valid, boundary_type == OPTICAL)
        active_rays = {f: input_rays[f][is_active] for f in input_ray_fields}
        if self.compile_active_rays:
            self._active_rays.append(active_rays)
        result["rays"]["active"] = active_rays

        is_finished = torch.logical_and(valid, boundary_type == TARGET)
        if self.compile_finished_rays:
            finished_rays = {f: input_rays[f][is_finished] for f in input_ray_fields}
            self._finished_rays.append(finished_rays)
            result["rays"]["finished"] = finished_rays

        is_stopped = torch.logical_and(valid, boundary_type == STOP)
        if self.compile_stopped_rays:
            stopped_rays = {f: input_rays[f][is_stopped] for f in input_ray_fields}
            self._stopped_rays.append(stopped_rays)
            result["rays"]["stopped"] = stopped_rays

        # engine.py:2114-2133
        gather_optical = intersect_result["gather_trig"][is_active]
        optical = {field: val[gather_optical] for field, val in system._amalgamated_optical.items()}
        optical_norm = intersect_result["norm"][is_active]
        if optical_norm.shape[0] > 0:
            optical["norm"] = optical_norm
        result["optical"] = optical

        # engine.py:2137-2189
        if self.compile_technical_intersections and self.compile_stopped_rays:
            gather_stop = (intersect_result["gather_trig"] - system._optical_count)[is_stopped]
            stop = {field: val[gather_stop] for field, val in system._amalgamated_stop.items()}
            stop_norm = intersect_result["norm"][is_stopped]
            if stop_norm.shape[0] > 0:
                stop["norm"] = stop_norm
            result["stop"] = stop
        if self.compile_technical_intersections and self.compile_finished_rays:
            gather_target = (intersect_result["gather_trig"] - system._optical_count
                             - system._stop_count)[is_finished]
            target = {field: val[gather_target] for field, val in system._amalgamated_target.items()}
            finished_norm = intersect_result["norm"][is_finished]
            if finished_norm.shape[0] > 0:
                target["norm"] = finished_norm
            result["target"] = target
        return result

    # --- operation.py:255-307
    def standard_reaction_main(self, proj_result):
        if "active" not in proj_result["rays"]:
            return {}
        rays = proj_result["rays"]["active"]
        if self.refractive_index_type == "index":
            mat_in = proj_result["optical"]["mat_in"].to(torch.int64)
            mat_out = proj_result["optical"]["mat_out"].to(torch.int64)
            wavelength = rays["wavelength"]
            n_stack = torch.stack([mat["n"](wavelength) for mat in self.optical_system.materials])
            ray_range = torch.arange(rays["x_start"].shape[0], dtype=torch.int64)
            n_in = n_stack[mat_in, ray_range]
            n_out = n_stack[mat_out, ray_range]
        else:
            n_in = proj_result["optical"]["n_in"]
            n_out = proj_result["optical"]["n_out"]
        new_rays = {}
        (new_rays["x_start"], new_rays["y_start"], new_rays["z_start"],
         new_rays["x_end"], new_rays["y_end"], new_rays["z_end"]) = geometry.snells_law_3D(
            rays["x_start"], rays["y_start"], rays["z_start"],
            rays["x_end"], rays["y_end"], rays["z_end"],
            proj_result["optical"]["norm"], n_in, n_out, self.new_ray_length)
        valid = torch.ones_like(rays["x_end"], dtype=torch.bool)
        return {"active": {"rays": new_rays, "valid": valid}}

    # --- engine.py:2193-2302
    def single_pass(self, input_rays):
        self.last_projection_result = self.process_projection_3D(input_rays)
        domain = self.last_projection_result["rays"]
        for t in list(domain.keys()):
            if domain[t]["x_start"].shape[0] == 0:
                domain.pop(t)
        op_result = self.standard_reaction_main(self.last_projection_result)
        if not bool(op_result):
            return {}
        for entry_type, entry in op_result.items():
            for sig in self.simple_ray_inheritance:
                entry["rays"][sig] = self.last_projection_result["rays"][entry_type][sig]
        new_ray_list = []
        for geo_entry in op_result.values():
            v = geo_entry["valid"]
            new_ray_list.append({k: f[v] for k, f in geo_entry["rays"].items()})
        return amalgamate(new_ray_list)

    # --- engine.py:2311-2330
    def ray_trace(self, max_iterations=25):
        self.clear_ray_history()
        self.last_intersect_result = []
        self.last_input_rays = []
        starting_rays = dict(self.optical_system._amalgamated_sources)
        self.live_counts = []
        for _ in range(max_iterations):
            self.live_counts.append(int(starting_rays["x_start"].shape[0]))
            result = self.single_pass(starting_rays)
            if bool(result):
                starting_rays = result
            else:
                break
