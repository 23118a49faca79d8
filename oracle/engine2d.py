"""
Oracle restatement of the 2D hot path of tfrt/engine.py + tfrt/operation.py (float64 torch CPU).
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

What is restated (reference file:line):
  OpticalSystem2D._merge_boundaries     tfrt/engine.py:425-521
  OpticalSystem2D.intersect             tfrt/engine.py:523-616   (segment norm :580-586)
  OpticalSystem2D._seg_or_arc           tfrt/engine.py:626-657
  OpticalSystem2D._get_arc_norm         tfrt/engine.py:666-670
  OpticalSystem2D._segment_intersection tfrt/engine.py:687-749
  OpticalSystem2D._arc_intersection     tfrt/engine.py:767-866
  OpticalEngine.process_projection_2D   tfrt/engine.py:1544-1986
  OpticalEngine.single_pass / ray_trace tfrt/engine.py:2193-2330
  StandardReaction.main (2D branch)     tfrt/operation.py:255-307

One documented departure: when a system holds BOTH segments and arcs, the reference builds the
reacting ray set seg-then-arc but the matching optical-boundary set arc-then-seg
(tfrt/engine.py:1958-1965), so rays are paired with the wrong boundaries whenever both kinds are
hit in one pass.  That is a latent bug, not a semantic to preserve (hard
part 9): here, and in the CUDA path, every ray reacts with the boundary it actually hit, and the
ray order stays the reference's (segment hits, then arc hits).  Systems with only segments or
only arcs -- every 2D script of the reference that runs -- are restated exactly.
"""

import math

import torch

from . import geometry
from .engine import OPTICAL, STOP, TARGET, amalgamate

PI = math.pi
SEGMENT_GEO_SIG = ("x_start", "y_start", "x_end", "y_end")
ARC_GEO_SIG = ("x_center", "y_center", "angle_start", "angle_end", "radius")


def _label(d, cat, key):
    if bool(d):
        d["catagory"] = cat * torch.ones_like(d[key], dtype=torch.int64)
        return d[key].shape[0]
    return 0


class OracleSystem2D:
    """Field-dict restatement of OpticalSystem2D (tfrt/engine.py:253-866)."""

    def __init__(self, sources, optical_segments=None, optical_arcs=None, stop_segments=None,
                 stop_arcs=None, target_segments=None, target_arcs=None, materials=None,
                 intersect_epsilion=1e-10, size_epsilion=1e-10, ray_start_epsilion=1e-10):
        self._inputs = dict(sources=sources, optical_segments=optical_segments,
                            optical_arcs=optical_arcs, stop_segments=stop_segments,
                            stop_arcs=stop_arcs, target_segments=target_segments,
                            target_arcs=target_arcs)
        self.materials = materials or []
        self.intersect_epsilion = intersect_epsilion
        self.size_epsilion = size_epsilion
        self.ray_start_epsilion = ray_start_epsilion
        self.update()

    def update(self):
        """OpticalSystem2D._update (tfrt/engine.py:303-321): re-amalgamate and re-merge."""
        def am(x):
            return amalgamate(x) if isinstance(x, list) else dict(x or {})

        for name, value in self._inputs.items():
            setattr(self, "_amalgamated_" + name, am(value))
        self._merge_boundaries()

    def _merge_boundaries(self):
        """tfrt/engine.py:425-521"""
        self._optical_seg_count = _label(self._amalgamated_optical_segments, OPTICAL, "x_start")
        self._stop_seg_count = _label(self._amalgamated_stop_segments, STOP, "x_start")
        self._target_seg_count = _label(self._amalgamated_target_segments, TARGET, "x_start")
        self._merged_segments = amalgamate(
            [self._amalgamated_optical_segments, self._amalgamated_stop_segments,
             self._amalgamated_target_segments], set(SEGMENT_GEO_SIG) | {"catagory"})
        self._optical_arc_count = _label(self._amalgamated_optical_arcs, OPTICAL, "x_center")
        self._stop_arc_count = _label(self._amalgamated_stop_arcs, STOP, "x_center")
        self._target_arc_count = _label(self._amalgamated_target_arcs, TARGET, "x_center")
        self._merged_arcs = amalgamate(
            [self._amalgamated_optical_arcs, self._amalgamated_stop_arcs,
             self._amalgamated_target_arcs], set(ARC_GEO_SIG) | {"catagory"})

    def intersect(self, rays):
        """tfrt/engine.py:523-616"""
        has_segments = bool(self._merged_segments)
        has_arcs = bool(self._merged_arcs)
        seg, arc = {}, {}
        if has_segments:
            m = self._merged_segments
            (seg["x"], seg["y"], seg["valid"], seg["ray_u"], seg["segment_u"], seg["gather_ray"],
             seg["gather_segment"]) = self._segment_intersection(
                rays["x_start"], rays["y_start"], rays["x_end"], rays["y_end"],
                m["x_start"], m["y_start"], m["x_end"], m["y_end"],
                self.intersect_epsilion, self.size_epsilion, self.ray_start_epsilion)
            seg["norm"] = (torch.atan2(m["y_end"] - m["y_start"], m["x_end"] - m["x_start"])
                           + PI / 2.0)[seg["gather_segment"]]
        if has_arcs:
            m = self._merged_arcs
            (arc["x"], arc["y"], arc["valid"], arc["ray_u"], arc["arc_u"], arc["gather_ray"],
             arc["gather_arc"]) = self._arc_intersection(
                rays["x_start"], rays["y_start"], rays["x_end"], rays["y_end"],
                m["x_center"], m["y_center"], m["angle_start"], m["angle_end"], m["radius"],
                self.intersect_epsilion, self.size_epsilion, self.ray_start_epsilion)
            arc["norm"] = self._get_arc_norm(m["radius"], arc["arc_u"], arc["gather_arc"])
        if has_segments and has_arcs:
            seg["valid"], arc["valid"] = self._seg_or_arc(
                seg["ray_u"], arc["ray_u"], seg["valid"], arc["valid"])
        return seg, arc

    @staticmethod
    def _seg_or_arc(seg_u, arc_u, seg_valid, arc_valid):
        """tfrt/engine.py:626-657 (the live variant: ties go to the arc)."""
        has_both = torch.logical_and(seg_valid, arc_valid)
        seg_less = seg_u < arc_u
        seg_valid = torch.where(has_both, seg_less, seg_valid)
        arc_valid = torch.where(has_both, torch.logical_not(seg_less), arc_valid)
        return seg_valid, arc_valid

    @staticmethod
    def _get_arc_norm(radius, arc_u, gather_arc):
        """tfrt/engine.py:666-670"""
        radius = radius[gather_arc]
        arc_norm = torch.where(radius < 0, arc_u + PI, arc_u)
        return geometry.tf_mod(arc_norm + PI, 2 * PI) - PI

    @staticmethod
    def _segment_intersection(rx1, ry1, rx2, ry2, sx1, sy1, sx2, sy2, intersect_epsilion,
                              size_epsilion, ray_start_epsilion):
        """tfrt/engine.py:687-749.  The dense (M, N) pass decides validity and argmin (not
        differentiable in the reference either); the gathered entries are re-evaluated 1-to-1
        with the same formula so autograd sees the reference's differentiable path."""
        n = rx1.shape[0]
        with torch.no_grad():
            _, _, valid, ray_u, seg_u = geometry.line_intersect(
                rx1, ry1, rx2, ry2, sx1, sy1, sx2, sy2, intersect_epsilion)
            valid = torch.logical_and(valid, seg_u >= -size_epsilion)
            valid = torch.logical_and(valid, seg_u <= 1 + size_epsilion)
            valid = torch.logical_and(valid, ray_u >= ray_start_epsilion)
            if ray_u.numel():
                inf = 2 * torch.max(ray_u) * torch.ones_like(ray_u)
                ray_u = torch.where(valid, ray_u, inf)
                closest = torch.argmin(ray_u, dim=0)
                valid = torch.any(valid, dim=0)
            else:
                closest = torch.zeros((n,), dtype=torch.int64)
                valid = torch.zeros((n,), dtype=torch.bool)
        gather_ray = torch.arange(n, dtype=torch.int64)
        x, y, _, ray_u, seg_u = geometry.raw_line_intersect(
            rx1, ry1, rx2, ry2, sx1[closest], sy1[closest], sx2[closest], sy2[closest],
            intersect_epsilion)
        return x, y, valid, ray_u, seg_u, gather_ray, closest

    @staticmethod
    def _arc_intersection(rx1, ry1, rx2, ry2, xc, yc, a1, a2, r, intersect_epsilion,
                          size_epsilion, ray_start_epsilion):
        """tfrt/engine.py:767-866"""
        n = rx1.shape[0]

        def solve(xc_, yc_, a1_, a2_, r_, dense):
            if dense:
                plus, minus = geometry.line_circle_intersect(
                    rx1, ry1, rx2, ry2, xc_, yc_, r_, intersect_epsilion)
                a1_, a2_ = a1_.reshape(-1, 1), a2_.reshape(-1, 1)
            else:
                plus, minus = geometry.raw_line_circle_intersect(
                    rx1, ry1, rx2, ry2, xc_, yc_, r_, intersect_epsilion)
            for d in (plus, minus):
                d["valid"] = torch.logical_and(d["valid"], d["u"] >= ray_start_epsilion)
                d["valid"] = torch.logical_and(
                    d["valid"], geometry.angle_in_interval(d["v"], a1_, a2_))
            # engine.py:831-834: `inf` only has to exceed every valid u
            inf = torch.full_like(plus["u"], float("inf"))
            pu = torch.where(plus["valid"], plus["u"], inf)
            mu = torch.where(minus["valid"], minus["u"], inf)
            choose_minus = mu < pu
            valid = torch.logical_or(torch.logical_and(minus["valid"], choose_minus),
                                     torch.logical_and(plus["valid"],
                                                       torch.logical_not(choose_minus)))
            x = torch.where(choose_minus, minus["x"], plus["x"])
            y = torch.where(choose_minus, minus["y"], plus["y"])
            ray_u = torch.where(choose_minus, mu, pu)
            arc_u = torch.where(choose_minus, minus["v"], plus["v"])
            return x, y, valid, ray_u, arc_u

        with torch.no_grad():
            _, _, valid, ray_u, _ = solve(xc, yc, a1, a2, r, True)
            if ray_u.numel():
                closest = torch.argmin(ray_u, dim=0)
                valid = torch.any(valid, dim=0)
            else:
                closest = torch.zeros((n,), dtype=torch.int64)
                valid = torch.zeros((n,), dtype=torch.bool)
        gather_ray = torch.arange(n, dtype=torch.int64)
        x, y, _, ray_u, arc_u = solve(xc[closest], yc[closest], a1[closest], a2[closest],
                                      r[closest], False)
        return x, y, valid, ray_u, arc_u, gather_ray, closest


class OracleEngine2D:
    """Restatement of OpticalEngine (2D) with a single StandardReaction operation."""

    def __init__(self, optical_system, compile_stopped_rays=False, compile_dead_rays=False,
                 compile_finished_rays=True, compile_active_rays=True, dead_ray_length=None,
                 new_ray_length=1.0, simple_ray_inheritance=("wavelength",),
                 refractive_index_type="index"):
        self.optical_system = optical_system
        self.compile_stopped_rays = compile_stopped_rays
        self.compile_dead_rays = compile_dead_rays
        self.compile_finished_rays = compile_finished_rays
        self.compile_active_rays = compile_active_rays
        self.dead_ray_length = dead_ray_length
        self.new_ray_length = new_ray_length
        self.simple_ray_inheritance = set(simple_ray_inheritance)
        self.refractive_index_type = refractive_index_type
        self.clear_ray_history()
        self.last_projection = []   # test helper: per iteration (seg_proj, arc_proj, input rays)

    def clear_ray_history(self):
        self._active_rays = []
        self._finished_rays = []
        self._stopped_rays = []
        self._dead_rays = []

    active_rays = property(lambda self: amalgamate(self._active_rays))
    finished_rays = property(lambda self: amalgamate(self._finished_rays))
    stopped_rays = property(lambda self: amalgamate(self._stopped_rays))
    dead_rays = property(lambda self: amalgamate(self._dead_rays))

    def process_projection_2D(self, input_rays):
        """tfrt/engine.py:1544-1986 (the parts a StandardReaction trace uses)."""
        system = self.optical_system
        fields = list(input_rays.keys())
        has_segments = bool(system._merged_segments)
        has_arcs = bool(system._merged_arcs)
        seg_proj, arc_proj = system.intersect(input_rays)
        self.last_projection.append((
            {k: v.detach() for k, v in seg_proj.items()},
            {k: v.detach() for k, v in arc_proj.items()},
            {k: v.detach().clone() for k, v in input_rays.items()}))

        if self.compile_dead_rays:
            if has_segments and has_arcs:
                is_dead = torch.logical_not(torch.logical_or(seg_proj["valid"], arc_proj["valid"]))
            elif has_segments:
                is_dead = torch.logical_not(seg_proj["valid"])
            else:
                is_dead = torch.logical_not(arc_proj["valid"])
            dead = {f: input_rays[f][is_dead] for f in fields}
            if bool(self.dead_ray_length):
                for c in "xy":
                    dead[c + "_end"] = dead[c + "_start"] + self.dead_ray_length * (
                        dead[c + "_end"] - dead[c + "_start"])
            self._dead_rays.append(dead)

        parts = []   # (active rays, optical boundary data) per geometry, seg first
        for proj, merged, optical, gather_key in (
                (seg_proj, system._merged_segments, system._amalgamated_optical_segments,
                 "gather_segment"),
                (arc_proj, system._merged_arcs, system._amalgamated_optical_arcs, "gather_arc")):
            if not bool(merged):
                continue
            input_rays["x_end"] = torch.where(proj["valid"], proj["x"], input_rays["x_end"])
            input_rays["y_end"] = torch.where(proj["valid"], proj["y"], input_rays["y_end"])
            boundary_type = merged["catagory"][proj[gather_key]]
            is_active = torch.logical_and(proj["valid"], boundary_type == OPTICAL)
            active = {f: input_rays[f][is_active] for f in fields}
            if self.compile_active_rays:
                self._active_rays.append(active)
            if self.compile_finished_rays:
                m = torch.logical_and(proj["valid"], boundary_type == TARGET)
                self._finished_rays.append({f: input_rays[f][m] for f in fields})
            if self.compile_stopped_rays:
                m = torch.logical_and(proj["valid"], boundary_type == STOP)
                self._stopped_rays.append({f: input_rays[f][m] for f in fields})
            g = proj[gather_key][is_active]
            opt = {f: v[g] for f, v in optical.items()}
            opt["norm"] = proj["norm"][is_active]
            parts.append((active, opt))
        return parts

    def standard_reaction_main(self, rays, optical):
        """tfrt/operation.py:255-307, dimension == 2."""
        if self.refractive_index_type == "index":
            mat_in = optical["mat_in"].to(torch.int64)
            mat_out = optical["mat_out"].to(torch.int64)
            wavelength = rays["wavelength"]
            n_stack = torch.stack([mat["n"](wavelength) for mat in self.optical_system.materials])
            ray_range = torch.arange(rays["x_start"].shape[0], dtype=torch.int64)
            n_in = n_stack[mat_in, ray_range]
            n_out = n_stack[mat_out, ray_range]
        else:
            n_in, n_out = optical["n_in"], optical["n_out"]
        new = {}
        new["x_start"], new["y_start"], new["x_end"], new["y_end"] = geometry.snells_law_2D(
            rays["x_start"], rays["y_start"], rays["x_end"], rays["y_end"], optical["norm"],
            n_in, n_out, self.new_ray_length)
        return new

    def single_pass(self, input_rays):
        """tfrt/engine.py:2193-2302.  Rays react per geometry with their own boundaries and are
        concatenated seg-then-arc (module docstring)."""
        new_sets = []
        for active, optical in self.process_projection_2D(input_rays):
            if active["x_start"].shape[0] == 0:
                continue
            new = self.standard_reaction_main(active, optical)
            for sig in self.simple_ray_inheritance:
                new[sig] = active[sig]
            new_sets.append(new)
        return amalgamate(new_sets)

    def ray_trace(self, max_iterations=25):
        """tfrt/engine.py:2311-2330"""
        self.clear_ray_history()
        self.last_projection = []
        self.live_counts = []
        starting_rays = dict(self.optical_system._amalgamated_sources)
        for _ in range(max_iterations):
            self.live_counts.append(int(starting_rays["x_start"].shape[0]))
            result = self.single_pass(starting_rays)
            if bool(result):
                starting_rays = result
            else:
                break
