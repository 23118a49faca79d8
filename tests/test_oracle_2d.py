"""Pins for the 2D oracle (oracle/engine2d.py): analytic known answers, the reference's
tie / ordering rules, and finite-difference gradients.  CPU only."""

import math

import numpy as np
import torch

import helpers
from oracle import engine2d as oe2
from oracle import geometry as og
from oracle import materials as om
from tensorflowraytrace_b200 import scenes

F64 = torch.float64
EPS0 = dict(intersect_epsilion=1e-10, size_epsilion=1e-10, ray_start_epsilion=1e-10)


def t(*v):
    return torch.tensor(v, dtype=F64)


def _src(xs, ys, ang, wl=575.0):
    n = len(xs)
    return {"x_start": t(*xs), "y_start": t(*ys),
            "x_end": t(*[x + math.cos(a) for x, a in zip(xs, ang)]),
            "y_end": t(*[y + math.sin(a) for y, a in zip(ys, ang)]),
            "wavelength": torch.full((n,), wl, dtype=F64)}


def _seg(rows, mat_in=None, mat_out=None):
    a = np.asarray(rows, dtype=np.float64).reshape(-1, 4)
    d = {k: torch.from_numpy(np.ascontiguousarray(a[:, i])) for i, k in enumerate(helpers.SEG_FIELDS)}
    if mat_in is not None:
        d["mat_in"] = torch.full((a.shape[0],), mat_in, dtype=torch.int64)
        d["mat_out"] = torch.full((a.shape[0],), mat_out, dtype=torch.int64)
    return d


def test_slab_refraction_matches_snell_analytically():
    """A ray crossing a parallel acrylic slab leaves parallel to itself, displaced by
    d sin(i - r)/cos r; inside, n sin r = sin i (operation.py:255-307 + geometry.py:565-653)."""
    inc = 0.4
    n = float(om.acrylic(t(575.0)))
    src = _src([-1.0], [0.0], [inc])
    slab = _seg([[0.0, -5.0, 0.0, 5.0], [1.0, 5.0, 1.0, -5.0]], 1, 0)   # norms -x and +x
    target = _seg([[3.0, -9.0, 3.0, 9.0]])
    system = oe2.OracleSystem2D([src], [slab], None, None, None, [target], None,
                                materials=[{"n": om.vacuum}, {"n": om.acrylic}], **EPS0)
    eng = oe2.OracleEngine2D(system)
    eng.ray_trace(4)
    act, fin = eng.active_rays, eng.finished_rays
    assert act["x_start"].shape[0] == 2 and fin["x_start"].shape[0] == 1
    r = math.asin(math.sin(inc) / n)
    # inside the slab (second active ray): direction angle r
    inside = math.atan2(float(act["y_end"][1] - act["y_start"][1]),
                        float(act["x_end"][1] - act["x_start"][1]))
    assert abs(inside - r) < 1e-12
    out = math.atan2(float(fin["y_end"][0] - fin["y_start"][0]),
                     float(fin["x_end"][0] - fin["x_start"][0]))
    assert abs(out - inc) < 1e-12
    y_exit = math.tan(inc) * 1.0 + math.tan(r) * 1.0
    assert abs(float(fin["y_start"][0]) - y_exit) < 1e-12
    assert abs(float(fin["y_end"][0]) - (y_exit + 2.0 * math.tan(inc))) < 1e-12


def test_normal_incidence_on_arc_is_undeviated_and_norm_flips_with_radius():
    for radius in (2.0, -2.0):
        src = _src([-5.0], [0.0], [0.0])
        arc = {"x_center": t(0.0), "y_center": t(0.0), "angle_start": t(2.0), "angle_end": t(-2.0),
               "radius": t(radius), "mat_in": torch.tensor([1]), "mat_out": torch.tensor([0])}
        system = oe2.OracleSystem2D([src], None, [arc], materials=[{"n": om.vacuum}, {"n": om.acrylic}], **EPS0)
        _, proj = system.intersect(system._amalgamated_sources)
        assert bool(proj["valid"][0]) and abs(float(proj["x"][0]) + 2.0) < 1e-12
        # hit at polar angle pi: outward normal for r > 0, flipped (0) for r < 0 (engine.py:666-670)
        want = -math.pi if radius > 0 else 0.0
        assert abs(float(proj["norm"][0]) - want) < 1e-12 or abs(abs(float(proj["norm"][0])) - math.pi) < 1e-12
        eng = oe2.OracleEngine2D(system)
        eng.ray_trace(1)
        a = eng.active_rays
        assert abs(float(a["y_end"][0])) < 1e-12 and abs(float(a["x_end"][0]) + 2.0) < 1e-12


def test_arc_angle_window_and_nearer_root():
    # the ray crosses the full circle twice; only the far side lies inside the angle window
    src = _src([-5.0], [0.1], [0.0])
    arc = {"x_center": t(0.0), "y_center": t(0.0), "angle_start": t(-0.5), "angle_end": t(0.5),
           "radius": t(1.0)}
    system = oe2.OracleSystem2D([src], None, None, None, None, None, [arc], **EPS0)
    _, proj = system.intersect(system._amalgamated_sources)
    assert bool(proj["valid"][0]) and float(proj["x"][0]) > 0.9
    arc["angle_start"], arc["angle_end"] = t(2.5), t(-2.5)      # wrapped window on the near side
    system = oe2.OracleSystem2D([src], None, None, None, None, None, [arc], **EPS0)
    _, proj = system.intersect(system._amalgamated_sources)
    assert bool(proj["valid"][0]) and float(proj["x"][0]) < -0.9


def test_seg_or_arc_picks_the_nearer_and_ties_go_to_the_arc():
    f = oe2.OracleSystem2D._seg_or_arc
    sv, av = f(t(1.0, 2.0, 3.0, 1.0), t(2.0, 1.0, 3.0, 5.0),
               torch.tensor([True, True, True, True]), torch.tensor([True, True, True, False]))
    assert sv.tolist() == [True, False, False, True]
    assert av.tolist() == [False, True, True, False]


def test_history_order_is_segments_then_arcs_per_iteration():
    scene = scenes.mixed_2d(ray_count=300)
    eng, _ = helpers.oracle2d_from_scene(scene, compile_stopped_rays=True, compile_dead_rays=True)
    eng.ray_trace(3)
    seg_proj, arc_proj, rays = eng.last_projection[0]
    # first iteration: active set = [segment hits in source order] + [arc hits in source order]
    n0 = int((seg_proj["valid"] & (seg_proj["gather_segment"] < 2)).sum())
    a0 = int((arc_proj["valid"] & (arc_proj["gather_arc"] < 2)).sum())
    first = eng._active_rays[0]["slot"], eng._active_rays[1]["slot"]
    assert first[0].shape[0] == n0 and first[1].shape[0] == a0
    assert torch.all(first[0][1:] > first[0][:-1]) and torch.all(first[1][1:] > first[1][:-1])
    assert eng.live_counts[1] == n0 + a0


def test_oracle_2d_gradient_matches_finite_differences():
    scene = scenes.mixed_2d(ray_count=120)
    eng, params = helpers.oracle2d_from_scene(scene, requires_grad=True, eps=EPS0)

    def loss():
        eng.optical_system.update()
        eng.ray_trace(4)
        fin = eng.finished_rays
        return ((fin["y_end"] - 0.2) ** 2).sum() + (fin["x_start"] * fin["y_start"]).sum()

    base = loss()
    grads = torch.autograd.grad(base, params, allow_unused=True)
    checked = 0
    h = 1e-6
    for p, g in zip(params, grads):
        if g is None:
            continue
        for i in range(p.shape[0]):
            with torch.no_grad():
                p[i] += h
                up = float(loss())
                p[i] -= 2 * h
                dn = float(loss())
                p[i] += h
            fd = (up - dn) / (2 * h)
            assert abs(fd - float(g[i])) <= 1e-5 * max(1.0, abs(fd)), (fd, float(g[i]))
            checked += 1
    assert checked >= 10


def test_snell_2d_total_internal_reflection_and_mirror_rule():
    # internal hit beyond the critical angle reflects: angle of reflection = angle of incidence
    n = 1.5
    inc = 1.0   # rad > asin(1/1.5)
    xs, ys = t(-math.sin(inc)), t(-math.cos(inc))
    out = og.snells_law_2D(xs, ys, t(0.0), t(0.0), t(math.pi / 2) * 1.0, t(n), t(1.0), 1.0)
    # surface along x with normal +y; ray travels from inside (below) upward
    d = (float(out[2] - out[0]), float(out[3] - out[1]))
    assert abs(d[0] - math.sin(inc)) < 1e-12 and abs(d[1] + math.cos(inc)) < 1e-12
    # n_in = 0: mirror only for rays arriving from the normal side
    out = og.snells_law_2D(t(-0.3), t(1.0), t(0.0), t(0.0), t(math.pi / 2), t(0.0), t(1.0), 1.0)
    assert float(out[3]) > 0.0      # reflected back up
