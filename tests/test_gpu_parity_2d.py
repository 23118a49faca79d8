"""Parity of the 2D CUDA path (K4: segments + arcs, angle-form Snell, backward) against the
float64 oracle (oracle/engine2d.py).  The 2D kernels compute in fp64 on fp32 inputs, so hit
indices must be identical (ties excepted: the kernel contracts a*b+c into FMAs) and values agree
to fp32 output rounding; tolerances are the north star's 1e-5 (points) and 1e-4 (gradients)."""

import numpy as np
import pytest
import torch

import helpers
from tensorflowraytrace_b200 import kernels, scenes

pytestmark = pytest.mark.gpu

POINT_TOL = 1e-5
GRAD_TOL = 1e-4
DEV = "cuda"
RAYS = helpers.RAY_FIELDS_2D
SETS = ("active", "finished", "stopped", "dead")


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype, device=DEV)


def _count(s):
    return s["x_start"].shape[0] if bool(s) else 0


def _compare_histories(engine, oracle, fields=RAYS + ("wavelength",)):
    worst = 0.0
    for name in SETS:
        e, o = getattr(engine, name + "_rays"), getattr(oracle, name + "_rays")
        assert _count(e) == _count(o), (name, _count(e), _count(o))
        if not _count(o):
            continue
        eslot = engine.last_history_slots[name].cpu().numpy()
        assert np.array_equal(eslot, o["slot"].numpy().astype(np.int64)), f"{name}: order differs"
        for f in fields:
            worst = max(worst, helpers.rel_err(e[f].detach().cpu().numpy(), o[f].detach().numpy()))
    return worst


# --------------------------------------------------------------------------- seam
@pytest.mark.parametrize("n_rays,n_seg,n_arc", [(1, 1, 0), (7, 0, 1), (300, 5, 4), (2049, 40, 33)])
def test_intersect2d_seam(n_rays, n_seg, n_arc):
    from oracle import engine2d as oe2
    rng = np.random.default_rng(n_rays + 10 * n_seg + n_arc)
    S = rng.uniform(-0.5, 0.5, size=(n_rays, 2))
    ang = rng.uniform(-0.7, 0.7, n_rays)
    rays = helpers.f32_round(np.stack([S[:, 0], S[:, 1], S[:, 0] + np.cos(ang), S[:, 1] + np.sin(ang)]))
    c = rng.uniform(-2, 2, size=(n_seg, 2)) + np.array([4.0, 0.0])
    d = rng.normal(size=(n_seg, 2))
    seg = helpers.f32_round(np.concatenate([c - d, c + d], axis=1).T.reshape(4, n_seg))
    a1 = rng.uniform(-np.pi, np.pi, n_arc)
    arc = helpers.f32_round(np.stack([
        rng.uniform(2, 7, n_arc), rng.uniform(-2, 2, n_arc), a1,
        np.clip(a1 + rng.uniform(0.5, 4.0, n_arc), None, None), rng.uniform(0.5, 2.0, n_arc)
        * rng.choice([-1.0, 1.0], n_arc)]).reshape(5, n_arc))
    if n_arc:   # angle_end wrapped into [-pi, pi] like the reference expects
        arc[3] = helpers.f32_round(np.mod(arc[3] + np.pi, 2 * np.pi) - np.pi)
    eps = helpers.EPS
    got_seg, got_arc = kernels.intersect2d(
        dev(rays), dev(seg) if n_seg else None, dev(arc) if n_arc else None,
        eps["intersect_epsilion"], eps["size_epsilion"], eps["ray_start_epsilion"])
    osys = oe2.OracleSystem2D(
        {k: torch.from_numpy(rays[i]) for i, k in enumerate(RAYS)},
        {k: torch.from_numpy(seg[i]) for i, k in enumerate(helpers.SEG_FIELDS)} if n_seg else None,
        {k: torch.from_numpy(arc[i]) for i, k in enumerate(helpers.ARC_FIELDS)} if n_arc else None,
        **eps)
    want_seg, want_arc = osys.intersect(osys._amalgamated_sources)
    for got, want, gkey, ukey in ((got_seg, want_seg, "gather_segment", "segment_u"),
                                  (got_arc, want_arc, "gather_arc", "arc_u")):
        assert bool(got) == bool(want)
        if not want:
            continue
        ov = want["valid"].numpy()
        assert np.array_equal(got["valid"].cpu().numpy(), ov)
        assert np.array_equal(got[gkey].cpu().numpy()[ov], want[gkey].numpy()[ov])
        for k in ("x", "y", "ray_u", ukey, "norm"):
            if ov.any():
                assert helpers.rel_err(got[k].cpu().numpy()[ov], want[k].numpy()[ov]) < POINT_TOL, k


# --------------------------------------------------------------------------- whole trace
@pytest.mark.parametrize("factory", ["config1", "arc_lens_2d", "mixed_2d"])
def test_trace2d_forward_parity(factory):
    scene = getattr(scenes, factory)()
    kw = dict(compile_dead_rays=True, compile_stopped_rays=True)
    engine, _, _ = scenes.build_engine_2d(scene, eps=helpers.EPS, **kw)
    engine.ray_trace(scene["max_iterations"])
    oracle, _ = helpers.oracle2d_from_scene(scene, **kw)
    oracle.ray_trace(scene["max_iterations"])
    live = engine.last_trace.live_counts()
    assert live[: len(oracle.live_counts)] == oracle.live_counts
    assert _compare_histories(engine, oracle) < POINT_TOL
    if factory == "config1":
        assert _count(engine.finished_rays) == 1000 and live == [1000, 1000, 1000]


def test_trace2d_dead_ray_length_and_flags():
    scene = scenes.mixed_2d(ray_count=500)
    kw = dict(compile_dead_rays=True, dead_ray_length=2.5, compile_active_rays=False)
    engine, _, _ = scenes.build_engine_2d(scene, eps=helpers.EPS, **kw)
    engine.ray_trace(4)
    oracle, _ = helpers.oracle2d_from_scene(scene, **kw)
    oracle.ray_trace(4)
    assert not bool(engine.active_rays) and not bool(engine.stopped_rays)
    for name in ("finished", "dead"):
        e, o = getattr(engine, name + "_rays"), getattr(oracle, name + "_rays")
        assert _count(e) == _count(o) > 0
        for f in RAYS:
            assert helpers.rel_err(e[f].cpu().numpy(), o[f].numpy()) < POINT_TOL


@pytest.mark.parametrize("factory", ["config1", "arc_lens_2d", "mixed_2d"])
def test_trace2d_gradient_parity(factory):
    """d(loss)/d(segment end points, arc centres and radii) vs oracle autograd; the loss mixes
    finished-ray positions and active-ray history points so every VJP branch is exercised."""
    scene = getattr(scenes, factory)()
    kw = dict(compile_stopped_rays=True)
    engine, params, _ = scenes.build_engine_2d(scene, eps=helpers.EPS, parametric=True, **kw)
    engine.ray_trace(scene["max_iterations"])
    oracle, oparams = helpers.oracle2d_from_scene(scene, requires_grad=True, **kw)
    oracle.ray_trace(scene["max_iterations"])

    def loss(fin, act):
        return ((fin["y_end"] - 0.3) ** 2).sum() + 0.1 * (fin["x_end"] * fin["y_start"]).sum() \
            + 0.05 * (act["x_end"] ** 2 + act["y_end"] * act["x_start"]).sum()

    grads = torch.autograd.grad(loss(engine.finished_rays, engine.active_rays), params,
                                allow_unused=True)
    ograds = torch.autograd.grad(loss(oracle.finished_rays, oracle.active_rays), oparams,
                                 allow_unused=True)
    assert len(grads) == len(ograds) > 0
    # angle_start / angle_end only enter comparisons: no gradient on either side
    scale = max(float(g.abs().max()) for g in ograds if g is not None)
    checked = 0
    for g, og in zip(grads, ograds):
        if og is None:
            assert g is None or float(g.abs().max()) == 0.0
            continue
        g = torch.zeros_like(og) if g is None else g.cpu().double()
        assert float((g - og).abs().max()) <= GRAD_TOL * scale
        checked += 1
    assert checked >= 3


def test_trace2d_source_gradient():
    """Cotangents reach the source rays (start and end) as in the reference's tape."""
    scene = scenes.config1(ray_count=64)
    engine, _, src = scenes.build_engine_2d(scene, eps=helpers.EPS)
    leaves = {}
    for k in RAYS:
        leaves[k] = src[k].clone().requires_grad_(True)
        src[k] = leaves[k]
    engine.optical_system.update()
    engine.ray_trace(3)
    fin = engine.finished_rays
    (fin["y_end"] ** 2).sum().backward()
    oracle, _ = helpers.oracle2d_from_scene(scene)
    osrc = oracle.optical_system._amalgamated_sources
    for k in RAYS:
        osrc[k] = osrc[k].clone().requires_grad_(True)
    oracle.ray_trace(3)
    (oracle.finished_rays["y_end"] ** 2).sum().backward()
    for k in RAYS:
        assert helpers.rel_err(leaves[k].grad.cpu().numpy(), osrc[k].grad.numpy()) < GRAD_TOL, k


def test_unfused_2d_matches_fused():
    """A custom RayOperation forces the projection-at-a-time path (intersect2d kernel + torch
    bookkeeping): same histories as the fused loop, values and gradients."""
    from tensorflowraytrace_b200 import engine as eng, operation

    class Tag(operation.RayOperation):
        pass

    scene = scenes.mixed_2d(ray_count=800)
    kw = dict(compile_dead_rays=True, compile_stopped_rays=True)
    fused, params, _ = scenes.build_engine_2d(scene, eps=helpers.EPS, parametric=True, **kw)
    fused.ray_trace(5)
    unf = eng.OpticalEngine(2, [operation.StandardReaction(), Tag()], **kw)
    unf.optical_system = fused.optical_system
    unf.ray_trace(5)
    assert fused._fusable() and not unf._fusable()
    for name in SETS:
        a, b = getattr(fused, name + "_rays"), getattr(unf, name + "_rays")
        assert _count(a) == _count(b), name
        for f in RAYS + ("wavelength",):
            if _count(a):
                assert helpers.rel_err(b[f].detach().cpu().numpy(), a[f].detach().cpu().numpy()) < 1e-4, (name, f)
    gf = torch.autograd.grad((fused.finished_rays["y_end"] ** 2).sum(), params, allow_unused=True)
    gu = torch.autograd.grad((unf.finished_rays["y_end"] ** 2).sum(), params, allow_unused=True)
    scale = max(float(g.abs().max()) for g in gf if g is not None)
    for a, b in zip(gf, gu):
        if a is None:
            continue
        b = torch.zeros_like(a) if b is None else b
        assert float((a - b).abs().max()) < 5e-3 * scale    # fp32 torch glue vs fp64 kernel


# --------------------------------------------------------------------------- 2D at scale
@pytest.mark.parametrize("beams,points", [(10, 11), (10, 1001), (2000, 1001), (10000, 11)])
def test_2d_at_scale_matches_the_oracle(beams, points):
    """the sizes of the reference's own scaling script
    (dev/performance.py:77-86: 10 / 10 000 beam points x 6 wavelengths against 10 / 1 000
    segments, 2 iterations).  Few rays over many segments are scanned by 32 threads per ray, the
    boundaries staged in shared memory; whatever the thread layout, the nearest hit (lowest index
    on ties) and the refracted rays must be the oracle's."""
    scene = scenes.performance_2d(beams, points)
    engine, _, _ = scenes.build_engine_2d(scene, eps=helpers.EPS, compile_dead_rays=True)
    engine.ray_trace(2)
    oracle, _ = helpers.oracle2d_from_scene(scene, compile_dead_rays=True)
    oracle.ray_trace(2)
    n = beams * 6
    assert engine.last_trace.live_counts()[0] == n == oracle.live_counts[0]
    assert engine.last_trace.live_counts() == oracle.live_counts
    for name in ("active", "finished", "dead"):
        e, o = getattr(engine, name + "_rays"), getattr(oracle, name + "_rays")
        n_o = o["x_start"].shape[0] if bool(o) else 0
        n_e = e["x_start"].shape[0] if bool(e) else 0
        assert n_e == n_o, name
        if n_o == 0:
            continue
        scale = max(float(np.max(np.abs(o[f].numpy()))) for f in helpers.RAY_FIELDS_2D)
        for f in helpers.RAY_FIELDS_2D + ("wavelength",):
            assert helpers.rel_err(e[f].detach().cpu().numpy(), o[f].numpy(), scale=scale) < 1e-5, (name, f)
    assert (engine.finished_rays["x_end"] - 10.0).abs().max() < 1e-4
