"""Edge cases of the trace path on the GPU, each checked against the float64 oracle: empty ray
sets, rays that all miss, systems without optical boundaries, a single iteration, rays that start
on a boundary, triangle counts around the tile size, and the largest mesh of BASELINE.json
(config 5's 99 374 triangles) with a ray subsample."""

import numpy as np
import pytest
import torch

import helpers
from tensorflowraytrace_b200 import mesh_tools, scenes

pytestmark = pytest.mark.gpu
RAYS = helpers.RAY_FIELDS


def _count(s):
    return s["x_start"].shape[0] if bool(s) else 0


def _sub(scene, sl):
    out = dict(scene)
    out["sources"] = {k: v[sl] for k, v in scene["sources"].items()}
    return out


def test_empty_ray_set_3d_and_2d():
    scene = _sub(scenes.config2(ray_count=64, step_count=3), slice(0, 0))
    engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_dead_rays=True)
    engine.ray_trace(3)
    for name in ("active", "finished", "dead"):
        assert _count(getattr(engine, name + "_rays")) == 0
    assert engine.last_trace.live_counts() == [0, 0, 0]
    scene2 = _sub(scenes.config1(ray_count=16), slice(0, 0))
    engine2, _, _ = scenes.build_engine_2d(scene2, eps=helpers.EPS, compile_dead_rays=True)
    engine2.ray_trace(3)
    for name in ("active", "finished", "dead"):
        assert _count(getattr(engine2, name + "_rays")) == 0


def test_all_rays_miss_and_dead_rays_are_the_inputs():
    scene = scenes.config2(ray_count=500, step_count=3)
    for k in ("x_end",):   # send every ray backwards, away from the lens
        scene["sources"][k] = 2 * scene["sources"]["x_start"] - scene["sources"][k]
    engine, _, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_dead_rays=True)
    engine.ray_trace(4)
    assert _count(engine.finished_rays) == 0 and _count(engine.active_rays) == 0
    dead = engine.dead_rays
    assert _count(dead) == 500
    for f in RAYS:   # unchanged rays, in source order
        want = helpers.f32_round(scene["sources"][f])
        assert helpers.rel_err(dead[f].detach().cpu().numpy(), want) < 1e-6, f
    assert engine.last_trace.live_counts()[:2] == [500, 0]


def test_target_only_system_and_single_iteration():
    scene = scenes.config2(ray_count=300, step_count=3)
    scene["surfaces"] = []
    engine, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    engine.ray_trace(1)
    oracle, _ = helpers.oracle_from_scene(scene)
    oracle.ray_trace(1)
    assert _count(engine.finished_rays) == _count(oracle.finished_rays) == 300
    assert _count(engine.active_rays) == 0
    for f in RAYS:
        assert helpers.rel_err(engine.finished_rays[f].detach().cpu().numpy(),
                               oracle.finished_rays[f].numpy()) < 1e-5
    # single iteration on the lens: every ray is still active, nothing finished
    scene = scenes.config2(ray_count=300, step_count=3)
    engine, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    engine.ray_trace(1)
    assert _count(engine.active_rays) == 300 and _count(engine.finished_rays) == 0


def test_rays_starting_on_a_boundary_do_not_rehit_it():
    """Rays launched from points ON the first lens surface (the state every refracted ray is
    in): the start triangle and its edge neighbours must be skipped via ray_start_epsilion."""
    scene = scenes.config2(ray_count=2000, step_count=5)
    engine, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    engine.ray_trace(3)
    act = engine.active_rays
    it = engine.last_history_iterations["active"].cpu().numpy()
    first = it == 0
    # restart the trace from the hit points of iteration 0, direction unchanged
    starts = np.stack([act[a + "_end"].detach().cpu().numpy()[first] for a in "xyz"], 1).astype(np.float64)
    dirs = starts - np.stack([act[a + "_start"].detach().cpu().numpy()[first] for a in "xyz"], 1)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    scene2 = scenes.config2(ray_count=8, step_count=5)
    scene2["sources"] = scenes.make_rays(starts, dirs, scenes.YELLOW)
    e2, _, _ = scenes.build_engine(scene2, eps=helpers.EPS)
    e2.ray_trace(1)
    o2, _ = helpers.oracle_from_scene(scene2)
    o2.ray_trace(1)
    cmp = helpers.compare_trace(e2, o2)
    assert cmp["max_rel"] < 1e-5
    # nothing re-hits at ray_u ~ 0: every first hit is on the SECOND surface, a lens thickness away
    ray_u = e2.last_trace.tape_f[0, 6].cpu().numpy()
    assert ray_u.min() > 0.05


@pytest.mark.parametrize("step_count", [1, 6, 7])   # 6*S^2 faces per surface: 14, 434, 590 triangles
def test_triangle_counts_around_the_tile_size(step_count):
    scene = scenes.config2(ray_count=3000, step_count=step_count)
    engine, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    engine.ray_trace(3)
    oracle, _ = helpers.oracle_from_scene(scene)
    oracle.ray_trace(3)
    cmp = helpers.compare_trace(engine, oracle)
    assert cmp["max_rel"] < 1e-5 and cmp["iterations"] == 3


def test_largest_mesh_with_a_ray_subsample():
    """config 5's mesh (99 374 triangles, 389 tiles, split over blockIdx.y) x 2 048 rays."""
    scene = scenes.config5(ray_count=2048)
    engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS)
    engine.ray_trace(3)
    assert engine.last_trace.T == 99374
    # forward: oracle on the same fp32 vertex fields the kernel sees
    oracle, _ = helpers.oracle_from_scene(scene)
    oracle.ray_trace(3)
    cmp = helpers.compare_trace(engine, oracle)
    assert cmp["max_rel"] < 1e-5
    # near ties only (compare_trace raises on anything else); 2 048 rays cross two lens surfaces
    assert cmp["mismatches"] <= helpers.near_tie_bound(2 * 2048)
    # gradients: oracle with fp64 vertices formed from the parameters (autograd needs the chain)
    oracle, oparams = helpers.oracle_from_scene(scene, requires_grad=True)
    oracle.ray_trace(3)
    err = scenes.imaging_error(scene)(engine)
    ofin = oracle.finished_rays
    goal = scene["goal_scale"] * ofin["rank"]
    oerr = (ofin["y_end"] - goal[:, 0]) ** 2 + (ofin["z_end"] - goal[:, 1]) ** 2
    if not cmp["exempt"]:
        grads = torch.autograd.grad(err.sum(), params)
        ograds = torch.autograd.grad(oerr.sum(), oparams)
        for g, og in zip(grads, ograds):
            assert helpers.rel_err(g.cpu().numpy(), og.numpy()) < 1e-4


@pytest.mark.parametrize("step_count", [13, 91])
def test_search_is_watertight_at_edges_and_vertices(step_count):
    """Rays aimed exactly at interior vertices and edge midpoints of the lens' first surface (the
    worst case for the fp32 search: every such ray lies in the inflation band of 2-6 triangles) must
    all hit it -- no ray may slip between two triangles -- and the hit triangle must be one of the
    triangles adjacent to the aimed feature.  Covers all three axis groups of the projected layout:
    the same mesh is also tested rotated so that its normals are mostly along y and along z."""
    from tensorflowraytrace_b200 import engine as eng
    rng = np.random.default_rng(step_count)
    surf = scenes.lens_surfaces(step_count)[0]
    verts = scenes.surface_vertices(surf)
    faces = surf["faces"]
    r = np.hypot(verts[:, 1], verts[:, 2])
    inner = np.nonzero(r < 0.8)[0]        # the hexagon's inradius is 0.866: stay off its rim
    pick = rng.choice(inner, size=min(3000, inner.size), replace=False)
    targets = [verts[pick]]
    f = faces[rng.choice(faces.shape[0], size=min(3000, faces.shape[0]), replace=False)]
    f = f[(r[f] < 0.8).all(axis=1)]
    for a, b in ((0, 1), (1, 2), (2, 0)):
        lam = rng.uniform(0.05, 0.95, size=(f.shape[0], 1))
        targets.append(lam * verts[f[:, a]] + (1 - lam) * verts[f[:, b]])
    aim = np.concatenate(targets)
    start = aim + np.stack([np.full(aim.shape[0], -3.0), rng.uniform(-0.3, 0.3, aim.shape[0]),
                            rng.uniform(-0.3, 0.3, aim.shape[0])], axis=1)
    dev = torch.device("cuda:0")
    for perm in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):       # normals mostly along x, then y, then z
        V = verts[:, perm].astype(np.float32)
        S, A = start[:, perm].astype(np.float32), aim[:, perm].astype(np.float32)
        E = (S + (A - S) * np.float32(0.25)).astype(np.float32)
        tri = [torch.tensor(np.ascontiguousarray(V[faces[:, c], k]), device=dev)
               for c in range(3) for k in range(3)]
        rays = [torch.tensor(np.ascontiguousarray(S[:, k]), device=dev) for k in range(3)] + \
               [torch.tensor(np.ascontiguousarray(E[:, k]), device=dev) for k in range(3)]
        out = eng.OpticalSystem3D._intersection(*rays, *tri, 1e-10, 1e-5, 1e-4)
        valid, gather_trig = out[3].cpu().numpy(), out[8].cpu().numpy()
        assert valid.all(), (perm, int((~valid).sum()))
        # the hit triangle touches the aimed point: distance from the aim to the triangle's plane
        # patch is ~0 (fp32), checked through the hit point itself
        hit = np.stack([out[0].detach().cpu().numpy(), out[1].detach().cpu().numpy(),
                        out[2].detach().cpu().numpy()], axis=1)
        assert np.abs(hit - A).max() < 2e-5, (perm, float(np.abs(hit - A).max()))
        corners = V[faces[gather_trig]]                                   # (n, 3, 3)
        lo, hi = corners.min(axis=1) - 2e-5, corners.max(axis=1) + 2e-5
        assert ((A >= lo) & (A <= hi)).all(), perm


def test_precompiled_source_feeds_the_trace():
    """A PrecompiledSource built from a trace's finished rays (`from_samples`, the reference's
    "static optic traced once" use, tfrt/sources.py:1099-1120) drives a second system."""
    from tensorflowraytrace_b200 import sources
    scene = scenes.config2(ray_count=2000, step_count=4)
    first, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    first.ray_trace(1)                                   # stop after the first surface
    act = first.active_rays
    assert _count(act) > 1500
    pre = sources.PrecompiledSource(3, do_downsample=False)
    pre.from_samples([{k: act[k].detach() for k in act.keys()}])
    second, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    second.optical_system.sources = [pre]
    second.optical_system.update()
    second.ray_trace(3)
    # the rays continue from where the first trace left them: same finished set as a full trace
    full, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    full.ray_trace(3)
    a, b = second.finished_rays, full.finished_rays
    assert _count(a) == _count(b) > 1500
    for f in RAYS:
        assert helpers.rel_err(a[f].detach().cpu().numpy(), b[f].detach().cpu().numpy()) < 1e-5, f
