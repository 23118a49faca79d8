"""Optional acceleration mode: `OpticalSystem3D(accel="grid")` walks a
device-built uniform grid instead of testing every triangle.  It must return the hit words the
brute-force search returns -- same triangle for every ray, near ties included, since every
candidate is tested with the brute-force kernel's own arithmetic and the winner is the same
lexicographic minimum -- hence the same histories and the same gradients."""

import numpy as np
import pytest
import torch

import helpers
from tensorflowraytrace_b200 import kernels, scenes

pytestmark = pytest.mark.gpu
RAYS = helpers.RAY_FIELDS
DEV = "cuda"


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype, device=DEV)


def _both(scene, iters, **kw):
    brute, bparams, _ = scenes.build_engine(scene, eps=helpers.EPS, **kw)
    brute.ray_trace(iters)
    grid, gparams, _ = scenes.build_engine(scene, eps=helpers.EPS, **kw)
    grid.optical_system.accel = "grid"
    grid.optical_system.update()
    assert grid.optical_system.grid is not None
    grid.ray_trace(iters)
    return brute, bparams, grid, gparams


def _assert_same_trace(brute, grid):
    a, b = brute.last_trace, grid.last_trace
    assert a.live_counts() == b.live_counts()
    for k in range(a.K):
        n = a.live_counts()[k]
        slots = a.live[k, :n].long()
        assert torch.equal(b.live[k, :n].long(), slots)
        assert torch.equal(a.tape_tri[k][slots], b.tape_tri[k][slots]), k        # same hit indices
        assert torch.equal(a.tape_status[k][slots], b.tape_status[k][slots])
        assert torch.equal(a.tape_f[k][:, slots], b.tape_f[k][:, slots])          # bit-identical tape


def test_grid_statistics_and_large_triangle_list():
    scene = scenes.config2(ray_count=1024, step_count=13)
    engine, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    engine.optical_system.accel = "grid"
    engine.optical_system.update()
    st = engine.optical_system.grid.stats()
    assert st["overflow"] == 0
    assert st["large"] == 2                      # the two target triangles stay out of the grid
    assert st["cells"] == st["dims"][0] * st["dims"][1] * st["dims"][2] > 8
    assert 2028 <= st["entries"] <= 64 * 2028    # every lens triangle is registered somewhere
    assert 0.01 < st["typical_extent"] < 0.3


@pytest.mark.parametrize("n_rays,n_tri", [(257, 40), (4097, 1500), (20000, 6000)])
def test_grid_seam_equals_brute_force_on_a_random_soup(n_rays, n_tri):
    rng = np.random.default_rng(n_tri)
    c = rng.uniform(-1, 1, size=(n_tri, 3)) + np.array([3.0, 0, 0])
    size = 0.4 if n_tri < 100 else 0.08
    tri = np.concatenate([c + size * rng.normal(size=(n_tri, 3)) for _ in range(3)], axis=1)
    tri[: max(1, n_tri // 200)] *= 1.0                     # keep as is
    big = np.array([[6, -9, -9, 6, 9, -9, 6, 0, 9]], dtype=np.float64)   # one huge triangle behind
    tri = np.concatenate([tri, big]).T.copy()
    S = rng.uniform(-0.3, 0.3, size=(n_rays, 3))
    D = np.stack([np.ones(n_rays), rng.uniform(-0.5, 0.5, n_rays), rng.uniform(-0.5, 0.5, n_rays)], 1)
    rays = np.concatenate([S, S + D], axis=1).T.copy()
    d_tri, d_rays = dev(tri), dev(rays)
    eps = helpers.EPS
    packed = kernels.pack_triangles(d_tri, eps["size_epsilion"])
    args = (eps["intersect_epsilion"], eps["size_epsilion"], eps["ray_start_epsilion"])
    brute = kernels.intersect3d(d_rays, d_tri, packed, *args)
    grid = kernels.TriangleGrid(d_tri, packed, eps["size_epsilion"])
    assert grid.stats()["overflow"] == 0 and grid.stats()["large"] >= 1
    fast = kernels.intersect3d(d_rays, d_tri, packed, *args, grid=grid)
    assert torch.equal(brute["valid"], fast["valid"])
    v = brute["valid"]
    assert int(v.sum()) > n_rays // 4
    assert torch.equal(brute["gather_trig"][v], fast["gather_trig"][v])
    for k in ("x", "y", "z", "ray_u", "trig_u", "trig_v"):
        assert torch.equal(brute[k][v], fast[k][v]), k


def test_grid_trace_equals_brute_force_lens_and_gradients():
    scene = scenes.lens_optimize_config(ray_count=30000, step_count=21, name="grid")
    brute, bparams, grid, gparams = _both(scene, 3)
    _assert_same_trace(brute, grid)
    errf = scenes.imaging_error(scene)
    gb = torch.autograd.grad(errf(brute).sum(), bparams)
    gg = torch.autograd.grad(errf(grid).sum(), gparams)
    for a, b in zip(gb, gg):
        assert float(a.abs().max()) > 0
        assert helpers.rel_err(b.cpu().numpy(), a.cpu().numpy()) < 1e-6     # same tape => same VJP
    for f in RAYS:
        assert torch.equal(brute.finished_rays[f], grid.finished_rays[f])


def test_grid_trace_equals_brute_force_light_guide():
    """Multi-bounce TIR inside a closed cylinder: rays start ON triangles, graze neighbours and
    bounce 8 times; the grid must never lose the brute-force winner."""
    scene = scenes.config4(ray_count=20000, theta_res=40, z_res=60, max_iterations=8)
    brute, _, grid, _ = _both(scene, 8, compile_dead_rays=True)
    _assert_same_trace(brute, grid)


def test_grid_full_size_cfg5_subsample_equals_brute_force():
    """BASELINE config 5's mesh (99 374 triangles) with a 262 144-ray subsample of its source:
    identical hit indices at every iteration."""
    scene = scenes.config5(ray_count=262144)
    brute, _, grid, _ = _both(scene, 3, compile_active_rays=False)
    assert brute.last_trace.T == 99374
    _assert_same_trace(brute, grid)
    st = grid.optical_system.grid.stats()
    assert st["overflow"] == 0 and st["large"] == 2
