"""The optimisation loop around the hot path (dev/simple_3d_optimize.py flow): multi-surface
parametric lens, constraints, vertex_update_map, sparse accumulator / smoother, SGD_Optimizer.
The mean imaging error must fall, the constraints must hold, the accumulator must equal its
dense form."""

import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_lens_optimisation_reduces_the_imaging_error():
    import simple_3d_optimize as ex
    engine, lens, opt, acc, smo = ex.build(ray_count=6000, rings=8, seed=3)
    first = None
    errors = []
    for step in range(40):
        err = opt.single_step([acc, acc], lr_scale=2e-5, momentum=0.8)
        for p in lens.parameters:
            opt.smooth(p, smo)
        errors.append(err)
        first = err if first is None else first
    assert np.isfinite(errors).all()
    assert np.mean(errors[-5:]) < 0.6 * np.mean(errors[:5]), (errors[:5], errors[-5:])
    engine.optical_system.update()
    p0, p1 = lens.parameters
    # ThicknessConstraint(0.0, "min") on the first surface, (0.2, "min") between the two
    assert abs(float((0.0 - p0).max())) < 1e-6
    assert abs(float((p0 - p1).max()) + 0.2) < 1e-6


def test_sparse_accumulator_matches_dense_matmul():
    from tensorflowraytrace_b200 import boundaries, mesh_tools, optimizer
    pts, faces = mesh_tools.hexagonal_mesh(1.0, 5)
    mesh = boundaries.TriangleMesh(np.stack([np.zeros(len(pts)), pts[:, 0], pts[:, 1]], 1), faces)
    top = mesh_tools.get_closest_point(mesh, (0, 0, 0))
    _, acc = mesh_tools.mesh_parametrization_tools(mesh, top)
    g = torch.randn(pts.shape[0], device="cuda")
    sparse = torch.sparse.mm(mesh_tools.to_torch_sparse(acc, "cuda"), g.reshape(-1, 1)).reshape(-1)
    dense = torch.as_tensor(acc.toarray(), dtype=torch.float32, device="cuda") @ g
    assert torch.allclose(sparse, dense, rtol=1e-5, atol=1e-5)
    p = torch.nn.Parameter(torch.randn(pts.shape[0], device="cuda"))
    before = p.detach().clone()
    sm = mesh_tools.mesh_smoothing_tool(mesh, [1.0, 1.0])
    optimizer.SGD_Optimizer.smooth(p, mesh_tools.to_torch_sparse(sm, "cuda"))
    want = torch.as_tensor(sm.toarray(), dtype=torch.float32, device="cuda") @ before
    assert torch.allclose(p.detach(), want, rtol=1e-5, atol=1e-5)


def _guide_system(guide, n_rays=4000, seed=2):
    import math
    from tensorflowraytrace_b200 import boundaries, engine as eng, materials, mesh_tools, operation, scenes, sources
    rng = np.random.default_rng(seed)
    rad = 0.3 * np.sqrt(rng.uniform(0, 1, n_rays))
    ang = rng.uniform(0, 2 * math.pi, n_rays)
    starts = np.stack([np.full(n_rays, 0.01), rad * np.cos(ang), rad * np.sin(ang)], 1)
    dirs = scenes.random_lambertian_sphere(math.pi / 3, n_rays, rng)
    src = sources.ManualSource(3).feed(scenes.make_rays(starts, dirs, 550.0, ray_length=0.1))
    tv, tf_ = mesh_tools.plane_mesh((4.05, 0.0, 0.0), 5.0, axis=0)
    target = boundaries.ManualTriangleBoundary(mesh=boundaries.TriangleMesh(tv, tf_))
    system = eng.OpticalSystem3D()
    system.optical, system.targets, system.sources = [guide], [target], [src]
    system.materials = [{"n": materials.vacuum}, {"n": materials.acrylic}]
    system.update()
    engine = eng.OpticalEngine(3, [operation.StandardReaction()], compile_active_rays=False)
    engine.optical_system = system
    engine.validate_system()
    return engine


def test_cylindrical_guide_parameters_and_symmetry():
    """ParametricCylindricalGuide: min-pinning constraint, cap vertices fixed, gradients reach
    the parameters, and the rotationally symmetric gradient is the per-layer sum of the
    per-vertex one (transpose of repeat_interleave)."""
    from tensorflowraytrace_b200 import boundaries
    mat = {"mat_in": np.int64(1), "mat_out": np.int64(0)}
    kw = dict(theta_res=12, z_res=9, initial_taper=(0.0, 0.15), material_dict=mat)
    full = boundaries.ParametricCylindricalGuide((0, 0, 0), (4, 0, 0), 0.5, **kw)
    sym = boundaries.ParametricCylindricalGuide((0, 0, 0), (4, 0, 0), 0.5,
                                                rotationally_symmetric=True, **kw)
    assert full.parameters.shape == (108,) and sym.parameters.shape == (9,)
    assert float(full.parameters.min()) == 0.0
    assert torch.equal(full.vertices[[0, -1]], full._zero_points[[0, -1]])   # caps (first, last vertex) stay on the axis
    assert torch.allclose(full.vertices, sym.vertices)
    grads = []
    for guide in (full, sym):
        engine = _guide_system(guide)
        engine.ray_trace(12)
        fin = engine.finished_rays
        assert fin["x_start"].shape[0] > 100
        loss = (fin["y_end"] ** 2 + fin["z_end"] ** 2).sum()
        (g,) = torch.autograd.grad(loss, [guide.parameters])
        assert torch.isfinite(g).all() and float(g.abs().max()) > 0
        grads.append(g)
    assert torch.allclose(grads[0].reshape(9, 12).sum(1), grads[1], rtol=1e-3, atol=1e-4 * float(grads[1].abs().max()))
    assert full.accumulator.shape == (108, 108)


def test_master_slave_boundary_shares_parameters():
    from tensorflowraytrace_b200 import boundaries, mesh_tools
    pts, faces = mesh_tools.hexagonal_mesh(1.0, 4)
    zero = boundaries.TriangleMesh(np.stack([np.zeros(len(pts)), pts[:, 0], pts[:, 1]], 1), faces)
    mat = {"mat_in": np.int64(1), "mat_out": np.int64(0)}

    def masters(v):          # mirror symmetry in z: vertices with z >= 0 own a parameter
        return v[:, 2] >= -1e-9

    def slaves(v, m, free):  # the mirror image of master m
        return {s for s in free if abs(v[s, 1] - v[m, 1]) < 1e-6 and abs(v[s, 2] + v[m, 2]) < 1e-6}

    init = 0.1 * (1 - pts[:, 0] ** 2 - pts[:, 1] ** 2)
    ms = boundaries.MasterSlaveParametricTriangleBoundary(
        masters, slaves, zero, boundaries.FromVectorVG((1, 0, 0)), initial_parameters=init,
        material_dict=mat)
    plain = boundaries.ParametricTriangleBoundary(
        zero, boundaries.FromVectorVG((1, 0, 0)), initial_parameters=init, material_dict=mat)
    assert ms.parameters.shape[0] == int(masters(zero.points).sum()) < pts.shape[0]
    assert torch.allclose(ms.vertices, plain.vertices, atol=1e-7)
    w = torch.randn_like(plain["xp"])
    (g_ms,) = torch.autograd.grad((ms["xp"] * w).sum() + (ms["norm"][:, 1] * w).sum(), [ms.parameters])
    (g_pl,) = torch.autograd.grad((plain["xp"] * w).sum() + (plain["norm"][:, 1] * w).sum(), [plain.parameters])
    want = torch.zeros_like(g_ms).index_add_(0, ms._gather, g_pl)
    assert torch.allclose(g_ms, want, rtol=1e-4, atol=1e-6)
