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
