"""The RayOperation hook API around the hot path (tfrt/operation.py:25-162, 166-196, 311-333;
engine.single_pass tfrt/engine.py:2193-2302): OldestAncestor in the fused loop, GhostThrough and
custom operations on the projection-at-a-time path, hook order, exclusions, technical
intersections."""

import numpy as np
import pytest
import torch

import helpers
from tensorflowraytrace_b200 import engine as eng, operation, scenes

pytestmark = pytest.mark.gpu


def test_oldest_ancestor_rides_the_fused_loop():
    scene = scenes.config2(ray_count=3000, step_count=4)
    fused, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    ops = [operation.StandardReaction(), operation.OldestAncestor()]
    e = eng.OpticalEngine(3, ops, simple_ray_inheritance={"wavelength"})
    e.optical_system = fused.optical_system
    e.annotate()
    e.optical_system.update()
    e.validate_system()
    assert e._fusable()
    e.ray_trace(3)
    fin = e.finished_rays
    assert torch.equal(fin["oldest_ancestor"].long(), e.last_history_slots["finished"].long())
    act = e.active_rays
    assert torch.equal(act["oldest_ancestor"].long(), e.last_history_slots["active"].long())


def test_ghost_through_goes_straight_and_hooks_run_in_order():
    calls = []

    class Recorder(operation.RayOperation):
        def preprocess(self, engine, proj_result):
            calls.append("pre")

        def main(self, engine, proj_result):
            calls.append("main")
            return {}

        def postprocess(self, engine, proj_result, new_rays):
            calls.append(("post", len(new_rays)))

    scene = scenes.config2(ray_count=500, step_count=3)
    base, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    e = eng.OpticalEngine(3, [operation.GhostThrough(), Recorder()],
                          simple_ray_inheritance={"wavelength"})
    e.optical_system = base.optical_system
    e.ray_trace(3)
    assert not e._fusable()
    assert calls[:3] == ["pre", "main", ("post", 1)]
    fin = e.finished_rays
    # straight lines: the finished rays' direction equals the source direction
    src = scene["sources"]
    n = fin["x_start"].shape[0]
    assert n == 500
    d_fin = np.stack([(fin[a + "_end"] - fin[a + "_start"]).detach().cpu().numpy() for a in "xyz"], 1)
    d_fin /= np.linalg.norm(d_fin, axis=1, keepdims=True)
    d_src = np.stack([src[a + "_end"] - src[a + "_start"] for a in "xyz"], 1)
    d_src /= np.linalg.norm(d_src, axis=1, keepdims=True)
    assert np.max(np.abs(d_fin - d_src)) < 1e-4
    assert abs(float(fin["x_end"].mean()) - 8.0) < 1e-4


def test_exclusions_and_signature_validation():
    class A(operation.RayOperation):
        pass

    class B(operation.RayOperation):
        @property
        def exclusions(self):
            return {A}

    with pytest.raises(RuntimeError, match="exclusive"):
        eng.OpticalEngine(3, [A(), B()])
    scene = scenes.config2(ray_count=64, step_count=2)
    base, _, _ = scenes.build_engine(scene, eps=helpers.EPS)

    class NeedsField(operation.RayOperation):
        @property
        def input_signature(self):
            return {"intensity"}

    e = eng.OpticalEngine(3, [operation.StandardReaction(), NeedsField()])
    e.optical_system = base.optical_system
    with pytest.raises(RuntimeError, match="sources signature"):
        e.validate_system()
    with pytest.raises(ValueError):
        eng.OpticalEngine(4, [])


def test_technical_intersections_expose_the_boundaries_hit():
    scene = scenes.config2(ray_count=800, step_count=3)
    base, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    e = eng.OpticalEngine(3, [operation.StandardReaction()], compile_technical_intersections=True,
                          compile_stopped_rays=True, simple_ray_inheritance={"wavelength"})
    e.optical_system = base.optical_system
    e.ray_trace(3)
    res = e.last_projection_result
    assert "target" in res and "norm" in res["target"]
    n_fin = res["rays"]["finished"]["x_start"].shape[0]
    assert res["target"]["xp"].shape[0] == n_fin > 0
    # the target plane's normal is +x
    assert torch.allclose(res["target"]["norm"][:, 0].abs(), torch.ones(n_fin, device="cuda"), atol=1e-6)
    # same histories as the fused loop
    base.ray_trace(3)
    for f in helpers.RAY_FIELDS:
        assert helpers.rel_err(e.finished_rays[f].detach().cpu().numpy(),
                               base.finished_rays[f].detach().cpu().numpy()) < 1e-5
