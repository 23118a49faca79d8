"""Golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the float64
oracle -- the reference itself cannot run here).  CPU: the oracle still
reproduces them exactly-ish (regression pin).  GPU: the CUDA path reproduces them within the
north-star tolerances through the public API."""

import os
import sys

import numpy as np
import pytest
import torch

import helpers

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

SCENES = make_golden.golden_scenes()


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", f"{name}.npz")))


@pytest.mark.parametrize("name", sorted(SCENES))
def test_oracle_reproduces_golden(name):
    want, got = load(name), make_golden.run_oracle(name, SCENES[name])
    assert set(want) == set(got)
    for k, v in want.items():
        if v.dtype.kind in "iu":
            assert np.array_equal(v, got[k]), k
        else:
            assert np.allclose(v, got[k], rtol=1e-11, atol=1e-13), k


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENES))
def test_cuda_path_reproduces_golden(name):
    from tensorflowraytrace_b200 import scenes
    scene, want = SCENES[name], load(name)
    if scene["dimension"] == 3:
        engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_dead_rays=True)
        fields = helpers.RAY_FIELDS
    else:
        engine, params, _ = scenes.build_engine_2d(scene, eps=helpers.EPS, compile_dead_rays=True,
                                                   compile_stopped_rays=True)
        fields = helpers.RAY_FIELDS_2D
    engine.ray_trace(scene["max_iterations"])
    live = engine.last_trace.live_counts()
    assert live[: len(want["live_counts"])] == want["live_counts"].tolist()
    for set_name in ("active", "finished", "stopped", "dead"):
        if f"{set_name}_slot" not in want:
            continue
        got = getattr(engine, set_name + "_rays")
        slot = engine.last_history_slots[set_name].cpu().numpy().astype(np.int64)
        assert np.array_equal(slot, want[f"{set_name}_slot"]), set_name
        for f in fields:
            assert helpers.rel_err(got[f].detach().cpu().numpy(), want[f"{set_name}_{f}"]) < 1e-5, (set_name, f)
    if scene["dimension"] == 3:
        tape_tri = engine.last_trace.tape_tri.cpu().numpy()
        k = 0
        while f"hit_tri_{k}" in want:
            assert np.array_equal(tape_tri[k][want[f"hit_slot_{k}"]], want[f"hit_tri_{k}"]), k
            k += 1
    if "grad_0" in want:
        err = scenes.imaging_error(scene)(engine)
        grads = torch.autograd.grad(err.sum(), params)
        assert abs(float(err.sum().detach()) - float(want["loss"])) < 1e-5 * abs(float(want["loss"]))
        for i, g in enumerate(grads):
            assert helpers.rel_err(g.cpu().numpy(), want[f"grad_{i}"]) < 1e-4, i
