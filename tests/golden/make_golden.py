"""Generate the committed golden fixtures under tests/golden/.

The reference (TensorFlow) cannot run in this environment, so these vectors are produced by the
float64 oracle (oracle/), not by the reference itself: they pin the oracle -- and through it the
CUDA path -- against regressions; they are NOT reference output.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import helpers  # noqa: E402
from tensorflowraytrace_b200 import scenes  # noqa: E402

RAYS3 = helpers.RAY_FIELDS
RAYS2 = helpers.RAY_FIELDS_2D


def golden_scenes():
    return {
        "lens3d": scenes.lens_optimize_config(ray_count=192, step_count=3, name="golden"),
        "guide3d": scenes.config4(ray_count=150, theta_res=8, z_res=5, max_iterations=6),
        "wedge2d": scenes.config1(ray_count=64),
        "mixed2d": scenes.mixed_2d(ray_count=160),
    }


def run_oracle(name, scene):
    """Returns a flat dict of float64 / int64 arrays describing the oracle's trace."""
    out = {}
    if scene["dimension"] == 3:
        grad = "goal_scale" in scene
        eng, params = helpers.oracle_from_scene(scene, requires_grad=grad, compile_dead_rays=True)
        eng.ray_trace(scene["max_iterations"])
        fields = RAYS3
        for k, res in enumerate(eng.last_intersect_result):
            out[f"hit_tri_{k}"] = np.where(res["valid"].numpy(), res["gather_trig"].numpy(), -1)
            out[f"hit_slot_{k}"] = eng.last_input_rays[k]["slot"].numpy().astype(np.int64)
        if grad:
            fin = eng.finished_rays
            goal = scene["goal_scale"] * fin["rank"]
            err = (fin["y_end"] - goal[:, 0]) ** 2 + (fin["z_end"] - goal[:, 1]) ** 2
            for i, g in enumerate(torch.autograd.grad(err.sum(), params)):
                out[f"grad_{i}"] = g.numpy()
            out["loss"] = np.array(float(err.sum().detach()))
    else:
        eng, _ = helpers.oracle2d_from_scene(scene, compile_dead_rays=True, compile_stopped_rays=True)
        eng.ray_trace(scene["max_iterations"])
        fields = RAYS2
    out["live_counts"] = np.array(eng.live_counts, dtype=np.int64)
    for set_name in ("active", "finished", "stopped", "dead"):
        s = getattr(eng, set_name + "_rays")
        if not bool(s):
            continue
        out[f"{set_name}_slot"] = s["slot"].detach().numpy().astype(np.int64)
        for f in fields:
            out[f"{set_name}_{f}"] = s[f].detach().numpy()
    return out


def main():
    for name, scene in golden_scenes().items():
        data = run_oracle(name, scene)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **data)
        print(f"{path}: {len(data)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
