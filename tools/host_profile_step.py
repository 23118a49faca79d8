"""Host-side cost of one optimisation step (update + trace + loss + backward) of a 3D config, or
of update + trace of a 2D one: wall clock per step, then cProfile by own time (GPU box)."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflowraytrace_b200 import optimizer, scenes  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
scene = getattr(scenes, cfg)()
K = scene["max_iterations"]
if scene["dimension"] == 2 or "goal_scale" not in scene:
    build = scenes.build_engine_2d if scene["dimension"] == 2 else scenes.build_engine
    engine, params, _ = build(scene, compile_active_rays=False, compile_finished_rays=True)

    def step():
        engine.optical_system.update()
        engine.ray_trace(K)
        engine.finished_rays

    def trace_only():
        engine.ray_trace(K)
else:
    engine, params, _ = scenes.build_engine(scene, compile_active_rays=False,
                                            compile_finished_rays=True)
    opt = optimizer.SGD_Optimizer(engine, params, scenes.imaging_error(scene), K)
    acc = [None] * len(params)

    def step():
        opt.process_gradient(acc)

    def trace_only():
        engine.ray_trace(K)


def wall(fn, n):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return 1e6 * host / n, 1e6 * (time.perf_counter() - t0) / n


print(cfg, "step: host us %.1f, with GPU drain us %.1f" % wall(step, reps))
print(cfg, "ray_trace only: host us %.1f, with GPU drain us %.1f" % wall(trace_only, reps))
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(40)
