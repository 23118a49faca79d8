"""cProfile of bench.py's end-to-end step (pinned host rays -> device, update, trace, read back) for a
small config (GPU box): where do the milliseconds of `e2e` go when the kernels take microseconds?"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflowraytrace_b200 import scenes  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "config2"
scene = getattr(scenes, cfg)()
build = scenes.build_engine_2d if scene["dimension"] == 2 else scenes.build_engine
engine, _, src = build(scene, compile_active_rays=False, compile_finished_rays=True)
K = scene["max_iterations"]
dev = torch.device("cuda:0")
host = {k: torch.from_numpy(np.ascontiguousarray(v).astype(np.float32)).pin_memory()
        for k, v in scene["sources"].items()}


def step():
    for k, t in host.items():
        src[k] = t.to(dev, non_blocking=True)
    engine.optical_system.update()
    engine.ray_trace(K)
    return engine.finished_rays["x_end"][:1].cpu()


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    step()
torch.cuda.synchronize()
print(f"{cfg}: {1e3 * (time.perf_counter() - t0) / 100:.3f} ms per e2e step (wall)")
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
