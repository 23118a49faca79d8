"""cProfile of the optimisation step of examples/simple_3d_optimize.py (host overhead study)."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import simple_3d_optimize as ex  # noqa: E402

engine, lens, opt, acc, smo = ex.build(ray_count=20000, rings=12)
for _ in range(10):
    opt.single_step([acc, acc], lr_scale=2e-5, momentum=0.8)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(100):
    opt.single_step([acc, acc], lr_scale=2e-5, momentum=0.8)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) * 10)
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    opt.single_step([acc, acc], lr_scale=2e-5, momentum=0.8)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
