"""Where does the host time of a small trace go?  cProfile of N steps of a config (GPU box)."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from tensorflowraytrace_b200 import scenes  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "config1"
scene = getattr(scenes, cfg)()
if scene["dimension"] == 2:
    engine, _, _ = scenes.build_engine_2d(scene, compile_active_rays=False)
else:
    engine, _, _ = scenes.build_engine(scene, compile_active_rays=False)
K = scene["max_iterations"]


def step():
    engine.optical_system.update()
    engine.ray_trace(K)


for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
