"""Short single-GPU run for ncu: one forward trace of a cfg3-shaped scene (reduced ray count)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from tensorflowraytrace_b200 import scenes  # noqa: E402
import gpu_check  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
sc = scenes.config3(ray_count=n)
tr = gpu_check.make_trace(sc, torch.device("cuda:0"))
tr.forward()
tr.forward()
torch.cuda.synchronize()
print("live", tr.live_counts(), "T", tr.T)
