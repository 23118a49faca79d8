"""Print the fp32 FMA-pipe micro-benchmarks (roofline denominators) of the current GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflowraytrace_b200 import kernels  # noqa: E402

for mode, name in ((0, "scalar FFMA"), (1, "FFMA2"), (2, "FFMA2+FFMA 1:1"), (3, "FFMA2+FFMA 2:1"),
                   (4, "8 FFMA2 + 4 ALU"), (5, "8 FFMA2 + 1 MUFU"), (6, "8 FFMA2 + 4 ALU + 1 MUFU"),
                   (7, "8 FFMA2 + 8 ALU")):
    print(f"{name:18s}: {kernels.fp32_peak(mode, 8192) / 1e12:.2f} TFLOP/s")
