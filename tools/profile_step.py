"""Short single-GPU run for ncu: one fwd+bwd optimisation step of a cfg3-shaped scene through the
public API with the brute-force search, then one with the acceleration grid (every kernel of a
step launches at least once: search, validation, reaction, materialisation, backward, grid build
and walk)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tensorflowraytrace_b200 import optimizer, scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
scene = scenes.config3(ray_count=n, with_rays=False)
source = scenes.lens_optimize_device_source(n, scene["seed"])
engine, params, _ = scenes.build_engine(scene, source=source, compile_active_rays=False)
opt = optimizer.SGD_Optimizer(engine, params, scenes.imaging_error(scene), 3)
for mode in (None, "grid"):
    engine.optical_system.accel = mode
    for _ in range(2):
        grads, err = opt.process_gradient([None] * len(params))
    torch.cuda.synchronize()
    print(mode, "mean error", float(err), "live", engine.last_trace.live_counts(),
          engine.last_trace.diagnostics)
