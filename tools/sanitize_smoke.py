"""Tiny end-to-end run for compute-sanitizer: 3D fwd+bwd (splits, atomics, compaction, histories)
and 2D fwd+bwd on small scenes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflowraytrace_b200 import scenes  # noqa: E402

scene = scenes.lens_optimize_config(ray_count=3000, step_count=5, name="san")
engine, params, _ = scenes.build_engine(scene, compile_dead_rays=True, compile_stopped_rays=True)
engine.ray_trace(3)
err = scenes.imaging_error(scene)(engine)
g = torch.autograd.grad(err.sum(), params)
scene4 = scenes.config4(ray_count=2000, theta_res=12, z_res=8, max_iterations=6)
e4, _, _ = scenes.build_engine(scene4, compile_dead_rays=True)
e4.ray_trace(6)
s2 = scenes.mixed_2d(ray_count=700)
e2, p2, _ = scenes.build_engine_2d(s2, parametric=True, compile_dead_rays=True, compile_stopped_rays=True)
e2.ray_trace(5)
g2 = torch.autograd.grad((e2.finished_rays["y_end"] ** 2).sum(), p2, allow_unused=True)
# rays born on the device (sampling kernels + the fused directed-source kernel) and the grid mode
src = scenes.lens_optimize_device_source(5000, 7)
scene5 = scenes.lens_optimize_config(ray_count=5000, step_count=5, name="san_dev", with_rays=False)
e5, p5, _ = scenes.build_engine(scene5, source=src, compile_finished_rays=True)
e5.optical_system.accel = "grid"
e5.optical_system.update()
e5.ray_trace(3)
g5 = torch.autograd.grad(scenes.imaging_error(scene5)(e5).sum(), p5)
torch.cuda.synchronize()
print("ok", float(err.sum()), e4.last_trace.live_counts(), e2.last_trace.live_counts())
