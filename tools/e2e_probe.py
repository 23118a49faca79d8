import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflowraytrace_b200 import scenes
scene = scenes.config2()
dev = torch.device("cuda")
host = {k: torch.from_numpy(np.ascontiguousarray(v).astype(np.float32)).pin_memory() for k, v in scene["sources"].items()}
engine, params, src = scenes.build_engine(scene, compile_active_rays=False)
def sync(): torch.cuda.synchronize()
def t(fn, n=20):
    fn(); sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    sync(); return (time.perf_counter() - t0) / n * 1e3
def h2d():
    for k, v in host.items(): src[k] = v.to(dev, non_blocking=True)
def resident():
    engine.optical_system.update(); engine.ray_trace(3)
def d2h():
    return engine.finished_rays["x_end"][:1].cpu()
def e2e():
    h2d(); resident(); d2h()
print("h2d ms", t(h2d)); print("resident ms", t(resident)); print("d2h ms", t(d2h)); print("e2e ms", t(e2e))
print("pinned?", [v.is_pinned() for v in host.values()])
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
def copy_big(): big.to(dev, non_blocking=True)
ms = t(copy_big, 5); print("256MB pinned h2d ms", ms, "GB/s", 0.268 / ms * 1e3)
