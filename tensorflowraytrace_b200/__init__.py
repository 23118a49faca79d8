"""
tensorflowraytrace_b200 -- B200-native implementation of the tfrt differentiable ray-trace hot
path (OpticalEngine.ray_trace and its backward pass), behind the tfrt public surface.

Layout
------
csrc/ + libtfrt_b200.so   hand-written sm_100a CUDA kernels behind a C ABI (include/tfrt_b200.h)
_lib.py, kernels.py       ctypes binding and thin launchers (torch tensors = device memory only)
autograd.py               torch.autograd.Function wrappers (the tf.GradientTape seam)
update / materials / operation / boundaries / sources / engine / optimizer
                          host-side mirror of the tfrt modules of the same name
parallel.py               ray sharding + parameter-gradient all-reduce (one process per GPU)
scenes.py, mesh_tools.py  deterministic synthetic scenes / meshes for the BASELINE configs

There is no CPU fallback: every compute entry point needs a CUDA device and raises otherwise.
"""

from . import materials, mesh_tools, scenes  # noqa: F401  (torch-free or light imports)

__all__ = ["materials", "mesh_tools", "scenes", "kernels", "engine", "boundaries", "sources",
           "operation", "optimizer", "parallel", "update", "autograd"]


def __getattr__(name):
    # heavier modules (torch + the CUDA library binding) are imported on first use
    if name in __all__:
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
