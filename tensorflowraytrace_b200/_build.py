"""
In-tree build of libtfrt_b200.so with nvcc for sm_100a (B200).

The library is plain CUDA C++ behind a C ABI (include/tfrt_b200.h): no torch headers, no
pybind, so one nvcc invocation builds it in seconds and the resulting .so travels with the
source tree.  `python -m tensorflowraytrace_b200._build` or `__graft_entry__.build()`.
"""

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libtfrt_b200.so")

SOURCES = ["abi.cu", "pack.cu", "intersect3d.cu", "trace3d.cu", "backward3d.cu", "trace2d.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libtfrt_b200.so cannot be built")
    return exe


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    lib_m = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, "common.cuh"),
                        os.path.join(HERE, "..", "include", "tfrt_b200.h")]
    return any(os.path.getmtime(d) > lib_m for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """Compile every CUDA source into libtfrt_b200.so (sm_100a).  Returns the library path."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
