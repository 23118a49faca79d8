"""
In-tree build of libtfrt_b200.so with nvcc for sm_100a (B200).

The library is plain CUDA C++ behind a C ABI (include/tfrt_b200.h): no torch headers, no
pybind, so one nvcc invocation builds it in seconds and the resulting .so travels with the
source tree.  `python -m tensorflowraytrace_b200._build` or `__graft_entry__.build()`.
"""

import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libtfrt_b200.so")
# the same sources as ptxas wrote them (no post-link steps): only the A/B parity test loads it
PTXAS_LIB_PATH = os.path.join(HERE, "libtfrt_b200_ptxas.so")

SOURCES = ["abi.cu", "pack.cu", "intersect3d.cu", "trace3d.cu", "backward3d.cu", "trace2d.cu",
           "sources.cu", "grid3d.cu"]

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


def can_build():
    """nvcc and the CUDA sources are both present."""
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return os.path.exists(exe) and len(sources()) == len(SOURCES)


def _deps():
    return sources() + [os.path.join(CSRC, "common.cuh"),
                        os.path.join(HERE, "..", "include", "tfrt_b200.h"),
                        os.path.join(HERE, "_reuse_patch.py"), os.path.join(HERE, "_sass_sched.py")]


def source_hash():
    """Hash of every file the library is built from; compiled into the library
    (tfrt_source_hash) so that a stale .so is detected whatever the file times say."""
    h = hashlib.sha256()
    for path in _deps():
        if os.path.exists(path):
            h.update(os.path.basename(path).encode())
            with open(path, "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]


def needs_build():
    """True when the library is missing (file-time checks are not used: a copied tree has
    arbitrary times; `_lib.load` compares the embedded source hash instead)."""
    return not os.path.exists(LIB_PATH)


def build(force=False, verbose=False, out_path=None, post_link=True):
    """Compile every CUDA source into libtfrt_b200.so (sm_100a).  Returns the library path.
    The library is written under a temporary name and renamed, so concurrent processes (one per
    GPU under torchrun) never load a half-written file.  out_path / post_link=False build the
    comparison library of tests/test_gpu_round2.py (same sources, instruction order as ptxas
    wrote it)."""
    out_path = out_path or LIB_PATH
    if not force and os.path.exists(out_path):
        return out_path
    fd, tmp = tempfile.mkstemp(prefix=".libtfrt_b200.", suffix=".so", dir=HERE)
    os.close(fd)
    cmd = ([_nvcc()] + NVCC_FLAGS + [f'-DTFRT_SOURCE_HASH="{source_hash()}"']
           + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + sources())
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    # Post-link steps on the search kernel's packed FP32 loops (same bits out, profiles/
    # r2_k1_tuning.md sections 5 and 6): the loop bodies of the 8-rays-per-thread kernels are
    # re-ordered so that runs of instructions share their coefficient operand (_sass_sched.py),
    # then the operand-reuse flags ptxas leaves unset elsewhere are set (_reuse_patch.py).
    # TFRT_SASS_SCHED=0 / TFRT_REUSE_PATCH=0 skip them; without cuobjdump the library is used as
    # ptxas wrote it.
    if post_link and os.environ.get("TFRT_REUSE_PATCH", "1") != "0":
        if shutil.which("cuobjdump"):
            from . import _reuse_patch, _sass_sched
            # each step rewrites the file only after it has succeeded on every block; a step that
            # fails leaves a correct library behind (as ptxas wrote it, or with the earlier step)
            try:
                if os.environ.get("TFRT_SASS_SCHED", "1") != "0":
                    rep = _sass_sched.reschedule(tmp, min_packed=100)
                    if verbose:
                        print(f"loop bodies re-scheduled: {len(rep)}")
                n = _reuse_patch.patch(tmp)
                if verbose:
                    print(f"reuse flags set: {n}")
            except Exception as e:      # noqa: BLE001  (an optimisation must not break the build)
                print(f"tensorflowraytrace_b200: post-link step skipped ({type(e).__name__}: {e})",
                      file=sys.stderr)
        else:
            print("tensorflowraytrace_b200: cuobjdump not found, operand-reuse patch skipped",
                  file=sys.stderr)
    os.chmod(tmp, 0o755)
    os.replace(tmp, out_path)
    if verbose:
        print(proc.stderr)
    return out_path


def build_unpatched(force=False):
    """libtfrt_b200_ptxas.so: the library without the post-link steps, rebuilt when the sources
    changed (its embedded source hash is compared)."""
    if not force and os.path.exists(PTXAS_LIB_PATH):
        with open(PTXAS_LIB_PATH, "rb") as f:
            data = f.read()
        at = data.find(b"TFRT_SOURCE_HASH=")
        if at >= 0 and data[at + 17:at + 33].decode(errors="replace") == source_hash():
            return PTXAS_LIB_PATH
    return build(force=True, out_path=PTXAS_LIB_PATH, post_link=False)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
