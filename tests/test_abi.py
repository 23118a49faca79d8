"""The C-ABI library builds, loads and exports every symbol include/tfrt_b200.h declares
(no compute calls: runs without a GPU)."""

import ctypes
import os
import re

from tensorflowraytrace_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tfrt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tfrt_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = _build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    assert lib.tfrt_abi_version() == _lib.ABI_VERSION
    assert lib.tfrt_last_error() is not None


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_build.build())
    names = declared_symbols()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in tfrt_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)


def test_struct_layouts_match_header_sizes():
    # pointer-size sanity of the ctypes mirrors (a mismatch would corrupt every call)
    assert ctypes.sizeof(_lib.TfrtMaterial) == 8 + 7 * 8
    a = _lib.TfrtTrace3DArgs()
    assert ctypes.sizeof(a) % 8 == 0
    assert _lib.TfrtTrace3DArgs.src.offset == 56     # 13 leading 4-byte scalars, padded to 8
    lib = _lib.load()
    for code, mirror in enumerate(_lib.STRUCT_MIRRORS):
        assert lib.tfrt_struct_size(code) == ctypes.sizeof(mirror), mirror.__name__
    # a tile of 256 triangles = 130 pair records of 24 floats (two spare pair slots complete
    # axis groups of odd size)
    assert _lib.load().tfrt_packed_floats(1) == 130 * 24
    assert _lib.load().tfrt_packed_floats(257) == 2 * 130 * 24
    assert _lib.load().tfrt_trace3d_sync_bytes(1000, 3) >= 4 * (2 * 3 + 4 * 4 * 3)
    assert _lib.load().tfrt_source_hash().decode() == _build.source_hash()


def test_sass_is_blackwell_native():
    """The intersect kernel must use the TMA bulk copy (UBLKCP) and packed FFMA2 on sm_100a."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", _build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out
    assert "FFMA2" in out


def test_operand_reuse_patch_is_applied_and_idempotent(tmp_path):
    """The build sets the operand-reuse flags ptxas leaves unset in the search kernel's packed
    loops (_reuse_patch.py): on the shipped library nothing is left to set, on an unpatched build
    of the same sources the tool finds the flags, sets them, and changes nothing but those bits."""
    import shutil
    import subprocess
    import pytest
    from tensorflowraytrace_b200 import _reuse_patch
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    if os.environ.get("TFRT_REUSE_PATCH", "1") == "0":
        pytest.skip("reuse patch disabled by TFRT_REUSE_PATCH=0")
    shipped = str(tmp_path / "shipped.so")
    shutil.copy(_build.build(), shipped)
    assert _reuse_patch.patch(shipped, dry=True) == 0
    # an unpatched object of the search kernel alone (seconds to compile)
    raw = str(tmp_path / "raw.so")
    src = os.path.join(_build.CSRC, "intersect3d.cu")
    abi = os.path.join(_build.CSRC, "abi.cu")
    subprocess.run([_build._nvcc()] + _build.NVCC_FLAGS + ["-o", raw, src, abi], check=True)
    before = open(raw, "rb").read()
    n = _reuse_patch.patch(raw)
    after = open(raw, "rb").read()
    assert n > 0 and len(before) == len(after)
    mask = (0xf << 58) | (1 << 45)
    changed = [i for i in range(0, len(before) - 7, 8) if before[i:i + 8] != after[i:i + 8]]
    assert 0 < len(changed) <= n
    for i in changed:
        a, b = int.from_bytes(before[i:i + 8], "little"), int.from_bytes(after[i:i + 8], "little")
        assert (a ^ b) & ~mask == 0 and b & a == a
    assert _reuse_patch.patch(raw, dry=True) == 0


def test_rescheduled_loop_bodies_respect_the_dependency_graph(tmp_path):
    """_sass_sched.py re-orders the loop bodies of the 8-rays-per-thread search kernels: every
    block must be a permutation of ptxas's instructions (only stall / yield / reuse bits differ),
    in an order that respects ptxas's dependency graph, with stall counts that keep ptxas's own
    minimum distances."""
    import shutil
    import subprocess
    import pytest
    from tensorflowraytrace_b200 import _sass_sched
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    raw = str(tmp_path / "raw.so")
    new = str(tmp_path / "new.so")
    subprocess.run([_build._nvcc()] + _build.NVCC_FLAGS +
                   ["-o", raw, os.path.join(_build.CSRC, "intersect3d.cu"),
                    os.path.join(_build.CSRC, "abi.cu")], check=True)
    shutil.copy(raw, new)
    report = _sass_sched.reschedule(new, min_packed=100)
    assert len(report) == 6            # three axis loops x (queue, no queue)
    for _, _, n, before, after, _, _ in report:
        assert n > 100 and after < before
    assert _sass_sched.verify(raw, new) == 6
    assert len(open(raw, "rb").read()) == len(open(new, "rb").read())
