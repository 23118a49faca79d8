"""
ctypes binding of libtfrt_b200.so (the C ABI declared in include/tfrt_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if it cannot be
loaded, or a call fails, a RuntimeError is raised.  Calling a compute entry point without a
CUDA device raises as well (the library returns a CUDA error).
"""

import ctypes
import os

from . import _build

c_float_p = ctypes.POINTER(ctypes.c_float)
c_void_p = ctypes.c_void_p


class TfrtMaterial(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("_pad", ctypes.c_int32), ("c", ctypes.c_double * 7)]


class TfrtTrace3DArgs(ctypes.Structure):
    _fields_ = [
        ("n_rays", ctypes.c_int32),
        ("n_tri", ctypes.c_int32),
        ("n_optical", ctypes.c_int32),
        ("max_iterations", ctypes.c_int32),
        ("intersect_eps", ctypes.c_float),
        ("size_eps", ctypes.c_float),
        ("ray_start_eps", ctypes.c_float),
        ("new_ray_length", ctypes.c_float),
        ("margin_ulps", ctypes.c_float),
        ("index_mode", ctypes.c_int32),
        ("n_materials", ctypes.c_int32),
        ("n_lambda", ctypes.c_int32),
        ("grid_max_cells", ctypes.c_int32),
        ("src", c_void_p * 7),
        ("tri", c_void_p * 9),
        ("norm", c_void_p),
        ("category", c_void_p),
        ("mat_in", c_void_p),
        ("mat_out", c_void_p),
        ("n_in", c_void_p),
        ("n_out", c_void_p),
        ("materials", c_void_p),
        ("n_table", c_void_p),
        ("lambda_index", c_void_p),
        ("packed", c_void_p),
        ("grid", c_void_p),
        ("cur", c_void_p * 6),
        ("cur_d", c_void_p),
        ("hit", c_void_p),
        ("tape_f", c_void_p),
        ("tape_tri", c_void_p),
        ("tape_status", c_void_p),
        ("tape_outpos", c_void_p),
        ("live", c_void_p),
        ("cat_ids", c_void_p * 3),
        ("sync_ws", c_void_p),
        ("counters", c_void_p),
        ("diag", c_void_p),
    ]


class TfrtTrace3DBwdArgs(ctypes.Structure):
    _fields_ = [
        ("g_hist", (c_void_p * 6) * 4),
        ("g_acc", c_void_p),
        ("g_ray_ws", c_void_p),
        ("g_tri_out", c_void_p * 9),
        ("g_norm_out", c_void_p),
        ("g_src", c_void_p * 6),
        ("dead_ray_length", ctypes.c_float),
        ("n_grad_tri", ctypes.c_int32),
    ]


class TfrtTrace2DArgs(ctypes.Structure):
    _fields_ = [
        ("n_rays", ctypes.c_int32),
        ("max_iterations", ctypes.c_int32),
        ("n_seg", ctypes.c_int32),
        ("n_seg_optical", ctypes.c_int32),
        ("n_seg_stop", ctypes.c_int32),
        ("n_arc", ctypes.c_int32),
        ("n_arc_optical", ctypes.c_int32),
        ("n_arc_stop", ctypes.c_int32),
        ("intersect_eps", ctypes.c_float),
        ("size_eps", ctypes.c_float),
        ("ray_start_eps", ctypes.c_float),
        ("new_ray_length", ctypes.c_float),
        ("index_mode", ctypes.c_int32),
        ("n_materials", ctypes.c_int32),
        ("src", c_void_p * 5),
        ("seg", c_void_p * 4),
        ("arc", c_void_p * 5),
        ("seg_mat_in", c_void_p),
        ("seg_mat_out", c_void_p),
        ("seg_n_in", c_void_p),
        ("seg_n_out", c_void_p),
        ("arc_mat_in", c_void_p),
        ("arc_mat_out", c_void_p),
        ("arc_n_in", c_void_p),
        ("arc_n_out", c_void_p),
        ("materials", c_void_p),
        ("tape_f", c_void_p),
        ("tape_d", c_void_p),
        ("tape_hit", c_void_p),
        ("tape_status", c_void_p),
        ("live_counts", c_void_p),
    ]


# every symbol include/tfrt_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "tfrt_last_error", "tfrt_abi_version", "tfrt_source_hash", "tfrt_struct_size",
    "tfrt_faces_forward", "tfrt_faces_backward",
    "tfrt_packed_floats", "tfrt_pack_triangles",
    "tfrt_intersect3d", "tfrt_grid_bytes", "tfrt_grid_build", "tfrt_grid_stats",
    "tfrt_trace3d_sync_bytes", "tfrt_trace3d_forward", "tfrt_history_count", "tfrt_materialize3d",
    "tfrt_trace3d_backward",
    "tfrt_intersect2d", "tfrt_trace2d_forward", "tfrt_trace2d_backward",
    "tfrt_sample_uniform", "tfrt_sample_square", "tfrt_sample_sphere", "tfrt_rays_from_directions",
    "tfrt_fp32_peak", "tfrt_launch_count", "tfrt_profile_enable", "tfrt_profile_read",
]

ABI_VERSION = 4
STRUCT_MIRRORS = (TfrtMaterial, TfrtTrace3DArgs, TfrtTrace3DBwdArgs, TfrtTrace2DArgs)

_lib = None


def lib_path():
    return _build.LIB_PATH


def _stale(path):
    """The library's embedded source hash differs from the hash of the sources on disk.  Read
    from the file's bytes: dlopen() here would keep the old mapping alive, and loading the
    rebuilt library under the same name would return it again."""
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError:
        return True
    at = data.find(b"TFRT_SOURCE_HASH=")
    if at < 0:
        return True
    return data[at + 17:at + 33].decode(errors="replace") != _build.source_hash()


def load():
    """Load (building first if needed) libtfrt_b200.so; raises RuntimeError on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB_PATH):
        _build.build()
    elif _build.can_build() and _stale(_build.LIB_PATH):
        # the sources changed since the library was built (a kernel edit without an ABI bump):
        # never measure or test old code silently
        _build.build(force=True)
    _lib = open_library(_build.LIB_PATH)
    return _lib


def open_library(path):
    """ctypes handle of a build of the library, prototypes declared and ABI checked."""
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise RuntimeError(f"tensorflowraytrace_b200: cannot load {path}: {e}") from e
    lib.tfrt_last_error.restype = ctypes.c_char_p
    lib.tfrt_abi_version.restype = ctypes.c_int
    lib.tfrt_source_hash.restype = ctypes.c_char_p
    lib.tfrt_struct_size.restype = ctypes.c_size_t
    lib.tfrt_struct_size.argtypes = [ctypes.c_int]
    lib.tfrt_packed_floats.restype = ctypes.c_size_t
    lib.tfrt_packed_floats.argtypes = [ctypes.c_int32]
    lib.tfrt_grid_bytes.restype = ctypes.c_size_t
    lib.tfrt_grid_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32]
    lib.tfrt_trace3d_sync_bytes.restype = ctypes.c_size_t
    lib.tfrt_trace3d_sync_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32]
    u32, i32, i64, f64 = ctypes.c_uint32, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    lib.tfrt_sample_uniform.argtypes = [u32, u32, u32, u32, i64, i32, f64, f64, c_void_p, c_void_p]
    lib.tfrt_sample_square.argtypes = [u32, u32, u32, i64, i32, f64, f64, c_void_p, c_void_p, c_void_p]
    lib.tfrt_sample_sphere.argtypes = [u32, u32, u32, i32, i64, i32, i64, f64, f64, c_void_p,
                                       c_void_p, c_void_p]
    lib.tfrt_rays_from_directions.argtypes = [i32, c_void_p, c_void_p, i32, c_void_p, c_void_p, f64,
                                              i32, c_void_p, c_void_p, c_void_p]
    lib.tfrt_launch_count.restype = ctypes.c_longlong
    lib.tfrt_launch_count.argtypes = [ctypes.c_int]
    if lib.tfrt_abi_version() != ABI_VERSION:
        raise RuntimeError(f"tensorflowraytrace_b200: {path} has ABI version "
                           f"{lib.tfrt_abi_version()}, this package needs {ABI_VERSION}; rebuild "
                           f"with `python -m tensorflowraytrace_b200._build --force`")
    for code, mirror in enumerate(STRUCT_MIRRORS):
        if lib.tfrt_struct_size(code) != ctypes.sizeof(mirror):
            raise RuntimeError(f"tensorflowraytrace_b200: ctypes mirror of {mirror.__name__} is "
                               f"{ctypes.sizeof(mirror)} bytes, the library's is "
                               f"{lib.tfrt_struct_size(code)}")
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().tfrt_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def ptr_array(tensors, n):
    arr = (c_void_p * n)()
    for i in range(n):
        t = tensors[i] if i < len(tensors) else None
        arr[i] = 0 if t is None else t.data_ptr()
    return arr
