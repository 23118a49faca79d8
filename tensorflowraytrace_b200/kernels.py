"""
Thin Python launchers over the C ABI (include/tfrt_b200.h).  Tensors are torch CUDA tensors used
purely as device memory; all arithmetic happens in libtfrt_b200.so on the current CUDA stream.
"""

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import TfrtMaterial, TfrtTrace3DArgs, TfrtTrace3DBwdArgs, check, ptr, ptr_array

TRI_FIELDS = ("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2")
RAY_FIELDS = ("x_start", "y_start", "z_start", "x_end", "y_end", "z_end")
HISTORY = {"active": 0, "stopped": 1, "finished": 2, "dead": 3}


def _stream():
    """The current CUDA stream as a raw handle (torch's fast accessor: this is called for every
    launch, and torch.cuda.current_stream() costs ~15 us of Python)."""
    try:
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    except AttributeError:  # pragma: no cover
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "tensorflowraytrace_b200: kernels need CUDA tensors (there is no CPU fallback)")


def _f32(t):
    return t.contiguous() if t.dtype == torch.float32 else t.to(torch.float32).contiguous()


def fp32_peak(packed=1, iters=4096):
    """Measured fp32 FMA throughput (FLOP/s) of the current device: the roofline denominator."""
    out = ctypes.c_double(0.0)
    check(_lib.load().tfrt_fp32_peak(int(packed), int(iters), ctypes.byref(out), _stream()),
          "tfrt_fp32_peak")
    return out.value


def launch_count(reset=False):
    """Kernel launches issued by libtfrt_b200 in this process (bench.py: gpu_launches)."""
    return int(_lib.load().tfrt_launch_count(1 if reset else 0))


def profile_enable(on=True):
    """Bracket every launch of the intersect kernel with CUDA events on its stream."""
    check(_lib.load().tfrt_profile_enable(1 if on else 0), "tfrt_profile_enable")


def profile_read():
    """(total milliseconds, launches) of the intersect kernel since the last read."""
    ms, n = ctypes.c_double(0.0), ctypes.c_longlong(0)
    check(_lib.load().tfrt_profile_read(ctypes.byref(ms), ctypes.byref(n)), "tfrt_profile_read")
    return ms.value, int(n.value)


def faces_forward(vertices, faces):
    """K3 forward.  vertices (V,3) f32 cuda, faces (F,3) int32 cuda -> (tri (9,F), norm (F,3))."""
    _require_cuda(vertices, faces)
    vertices = _f32(vertices)
    F = faces.shape[0]
    tri = torch.empty((9, F), dtype=torch.float32, device=vertices.device)
    norm = torch.empty((F, 3), dtype=torch.float32, device=vertices.device)
    check(_lib.load().tfrt_faces_forward(ptr(vertices), ptr(faces), F,
                                         ptr_array([tri[i] for i in range(9)], 9), ptr(norm),
                                         _stream()), "tfrt_faces_forward")
    return tri, norm


def faces_backward(vertices, faces, update_map, g_tri, g_norm):
    """K3 backward -> g_vertices (V,3) f32."""
    _require_cuda(vertices, faces)
    vertices = _f32(vertices)
    V, F = vertices.shape[0], faces.shape[0]
    g_vertices = torch.empty((V, 3), dtype=torch.float32, device=vertices.device)
    g_list = [None] * 9 if g_tri is None else [_f32(g_tri[i]) for i in range(9)]
    g_norm = None if g_norm is None else _f32(g_norm)
    check(_lib.load().tfrt_faces_backward(ptr(vertices), ptr(faces), ptr(update_map), F, V,
                                          ptr_array(g_list, 9), ptr(g_norm), ptr(g_vertices),
                                          _stream()), "tfrt_faces_backward")
    return g_vertices


DEFAULT_MARGIN_ULPS = 4.0


def pack_triangles(tri, size_eps, margin_ulps=DEFAULT_MARGIN_ULPS):
    """tri: (9,T) f32 cuda (rows xp..z2) -> packed plane coefficients for the intersect kernel.
    margin_ulps: fp32 search margin (include/tfrt_b200.h); the same value must be given to the
    calls that use the packed records."""
    _require_cuda(tri)
    tri = _f32(tri)
    T = tri.shape[1]
    lib = _lib.load()
    packed = torch.empty((lib.tfrt_packed_floats(T),), dtype=torch.float32, device=tri.device)
    check(lib.tfrt_pack_triangles(ptr_array([tri[i] for i in range(9)], 9), T,
                                  ctypes.c_float(size_eps), ctypes.c_float(margin_ulps),
                                  ptr(packed), _stream()),
          "tfrt_pack_triangles")
    return packed


DEFAULT_GRID_MAX_CELLS = 1 << 21


class TriangleGrid:
    """Uniform grid over a packed triangle set (csrc/grid3d.cu), built on the device: the
    optional acceleration structure of the nearest-hit search.  Same hits as the brute-force
    search; not the path the tests/s metric is quoted on."""

    def __init__(self, tri, packed, size_eps, margin_ulps=DEFAULT_MARGIN_ULPS,
                 max_cells=DEFAULT_GRID_MAX_CELLS):
        _require_cuda(tri, packed)
        tri = _f32(tri)
        self.n_tri, self.max_cells = int(tri.shape[1]), int(max_cells)
        lib = _lib.load()
        self.workspace = torch.empty((lib.tfrt_grid_bytes(self.n_tri, self.max_cells),),
                                     dtype=torch.uint8, device=tri.device)
        check(lib.tfrt_grid_build(ptr_array([tri[i] for i in range(9)], 9), ptr(packed),
                                  self.n_tri, ctypes.c_float(size_eps),
                                  ctypes.c_float(margin_ulps), self.max_cells,
                                  ptr(self.workspace), _stream()), "tfrt_grid_build")

    def data_ptr(self):
        return self.workspace.data_ptr()

    def stats(self):
        """dict: dims, cells, entries, large triangles, overflow, typical extent (waits)."""
        out = (ctypes.c_int64 * 8)()
        check(_lib.load().tfrt_grid_stats(ptr(self.workspace), self.n_tri, self.max_cells, out,
                                          _stream()), "tfrt_grid_stats")
        return {"dims": (int(out[0]), int(out[1]), int(out[2])), "cells": int(out[3]),
                "entries": int(out[4]), "large": int(out[5]), "overflow": int(out[6]),
                "typical_extent": out[7] * 1e-6}


def intersect3d(rays, tri, packed, intersect_eps, size_eps, ray_start_eps,
                margin_ulps=DEFAULT_MARGIN_ULPS, diag=None, grid=None):
    """b1 seam: nearest valid hit per ray.  rays (6,N) f32, tri (9,T) f32.  Returns a dict with the
    reference's field names (tfrt/engine.py:1026-1043, without `norm`).  diag: optional (4,) int32
    cuda tensor receiving the re-validation counters (include/tfrt_b200.h).  grid: optional
    TriangleGrid of the same triangles (the search walks it instead of testing them all)."""
    _require_cuda(rays, tri, packed)
    rays, tri = _f32(rays), _f32(tri)
    N, T = rays.shape[1], tri.shape[1]
    dev = rays.device
    hit = torch.empty((N + 1,), dtype=torch.int64, device=dev)
    out = {k: torch.empty((N,), dtype=torch.float32, device=dev)
           for k in ("x", "y", "z", "ray_u", "trig_u", "trig_v")}
    valid = torch.empty((N,), dtype=torch.uint8, device=dev)
    gather_trig = torch.empty((N,), dtype=torch.int32, device=dev)
    check(_lib.load().tfrt_intersect3d(
        ptr_array([rays[i] for i in range(6)], 6), N, ptr_array([tri[i] for i in range(9)], 9),
        ptr(packed), T, ctypes.c_float(intersect_eps), ctypes.c_float(size_eps),
        ctypes.c_float(ray_start_eps), ctypes.c_float(margin_ulps), ptr(hit), ptr(out["x"]),
        ptr(out["y"]), ptr(out["z"]), ptr(valid), ptr(out["ray_u"]), ptr(out["trig_u"]),
        ptr(out["trig_v"]), ptr(gather_trig), ptr(diag),
        ctypes.c_void_p(0 if grid is None else grid.data_ptr()),
        0 if grid is None else grid.max_cells, _stream()), "tfrt_intersect3d")
    out["valid"] = valid.bool()
    out["gather_trig"] = gather_trig.long()
    out["gather_ray"] = torch.arange(N, device=dev)
    return out


def material_table(descriptors, device):
    """descriptors: list of (kind, [coefficients]) -> device byte tensor of TfrtMaterial."""
    arr = (TfrtMaterial * max(len(descriptors), 1))()
    for i, (kind, coef) in enumerate(descriptors):
        arr[i].kind = int(kind)
        for j in range(7):
            arr[i].c[j] = float(coef[j]) if j < len(coef) else 0.0
    raw = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
    return torch.from_numpy(raw).to(device)


# ==============================================================================================
# rays born on the device (csrc/sources.cu)

def sample_uniform(seed, stream, epoch, which, first, n, lo, hi, device):
    """(n,) f32: lo + (hi - lo) U for the global sample indices first .. first + n - 1."""
    out = torch.empty((int(n),), dtype=torch.float32, device=device)
    check(_lib.load().tfrt_sample_uniform(seed, stream, epoch, which, int(first), int(n),
                                          float(lo), float(hi), ptr(out), _stream()),
          "tfrt_sample_uniform")
    return out


def sample_square(seed, stream, epoch, first, n, x_size, y_size, device):
    """RandomUniformSquare on the device: (points (n,2), ranks (n,2))."""
    points = torch.empty((int(n), 2), dtype=torch.float32, device=device)
    ranks = torch.empty((int(n), 2), dtype=torch.float32, device=device)
    check(_lib.load().tfrt_sample_square(seed, stream, epoch, int(first), int(n), float(x_size),
                                         float(y_size), ptr(points), ptr(ranks), _stream()),
          "tfrt_sample_square")
    return points, ranks


SPHERE_MODES = {"random_uniform": 0, "random_lambertian": 1, "static_uniform": 2,
                "static_lambertian": 3}


def sample_sphere(seed, stream, epoch, mode, first, n, total, angular_size, radius, device):
    """Sphere distributions on the device: (points (n,3), ranks (n,2) = (phi, theta))."""
    points = torch.empty((int(n), 3), dtype=torch.float32, device=device)
    ranks = torch.empty((int(n), 2), dtype=torch.float32, device=device)
    check(_lib.load().tfrt_sample_sphere(seed, stream, epoch, SPHERE_MODES[mode], int(first),
                                         int(n), int(total), float(angular_size), float(radius),
                                         ptr(points), ptr(ranks), _stream()),
          "tfrt_sample_sphere")
    return points, ranks


def rays_from_directions(dirs, base, quat, center, ray_length, forward, wavelength):
    """(7,N) f32 block x_start .. z_end, wavelength of a directed 3D source in one launch:
    dirs (N,3) unit vectors about +x, base (N,2 | N,3) or None, quat (w,x,y,z), center (3,)."""
    _require_cuda(dirs, base, wavelength)
    dirs = _f32(dirs)
    n = dirs.shape[0]
    base = None if base is None else _f32(base)
    wavelength = None if wavelength is None else _f32(wavelength)
    rows = torch.empty((7, n), dtype=torch.float32, device=dirs.device)
    q = (ctypes.c_double * 4)(*[float(v) for v in quat])
    c = (ctypes.c_double * 3)(*[float(v) for v in center])
    check(_lib.load().tfrt_rays_from_directions(
        n, ptr(dirs), ptr(base), 0 if base is None else int(base.shape[1]), q, c,
        float(ray_length), 1 if forward else 0, ptr(wavelength), ptr(rows), _stream()),
        "tfrt_rays_from_directions")
    return rows


class TraceWorkspace:
    """Tape and scratch of a whole-loop trace for one (N rays, K iterations) shape: one device
    allocation carved into the arrays of TfrtTrace3DArgs, plus a pinned host mirror of the
    counters.  Reused from trace to trace by the engine, so a steady-state trace allocates
    nothing (engine.OpticalEngine keeps a workspace only while no autograd graph still reads
    its tape)."""

    def __init__(self, N, K, device):
        self.N, self.K, self.device = int(N), int(K), device
        self.generation = 0           # bumped by every trace that binds this workspace
        n1 = max(self.N, 1)
        lib = _lib.load()
        sync_bytes = int(lib.tfrt_trace3d_sync_bytes(self.N, self.K))
        spec = [   # name, dtype, shape   (8-byte types first: every block stays aligned)
            ("cur_d", torch.float64, (6, n1)),
            ("hit", torch.int64, (n1,)),
            ("sync_ws", torch.int64, ((sync_bytes + 7) // 8 + 2,)),
            ("cur", torch.float32, (6, n1)),
            ("tape_f", torch.float32, (K, 9, n1)),
            ("tape_tri", torch.int32, (K, n1)),
            ("tape_outpos", torch.int32, (K, n1)),
            ("live", torch.int32, (K + 1, n1)),
            ("cat_ids", torch.int32, (3, n1)),
            ("counters", torch.int32, ((K + 1) * 4 + 4,)),      # + the 4 diag words
            ("tape_status", torch.uint8, (K, n1)),
        ]
        total, offsets = 0, []
        for _, dtype, shape in spec:
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            offsets.append(total)
            total += (nbytes + 15) // 16 * 16
        self._buffer = torch.empty((total,), dtype=torch.uint8, device=device)
        for (name, dtype, shape), off in zip(spec, offsets):
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
            setattr(self, name, self._buffer[off:off + nbytes].view(dtype).view(shape))
        self.diag = self.counters[(K + 1) * 4:]
        self.counters_pinned = torch.empty(((K + 1) * 4 + 4,), dtype=torch.int32).pin_memory()
        self.copied = torch.cuda.Event()


class Trace3D:
    """One whole-loop trace (b4 seam): binds inputs to a workspace and exposes the histories."""

    def __init__(self, src, tri, norm, category, n_optical, max_iterations, *, packed=None,
                 mat_in=None, mat_out=None, n_in=None, n_out=None, materials=None, n_materials=0,
                 n_table=None, lambda_index=None,
                 intersect_eps=1e-10, size_eps=1e-10, ray_start_eps=1e-10, new_ray_length=1.0,
                 margin_ulps=DEFAULT_MARGIN_ULPS, workspace=None, grid=None):
        _require_cuda(src, tri, norm, category)   # norm may be None: geometric normals
        self.grid = grid
        self.src = _f32(src)          # (7,N): x_start..z_end, wavelength
        self.tri = _f32(tri)          # (9,T)
        self.norm = None if norm is None else _f32(norm)        # (T,3)
        self.category = category if (category.dtype == torch.int32 and category.is_contiguous()) \
            else category.to(torch.int32).contiguous()
        dev = self.src.device
        N, T, K = self.src.shape[1], self.tri.shape[1], int(max_iterations)
        self.N, self.T, self.K = N, T, K
        self.size_eps = float(size_eps)
        self.margin_ulps = float(margin_ulps)
        self.packed = packed if packed is not None else pack_triangles(self.tri, size_eps, margin_ulps)
        self.index_mode = 1 if (materials is not None or n_table is not None) else 0

        def i32(t):
            return None if t is None else (
                t if (t.dtype == torch.int32 and t.is_contiguous()) else t.to(torch.int32).contiguous())

        self.mat_in, self.mat_out = i32(mat_in), i32(mat_out)
        self.n_in = None if n_in is None else _f32(n_in)
        self.n_out = None if n_out is None else _f32(n_out)
        self.materials = materials
        self.n_table = None if n_table is None else n_table.to(torch.float64).contiguous()
        self.lambda_index = i32(lambda_index)
        ws = workspace
        if ws is None or ws.N != N or ws.K != K or ws.device != dev:
            ws = TraceWorkspace(N, K, dev)
        self.ws = ws
        ws.generation += 1
        self.ws_generation = ws.generation
        for name in ("cur", "cur_d", "hit", "tape_f", "tape_tri", "tape_status", "tape_outpos",
                     "live", "cat_ids", "counters", "diag"):
            setattr(self, name, getattr(ws, name))
        self._counters_host = None

        a = TfrtTrace3DArgs()
        a.n_rays, a.n_tri, a.n_optical, a.max_iterations = N, T, int(n_optical), K
        a.intersect_eps, a.size_eps = float(intersect_eps), float(size_eps)
        a.ray_start_eps, a.new_ray_length = float(ray_start_eps), float(new_ray_length)
        a.margin_ulps = self.margin_ulps
        a.index_mode, a.n_materials = self.index_mode, int(n_materials)
        a.n_lambda = 0 if self.n_table is None else int(self.n_table.shape[1])
        src0, stride = self.src.data_ptr(), self.src.stride(0) * 4
        for i in range(7):
            a.src[i] = src0 + i * stride
        tri0, stride = self.tri.data_ptr(), self.tri.stride(0) * 4
        for i in range(9):
            a.tri[i] = tri0 + i * stride
        p = lambda t: 0 if t is None else t.data_ptr()   # noqa: E731
        a.norm, a.category = p(self.norm), self.category.data_ptr()
        a.mat_in, a.mat_out, a.n_in, a.n_out = p(self.mat_in), p(self.mat_out), p(self.n_in), p(self.n_out)
        a.materials, a.n_table, a.lambda_index = p(materials), p(self.n_table), p(self.lambda_index)
        a.packed = self.packed.data_ptr()
        a.grid = 0 if grid is None else grid.data_ptr()
        a.grid_max_cells = 0 if grid is None else grid.max_cells
        n1 = max(N, 1)
        cur0, curd = ws.cur.data_ptr(), ws.cur_d.data_ptr()
        for i in range(6):
            a.cur[i] = cur0 + i * n1 * 4
        a.cur_d = curd
        a.hit, a.tape_f = ws.hit.data_ptr(), ws.tape_f.data_ptr()
        a.tape_tri, a.tape_status = ws.tape_tri.data_ptr(), ws.tape_status.data_ptr()
        a.tape_outpos, a.live = ws.tape_outpos.data_ptr(), ws.live.data_ptr()
        cat0 = ws.cat_ids.data_ptr()
        for i in range(3):
            a.cat_ids[i] = cat0 + i * n1 * 4
        a.sync_ws = (ws.sync_ws.data_ptr() + 15) // 16 * 16
        a.counters, a.diag = ws.counters.data_ptr(), ws.diag.data_ptr()
        self.args = a

    def forward(self):
        """Enqueue the whole trace loop, then the copy of the counters into pinned host memory
        (all asynchronous: nothing here waits for the GPU)."""
        self._counters_host = None
        check(_lib.load().tfrt_trace3d_forward(ctypes.byref(self.args), _stream()),
              "tfrt_trace3d_forward")
        self.ws.counters_pinned.copy_(self.ws.counters, non_blocking=True)
        self.ws.copied.record()
        return self

    @property
    def counters_host(self):
        """(K+1, 4) int32 on the host.  Waits for the trace: the live / history counts are data
        dependent, so this is the one synchronisation of a trace, paid when the histories are
        materialised (not when the trace is enqueued)."""
        if self._counters_host is None:
            self.ws.copied.synchronize()
            both = self.ws.counters_pinned.numpy()
            self._counters_host = both[: (self.K + 1) * 4].reshape(self.K + 1, 4).copy()
            self._diag_host = both[(self.K + 1) * 4:].copy()
        return self._counters_host

    @property
    def diagnostics(self):
        """Re-validation counters of the trace (include/tfrt_b200.h): dict with `revalidated`
        (fp32 winners rejected by the exact fp64 rule and searched again), `gross` (rejections
        far beyond fp32 rounding: a search defect, expected 0) and `research_hits`."""
        self.counters_host
        d = self._diag_host
        return {"revalidated": int(d[0]), "gross": int(d[1]), "research_hits": int(d[2])}

    def live_counts(self):
        """Live rays entering each executed iteration (for the tests/s metric)."""
        return [int(c) for c in self.counters_host[: self.K, 0]]

    def history_count(self, which):
        w = HISTORY[which] if isinstance(which, str) else int(which)
        ch = self.counters_host
        if w == 0:
            return int(ch[1:, 0].sum())
        return int(ch[self.K, w])

    def history(self, which, dead_ray_length=None):
        """Materialise one ray set in the reference's order.  Returns (rays (6,n) f32,
        slot (n,) int32 source index, iteration (n,) int32)."""
        w = HISTORY[which] if isinstance(which, str) else int(which)
        n = self.history_count(w)
        dev = self.src.device
        rays = torch.empty((6, n), dtype=torch.float32, device=dev)
        meta = torch.empty((2, n), dtype=torch.int32, device=dev)
        slot, it = meta[0], meta[1]
        if n:
            ch = np.ascontiguousarray(self.counters_host.reshape(-1))
            drl = -1.0 if dead_ray_length is None else float(dead_ray_length)
            r0 = rays.data_ptr()
            out = (ctypes.c_void_p * 6)(*[r0 + i * n * 4 for i in range(6)])
            check(_lib.load().tfrt_materialize3d(
                ctypes.byref(self.args), ch.ctypes.data_as(ctypes.c_void_p), w,
                ctypes.c_float(drl), out, ptr(slot), ptr(it), _stream()), "tfrt_materialize3d")
        return rays, slot, it

    def backward(self, g_hist, want_src_grad=False, dead_ray_length=None, n_grad_tri=0):
        """b5 seam.  g_hist: dict which -> (6,n) f32 cotangents of `history(which)` rays (or
        None).  Returns (g_tri (9,T), g_norm (T,3), g_src (6,N) or None)."""
        dev = self.src.device
        b = TfrtTrace3DBwdArgs()
        keep = []
        for name, w in HISTORY.items():
            g = g_hist.get(name)
            if g is None:
                continue
            g = _f32(g)
            if g.shape != (6, self.history_count(w)):
                raise ValueError(f"cotangent of {name} rays has shape {tuple(g.shape)}")
            keep.append(g)
            for i in range(6):
                b.g_hist[w][i] = g[i].data_ptr()
        g_acc = torch.empty((self.T * 16,), dtype=torch.float64, device=dev)   # 128-byte rows
        g_tri = torch.empty((9, self.T), dtype=torch.float32, device=dev)
        g_norm = torch.empty((self.T, 3), dtype=torch.float32, device=dev)
        g_src = torch.empty((6, max(self.N, 1)), dtype=torch.float32, device=dev) if want_src_grad else None
        b.dead_ray_length = -1.0 if dead_ray_length is None else float(dead_ray_length)
        b.n_grad_tri = int(n_grad_tri)
        b.g_acc = g_acc.data_ptr()
        for i in range(9):
            b.g_tri_out[i] = g_tri[i].data_ptr()
        b.g_norm_out = g_norm.data_ptr()
        if g_src is not None:
            for i in range(6):
                b.g_src[i] = g_src[i].data_ptr()
        ch = np.ascontiguousarray(self.counters_host.reshape(-1))
        check(_lib.load().tfrt_trace3d_backward(
            ctypes.byref(self.args), ch.ctypes.data_as(ctypes.c_void_p), ctypes.byref(b),
            _stream()), "tfrt_trace3d_backward")
        del keep
        return g_tri, g_norm, (None if g_src is None else g_src[:, : self.N])


# ==============================================================================================
# 2D (K4)

SEG_FIELDS = ("x_start", "y_start", "x_end", "y_end")
ARC_FIELDS = ("x_center", "y_center", "angle_start", "angle_end", "radius")
RAY_FIELDS_2D = ("x_start", "y_start", "x_end", "y_end")


def _rows(t, n):
    return None if t is None else [t[i] for i in range(n)]


def intersect2d(rays, seg, arc, intersect_eps, size_eps, ray_start_eps):
    """2D intersection seam (tfrt/engine.py:523-616).  rays (4,N) f32; seg (4,S) f32 or None;
    arc (5,A) f32 or None.  Returns (seg_result, arc_result) dicts with the reference's field
    names ({} for an absent geometry)."""
    _require_cuda(rays, seg, arc)
    rays = _f32(rays)
    seg = None if seg is None or seg.shape[1] == 0 else _f32(seg)
    arc = None if arc is None or arc.shape[1] == 0 else _f32(arc)
    if seg is None and arc is None:
        raise ValueError("intersect2d: no boundaries")
    N, dev = rays.shape[1], rays.device

    def outputs(present):
        if not present:
            return None, None, None
        return (torch.empty((5, N), dtype=torch.float32, device=dev),
                torch.empty((N,), dtype=torch.uint8, device=dev),
                torch.empty((N,), dtype=torch.int32, device=dev))

    so, sv, sg = outputs(seg is not None)
    ao, av, ag = outputs(arc is not None)
    check(_lib.load().tfrt_intersect2d(
        ptr_array(_rows(rays, 4), 4), N,
        ptr_array(_rows(seg, 4) or [], 4), 0 if seg is None else seg.shape[1],
        ptr_array(_rows(arc, 5) or [], 5), 0 if arc is None else arc.shape[1],
        ctypes.c_float(intersect_eps), ctypes.c_float(size_eps), ctypes.c_float(ray_start_eps),
        ptr_array(_rows(so, 5) or [], 5), ptr(sv), ptr(sg),
        ptr_array(_rows(ao, 5) or [], 5), ptr(av), ptr(ag), _stream()), "tfrt_intersect2d")
    ray_range = torch.arange(N, device=dev)
    seg_res, arc_res = {}, {}
    if seg is not None:
        seg_res = {"x": so[0], "y": so[1], "ray_u": so[2], "segment_u": so[3], "norm": so[4],
                   "valid": sv.bool(), "gather_ray": ray_range, "gather_segment": sg.long()}
    if arc is not None:
        arc_res = {"x": ao[0], "y": ao[1], "ray_u": ao[2], "arc_u": ao[3], "norm": ao[4],
                   "valid": av.bool(), "gather_ray": ray_range, "gather_arc": ag.long()}
    return seg_res, arc_res


class Trace2D:
    """One whole-loop 2D trace: owns the per-(iteration, slot) tape."""

    def __init__(self, src, seg, arc, counts, max_iterations, *, seg_mat=None, arc_mat=None,
                 materials=None, n_materials=0, intersect_eps=1e-10, size_eps=1e-10,
                 ray_start_eps=1e-10, new_ray_length=1.0, reuse=None):
        """src (5,N): x_start,y_start,x_end,y_end,wavelength.  seg (4,S) / arc (5,A) merged
        optical -> stop -> target (either may be None).  counts = dict(seg_optical, seg_stop,
        arc_optical, arc_stop).  seg_mat / arc_mat: (in, out) tensors over the OPTICAL
        boundaries -- int material indices when `materials` is given, else float n values.
        reuse: an earlier Trace2D whose tape may be overwritten (taken over when the shapes
        agree; the caller guarantees no autograd graph still reads it)."""
        _require_cuda(src, seg, arc)
        self.src = _f32(src)
        self.seg = None if seg is None or seg.shape[1] == 0 else _f32(seg)
        self.arc = None if arc is None or arc.shape[1] == 0 else _f32(arc)
        dev = self.src.device
        N, K = self.src.shape[1], int(max_iterations)
        self.N, self.K = N, K
        self.S = 0 if self.seg is None else self.seg.shape[1]
        self.A = 0 if self.arc is None else self.arc.shape[1]
        self.index_mode = 1 if materials is not None else 0
        self.materials = materials
        n1 = max(N, 1)
        if reuse is not None and reuse.N == N and reuse.K == K and reuse.tape_f.device == dev:
            self.tape_f, self.tape_d = reuse.tape_f, reuse.tape_d
            self.tape_hit, self.tape_status, self.live = reuse.tape_hit, reuse.tape_status, reuse.live
            reuse.generation += 1      # the old trace's backward would read the wrong tape
        else:
            # entries of slots that are not live at an iteration are never read (status 255)
            self.tape_f = torch.empty((K, 4, n1), dtype=torch.float32, device=dev)
            self.tape_d = torch.empty((K, 2, n1), dtype=torch.float64, device=dev)
            self.tape_hit = torch.empty((K, n1), dtype=torch.int32, device=dev)
            self.tape_status = torch.empty((K, n1), dtype=torch.uint8, device=dev)
            self.live = torch.empty((K,), dtype=torch.int32, device=dev)
        self.generation = 0
        self._live_host = None

        def conv(pair):
            if pair is None:
                return None, None
            if self.index_mode:
                return (pair[0].to(torch.int32).contiguous(), pair[1].to(torch.int32).contiguous())
            return _f32(pair[0]), _f32(pair[1])

        self._seg_mat, self._arc_mat = conv(seg_mat), conv(arc_mat)
        a = _lib.TfrtTrace2DArgs()
        a.n_rays, a.max_iterations = N, K
        a.n_seg, a.n_seg_optical, a.n_seg_stop = self.S, int(counts.get("seg_optical", 0)), int(counts.get("seg_stop", 0))
        a.n_arc, a.n_arc_optical, a.n_arc_stop = self.A, int(counts.get("arc_optical", 0)), int(counts.get("arc_stop", 0))
        a.intersect_eps, a.size_eps = float(intersect_eps), float(size_eps)
        a.ray_start_eps, a.new_ray_length = float(ray_start_eps), float(new_ray_length)
        a.index_mode, a.n_materials = self.index_mode, int(n_materials)
        for i in range(5):
            a.src[i] = self.src[i].data_ptr()
        for i in range(4):
            a.seg[i] = 0 if self.seg is None else self.seg[i].data_ptr()
        for i in range(5):
            a.arc[i] = 0 if self.arc is None else self.arc[i].data_ptr()
        p = lambda t: 0 if t is None else t.data_ptr()   # noqa: E731
        if self.index_mode:
            a.seg_mat_in, a.seg_mat_out = p(self._seg_mat[0]), p(self._seg_mat[1])
            a.arc_mat_in, a.arc_mat_out = p(self._arc_mat[0]), p(self._arc_mat[1])
        else:
            a.seg_n_in, a.seg_n_out = p(self._seg_mat[0]), p(self._seg_mat[1])
            a.arc_n_in, a.arc_n_out = p(self._arc_mat[0]), p(self._arc_mat[1])
        a.materials = p(materials)
        a.tape_f, a.tape_d = self.tape_f.data_ptr(), self.tape_d.data_ptr()
        a.tape_hit, a.tape_status = self.tape_hit.data_ptr(), self.tape_status.data_ptr()
        a.live_counts = self.live.data_ptr()
        self.args = a

    def forward(self):
        self._live_host = None
        check(_lib.load().tfrt_trace2d_forward(ctypes.byref(self.args), _stream()),
              "tfrt_trace2d_forward")
        return self

    def live_counts(self):
        if self._live_host is None:
            self._live_host = [int(v) for v in self.live.cpu().tolist()]
        return self._live_host

    def backward(self, g_tape, want_src_grad=False):
        """g_tape (K,4,N) f32 -> (g_seg (4,S) | None, g_arc (5,A) | None, g_src (4,N) | None)."""
        dev = self.src.device
        g_tape = _f32(g_tape)
        if tuple(g_tape.shape) != tuple(self.tape_f.shape):
            raise ValueError(f"cotangent of the 2D tape has shape {tuple(g_tape.shape)}")
        g_acc = torch.empty((4 * self.S + 3 * self.A + 1,), dtype=torch.float64, device=dev)
        g_seg = torch.empty((4, self.S), dtype=torch.float32, device=dev) if self.S else None
        g_arc = torch.empty((5, self.A), dtype=torch.float32, device=dev) if self.A else None
        g_src = torch.empty((4, max(self.N, 1)), dtype=torch.float32, device=dev) if want_src_grad else None
        check(_lib.load().tfrt_trace2d_backward(
            ctypes.byref(self.args), ptr(g_tape), ptr(g_acc),
            ptr_array(_rows(g_seg, 4) or [], 4), ptr_array(_rows(g_arc, 5) or [], 5),
            ptr_array(_rows(g_src, 4) or [], 4), _stream()), "tfrt_trace2d_backward")
        return g_seg, g_arc, (None if g_src is None else g_src[:, : self.N])
