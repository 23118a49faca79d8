"""
torch.autograd wrappers around the CUDA kernels: the replacement of the reference's
`tf.GradientTape` seam (tfrt/optimizer.py:216-220).

FacesFunction : vertices -> triangle fields + unit normals (K3, tfrt/boundaries.py:890-923)
TraceFunction : (triangle fields, normals, source rays) -> ray histories of the whole
                OpticalEngine.ray_trace loop (K1 + K2, tfrt/engine.py:2311-2330)

Gradient semantics follow the reference exactly: nothing flows through
argmin / masks / barycentrics / material indices; `vertex_update_map` blocks per-(face,corner)
contributions.
"""

import torch

from . import kernels


class FacesFunction(torch.autograd.Function):
    """vertices (V,3) f32 -> (tri (9,F), norm (F,3)); faces (F,3) int32; update_map (F,3) uint8."""

    @staticmethod
    def forward(ctx, vertices, faces, update_map):
        vertices = vertices.contiguous()
        tri, norm = kernels.faces_forward(vertices, faces)
        ctx.save_for_backward(vertices)
        ctx.faces, ctx.update_map = faces, update_map
        return tri, norm

    @staticmethod
    def backward(ctx, g_tri, g_norm):
        (vertices,) = ctx.saved_tensors
        g_vertices = kernels.faces_backward(
            vertices, ctx.faces, ctx.update_map,
            None if g_tri is None else g_tri.contiguous(),
            None if g_norm is None else g_norm.contiguous())
        return g_vertices, None, None


def run_trace(tri, norm, src, plan, workspace=None):
    """Enqueue the whole-loop trace described by `plan` (asynchronous) and leave the
    kernels.Trace3D object (tape, counters) in plan.trace.  No autograd here: the graph node is
    created by TraceFunction when the histories are materialised."""
    tr = kernels.Trace3D(
        src.detach(), tri.detach(), None if norm is None else norm.detach(), plan.category,
        plan.n_optical, plan.max_iterations,
        packed=plan.packed, mat_in=plan.mat_in, mat_out=plan.mat_out, n_in=plan.n_in,
        n_out=plan.n_out, materials=plan.materials, n_materials=plan.n_materials,
        n_table=plan.n_table, lambda_index=plan.lambda_index,
        intersect_eps=plan.intersect_eps, size_eps=plan.size_eps,
        ray_start_eps=plan.ray_start_eps, new_ray_length=plan.new_ray_length,
        margin_ulps=plan.margin_ulps, workspace=workspace, grid=plan.grid)
    tr.forward()
    plan.trace = tr
    return tr


class TraceFunction(torch.autograd.Function):
    """Autograd node of a whole-loop trace.  Inputs: tri (9,T), norm (T,3) or None (geometric
    normals evaluated in the kernel from the corners, their gradient then arrives through
    `tri`), src (7,N) [x_start..z_end, wavelength], and a `plan` (TracePlan) carrying the
    non-tensor configuration.  If plan.trace is unset the trace runs here; otherwise (the
    engine enqueued it earlier with run_trace) only the histories are materialised.  Outputs:
    the four history ray sets (6,n_which) in the reference's order -- active, stopped, finished,
    dead -- with (6,0) for sets the engine does not compile."""

    @staticmethod
    def forward(ctx, tri, norm, src, plan):
        tr = plan.trace if plan.trace is not None else run_trace(tri, norm, src, plan)
        outs, plan.slots, plan.iters = [], {}, {}
        for name in ("active", "stopped", "finished", "dead"):
            if name in plan.compile:
                rays, slot, it = tr.history(
                    name, dead_ray_length=plan.dead_ray_length if name == "dead" else None)
            else:
                rays = torch.empty((6, 0), dtype=torch.float32, device=src.device)
                slot = it = torch.empty((0,), dtype=torch.int32, device=src.device)
            outs.append(rays)
            plan.slots[name], plan.iters[name] = slot, it
        ctx.plan = plan
        ctx.trace = tr
        ctx.lease = plan.lease      # keeps the workspace from being reused while this graph lives
        ctx.src_shape = src.shape
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_active, g_stopped, g_finished, g_dead):
        plan, tr = ctx.plan, ctx.trace
        if tr.ws.generation != tr.ws_generation:
            raise RuntimeError(
                "tensorflowraytrace_b200: the tape of this trace was overwritten by a later trace "
                "that reused its workspace")
        g_hist = {}
        for name, g in (("active", g_active), ("stopped", g_stopped), ("finished", g_finished),
                        ("dead", g_dead)):
            if g is not None and name in plan.compile and g.shape[1] > 0:
                g_hist[name] = g.contiguous()
        want_src = ctx.needs_input_grad[2]
        g_tri, g_norm, g_src = tr.backward(g_hist, want_src_grad=want_src,
                                           dead_ray_length=plan.dead_ray_length,
                                           n_grad_tri=plan.n_grad_tri)
        g_src_full = None
        if want_src:
            g_src_full = torch.zeros(ctx.src_shape, dtype=torch.float32, device=g_tri.device)
            g_src_full[:6] = g_src
        return (g_tri if ctx.needs_input_grad[0] else None,
                g_norm if (ctx.needs_input_grad[1] and tr.norm is not None) else None,
                g_src_full, None)


class WorkspaceLease:
    """Marks a trace workspace as in use: held (strongly) by the plan of the engine's pending
    trace and by the autograd node of its histories, and weakly by the engine.  The engine
    reuses the workspace for its next trace only when the lease has died, i.e. when no autograd
    graph can still read the tape."""

    __slots__ = ("workspace", "__weakref__")

    def __init__(self, workspace):
        self.workspace = workspace


class TracePlan:
    """Non-tensor configuration of one fused trace (filled by OpticalEngine.ray_trace)."""

    def __init__(self, **kw):
        self.category = kw["category"]
        self.n_optical = kw["n_optical"]
        self.max_iterations = kw["max_iterations"]
        self.mat_in, self.mat_out = kw.get("mat_in"), kw.get("mat_out")
        self.n_in, self.n_out = kw.get("n_in"), kw.get("n_out")
        self.materials, self.n_materials = kw.get("materials"), kw.get("n_materials", 0)
        self.n_table, self.lambda_index = kw.get("n_table"), kw.get("lambda_index")
        self.margin_ulps = kw.get("margin_ulps", kernels.DEFAULT_MARGIN_ULPS)
        self.grid = kw.get("grid")     # optional kernels.TriangleGrid (acceleration mode)
        self.lease = None
        self.intersect_eps = kw["intersect_eps"]
        self.size_eps = kw["size_eps"]
        self.ray_start_eps = kw["ray_start_eps"]
        self.new_ray_length = kw["new_ray_length"]
        self.dead_ray_length = kw.get("dead_ray_length")
        self.compile = set(kw["compile"])
        self.packed = kw.get("packed")   # tfrt_pack_triangles output (built when None)
        self.n_grad_tri = kw.get("n_grad_tri", 0)   # leading triangles that need cotangents
        self.trace = None
        self.slots, self.iters = {}, {}


class Trace2DFunction(torch.autograd.Function):
    """Whole-loop 2D trace (K4).  Inputs: seg (4,S) or None, arc (5,A) or None, src (5,N), and a
    `plan` dict: either {"trace": kernels.Trace2D already run (the engine enqueues it in
    ray_trace), "lease": WorkspaceLease} or the non-tensor configuration (kernels.Trace2D keyword
    arguments + counts + max_iterations), in which case the trace runs here.  Output: the tape
    (K,4,N) f32 -- ray start and projected end per (iteration, source ray) -- from which the
    engine gathers the reference-ordered history sets with plain torch indexing (so their
    cotangents scatter back into this tape's cotangent)."""

    @staticmethod
    def forward(ctx, seg, arc, src, plan):
        tr = plan.get("trace")
        if tr is None:
            tr = kernels.Trace2D(src, seg, arc, plan["counts"], plan["max_iterations"],
                                 **plan["kwargs"])
            tr.forward()
            plan["trace"] = tr
        ctx.trace = tr
        ctx.generation = tr.generation
        ctx.lease = plan.get("lease")   # keeps the tape from being reused while this graph lives
        ctx.src_shape = src.shape
        return tr.tape_f[:, :, : tr.N]

    @staticmethod
    def backward(ctx, g_tape):
        tr = ctx.trace
        if tr.generation != ctx.generation:
            raise RuntimeError(
                "tensorflowraytrace_b200: the tape of this 2D trace was overwritten by a later "
                "trace that reused it")
        want_src = ctx.needs_input_grad[2]
        g_full = g_tape
        if g_full.shape[2] != tr.tape_f.shape[2]:     # N == 0 padding
            g_full = torch.zeros_like(tr.tape_f)
        g_seg, g_arc, g_src = tr.backward(g_full.contiguous(), want_src_grad=want_src)
        g_src_full = None
        if want_src:
            g_src_full = torch.zeros(ctx.src_shape, dtype=torch.float32, device=g_tape.device)
            g_src_full[:4] = g_src
        return (g_seg if ctx.needs_input_grad[0] else None,
                g_arc if ctx.needs_input_grad[1] else None, g_src_full, None)
