"""Shared test helpers: build the float64 oracle and the CUDA inputs from one scene."""

import numpy as np
import torch

from oracle import boundaries as ob
from oracle import engine as oe
from oracle import materials as om
from tensorflowraytrace_b200 import scenes

TRI_FIELDS = ("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2")
RAY_FIELDS = ("x_start", "y_start", "z_start", "x_end", "y_end", "z_end")

# fp32-appropriate epsilons shared by oracle and kernel in parity tests
EPS = dict(intersect_epsilion=1e-10, size_epsilion=1e-5, ray_start_epsilion=1e-4)


def f32_round(x):
    """Round float64 data through float32: the oracle then sees exactly the kernel's inputs."""
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def oracle_from_scene(scene, ray_slice=None, requires_grad=False, eps=EPS, round_inputs=True,
                      fp32_vertices=True, **engine_kwargs):
    """Returns (engine, parameters list).  Inputs are rounded through fp32 (vertices are formed in
    fp64 from fp32-rounded zero points / vectors / parameters, then the *fields* are what they
    are: the kernel computes vertices in fp32, see `round_vertices`)."""
    rnd = f32_round if round_inputs else (lambda a: np.asarray(a, dtype=np.float64))
    src = {}
    for k, v in scene["sources"].items():
        v = v if ray_slice is None else v[ray_slice]
        src[k] = torch.from_numpy(np.ascontiguousarray(rnd(v)))
    n_src = src["x_start"].shape[0]
    src["slot"] = torch.arange(n_src, dtype=torch.float64)   # source-ray id (local)
    params, optical = [], []
    for s in scene["surfaces"]:
        p = torch.from_numpy(rnd(s["parameters"])).clone().requires_grad_(requires_grad)
        params.append(p)
        zero = torch.from_numpy(rnd(s["zero_points"]))
        vec = torch.from_numpy(rnd(s["vectors"]))
        verts = ob.parametric_vertices(zero, vec, p)
        if round_inputs and fp32_vertices:
            # The CUDA path stores vertices as fp32 fields, formed in fp32 (one multiply, one
            # add, no contraction).  Give the oracle exactly those values -- with a
            # straight-through gradient, so autograd still reaches the parameters.
            v32 = (zero.float() + p.detach().float().reshape(-1, 1) * vec.float()).double()
            verts = verts + (v32 - verts.detach())
        fields = ob.fields_from_vertices(verts, s["faces"], s.get("vertex_update_map"))
        F = s["faces"].shape[0]
        fields["mat_in"] = torch.full((F,), s["mat_in"], dtype=torch.int64)
        fields["mat_out"] = torch.full((F,), s["mat_out"], dtype=torch.int64)
        optical.append(fields)

    def static(lst):
        out = []
        for m in lst:
            out.append(ob.fields_from_vertices(torch.from_numpy(rnd(m["vertices"])), m["faces"]))
        return out

    system = oe.OracleSystem3D(
        [src], optical, static(scene["stops"]), static(scene["targets"]),
        materials=[{"n": om.BY_NAME[name]} for name in scene["materials"]], **eps)
    inherit = tuple(k for k in src if k not in RAY_FIELDS)
    engine = oe.OracleEngine3D(system, simple_ray_inheritance=inherit, **engine_kwargs)
    return engine, params


ATOL_FRACTION = 0.1


def rel_err(a, b, floor=ATOL_FRACTION, scale=None):
    """Elementwise relative error with an absolute floor:  max_i |a_i - b_i| / (|b_i| + atol),
    atol = `floor` x the largest |b| (or x `scale`, the coordinate scale of the scene, when the
    array is one coordinate of points whose other coordinates are larger: fp32 fields carry an
    absolute error of an ulp of the scene's coordinates).  `rel_err(a, b) < rtol` is therefore the elementwise bound
    |a_i - b_i| <= rtol |b_i| + rtol x 0.1 x max|b|: every entry is held to the north star's
    relative tolerance against its own magnitude, and entries much smaller than the array's
    scale (a coordinate that happens to be ~0, a gradient entry that cancels) against a tenth
    of that scale."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if not a.size:
        return 0.0
    atol = max(floor * max(np.max(np.abs(b)), scale or 0.0), 1e-30)
    return float(np.max(np.abs(a - b) / (np.abs(b) + atol)))


# ----------------------------------------------------------------------------------------------
# CUDA-vs-oracle comparison of a whole trace

def near_tie_bound(ray_hits, size_eps=EPS["size_epsilion"], p=1e-6):
    """Upper bound on the number of near-tie index mismatches among `ray_hits` ray-surface hits.
    A near tie needs the hit inside the strip where the size_epsilion inflation makes two
    neighbouring triangles both valid: barycentric width size_eps on either side of each of the
    three edges, i.e. a fraction 6 * size_eps of the triangle's area (every one of them counted
    as a mismatch, although only about half flip the winner).  The count is Poisson with that
    mean; the bound is its (1 - p) quantile."""
    import math
    mean = ray_hits * 6.0 * size_eps
    k, term, cdf = 0, math.exp(-mean), math.exp(-mean)
    while 1.0 - cdf > p:
        k += 1
        term *= mean / k
        cdf += term
    return k


def compare_trace(engine, oracle, near_tie_rel=1e-6):
    """Compare the fused CUDA trace held by `engine.last_trace` with a traced oracle engine,
    iteration by iteration and slot by slot.

    Returns a dict: mismatches (total hit-index mismatches), near_ties (how many of those were
    near ties: gap to the runner-up < near_tie_rel relative), exempt (set of slots whose path
    diverged after a near tie), max_rel (worst relative error of S / hit point / ray_u over
    matching slots).  Raises AssertionError on a mismatch that is not a near tie."""
    tr = engine.last_trace
    K, exempt = tr.K, set()
    out = {"mismatches": 0, "near_ties": 0, "max_rel": 0.0, "iterations": 0}
    tape_tri = tr.tape_tri.cpu().numpy()
    tape_status = tr.tape_status.cpu().numpy()
    tape_f = tr.tape_f.cpu().numpy()
    for k in range(min(K, len(oracle.last_intersect_result))):
        ores, oin = oracle.last_intersect_result[k], oracle.last_input_rays[k]
        slots = oin["slot"].numpy().astype(np.int64)
        keep = np.array([s not in exempt for s in slots], dtype=bool)
        ovalid = ores["valid"].numpy()
        otri = np.where(ovalid, ores["gather_trig"].numpy(), -1)
        # every non-exempt oracle-live slot must be live in the kernel at iteration k
        assert np.all(tape_status[k][slots[keep]] != 255), f"iteration {k}: live sets differ"
        ktri = tape_tri[k][slots]
        bad = keep & (ktri != otri)
        if bad.any():
            idx = np.nonzero(bad)[0]
            rays = {f: v[idx] for f, v in oin.items()}
            excl = torch.from_numpy(np.where(otri[idx] >= 0, otri[idx], 0))
            second = oracle.optical_system.second_best(rays, excl).numpy()
            best = np.where(ovalid[idx], ores["ray_u"].numpy()[idx], np.inf)
            for j, i in enumerate(idx):
                both_inf = not np.isfinite(best[j]) and not np.isfinite(second[j])
                gap = 0.0 if both_inf else abs(second[j] - best[j]) / max(abs(best[j]), 1e-30)
                # the kernel's pick must be (numerically) as near as the oracle's
                assert np.isfinite(gap) and gap < near_tie_rel, (
                    f"iteration {k} slot {slots[i]}: kernel tri {ktri[i]} vs oracle {otri[i]}, "
                    f"gap to runner-up {gap:.3e} is not a near tie")
                exempt.add(int(slots[i]))
            out["mismatches"] += int(bad.sum())
            out["near_ties"] += int(bad.sum())
        good = keep & ~bad & ovalid
        if good.any():
            s = slots[good]
            for c, f in enumerate(("x_start", "y_start", "z_start")):
                out["max_rel"] = max(out["max_rel"], rel_err(tape_f[k][c][s], oin[f].numpy()[good]))
            for c, f in enumerate(("x", "y", "z")):
                out["max_rel"] = max(out["max_rel"], rel_err(tape_f[k][3 + c][s], ores[f].numpy()[good]))
            out["max_rel"] = max(out["max_rel"], rel_err(tape_f[k][6][s], ores["ray_u"].numpy()[good]))
        out["iterations"] += 1
    out["exempt"] = exempt
    return out


def compare_sets(engine_set, oracle_set, exempt, fields, slot_field="slot", engine_slots=None):
    """Worst relative error over `fields` between an engine history set and the oracle's, matched
    by source slot, skipping exempt slots.  engine_slots: (n,) source slot of each engine row."""
    oslots = oracle_set[slot_field].numpy().astype(np.int64)
    eslots = np.asarray(engine_slots, dtype=np.int64)
    okeep = np.array([s not in exempt for s in oslots], dtype=bool)
    ekeep = np.array([s not in exempt for s in eslots], dtype=bool)
    assert np.array_equal(oslots[okeep], eslots[ekeep]), "history order / membership differs"
    worst = 0.0
    for f in fields:
        e = engine_set[f].detach().cpu().numpy()[ekeep]
        o = oracle_set[f].detach().numpy()[okeep]
        worst = max(worst, rel_err(e, o))
    return worst


# ----------------------------------------------------------------------------------------------
# 2D

SEG_FIELDS = ("x_start", "y_start", "x_end", "y_end")
ARC_FIELDS = ("x_center", "y_center", "angle_start", "angle_end", "radius")
RAY_FIELDS_2D = ("x_start", "y_start", "x_end", "y_end")


def oracle2d_from_scene(scene, requires_grad=False, eps=EPS, **engine_kwargs):
    """(OracleEngine2D, parameter tensors) for a scenes.*_2d scene; inputs rounded through fp32.
    Parameters (when requires_grad) are the geometry columns of the optical boundaries, in the
    order scenes.build_engine_2d(parametric=True) returns them."""
    from oracle import engine2d as oe2
    src = {k: torch.from_numpy(np.ascontiguousarray(f32_round(v))) for k, v in scene["sources"].items()}
    src["slot"] = torch.arange(src["x_start"].shape[0], dtype=torch.float64)
    params = []

    def make(entries, kind):
        out = []
        for e in entries:
            names = SEG_FIELDS if kind == "segments" else ARC_FIELDS
            data = f32_round(np.asarray(e[kind]).reshape(-1, len(names)))
            d = {}
            for i, k in enumerate(names):
                col = torch.from_numpy(np.ascontiguousarray(data[:, i]))
                if requires_grad and "mat_in" in e:
                    col = col.clone().requires_grad_(True)
                    params.append(col)
                d[k] = col
            if "mat_in" in e:
                n = data.shape[0]
                d["mat_in"] = torch.full((n,), e["mat_in"], dtype=torch.int64)
                d["mat_out"] = torch.full((n,), e["mat_out"], dtype=torch.int64)
            out.append(d)
        return out

    # same creation order as scenes.build_engine_2d: optical segments, then optical arcs
    opt_seg = make(scene["optical_segments"], "segments")
    opt_arc = make(scene["optical_arcs"], "arcs")
    system = oe2.OracleSystem2D(
        [src], opt_seg, opt_arc, make(scene["stop_segments"], "segments"),
        make(scene["stop_arcs"], "arcs"), make(scene["target_segments"], "segments"),
        make(scene["target_arcs"], "arcs"),
        materials=[{"n": om.BY_NAME[name]} for name in scene["materials"]], **eps)
    engine = oe2.OracleEngine2D(system, simple_ray_inheritance=("wavelength", "slot"),
                                **engine_kwargs)
    return engine, params
