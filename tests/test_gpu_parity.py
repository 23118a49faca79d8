"""Parity of the CUDA path (through the C ABI and the public engine API) against the float64
oracle.  Tolerances are the north star's: hit indices bit-exact except near ties (< 1e-6
relative gap to the runner-up), points / directions <= 1e-5 relative, gradients <= 1e-4."""

import numpy as np
import pytest
import torch

import helpers
from oracle import boundaries as ob
from oracle import engine as oe
from oracle import geometry as og
from oracle import materials as om
from tensorflowraytrace_b200 import kernels, scenes

pytestmark = pytest.mark.gpu

POINT_TOL = 1e-5      # relative, intersection points / ray ends (north star)
GRAD_TOL = 1e-4       # relative, parameter / vertex gradients (north star)
RAYS = helpers.RAY_FIELDS
DEV = "cuda"


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype, device=DEV)


# --------------------------------------------------------------------------- no CPU fallback
def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kernels.faces_forward(torch.zeros(3, 3), torch.zeros(1, 3, dtype=torch.int32))


# --------------------------------------------------------------------------- K3
@pytest.mark.parametrize("n_side", [1, 7])
def test_faces_forward_backward(n_side):
    rng = np.random.default_rng(10 + n_side)
    from tensorflowraytrace_b200 import mesh_tools
    pts, faces = mesh_tools.hexagonal_mesh(1.0, n_side)
    pts = helpers.f32_round(pts + 0.05 * rng.normal(size=pts.shape))
    umap = rng.uniform(size=(faces.shape[0], 3)) > 0.3
    v = dev(pts).requires_grad_(True)
    from tensorflowraytrace_b200.autograd import FacesFunction
    tri, norm = FacesFunction.apply(v, dev(faces, torch.int32), dev(umap, torch.uint8))
    ov = torch.from_numpy(pts).requires_grad_(True)
    of = ob.fields_from_vertices(ov, faces, umap)
    for i, k in enumerate(helpers.TRI_FIELDS):
        assert np.array_equal(tri[i].detach().cpu().numpy().astype(np.float64), of[k].detach().numpy())
    assert helpers.rel_err(norm.detach().cpu().numpy(), of["norm"].detach().numpy()) < 1e-6
    g_tri = rng.normal(size=(9, faces.shape[0]))
    g_norm = rng.normal(size=(faces.shape[0], 3))
    (tri * dev(g_tri)).sum().add((norm * dev(g_norm)).sum()).backward()
    oloss = sum((of[k] * torch.from_numpy(helpers.f32_round(g_tri[i]))).sum()
                for i, k in enumerate(helpers.TRI_FIELDS))
    oloss = oloss + (of["norm"] * torch.from_numpy(helpers.f32_round(g_norm))).sum()
    (og_v,) = torch.autograd.grad(oloss, ov)
    assert helpers.rel_err(v.grad.cpu().numpy(), og_v.numpy()) < GRAD_TOL


# --------------------------------------------------------------------------- b1 seam
def _soup(n_rays, n_tri, seed):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, size=(n_tri, 3)) + np.array([3.0, 0, 0])
    tri = np.concatenate([c + 0.4 * rng.normal(size=(n_tri, 3)) for _ in range(3)], axis=1)  # (T,9)
    S = rng.uniform(-0.3, 0.3, size=(n_rays, 3))
    D = np.stack([np.ones(n_rays), rng.uniform(-0.5, 0.5, n_rays), rng.uniform(-0.5, 0.5, n_rays)], 1)
    rays = np.concatenate([S, S + D], axis=1)
    return helpers.f32_round(rays.T.copy()), helpers.f32_round(tri.T.copy())


@pytest.mark.parametrize("n_rays,n_tri", [(1, 1), (5, 3), (129, 255), (1000, 257), (4097, 1500)])
def test_intersect3d_random_soup(n_rays, n_tri):
    rays, tri = _soup(n_rays, n_tri, 100 + n_rays)
    eps = helpers.EPS
    d_tri = dev(tri)
    packed = kernels.pack_triangles(d_tri, eps["size_epsilion"])
    got = kernels.intersect3d(dev(rays), d_tri, packed, eps["intersect_epsilion"],
                              eps["size_epsilion"], eps["ray_start_epsilion"])
    sys3 = oe.OracleSystem3D(
        {k: torch.from_numpy(rays[i]) for i, k in enumerate(RAYS)},
        {**{k: torch.from_numpy(tri[i]) for i, k in enumerate(helpers.TRI_FIELDS)},
         "norm": torch.zeros(n_tri, 3, dtype=torch.float64)}, **eps)
    want = sys3.intersect(sys3._amalgamated_sources)
    ovalid = want["valid"].numpy()
    kvalid = got["valid"].cpu().numpy()
    ktri = np.where(kvalid, got["gather_trig"].cpu().numpy(), -1)
    otri = np.where(ovalid, want["gather_trig"].numpy(), -1)
    bad = np.nonzero(ktri != otri)[0]
    if bad.size:   # every disagreement must be a near tie
        second = sys3.second_best({k: v[bad] for k, v in sys3._amalgamated_sources.items()},
                                  torch.from_numpy(np.maximum(otri[bad], 0))).numpy()
        best = np.where(ovalid[bad], want["ray_u"].numpy()[bad], np.inf)
        gap = np.abs(second - best) / np.maximum(np.abs(best), 1e-30)
        assert np.all(gap < 1e-6), (bad, gap)
    ok = (ktri == otri) & ovalid
    assert ok.sum() > 0 or n_rays < 10
    for k in ("x", "y", "z", "ray_u", "trig_u", "trig_v"):
        assert helpers.rel_err(got[k].cpu().numpy()[ok], want[k].numpy()[ok]) < POINT_TOL, k
    assert got["gather_ray"].tolist() == list(range(n_rays))


def test_intersect3d_empty_ray_set():
    rays, tri = _soup(4, 9, 1)
    d_tri = dev(tri)
    packed = kernels.pack_triangles(d_tri, 1e-5)
    got = kernels.intersect3d(dev(rays[:, :0]), d_tri, packed, 1e-10, 1e-5, 1e-4)
    assert got["valid"].shape == (0,) and got["x"].shape == (0,)


def test_intersect3d_tie_takes_lowest_index_and_degenerates_are_skipped():
    # two coincident triangles + one degenerate (zero-area) + one parallel to the ray
    tri = np.array([
        [1, -1, -1, 1, 1, -1, 1, 0, 1],      # hit at x=1
        [1, -1, -1, 1, 1, -1, 1, 0, 1],      # exact duplicate: tie -> index 0 wins
        [0.5, 0, 0, 0.5, 0, 0, 0.5, 0, 0],   # degenerate
        [0, -1, 0, 2, -1, 0, 1, 1, 0],       # contains the ray's line (parallel, det = 0)
    ], dtype=np.float64).T
    rays = np.array([[0, 0, 0, 1, 0, 0]], dtype=np.float64).T
    d_tri = dev(tri)
    packed = kernels.pack_triangles(d_tri, 1e-5)
    got = kernels.intersect3d(dev(rays), d_tri, packed, 1e-10, 1e-5, 1e-4)
    assert bool(got["valid"][0]) and int(got["gather_trig"][0]) == 0
    assert abs(float(got["x"][0]) - 1.0) < 1e-6 and abs(float(got["ray_u"][0]) - 1.0) < 1e-6
    # a ray starting ON the triangle must not re-hit it (ray_start_epsilion)
    rays2 = np.array([[1, 0, 0, 2, 0, 0]], dtype=np.float64).T
    got2 = kernels.intersect3d(dev(rays2), d_tri, packed, 1e-10, 1e-5, 1e-4)
    assert not bool(got2["valid"][0])


# --------------------------------------------------------------------------- whole trace
def _trace_both(scene, **engine_kwargs):
    engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS, **engine_kwargs)
    engine.ray_trace(scene["max_iterations"])
    oracle, oparams = helpers.oracle_from_scene(scene, **engine_kwargs)
    oracle.ray_trace(scene["max_iterations"])
    return engine, params, oracle, oparams


def test_trace_cfg2_forward_parity():
    scene = scenes.config2(ray_count=8192)
    engine, _, oracle, _ = _trace_both(scene)
    cmp = helpers.compare_trace(engine, oracle)
    assert cmp["iterations"] == 3
    assert cmp["max_rel"] < POINT_TOL, cmp
    assert engine.last_trace.live_counts()[: len(oracle.live_counts)] == oracle.live_counts or cmp["exempt"]
    fin, ofin = engine.finished_rays, oracle.finished_rays
    slots = engine.last_trace  # noqa: F841
    worst = helpers.compare_sets(fin, ofin, cmp["exempt"], RAYS + ("wavelength",),
                                 engine_slots=_slots(engine, "finished"))
    assert worst < POINT_TOL, worst
    act, oact = engine.active_rays, oracle.active_rays
    worst = helpers.compare_sets(act, oact, cmp["exempt"], RAYS, engine_slots=_slots(engine, "active"))
    assert worst < POINT_TOL, worst
    # output ray directions (unit vectors) to 1e-5
    d = np.stack([(fin[a + "_end"] - fin[a + "_start"]).detach().cpu().numpy() for a in "xyz"], 1)
    od = np.stack([(ofin[a + "_end"] - ofin[a + "_start"]).numpy() for a in "xyz"], 1)
    keep = np.array([s not in cmp["exempt"] for s in _slots(engine, "finished")])
    okeep = np.array([s not in cmp["exempt"] for s in ofin["slot"].numpy().astype(int)])
    d, od = d[keep], od[okeep]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    od /= np.linalg.norm(od, axis=1, keepdims=True)
    assert np.max(np.abs(d - od)) < POINT_TOL


def _slots(engine, name):
    """Source slot of every row of a history set of the last fused trace."""
    _, slot, _ = engine.last_trace.history(name, dead_ray_length=(
        engine.dead_ray_length if name == "dead" and engine.dead_ray_length else None))
    return slot.cpu().numpy().astype(np.int64)


def test_trace_gradient_parity_lens():
    """fwd+bwd to the surface parameters (cfg3 topology, sub-sampled) vs oracle autograd."""
    scene = scenes.lens_optimize_config(ray_count=6000, step_count=9, name="cfg3s")
    engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS)
    engine.ray_trace(3)
    err = scenes.imaging_error(scene)(engine)
    grads = torch.autograd.grad(err.sum(), params)
    oracle, oparams = helpers.oracle_from_scene(scene, requires_grad=True)
    oracle.ray_trace(3)
    ofin = oracle.finished_rays
    goal = scene["goal_scale"] * ofin["rank"]
    oerr = (ofin["y_end"] - goal[:, 0]) ** 2 + (ofin["z_end"] - goal[:, 1]) ** 2
    cmp = helpers.compare_trace(engine, oracle)
    assert cmp["max_rel"] < POINT_TOL
    keep = torch.tensor([int(s) not in cmp["exempt"] for s in ofin["slot"].tolist()])
    ograds = torch.autograd.grad(oerr[keep].sum(), oparams)
    if cmp["exempt"]:   # drop the same rays from the CUDA loss
        ekeep = torch.tensor([int(s) not in cmp["exempt"] for s in _slots(engine, "finished")], device=DEV)
        engine.optical_system.update()     # fresh graph (the reference updates inside the tape)
        engine.ray_trace(3)
        err = scenes.imaging_error(scene)(engine)
        grads = torch.autograd.grad(err[ekeep].sum(), params)
    assert abs(float(err.sum()) - float(oerr.sum())) / float(oerr.sum()) < 1e-3 or cmp["exempt"]
    for g, og_ in zip(grads, ograds):
        assert float(og_.abs().max()) > 0
        assert helpers.rel_err(g.cpu().numpy(), og_.numpy()) < GRAD_TOL


def test_trace_light_guide_tir_multiwavelength():
    """cfg4 topology, reduced: TIR bounces in an acrylic guide, six wavelengths, dead and
    finished rays, 8 iterations with a decaying live set."""
    scene = scenes.config4(ray_count=3000, theta_res=16, z_res=12, max_iterations=8)
    engine, _, oracle, _ = _trace_both(scene, compile_dead_rays=True, compile_stopped_rays=True)
    cmp = helpers.compare_trace(engine, oracle)
    assert cmp["max_rel"] < POINT_TOL, cmp
    assert cmp["iterations"] >= 4
    live = engine.last_trace.live_counts()
    assert live[0] == 3000 and live[-1] < live[0]
    for name in ("finished", "dead", "active"):
        e, o = getattr(engine, name + "_rays"), getattr(oracle, name + "_rays")
        if not bool(o) or o["x_start"].shape[0] == 0:
            assert not bool(e) or e["x_start"].shape[0] == 0
            continue
        worst = helpers.compare_sets(e, o, cmp["exempt"], RAYS + ("wavelength",),
                                     engine_slots=_slots(engine, name))
        assert worst < POINT_TOL, (name, worst)
    assert len(set(engine.finished_rays["wavelength"].cpu().tolist())) == 6


def _mirror_scene(n=2000):
    """Two facing mirrors (n_in = 0 on the reflective side) + a target: value-mode materials."""
    from tensorflowraytrace_b200 import mesh_tools
    rng = np.random.default_rng(77)
    pv, pf = mesh_tools.plane_mesh((0.0, 0.0, 2.0), 3.0, axis=2)
    qv, qf = mesh_tools.plane_mesh((0.0, 0.0, -2.0), 3.0, axis=2)
    tv, tf_ = mesh_tools.plane_mesh((4.0, 0.0, 0.0), 10.0, axis=0)
    starts = np.stack([np.full(n, -2.5), rng.uniform(-0.5, 0.5, n), rng.uniform(-0.5, 0.5, n)], 1)
    dirs = np.stack([np.ones(n), rng.uniform(-0.2, 0.2, n), rng.uniform(0.6, 1.4, n)], 1)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    return dict(starts=starts, dirs=dirs, mirrors=[(pv, pf), (qv, qf)], target=(tv, tf_))


def test_value_mode_mirrors_and_stop():
    from tensorflowraytrace_b200 import boundaries, engine as eng, operation, sources
    sc = _mirror_scene()
    src_np = scenes.make_rays(sc["starts"], sc["dirs"], 500.0)
    src_np = {k: helpers.f32_round(v) for k, v in src_np.items()}
    # CUDA side through the public API, StandardReaction("value")
    src = sources.ManualSource(3).feed(src_np)
    mirrors = [boundaries.ManualTriangleBoundary(
        mesh=boundaries.TriangleMesh(v, f), material_dict={"n_in": 0.0, "n_out": 1.0})
        for v, f in sc["mirrors"]]
    target = boundaries.ManualTriangleBoundary(mesh=boundaries.TriangleMesh(*sc["target"]))
    sv, sf = __import__("tensorflowraytrace_b200").mesh_tools.plane_mesh((1.0, 0.0, 0.0), 0.4, axis=0)
    stop = boundaries.ManualTriangleBoundary(mesh=boundaries.TriangleMesh(sv, sf))
    system = eng.OpticalSystem3D(**helpers.EPS)
    system.optical, system.stops, system.targets, system.sources = mirrors, [stop], [target], [src]
    system.update()
    engine = eng.OpticalEngine(3, [operation.StandardReaction("value")], compile_stopped_rays=True,
                               compile_dead_rays=True, dead_ray_length=3.0)
    engine.optical_system = system
    engine.validate_system()
    engine.ray_trace(6)
    # oracle
    osrc = {k: torch.from_numpy(v) for k, v in src_np.items()}
    osrc["slot"] = torch.arange(len(src_np["x_start"]), dtype=torch.float64)
    oopt = []
    for v, f in sc["mirrors"]:
        d = ob.fields_from_vertices(torch.from_numpy(helpers.f32_round(v)), f)
        d["n_in"] = torch.zeros(f.shape[0], dtype=torch.float64)
        d["n_out"] = torch.ones(f.shape[0], dtype=torch.float64)
        oopt.append(d)
    osys = oe.OracleSystem3D([osrc], oopt,
                             [ob.fields_from_vertices(torch.from_numpy(helpers.f32_round(sv)), sf)],
                             [ob.fields_from_vertices(torch.from_numpy(helpers.f32_round(sc["target"][0])), sc["target"][1])],
                             **helpers.EPS)
    oracle = oe.OracleEngine3D(osys, compile_stopped_rays=True, compile_dead_rays=True,
                               dead_ray_length=3.0, refractive_index_type="value",
                               simple_ray_inheritance=("wavelength", "slot"))
    oracle.ray_trace(6)
    cmp = helpers.compare_trace(engine, oracle)
    assert cmp["max_rel"] < POINT_TOL
    seen = 0
    for name in ("finished", "stopped", "dead", "active"):
        e, o = getattr(engine, name + "_rays"), getattr(oracle, name + "_rays")
        assert bool(e) == bool(o), name           # compiled sets exist even when empty
        n_o = o["x_start"].shape[0] if bool(o) else 0
        n_e = e["x_start"].shape[0] if bool(e) else 0
        assert n_e == n_o or cmp["exempt"], name
        if n_o:
            seen += 1
            assert helpers.compare_sets(e, o, cmp["exempt"], RAYS, engine_slots=_slots(engine, name)) < POINT_TOL
    assert seen >= 3
    assert max(engine.last_trace.live_counts()[1:]) > 0     # reflections happened


def test_unfused_single_pass_path_matches_fused():
    """A custom no-op RayOperation forces the projection-at-a-time path (intersection kernel +
    torch bookkeeping); histories must equal the fused loop's."""
    from tensorflowraytrace_b200 import engine as eng, operation

    class Tag(operation.RayOperation):
        pass

    scene = scenes.config2(ray_count=4096, step_count=5)
    fused, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    fused.ray_trace(3)
    unf = eng.OpticalEngine(3, [operation.StandardReaction(), Tag()],
                            simple_ray_inheritance={"wavelength"})
    unf.optical_system = fused.optical_system
    unf.ray_trace(3)
    assert not unf._fusable() and fused._fusable()
    for name in ("finished", "active"):
        a, b = getattr(fused, name + "_rays"), getattr(unf, name + "_rays")
        assert a["x_start"].shape == b["x_start"].shape
        for f in RAYS + ("wavelength",):
            assert helpers.rel_err(b[f].detach().cpu().numpy(), a[f].detach().cpu().numpy()) < POINT_TOL, (name, f)


def test_unfused_gradient_matches_fused():
    from tensorflowraytrace_b200 import engine as eng, operation

    class Tag(operation.RayOperation):
        pass

    scene = scenes.lens_optimize_config(ray_count=2000, step_count=5, name="g")
    fused, params, _ = scenes.build_engine(scene, eps=helpers.EPS)
    errf = scenes.imaging_error(scene)
    fused.ray_trace(3)
    g_f = torch.autograd.grad(errf(fused).sum(), params)
    unf = eng.OpticalEngine(3, [operation.StandardReaction(), Tag()],
                            simple_ray_inheritance={"wavelength", "rank"})
    unf.optical_system = fused.optical_system
    unf.optical_system.update()
    unf.ray_trace(3)
    g_u = torch.autograd.grad(errf(unf).sum(), params)
    for a, b in zip(g_f, g_u):
        assert helpers.rel_err(b.cpu().numpy(), a.cpu().numpy()) < 1e-3   # fp32 torch glue vs fp64 kernel


# --------------------------------------------------------------------------- full-size properties
def test_full_size_cfg3_properties():
    """1M rays x 20k triangles (BASELINE config 3): size-independent properties.
    (a) every ray is accounted for; (b) finished directions are unit length; (c) the trace is
    deterministic; (d) tracing two ray shards and concatenating equals tracing the whole set
    (the multi-GPU sharding invariant); (e) parameter gradients of the shards add up."""
    scene = scenes.config3()
    errf = scenes.imaging_error(scene)
    whole, params, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_active_rays=False,
                                           compile_dead_rays=True, compile_stopped_rays=True)
    whole.ray_trace(3)
    tr = whole.last_trace
    ch = tr.counters_host
    n = tr.N
    assert n == 1048576 and tr.T == 20174
    assert ch[3, 0] + ch[3, 1] + ch[3, 2] + ch[3, 3] == n
    fin = whole.finished_rays
    assert fin["x_start"].shape[0] == ch[3, 2] > 0.9 * n
    d = torch.stack([fin[a + "_end"] - fin[a + "_start"] for a in "xyz"])
    assert torch.all(fin["x_end"] == 8.0) or float((fin["x_end"] - 8.0).abs().max()) < 1e-5
    assert float(d[0].min()) > 0
    g_whole = torch.autograd.grad(errf(whole).sum(), params)
    ref = {k: fin[k].clone() for k in RAYS}
    whole.optical_system.update()
    whole.ray_trace(3)
    for k in RAYS:
        assert torch.equal(whole.finished_rays[k], ref[k])       # (c)
    # (d): histories are iteration-major, so shards are compared per source ray (slot), which is
    # also how a multi-GPU caller reassembles them
    def per_slot(engine, offset):
        slot = engine.last_history_slots["finished"].long() + offset
        return slot, {k: engine.finished_rays[k].detach() for k in RAYS}

    wslot, wrays = per_slot(whole, 0)
    dense = {k: torch.full((n,), float("nan"), device=DEV) for k in RAYS}
    for k in RAYS:
        dense[k][wslot] = wrays[k]
    g_parts, seen = None, 0
    from tensorflowraytrace_b200.parallel import shard_range
    for r in range(2):
        sh, p, _ = scenes.build_engine(scene, eps=helpers.EPS, ray_shard=(r, 2), compile_active_rays=False)
        sh.ray_trace(3)
        sslot, srays = per_slot(sh, shard_range(n, r, 2)[0])
        for k in RAYS:
            assert torch.equal(srays[k], dense[k][sslot])                    # (d)
        seen += sslot.numel()
        g = torch.autograd.grad(errf(sh).sum(), p)
        g_parts = g if g_parts is None else [a + b for a, b in zip(g_parts, g)]
    assert seen == wslot.numel()
    for a, b in zip(g_whole, g_parts):
        assert helpers.rel_err(b.cpu().numpy(), a.cpu().numpy()) < GRAD_TOL   # (e)


def test_full_size_cfg5_subsample_parity():
    """BASELINE config 5 at full size (16 777 216 rays x 99 374 triangles, the headline bench
    workload, through the persistent / split / atomic-merge kernel path) checked against the
    oracle on a random subsample of the SAME rays: the oracle traces 3 072 of them against the
    full mesh, and their hit indices, hit points and ray parameters must agree slot for slot."""
    from types import SimpleNamespace
    scene = scenes.config5()
    engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_active_rays=False)
    engine.ray_trace(3)
    tr = engine.last_trace
    assert tr.N == 16777216 and tr.T == 99374
    ch = tr.counters_host
    assert int(ch[3].sum()) == tr.N                      # every ray accounted for
    rng = np.random.default_rng(99)
    idx = np.sort(rng.choice(tr.N, size=3072, replace=False))
    didx = torch.from_numpy(idx).to(DEV)
    sub = SimpleNamespace(last_trace=SimpleNamespace(
        K=tr.K, tape_tri=tr.tape_tri[:, didx], tape_status=tr.tape_status[:, didx],
        tape_f=tr.tape_f[:, :, didx]))
    oracle, _ = helpers.oracle_from_scene(scene, ray_slice=idx)
    oracle.ray_trace(3)
    cmp = helpers.compare_trace(sub, oracle)
    assert cmp["iterations"] == 3
    assert cmp["max_rel"] < POINT_TOL, cmp
    # near ties only (anything else raises); 3 072 rays cross the two lens surfaces
    assert cmp["mismatches"] <= helpers.near_tie_bound(2 * 3072)


def test_full_size_cfg4_subsample_parity():
    """BASELINE config 4 at full size (4 194 304 rays x 50 002 triangles x 16 iterations of TIR
    bounces, six wavelengths, decaying live set): a random subsample of the same rays traced by
    the oracle must agree slot for slot at every iteration."""
    from types import SimpleNamespace
    scene = scenes.config4()
    engine, _, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_active_rays=False)
    engine.ray_trace(16)
    tr = engine.last_trace
    assert tr.N == 4194304 and tr.T == 50002
    live = tr.live_counts()
    assert live[0] == tr.N and live[-1] < live[0]
    rng = np.random.default_rng(7)
    idx = np.sort(rng.choice(tr.N, size=1024, replace=False))
    didx = torch.from_numpy(idx).to(DEV)
    sub = SimpleNamespace(last_trace=SimpleNamespace(
        K=tr.K, tape_tri=tr.tape_tri[:, didx], tape_status=tr.tape_status[:, didx],
        tape_f=tr.tape_f[:, :, didx]))
    oracle, _ = helpers.oracle_from_scene(scene, ray_slice=idx)
    oracle.ray_trace(16)
    cmp = helpers.compare_trace(sub, oracle)
    assert cmp["iterations"] >= 8
    assert cmp["max_rel"] < POINT_TOL, cmp


def test_full_size_cfg3_subsample_gradient():
    """Backward at BASELINE config 3's launch shape (1 048 576 rays): the gradient of the loss
    restricted to a random subsample of rays equals the oracle's gradient for those rays."""
    scene = scenes.config3()
    engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_active_rays=False)
    engine.ray_trace(3)
    rng = np.random.default_rng(5)
    idx = np.sort(rng.choice(1048576, size=4096, replace=False))
    chosen = torch.zeros(1048576, dtype=torch.bool, device=DEV)
    chosen[torch.from_numpy(idx).to(DEV)] = True
    err = scenes.imaging_error(scene)(engine)
    mask = chosen[engine.last_history_slots["finished"].long()]
    grads = torch.autograd.grad(err[mask].sum(), params)
    oracle, oparams = helpers.oracle_from_scene(scene, ray_slice=idx, requires_grad=True)
    oracle.ray_trace(3)
    ofin = oracle.finished_rays
    goal = scene["goal_scale"] * ofin["rank"]
    oerr = (ofin["y_end"] - goal[:, 0]) ** 2 + (ofin["z_end"] - goal[:, 1]) ** 2
    assert int(mask.sum()) == oerr.shape[0]
    assert abs(float(err[mask].sum().detach()) - float(oerr.sum().detach())) < 1e-4 * float(oerr.sum().detach())
    ograds = torch.autograd.grad(oerr.sum(), oparams)
    for g, og_ in zip(grads, ograds):
        assert float(og_.abs().max()) > 0
        assert helpers.rel_err(g.cpu().numpy(), og_.numpy()) < GRAD_TOL
