"""Round-2 parity tests of the CUDA path (through the C ABI and the public engine API):

* the configurations nobody had checked: BASELINE config 2 at its real 65 536 rays, the BACKWARD
  pass of config 5 (16.7 M rays x 99 374 triangles) on a ray subsample against oracle autograd,
  and the all-reduced gradient on real NCCL (2 GPUs) against the single-GPU gradient;
* the fp64 re-validation of the fp32 search's winner: its counters (no gross failure on any
  config), and the exact re-search forced on a random soup;
* arbitrary material callables on the fused path;
* scale awareness: a lens in millimetres and a lens far from the origin.

Tolerances are the north star's: hit indices bit-exact except near ties (< 1e-6 relative gap to
the runner-up), points / directions 1e-5, gradients 1e-4 -- elementwise relative with the
absolute floor documented in helpers.rel_err."""

import os
import sys

import numpy as np
import pytest
import torch

import helpers
from oracle import engine as oe
from tensorflowraytrace_b200 import kernels, scenes

pytestmark = pytest.mark.gpu

POINT_TOL = 1e-5
GRAD_TOL = 1e-4
RAYS = helpers.RAY_FIELDS
DEV = "cuda"


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype, device=DEV)


def check_diagnostics(trace, what):
    """No winner of the fp32 search may fail the exact rule by more than rounding (a search
    defect); rounding-level rejections (searched again exactly) must stay rare."""
    d = trace.diagnostics
    live = max(sum(trace.live_counts()), 1)
    assert d["gross"] == 0, (what, d)
    assert d["revalidated"] <= max(16, 1e-2 * live), (what, d, live)
    return d


# --------------------------------------------------------------------------- config 2, full size
def test_cfg2_at_its_real_size_forward_parity():
    """BASELINE config 2 exactly as named: 65 536 rays (golden-spiral point source) x 2 030
    triangles x 3 iterations, every ray compared with the oracle."""
    scene = scenes.config2()
    assert scene["sources"]["x_start"].shape[0] == 65536
    engine, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    engine.ray_trace(3)
    assert engine.last_trace.N == 65536 and engine.last_trace.T == 2030
    oracle, _ = helpers.oracle_from_scene(scene)
    oracle.ray_trace(3)
    cmp = helpers.compare_trace(engine, oracle)
    assert cmp["iterations"] == 3
    assert cmp["max_rel"] < POINT_TOL, cmp
    # expected near ties: a ray within fp32 rounding of an edge shared by two triangles sees both
    # at (numerically) the same distance; compare_trace raises on any other disagreement
    assert cmp["mismatches"] == cmp["near_ties"]
    slots = engine.last_history_slots["finished"].cpu().numpy().astype(np.int64)
    worst = helpers.compare_sets(engine.finished_rays, oracle.finished_rays, cmp["exempt"],
                                 RAYS + ("wavelength",), engine_slots=slots)
    assert worst < POINT_TOL, worst
    check_diagnostics(engine.last_trace, "cfg2")


# --------------------------------------------------------------------------- config 5 backward
def test_full_size_cfg5_subsample_gradient():
    """Backward at BASELINE config 5's launch shape (16 777 216 rays x 99 374 triangles): the
    gradient of the loss restricted to a random subsample of rays equals the oracle's gradient
    for those rays (vertex gradients projected on the surface parameters)."""
    scene = scenes.config5()
    engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_active_rays=False)
    engine.ray_trace(3)
    n = 16777216
    assert engine.last_trace.N == n
    rng = np.random.default_rng(55)
    idx = np.sort(rng.choice(n, size=4096, replace=False))
    chosen = torch.zeros(n, dtype=torch.bool, device=DEV)
    chosen[torch.from_numpy(idx).to(DEV)] = True
    err = scenes.imaging_error(scene)(engine)
    mask = chosen[engine.last_history_slots["finished"].long()]
    grads = torch.autograd.grad(err[mask].sum(), params)
    check_diagnostics(engine.last_trace, "cfg5")
    oracle, oparams = helpers.oracle_from_scene(scene, ray_slice=idx, requires_grad=True)
    oracle.ray_trace(3)
    ofin = oracle.finished_rays
    goal = scene["goal_scale"] * ofin["rank"]
    oerr = (ofin["y_end"] - goal[:, 0]) ** 2 + (ofin["z_end"] - goal[:, 1]) ** 2
    # a subsampled ray whose path differs after a near tie would change the sums: compare the
    # finished sets first (same membership => same rays)
    eslots = engine.last_history_slots["finished"][mask].cpu().numpy().astype(np.int64)
    oslots = idx[ofin["slot"].numpy().astype(np.int64)]
    assert np.array_equal(eslots, oslots)
    total, ototal = float(err[mask].sum().detach()), float(oerr.sum().detach())
    assert abs(total - ototal) < 1e-4 * abs(ototal)
    ograds = torch.autograd.grad(oerr.sum(), oparams)
    for g, og_ in zip(grads, ograds):
        assert float(og_.abs().max()) > 0
        assert helpers.rel_err(g.cpu().numpy(), og_.numpy()) < GRAD_TOL


# --------------------------------------------------------------------------- NCCL, 2 GPUs
def _nccl_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    for p in (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
              os.path.dirname(os.path.abspath(__file__))):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from tensorflowraytrace_b200 import optimizer, parallel
    from tensorflowraytrace_b200 import scenes as sc
    parallel.init_from_env()
    scene = sc.lens_optimize_config(ray_count=40000, step_count=9, name="nccl")
    engine, params, _ = sc.build_engine(scene, eps=helpers.EPS, ray_shard=(rank, world),
                                        compile_active_rays=False)
    opt = optimizer.SGD_Optimizer(engine, params, sc.imaging_error(scene), 3)
    grads, err = opt.raw_gradient()
    grads = [torch.zeros_like(p) if g is None else g.contiguous() for g, p in zip(grads, params)]
    stats = torch.stack([err.sum().detach().to(torch.float32),
                         torch.tensor(float(err.numel()), device=err.device)])
    parallel.allreduce_sum_(grads + [stats])          # ONE flat NCCL all-reduce
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), *[g.cpu().numpy() for g in grads],
                 stats=stats.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                    reason="needs 2 GPUs")
def test_nccl_allreduced_gradient_equals_single_gpu(tmp_path):
    """Rays sharded over 2 ranks (one process per GPU), parameter gradients summed by one NCCL
    all-reduce: equal to the single-GPU gradient of the whole ray set to 1e-4."""
    import torch.multiprocessing as mp
    port = 29650 + os.getpid() % 200
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "reduced.npz"))
    scene = scenes.lens_optimize_config(ray_count=40000, step_count=9, name="nccl")
    engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_active_rays=False)
    engine.ray_trace(3)
    err = scenes.imaging_error(scene)(engine)
    want = torch.autograd.grad(err.sum(), params)
    for i, w in enumerate(want):
        assert helpers.rel_err(got[f"arr_{i}"], w.cpu().numpy()) < GRAD_TOL
    assert abs(float(got["stats"][0]) - float(err.sum())) < 1e-4 * abs(float(err.sum()))
    assert int(got["stats"][1]) == err.numel()


# --------------------------------------------------------------------------- re-validation
def _soup(n_rays, n_tri, seed):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, size=(n_tri, 3)) + np.array([3.0, 0, 0])
    tri = np.concatenate([c + 0.4 * rng.normal(size=(n_tri, 3)) for _ in range(3)], axis=1)
    S = rng.uniform(-0.3, 0.3, size=(n_rays, 3))
    D = np.stack([np.ones(n_rays), rng.uniform(-0.5, 0.5, n_rays), rng.uniform(-0.5, 0.5, n_rays)], 1)
    rays = np.concatenate([S, S + D], axis=1)
    return helpers.f32_round(rays.T.copy()), helpers.f32_round(tri.T.copy())


@pytest.mark.parametrize("margin_ulps", [0.0, 8.0, 2.0e4])
def test_rejected_winners_are_searched_again_exactly(margin_ulps):
    """The fp32 search proposes, the fp64 rule decides.  With an absurd search margin (2e4 ulps:
    the search band is percent-wide) most proposals lie outside the exact triangle; every one
    must be rejected and searched again in fp64, and the final answer must equal the oracle's
    for every ray (no near-tie exemption needed: the exact search is fp64)."""
    rays, tri = _soup(3000, 700, 4242)
    eps = dict(intersect_epsilion=1e-10, size_epsilion=1e-9, ray_start_epsilion=1e-6)
    d_tri = dev(tri)
    packed = kernels.pack_triangles(d_tri, eps["size_epsilion"], margin_ulps)
    diag = torch.zeros(4, dtype=torch.int32, device=DEV)
    got = kernels.intersect3d(dev(rays), d_tri, packed, eps["intersect_epsilion"],
                              eps["size_epsilion"], eps["ray_start_epsilion"], margin_ulps, diag)
    sys3 = oe.OracleSystem3D(
        {k: torch.from_numpy(rays[i]) for i, k in enumerate(RAYS)},
        {**{k: torch.from_numpy(tri[i]) for i, k in enumerate(helpers.TRI_FIELDS)},
         "norm": torch.zeros(tri.shape[1], 3, dtype=torch.float64)}, **eps)
    want = sys3.intersect(sys3._amalgamated_sources)
    ovalid = want["valid"].numpy()
    kvalid = got["valid"].cpu().numpy()
    ktri = np.where(kvalid, got["gather_trig"].cpu().numpy(), -1)
    otri = np.where(ovalid, want["gather_trig"].numpy(), -1)
    d = diag.cpu().numpy()
    assert d[1] == 0, d                                   # nothing beyond rounding ...
    if margin_ulps > 1e3:
        assert d[0] > 50, d                               # ... although many proposals fall
        assert np.array_equal(ktri, otri)                 # and the answers are the oracle's
    else:
        bad = np.nonzero(ktri != otri)[0]                 # fp32 ordering of near ties only
        if bad.size:
            second = sys3.second_best({k: v[bad] for k, v in sys3._amalgamated_sources.items()},
                                      torch.from_numpy(np.maximum(otri[bad], 0))).numpy()
            best = np.where(ovalid[bad], want["ray_u"].numpy()[bad], np.inf)
            assert np.all(np.abs(second - best) / np.maximum(np.abs(best), 1e-30) < 1e-6)
    ok = (ktri == otri) & ovalid
    for k in ("x", "y", "z", "ray_u"):
        assert helpers.rel_err(got[k].cpu().numpy()[ok], want[k].numpy()[ok]) < POINT_TOL, k


def test_degenerate_and_grazing_winners_never_react():
    """A zero-area triangle, a triangle containing the ray's line (den = 0) and a ray that
    starts on a triangle: none of them may come out as a valid hit, whatever the search
    proposes (the fp64 rule has the last word)."""
    tri = np.array([
        [0.5, 0, 0, 0.5, 0, 0, 0.5, 0, 0],        # degenerate
        [0, -1, 0, 2, -1, 0, 1, 1, 0],            # contains the ray's line (parallel)
        [0, -1, -1, 0, 1, -1, 0, 0, 1],           # the ray starts on this one
        [2, -1, -1, 2, 1, -1, 2, 0, 1],           # the real hit at x = 2
    ], dtype=np.float64).T
    rays = np.array([[0, 0, 0, 1, 0, 0]], dtype=np.float64).T
    d_tri = dev(tri)
    for margin in (0.0, 8.0, 1e6):
        packed = kernels.pack_triangles(d_tri, 1e-5, margin)
        diag = torch.zeros(4, dtype=torch.int32, device=DEV)
        got = kernels.intersect3d(dev(rays), d_tri, packed, 1e-10, 1e-5, 1e-4, margin, diag)
        assert bool(got["valid"][0]) and int(got["gather_trig"][0]) == 3, (margin, got)
        assert abs(float(got["x"][0]) - 2.0) < 1e-6


# --------------------------------------------------------------------------- diagnostics, full size
@pytest.mark.parametrize("name,iters", [("config3", 3), ("config4", 16)])
def test_revalidation_counters_at_full_size(name, iters):
    scene = getattr(scenes, name)()
    engine, _, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_active_rays=False)
    engine.ray_trace(iters)
    d = check_diagnostics(engine.last_trace, name)
    assert d["research_hits"] <= d["revalidated"]


# --------------------------------------------------------------------------- arbitrary materials
def test_user_material_callable_runs_on_the_fused_path():
    """Any {"n": callable} is a material in the reference (tfrt/operation.py:265-267).  A user
    lambda (Cauchy dispersion) must ride the fused loop and give what the projection-at-a-time
    path gives, histories and parameter gradients alike."""
    from tensorflowraytrace_b200 import engine as eng, operation

    class Tag(operation.RayOperation):
        pass

    scene = scenes.lens_optimize_config(ray_count=3000, step_count=5, name="usermat")
    rng = np.random.default_rng(3)
    scene["sources"]["wavelength"] = rng.choice([450.0, 550.0, 650.0], size=3000)
    fused, params, _ = scenes.build_engine(scene, eps=helpers.EPS)
    cauchy = lambda wl: 1.45 + 4.0e3 / (wl * wl)            # noqa: E731
    fused.optical_system.materials = [{"n": lambda wl: 1.0 + 0.0 * wl}, {"n": cauchy}]
    assert not fused.optical_system.materials_have_descriptors() and fused._fusable()
    errf = scenes.imaging_error(scene)
    fused.ray_trace(3)
    assert fused.last_trace.n_table is not None and fused.last_trace.n_table.shape == (2, 3)
    g_f = torch.autograd.grad(errf(fused).sum(), params)
    unf = eng.OpticalEngine(3, [operation.StandardReaction(), Tag()],
                            simple_ray_inheritance={"wavelength", "rank"})
    unf.optical_system = fused.optical_system
    unf.optical_system.update()
    unf.ray_trace(3)
    assert not unf._fusable()
    for name in ("finished", "active"):
        a, b = getattr(fused, name + "_rays"), getattr(unf, name + "_rays")
        assert a["x_start"].shape == b["x_start"].shape and a["x_start"].shape[0] > 0
        for f in RAYS + ("wavelength",):
            assert helpers.rel_err(b[f].detach().cpu().numpy(), a[f].detach().cpu().numpy()) < POINT_TOL, (name, f)
    g_u = torch.autograd.grad(errf(unf).sum(), params)
    for a, b in zip(g_f, g_u):
        assert float(a.abs().max()) > 0
        assert helpers.rel_err(b.cpu().numpy(), a.cpu().numpy()) < GRAD_TOL
    # dispersion really entered: the three wavelengths land in different places
    fin = fused.finished_rays
    r = torch.sqrt(fin["y_end"] ** 2 + fin["z_end"] ** 2)
    means = [float(r[fin["wavelength"] == w].mean()) for w in (450.0, 550.0, 650.0)]
    assert abs(means[0] - means[2]) > 1e-4


# --------------------------------------------------------------------------- scale awareness
def _scaled(scene, factor, offset):
    """The same scene in other units / elsewhere: x -> factor * x + offset."""
    out = dict(scene)
    off = np.asarray(offset, dtype=np.float64)
    out["sources"] = dict(scene["sources"])
    for i, a in enumerate("xyz"):
        for e in ("_start", "_end"):
            out["sources"][a + e] = scene["sources"][a + e] * factor + off[i]
    out["surfaces"] = []
    for s in scene["surfaces"]:
        s2 = dict(s)
        s2["zero_points"] = s["zero_points"] * factor + off
        s2["parameters"] = s["parameters"] * factor
        out["surfaces"].append(s2)
    for kind in ("stops", "targets"):
        out[kind] = [dict(m, vertices=m["vertices"] * factor + off) for m in scene[kind]]
    return out


@pytest.mark.parametrize("factor,offset", [(100.0, (0, 0, 0)), (1.0, (24.0, -16.0, 8.0))])
def test_scene_in_millimetres_and_far_from_the_origin(factor, offset):
    """The search's allowances follow the scene's coordinates (ulps of the triangle's largest
    coordinate), the exact rule is applied in fp64: a lens in millimetres (|P| ~ 100) and one far
    from the origin (|P| ~ 30, ten times the lens) trace like the unit lens -- same epsilons, no
    gross rejection, every disagreement with the oracle a near tie."""
    base = scenes.config2(ray_count=8192, step_count=9)
    scene = _scaled(base, factor, offset)
    engine, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    engine.ray_trace(3)
    oracle, _ = helpers.oracle_from_scene(scene)
    oracle.ray_trace(3)
    cmp = helpers.compare_trace(engine, oracle)
    assert cmp["iterations"] == 3
    assert cmp["max_rel"] < POINT_TOL, cmp
    assert cmp["mismatches"] == cmp["near_ties"]
    d = check_diagnostics(engine.last_trace, f"x{factor}+{offset}")
    # nearly every ray still reaches the target
    assert engine.last_trace.history_count("finished") > 0.95 * 8192, d


# --------------------------------------------------------------------------- post-link steps, A/B
@pytest.mark.parametrize("case", ["soup_8rays", "soup_short_items", "cfg5_mesh", "grazing"])
def test_rescheduled_search_kernel_is_bit_identical_to_the_ptxas_build(case, monkeypatch):
    """The build re-orders the search kernel's loop bodies and sets operand-reuse flags after
    ptxas (_sass_sched.py, _reuse_patch.py).  The same sources as ptxas wrote them are built
    beside the library (libtfrt_b200_ptxas.so); every output of the b1 seam -- hit index, hit
    point, ray_u, barycentrics, validity, re-validation counters -- must be bit-identical between
    the two builds, on inputs that go through every kernel instance (8 / 4 / 2 rays per thread,
    item queue or not, split triangle ranges with the atomic merge) and through the rare path
    (near ties, grazing rays)."""
    from tensorflowraytrace_b200 import _build, _lib
    if not os.path.exists(_build.PTXAS_LIB_PATH):
        if not _build.can_build():
            pytest.skip("comparison build not present and nvcc not available")
        _build.build_unpatched()
    plain = _lib.open_library(_build.PTXAS_LIB_PATH)
    shipped = _lib.load()
    assert plain.tfrt_source_hash() == shipped.tfrt_source_hash()

    rng = np.random.default_rng(7)
    if case == "cfg5_mesh":
        scene = scenes.config5(ray_count=300000)
        system = scenes.build_engine(scene, eps=helpers.EPS)[0].optical_system
        tri = system.merged_device_arrays()[0].detach()
        rays = torch.stack([system.sources[k] for k in RAYS]).float()
    else:
        n_rays, n_tri = {"soup_8rays": (400000, 3000), "soup_short_items": (5000, 40000),
                         "grazing": (200000, 2048)}[case]
        centre = rng.uniform(-1, 1, (n_tri, 3))
        tri_np = np.concatenate([centre + rng.normal(0, 0.08, (n_tri, 3)) for _ in range(3)], 1).T
        start = rng.uniform(-1, 1, (n_rays, 3)) + np.array([-3.0, 0, 0])
        end = rng.uniform(-1, 1, (n_rays, 3)) + np.array([3.0, 0, 0])
        if case == "grazing":
            # aim every ray at a point on an edge of some triangle: ties and edge rounding
            pick = rng.integers(0, n_tri, n_rays)
            a, b = tri_np[0:3, pick].T, tri_np[3:6, pick].T
            w = rng.uniform(0, 1, (n_rays, 1))
            end = start + 2.0 * ((1 - w) * a + w * b - start)
        tri = dev(tri_np)
        rays = dev(np.concatenate([start, end], 1).T)

    def run(lib):
        monkeypatch.setattr(_lib, "_lib", lib)
        packed = kernels.pack_triangles(tri, helpers.EPS["size_epsilion"])
        diag = torch.zeros(4, dtype=torch.int32, device=DEV)
        out = kernels.intersect3d(rays, tri, packed, helpers.EPS["intersect_epsilion"],
                                  helpers.EPS["size_epsilion"], helpers.EPS["ray_start_epsilion"],
                                  diag=diag)
        torch.cuda.synchronize()
        return out, diag

    a, da = run(shipped)
    b, db = run(plain)
    assert int(a["valid"].sum()) > 0.05 * rays.shape[1]
    for k in a:
        assert torch.equal(a[k], b[k]), (case, k)
    assert torch.equal(da, db)


# --------------------------------------------------------------------------- tape reuse
def _lens_engine():
    scene = scenes.lens_optimize_config(ray_count=4096, step_count=5, name="reuse")
    engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_active_rays=False,
                                            compile_finished_rays=True)
    return scene, engine, params


def test_trace_workspace_is_reused_only_when_no_graph_can_read_it():
    """The engine keeps the tape of a trace for the next one (no allocation in steady state), but a
    trace whose histories were read hangs on an autograd graph that will read the tape in its
    backward pass: the next trace must then get a tape of its own, the old graph must still give
    the right gradient, and a tape that is overwritten under a live graph must be refused."""
    scene, engine, params = _lens_engine()
    errf = scenes.imaging_error(scene)
    K = scene["max_iterations"]

    engine.ray_trace(K)
    first = engine.last_trace
    loss_a = errf(engine).sum()                      # reads finished_rays: graph A holds the tape
    engine.optical_system.update()
    engine.ray_trace(K)
    second = engine.last_trace
    assert second.ws is not first.ws                 # A is alive: no reuse
    loss_b = errf(engine).sum()
    g_a = torch.autograd.grad(loss_a, params)
    g_b = torch.autograd.grad(loss_b, params)
    for a, b in zip(g_a, g_b):       # same system, same rays: same gradient (atomic order aside)
        assert helpers.rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-6
    _, fresh_engine, fresh_params = _lens_engine()
    fresh_engine.ray_trace(K)
    g_ref = torch.autograd.grad(errf(fresh_engine).sum(), fresh_params)
    for a, r in zip(g_a, g_ref):
        assert helpers.rel_err(a.cpu().numpy(), r.cpu().numpy()) < 1e-6

    # histories never read: the pending trace holds nothing, the next trace takes its tape over
    del loss_a, loss_b, g_a, g_b
    engine.optical_system.update()
    engine.ray_trace(K)
    ws = engine.last_trace.ws
    engine.optical_system.update()
    engine.ray_trace(K)
    assert engine.last_trace.ws is ws
    fin = engine.finished_rays
    assert fin["x_end"].shape[0] == \
        engine.last_trace.history_count(kernels.HISTORY["finished"]) > 0

    # a tape overwritten under a live graph is detected, not differentiated
    loss = errf(engine).sum()
    engine.last_trace.ws.generation += 1
    with pytest.raises(RuntimeError, match="overwritten"):
        torch.autograd.grad(loss, params)


def test_2d_tape_reuse_follows_the_same_rule():
    scene = scenes.config1()
    engine, params, _ = scenes.build_engine_2d(scene, eps=helpers.EPS, parametric=True)
    K = scene["max_iterations"]

    def loss_of(e):
        fin = e.finished_rays
        return ((fin["y_end"] - 0.3) ** 2).sum() + 0.1 * (fin["x_end"] * fin["y_start"]).sum()

    engine.ray_trace(K)
    first = engine.last_trace
    loss_a = loss_of(engine)                         # graph A reads the tape in its backward
    engine.optical_system.update()
    engine.ray_trace(K)
    assert engine.last_trace.tape_f.data_ptr() != first.tape_f.data_ptr()
    g_a = torch.autograd.grad(loss_a, params, allow_unused=True)
    g_b = torch.autograd.grad(loss_of(engine), params, allow_unused=True)
    for a, b in zip(g_a, g_b):
        assert (a is None and b is None) or \
            helpers.rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-6
    del loss_a, g_a, g_b
    # unread traces hand their tape on
    engine.optical_system.update()
    engine.ray_trace(K)
    tape = engine.last_trace.tape_f.data_ptr()
    engine.optical_system.update()
    engine.ray_trace(K)
    assert engine.last_trace.tape_f.data_ptr() == tape
    assert engine.finished_rays["x_end"].shape[0] > 0
    # overwritten under a live graph: refused
    loss = loss_of(engine)
    engine.last_trace.generation += 1
    with pytest.raises(RuntimeError, match="overwritten"):
        torch.autograd.grad(loss, params, allow_unused=True)
