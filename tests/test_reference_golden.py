"""Pins against the REFERENCE'S OWN CODE.

tests/golden/ref_*.npz were produced by tests/golden/make_reference_golden.py, which imports the
unmodified reference (`tfrt.geometry / engine / operation / materials / boundaries / sources`) in the
build container with a stand-in `tensorflow` module (oracle/tf_shim: each tf op mapped to the
IEEE-equivalent torch float64 op, GradientTape -> torch autograd).  So every line of reference Python
on the hot path ran as written to make these vectors.

CPU (`not gpu`): the oracle restatement reproduces them to float64 rounding -- functions, whole traces
(history sets in reference order), the vertices -> fields map, and the GradientTape gradients of
dev/simple_3d_optimize.py's loss.  GPU: the CUDA path reproduces them through the public API within
the north-star tolerances (hit membership exact, points / directions 1e-5, gradients 1e-4)."""

import os
import sys

import numpy as np
import pytest
import torch

import helpers
from oracle import geometry as og
from oracle import materials as om

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_golden as mrg  # noqa: E402  (imports neither the shim nor the reference)

SCENES = mrg.golden_scenes()
SETS = ("active", "finished", "stopped", "dead")


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", f"{name}.npz")))


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def close(a, b, tol=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return bool(np.all(np.abs(a - b) <= tol * (1.0 + np.abs(b))))


# ------------------------------------------------------------------------------------ functions

def test_functions_line_triangle_intersect():
    ref = load("ref_functions")
    r, t = ref["tri_in_rays"], ref["tri_in_tris"]
    args = [T(r[:, i]) for i in range(6)] + [T(t[:, i]) for i in range(9)]
    got = og.line_triangle_intersect(*args, 1e-10)
    valid = ref["tri_valid"]
    assert np.array_equal(got[3].numpy(), valid)
    assert valid.shape == (t.shape[0], r.shape[0])          # (triangles, rays): geometry.py:238-251
    assert not valid[0].any() and not valid[3, 1]            # the degenerate / parallel cases
    for i, key in enumerate(("x", "y", "z", "valid", "ray_u", "trig_u", "trig_v")):
        assert close(got[i].numpy(), ref[f"tri_{key}"]), key    # garbage-where-invalid included


def test_functions_snells_law_3d():
    ref = load("ref_functions")
    a = ref["snell3_in"]
    got = og.snells_law_3D(*[T(a[:, i]) for i in range(6)], T(a[:, 6:9]), T(a[:, 9]), T(a[:, 10]),
                           torch.tensor(1.5, dtype=torch.float64))
    assert close(np.stack([g.numpy() for g in got], axis=1), ref["snell3_out"])


def test_functions_2d():
    ref = load("ref_functions")
    a, b = ref["seg_in_first"], ref["seg_in_second"]
    got = og.line_intersect(*[T(a[:, i]) for i in range(4)], *[T(b[:, i]) for i in range(4)], 1e-10)
    for i, key in enumerate(("x", "y", "valid", "u", "v")):
        assert close(got[i].numpy(), ref[f"seg_{key}"]), key
    a, c = ref["circ_in_lines"], ref["circ_in_circles"]
    plus, minus = og.line_circle_intersect(*[T(a[:, i]) for i in range(4)],
                                           *[T(c[:, i]) for i in range(3)], 1e-10)
    for tag, d in (("plus", plus), ("minus", minus)):
        for key in ("x", "y", "valid", "u", "v"):
            assert close(d[key].numpy(), ref[f"circ_{tag}_{key}"], 1e-10), (tag, key)
    a = ref["snell2_in"]
    got = og.snells_law_2D(*[T(a[:, i]) for i in range(7)], torch.tensor(0.75, dtype=torch.float64))
    assert close(np.stack([g.numpy() for g in got], axis=1), ref["snell2_out"], 1e-11)
    a = ref["angle_in"]
    assert np.array_equal(og.angle_in_interval(T(a[:, 0]), T(a[:, 1]), T(a[:, 2])).numpy(),
                          ref["angle_out"])


def test_functions_materials():
    ref = load("ref_functions")
    wl = T(ref["mat_wavelength"])
    for name in ("vacuum", "acrylic", "crown_glass", "flint_glass", "fused_silica", "polycarbonate",
                 "soda_lime", "reflective"):
        got = np.broadcast_to(np.asarray(om.BY_NAME[name](wl), dtype=np.float64), wl.shape)
        assert close(got, ref[f"mat_{name}"], 1e-14), name
    assert close(np.asarray(om.build_constant_material(1.7)(wl)), ref["mat_constant_1p7"])


# ------------------------------------------------------------------------------------ traces (CPU)

def _oracle_trace(name):
    scene = SCENES[name]
    if scene["dimension"] == 3:
        grad = "goal_scale" in scene
        eng, params = helpers.oracle_from_scene(scene, requires_grad=grad, fp32_vertices=False,
                                                compile_dead_rays=True, compile_stopped_rays=True)
        fields = helpers.RAY_FIELDS
    else:
        eng, params = helpers.oracle2d_from_scene(scene, compile_dead_rays=True,
                                                  compile_stopped_rays=True)
        fields = helpers.RAY_FIELDS_2D
    eng.ray_trace(scene["max_iterations"])
    return scene, eng, params, fields


@pytest.mark.parametrize("name", sorted(SCENES))
def test_oracle_matches_reference_trace(name):
    want = load(name)
    scene, eng, params, fields = _oracle_trace(name)
    seen = 0
    for set_name in SETS:
        s = getattr(eng, set_name + "_rays")
        n = int(s["slot"].shape[0]) if bool(s) else 0
        if f"{set_name}_slot" not in want:
            assert n == 0, set_name
            continue
        # same membership in the same (reference) order
        assert np.array_equal(s["slot"].detach().numpy().astype(np.int64), want[f"{set_name}_slot"]), set_name
        for f in fields:
            assert close(s[f].detach().numpy(), want[f"{set_name}_{f}"], 1e-10), (set_name, f)
        seen += n
    assert seen > 0
    if "merged_xp" in want:
        merged = eng.optical_system._merged if hasattr(eng.optical_system, "_merged") else None
        if merged is not None:
            for f in ("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2", "norm"):
                assert close(merged[f].detach().numpy(), want[f"merged_{f}"], 1e-13), f
            assert np.array_equal(merged["catagory"].numpy(), want["merged_catagory"])
    if "grad_0" in want:
        fin = eng.finished_rays
        goal = scene["goal_scale"] * fin["rank"]
        err = (fin["y_end"] - goal[:, 0]) ** 2 + (fin["z_end"] - goal[:, 1]) ** 2
        assert abs(float(err.sum().detach()) - float(want["loss"])) < 1e-11 * abs(float(want["loss"]))
        for i, g in enumerate(torch.autograd.grad(err.sum(), params)):
            assert helpers.rel_err(g.numpy(), want[f"grad_{i}"]) < 1e-9, i


# ------------------------------------------------------------------------------------ traces (GPU)

@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENES))
def test_cuda_path_matches_reference_trace(name):
    from tensorflowraytrace_b200 import scenes
    scene, want = SCENES[name], load(name)
    if scene["dimension"] == 3:
        engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS, compile_dead_rays=True,
                                                compile_stopped_rays=True)
        fields = helpers.RAY_FIELDS
    else:
        engine, params, _ = scenes.build_engine_2d(scene, eps=helpers.EPS, compile_dead_rays=True,
                                                   compile_stopped_rays=True)
        fields = helpers.RAY_FIELDS_2D
    engine.ray_trace(scene["max_iterations"])
    for set_name in SETS:
        got = getattr(engine, set_name + "_rays")
        if f"{set_name}_slot" not in want:
            assert not bool(got) or got["x_start"].shape[0] == 0, set_name
            continue
        slot = engine.last_history_slots[set_name].cpu().numpy().astype(np.int64)
        assert np.array_equal(slot, want[f"{set_name}_slot"]), set_name      # bit-exact membership / order
        for f in fields:
            assert helpers.rel_err(got[f].detach().cpu().numpy(), want[f"{set_name}_{f}"]) < 1e-5, (set_name, f)
    if "grad_0" in want:
        err = scenes.imaging_error(scene)(engine)
        grads = torch.autograd.grad(err.sum(), params)
        assert abs(float(err.sum().detach()) - float(want["loss"])) < 1e-5 * abs(float(want["loss"]))
        for i, g in enumerate(grads):
            assert helpers.rel_err(g.cpu().numpy(), want[f"grad_{i}"]) < 1e-4, i


@pytest.mark.gpu
def test_cuda_intersection_seam_matches_reference_function():
    """`OpticalSystem3D._intersection` on the reference's own line_triangle_intersect
    vectors: same winner per ray as argmin over the reference's (M, N) ray_u, same hit point."""
    from tensorflowraytrace_b200 import engine as eng
    ref = load("ref_functions")
    r, t = ref["tri_in_rays"], ref["tri_in_tris"]
    valid = ref["tri_valid"] & (ref["tri_trig_u"] >= -1e-5) & (ref["tri_trig_v"] >= -1e-5) \
        & (ref["tri_trig_u"] + ref["tri_trig_v"] <= 1 + 1e-5) & (ref["tri_ray_u"] >= 1e-4)
    u = np.where(valid, ref["tri_ray_u"], np.inf)
    want_tri = np.where(valid.any(axis=0), u.argmin(axis=0), -1)
    dev = torch.device("cuda:0")
    args = [torch.tensor(r[:, i], dtype=torch.float32, device=dev) for i in range(6)] + \
           [torch.tensor(t[:, i], dtype=torch.float32, device=dev) for i in range(9)]
    x, y, z, ok, ray_u, trig_u, trig_v, gather_ray, gather_trig = eng.OpticalSystem3D._intersection(
        *args, 1e-10, 1e-5, 1e-4)
    ok = ok.cpu().numpy()
    got_tri = np.where(ok, gather_trig.cpu().numpy(), -1)
    assert np.array_equal(got_tri, want_tri)
    hit = ok.nonzero()[0]
    assert hit.size > 3
    for name, arr in (("x", x), ("y", y), ("z", z), ("ray_u", ray_u)):
        want = ref[f"tri_{name}"][want_tri[hit], hit]
        assert helpers.rel_err(arr.cpu().numpy()[hit], want) < 1e-5, name


# ------------------------------------------------------------------------------------ sources

SOURCE_CASES = mrg.source_cases()


@pytest.mark.parametrize("name", sorted(SOURCE_CASES))
def test_sources_and_distributions_match_reference(name):
    """8(f) rank 2: our distributions / sources (fp32 fields) against the arrays the reference's own
    classes produced for the same constructor arguments (tfrt/distributions.py, tfrt/sources.py)."""
    from tensorflowraytrace_b200 import distributions as D
    from tensorflowraytrace_b200 import sources as S
    ref = load("ref_sources")
    obj = mrg.build_case(D, S, SOURCE_CASES[name])
    checked = 0
    for key, want in ref.items():
        case, attr = key.split("__")
        if case != name:
            continue
        got = getattr(obj, attr) if name.startswith("dist_") else obj[attr]
        got = got.detach().cpu().numpy().astype(np.float64)
        assert got.shape == want.shape, (key, got.shape, want.shape)
        assert np.allclose(got, want, rtol=2e-6, atol=2e-6), (key, np.abs(got - want).max())
        checked += 1
    assert checked >= 2


# ------------------------------------------------------------------------------------ optimiser

@pytest.mark.gpu
def test_optimizer_trajectory_matches_reference():
    """8(f) rank 1: `SGD_Optimizer.process_gradient / single_step` (lr scaling, common clip, dense vs
    sparse accumulator, Keras SGD(nesterov=True) at its default 0.01 step, momentum schedule) on the
    CUDA path against the parameter trajectory of the reference's own `tfrt.optimizer.SGD_Optimizer`
    driving the reference engine."""
    import scipy.sparse as sp
    from tensorflowraytrace_b200 import mesh_tools, optimizer, scenes
    want = load("ref_optimizer")
    scene = mrg.optimizer_scene()
    engine, params, _ = scenes.build_engine(scene, eps=helpers.EPS)
    goal_scale = float(scene["goal_scale"])

    def error_function(eng):
        fin = eng.finished_rays
        output = torch.stack([fin["y_end"], fin["z_end"]], dim=1)
        return (output - fin["rank"] * goal_scale) ** 2

    opt = optimizer.SGD_Optimizer(engine, params, error_function, scene["max_iterations"], **mrg.OPT_KW)
    acc0, _ = mrg.optimizer_accumulators(scene)
    accs = [mesh_tools.to_torch_sparse(sp.csr_matrix(acc0), "cuda"), None]
    for i in range(2):
        assert helpers.rel_err(params[i].detach().cpu().numpy(), want[f"initial_{i}"]) < 1e-6
    grads, err = opt.process_gradient(accs)
    assert abs(float(err) - float(want["mean_error_0"])) < 1e-5 * float(want["mean_error_0"])
    for i in range(2):
        assert helpers.rel_err(grads[i].cpu().numpy(), want[f"processed_{i}"]) < 1e-4, i
    assert float(np.abs(want["processed_1"]).max()) == pytest.approx(10 * mrg.OPT_KW["learning_rate"])  # clip hit
    for k, (lr_scale, momentum) in enumerate(mrg.OPT_STEPS):
        e = opt.single_step(accs, lr_scale=lr_scale, momentum=momentum)
        assert abs(float(e) - float(want["errors"][k])) < 1e-4 * float(want["errors"][k]), k
        for i in range(2):
            moved = np.abs(want[f"step{k}_{i}"] - want[f"initial_{i}"]).max()
            diff = np.abs(params[i].detach().cpu().numpy() - want[f"step{k}_{i}"]).max()
            assert diff < 2e-4 * moved + 1e-7, (k, i, diff, moved)


# ------------------------------------------------------------------------------------ multi-surface flow

def _multisurface_ours(device):
    """The fixture's scene on this package's classes (same constructor calls as the reference)."""
    import math
    from tensorflowraytrace_b200 import boundaries, distributions, engine, materials, mesh_tools, \
        operation, sources
    verts, faces, init = mrg.multisurface_inputs()
    zero = boundaries.TriangleMesh(verts, faces)
    top = mesh_tools.get_closest_point(zero, (0, 0, 0))
    vum, acc = mesh_tools.mesh_parametrization_tools(zero, top)
    smo = mesh_tools.mesh_smoothing_tool(zero, [1.0, 0.02])
    if device is None:
        return zero, top, vum, acc, smo
    side = mrg.MULTI_SIDE
    base_points = distributions.StaticUniformSquare(0.2, side)
    angles = distributions.StaticUniformSphere(math.pi / 16.0, side * side)
    source = sources.AngularSource(
        3, (-4.0, 0.0, 0.0), (1.0, 0.0, 0.0), angles, base_points, [575.0] * (side * side),
        dense=False, extra_fields={"rank": ("base_point", base_points, "ranks")})
    mat = {"mat_in": np.int64(1), "mat_out": np.int64(0)}
    lens = boundaries.ParametricMultiTriangleBoundary(
        zero, boundaries.FromVectorVG((1, 0, 0)),
        [boundaries.ThicknessConstraint(0.0, "min"), boundaries.ThicknessConstraint(0.2, "min")],
        [True, False], initial_parameters=init, material_list=[mat, mat], vertex_update_map=vum)
    tv = np.array([[8.0, -50.0, -50.0], [8.0, 50.0, -50.0], [8.0, 50.0, 50.0], [8.0, -50.0, 50.0]])
    target = boundaries.ManualTriangleBoundary(
        mesh=boundaries.TriangleMesh(tv, np.array([[0, 1, 2], [0, 2, 3]])))
    system = engine.OpticalSystem3D(**helpers.EPS)
    system.optical = lens.surfaces
    system.targets = [target]
    system.sources = [source]
    system.materials = [{"n": materials.vacuum}, {"n": materials.acrylic}]
    system.update()
    eng = engine.OpticalEngine(3, [operation.StandardReaction()],
                               simple_ray_inheritance={"wavelength", "rank"})
    eng.optical_system = system
    eng.validate_system()
    return eng, system, lens, acc, smo


def test_mesh_tools_match_reference_functions():
    """`get_closest_point`, `mesh_parametrization_tools`, `mesh_smoothing_tool` (sparse here) against
    the dense arrays the reference's own tfrt/mesh_tools.py:221-421 produced for the same mesh."""
    want = load("ref_multisurface")
    zero, top, vum, acc, smo = _multisurface_ours(None)
    assert int(top) == int(want["top_parent"])
    assert np.array_equal(np.asarray(vum).astype(bool), want["vertex_update_map"])
    assert np.abs(acc.toarray() - want["accumulator"]).max() == 0.0
    assert np.abs(smo.toarray() - want["smoother"]).max() < 1e-15


@pytest.mark.gpu
def test_multisurface_optimisation_flow_matches_reference():
    """dev/simple_3d_optimize.py's flow end to end on the CUDA path against the reference's own classes:
    constraints applied by update(), merged fields with vertex_update_map, three
    `single_step(accumulators)` + `smooth` rounds, constraints re-applied."""
    from tensorflowraytrace_b200 import mesh_tools, optimizer
    want = load("ref_multisurface")
    eng, system, lens, acc, smo = _multisurface_ours("cuda")
    for i, p in enumerate(lens.parameters):
        assert np.abs(p.detach().cpu().numpy() - want[f"constrained_{i}"]).max() < 1e-6, i
    tri, norm, _, _ = system.merged_device_arrays()
    for c, f in enumerate(("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2")):
        assert np.abs(tri[c].detach().cpu().numpy() - want[f"merged_{f}"]).max() < 1e-6, f
    assert np.abs(norm.detach().cpu().numpy() - want["merged_norm"]).max() < 1e-5

    def error_function(engine):
        fin = engine.finished_rays
        output = torch.stack([fin["y_end"], fin["z_end"]], dim=1)
        return (output - fin["rank"] * (-0.4)) ** 2

    opt = optimizer.SGD_Optimizer(eng, lens.parameters, error_function, 3, **mrg.MULTI_KW)
    accs = [mesh_tools.to_torch_sparse(acc, "cuda")] * 2
    smoother = mesh_tools.to_torch_sparse(smo, "cuda")
    for k, (lr_scale, momentum) in enumerate(mrg.MULTI_STEPS):
        e = opt.single_step(accs, lr_scale=lr_scale, momentum=momentum)
        assert abs(float(e) - float(want["errors"][k])) < 2e-4 * float(want["errors"][k]), k
        for p in lens.parameters:
            opt.smooth(p, smoother)
        for i, p in enumerate(lens.parameters):
            moved = np.abs(want[f"step{k}_{i}"] - want[f"constrained_{i}"]).max()
            diff = np.abs(p.detach().cpu().numpy() - want[f"step{k}_{i}"]).max()
            assert diff < 5e-4 * moved + 2e-7, (k, i, diff, moved)
    system.update()
    for i, p in enumerate(lens.parameters):
        assert np.abs(p.detach().cpu().numpy() - want[f"final_{i}"]).max() < 2e-6, i


# ------------------------------------------------------------------------------------ two sources, OldestAncestor

def test_oracle_matches_reference_two_sources():
    """Amalgamation order of two sources: the reference's `oldest_ancestor` (running index over the
    sources, tfrt/operation.py:186-193) equals the oracle's slot of the concatenated rays."""
    want = load("ref_two_sources")
    scene = mrg.two_source_scene()
    eng, _ = helpers.oracle_from_scene(scene, fp32_vertices=False, compile_dead_rays=True)
    eng.ray_trace(scene["max_iterations"])
    for set_name in ("active", "finished"):
        s = getattr(eng, set_name + "_rays")
        assert np.array_equal(s["slot"].numpy().astype(np.int64), want[f"{set_name}_oldest_ancestor"])
        for f in helpers.RAY_FIELDS:
            assert close(s[f].detach().numpy(), want[f"{set_name}_{f}"], 1e-10), (set_name, f)


@pytest.mark.gpu
def test_cuda_two_sources_and_oldest_ancestor_match_reference():
    from tensorflowraytrace_b200 import engine as eng_mod, operation, scenes, sources
    want = load("ref_two_sources")
    scene = mrg.two_source_scene()
    base, _, _ = scenes.build_engine(scene, eps=helpers.EPS)
    system = base.optical_system
    srcs = []
    for sl in (slice(0, mrg.TWO_SPLIT), slice(mrg.TWO_SPLIT, None)):
        src = sources.ManualSource(3)
        src.feed({k: v[sl] for k, v in scene["sources"].items()})
        srcs.append(src)
    system.sources = srcs
    e = eng_mod.OpticalEngine(3, [operation.StandardReaction(), operation.OldestAncestor()],
                              compile_dead_rays=True, simple_ray_inheritance={"wavelength", "rank"})
    e.optical_system = system
    e.annotate()
    system.update()
    e.validate_system()
    e.ray_trace(scene["max_iterations"])
    for set_name in ("active", "finished"):
        s = getattr(e, set_name + "_rays")
        assert np.array_equal(s["oldest_ancestor"].cpu().numpy().astype(np.int64),
                              want[f"{set_name}_oldest_ancestor"]), set_name
        for f in helpers.RAY_FIELDS:
            assert helpers.rel_err(s[f].detach().cpu().numpy(), want[f"{set_name}_{f}"]) < 1e-5, (set_name, f)


# ------------------------------------------------------------------------------------ a12 variants

@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ms", "guide", "guide_sym"])
def test_parametric_boundary_variants_match_reference(name):
    """parameters -> vertices -> fields -> unit normal and the gradient back to the
    parameters for `MasterSlaveParametricTriangleBoundary` (gather) and `ParametricCylindricalGuide`
    (min-pinning constraint, repeat for rotational symmetry), against the arrays and the
    GradientTape gradient of the reference's own classes (tfrt/boundaries.py:890-923, 1116-1224,
    1416-1611)."""
    from tensorflowraytrace_b200 import boundaries as B
    want = load("ref_boundaries")
    mat = {"mat_in": np.int64(1), "mat_out": np.int64(0)}
    if name == "ms":
        verts, faces, masters, init = mrg.boundary_inputs()
        b = B.MasterSlaveParametricTriangleBoundary(
            masters, mrg.attach_mirror_slaves, B.TriangleMesh(verts, faces), B.FromVectorVG((1, 0, 0)),
            initial_parameters=init, material_dict=mat)
        assert np.array_equal(b._gather.cpu().numpy().astype(np.int64), want["ms_gather"])
    else:
        b = B.ParametricCylindricalGuide(*mrg.GUIDE_ARGS, rotationally_symmetric=(name == "guide_sym"),
                                         material_dict=mat, **mrg.GUIDE_KW)
        n = b.parameters.shape[0]
        bump = 0.03 * torch.sin(torch.arange(n, dtype=torch.float64) * 0.7)
        with torch.no_grad():
            b.parameters.add_(torch.tensor(helpers.f32_round(bump.numpy()), dtype=b.parameters.dtype,
                                           device=b.parameters.device))
        assert np.array_equal(np.asarray(b.vertex_update_map).astype(bool), want[f"{name}_vertex_update_map"])
        acc = b.accumulator
        acc = acc.toarray() if hasattr(acc, "toarray") else np.asarray(acc.to_dense().cpu() if hasattr(acc, "to_dense") else acc)
        assert np.abs(acc - want[f"{name}_accumulator"]).max() == 0.0
    b.update()
    assert np.abs(b.parameters.detach().cpu().numpy() - want[f"{name}_parameters"]).max() < 1e-6
    assert np.array_equal(b.faces.cpu().numpy().astype(np.int64), want[f"{name}_faces"])
    assert np.abs(b.vertices.detach().cpu().numpy() - want[f"{name}_vertices"]).max() < 1e-6
    fields = ("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2")
    for k in fields:
        assert np.abs(b[k].detach().cpu().numpy() - want[f"{name}_{k}"]).max() < 1e-6, k
    assert np.abs(b["norm"].detach().cpu().numpy() - want[f"{name}_norm"]).max() < 1e-5
    n_faces = b["xp"].shape[0]
    w = torch.tensor(mrg.boundary_weights(n_faces, n_faces), dtype=torch.float32, device=b["xp"].device)
    loss = sum((b[k] * w[:, i]).sum() for i, k in enumerate(fields)) + (b["norm"] * w[:, 9:12]).sum()
    (g,) = torch.autograd.grad(loss, [b.parameters])
    assert abs(float(loss) - float(want[f"{name}_loss"])) < 1e-4 * max(abs(float(want[f"{name}_loss"])), 1.0)
    assert helpers.rel_err(g.cpu().numpy(), want[f"{name}_grad"]) < 1e-4


# ------------------------------------------------------------------------------------ 2D parametric lens

@pytest.mark.gpu
def test_2d_parametric_lens_gradient_matches_reference():
    """`ParametricSegmentBoundary` in a 2D system: trace + imaging loss + gradient to the boundary
    parameters through the fused 2D kernels, against the reference engine's GradientTape."""
    from tensorflowraytrace_b200 import boundaries as B, distributions as D, engine as eng_mod, \
        materials, operation, sources
    want = load("ref_param2d")
    d = mrg.param2d_inputs()
    zero = D.StaticUniformAperaturePoints(*d["zero"], mrg.PARAM2D_POINTS)
    one = D.StaticUniformAperaturePoints(*d["one"], mrg.PARAM2D_POINTS)
    front = B.ParametricSegmentBoundary(zero, one, initial_parameters=d["init"])
    back = B.ManualSegmentBoundary()
    back.feed_segments(d["back"])
    for b in (front, back):
        eng_mod.annotation_helper(b, "mat_in", 1, "x_start", dtype=torch.int64)
        eng_mod.annotation_helper(b, "mat_out", 0, "x_start", dtype=torch.int64)
    target = B.ManualSegmentBoundary()
    target.feed_segments(d["target"])
    src = sources.ManualSource(2)
    src.feed(d["rays"])
    system = eng_mod.OpticalSystem2D(**helpers.EPS)
    system.optical_segments = [front, back]
    system.target_segments = [target]
    system.sources = [src]
    system.materials = [{"n": materials.vacuum}, {"n": materials.acrylic}]
    system.update()
    e = eng_mod.OpticalEngine(2, [operation.StandardReaction()], compile_dead_rays=True,
                              simple_ray_inheritance={"wavelength", "rank"})
    e.optical_system = system
    e.validate_system()
    for k in ("x_start", "y_start", "x_end", "y_end"):
        assert np.abs(front[k].detach().cpu().numpy() - want[f"front_{k}"]).max() < 1e-6, k
    e.ray_trace(3)
    fin = e.finished_rays
    slot = e.last_history_slots["finished"].cpu().numpy().astype(np.int64)
    assert np.array_equal(slot, want["finished_slot"])
    for f in helpers.RAY_FIELDS_2D:
        assert helpers.rel_err(fin[f].detach().cpu().numpy(), want[f"finished_{f}"]) < 1e-5, f
    err = (fin["y_end"] - fin["rank"] * d["goal_scale"]) ** 2
    assert abs(float(err.sum()) - float(want["loss"])) < 1e-4 * float(want["loss"])
    (g,) = torch.autograd.grad(err.sum(), [front.parameters])
    assert helpers.rel_err(g.cpu().numpy(), want["grad"]) < 1e-4


# ------------------------------------------------------------------------------------ mesh generators

@pytest.mark.parametrize("name", sorted(mrg.mesh_cases()))
def test_mesh_generators_match_reference_functions(name):
    """`mesh_tools.hexagonal_mesh` / `cylindrical_mesh`: same vertices (bit-exact) in the same order
    and the same faces in the same order and winding as the reference's generators, so vertex
    indices (parameters, accumulators, constraints) and triangle indices mean the same thing."""
    from tensorflowraytrace_b200 import mesh_tools
    want = load("ref_meshes")
    fn, kw = mrg.mesh_cases()[name]
    pts, faces = getattr(mesh_tools, fn)(**kw)
    assert np.array_equal(faces, want[f"{name}_faces"])
    assert np.abs(pts - want[f"{name}_points"]).max() < 1e-15


def test_base_point_transformation_matches_reference():
    from tensorflowraytrace_b200 import distributions as D
    want = load("ref_sources")["transformed_square__points"]
    got = mrg.transformation_case(D).points.detach().cpu().numpy().astype(np.float64)
    assert got.shape == want.shape == (6, 3)
    assert np.allclose(got, want, atol=2e-6)


# ------------------------------------------------------------------------------------ projection result

def _check_projection(result, want, tol):
    flat = {}

    def walk(prefix, d):
        for k, v in d.items():
            if isinstance(v, dict):
                walk(f"{prefix}{k}.", v)
            else:
                flat[f"{prefix}{k}"] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    walk("", result)
    missing = set(want) - set(flat)
    assert not missing, sorted(missing)
    counts = {c: want[f"rays.{c}.x_start"].shape[0] for c in ("active", "finished", "stopped", "dead")}
    assert min(counts.values()) > 10, counts                      # every category is populated
    for key, w in want.items():
        g = flat[key]
        assert g.shape == w.shape, (key, g.shape, w.shape)
        if w.dtype.kind in "iub":
            assert np.array_equal(g.astype(np.int64), w.astype(np.int64)), key
        else:
            assert helpers.rel_err(g, w) < tol, (key, helpers.rel_err(g, w))


def test_oracle_projection_result_matches_reference():
    """One `process_projection_3D` pass with technical intersections, every compile flag and
    dead-ray lengthening: the whole nested result (tfrt/engine.py:1988-2191) against the reference."""
    want = load("ref_projection")
    scene = mrg.projection_scene()
    eng, _ = helpers.oracle_from_scene(scene, fp32_vertices=False, compile_dead_rays=True,
                                       compile_stopped_rays=True, compile_technical_intersections=True,
                                       dead_ray_length=5.0)
    rays = dict(eng.optical_system._amalgamated_sources)
    _check_projection(eng.process_projection_3D(rays), want, 1e-10)


@pytest.mark.gpu
def test_cuda_projection_result_matches_reference():
    from tensorflowraytrace_b200 import scenes
    want = load("ref_projection")
    scene = mrg.projection_scene()
    engine, _, src = scenes.build_engine(scene, eps=helpers.EPS, compile_dead_rays=True,
                                         compile_stopped_rays=True,
                                         compile_technical_intersections=True, dead_ray_length=5.0)
    src["slot"] = torch.arange(scene["sources"]["x_start"].shape[0], dtype=torch.float32)
    engine.optical_system.update()
    rays = dict(engine.optical_system._amalgamated_sources)
    _check_projection(engine.process_projection_3D(rays), want, 1e-5)


@pytest.mark.gpu
def test_cuda_projection_result_2d_matches_reference():
    """One `process_projection_2D` pass (segments only) with technical intersections, every compile
    flag and dead-ray lengthening against the reference's result dict (tfrt/engine.py:1544-1986)."""
    from tensorflowraytrace_b200 import scenes
    want = load("ref_projection2d")
    scene = mrg.projection_scene_2d()
    engine, _, src = scenes.build_engine_2d(scene, eps=helpers.EPS, compile_dead_rays=True,
                                            compile_stopped_rays=True,
                                            compile_technical_intersections=True, dead_ray_length=3.0)
    src["slot"] = torch.arange(scene["sources"]["x_start"].shape[0], dtype=torch.float32)
    engine.optical_system.update()
    rays = dict(engine.optical_system._amalgamated_sources)
    _check_projection(engine.process_projection_2D(rays), want, 1e-5)


@pytest.mark.gpu
def test_cuda_ghost_through_hook_path_matches_reference():
    """The operation-hook path (`single_pass` with a non-StandardReaction operation): the
    segments-only 2D scene traced with `[GhostThrough()]` against the reference engine's own run --
    same membership and order of every history set, same rays."""
    from tensorflowraytrace_b200 import engine as eng_mod, operation, scenes
    want = load("ref_ghost2d")
    scene = SCENES["ref_segments2d"]
    base, _, src = scenes.build_engine_2d(scene, eps=helpers.EPS)
    src["slot"] = torch.arange(scene["sources"]["x_start"].shape[0], dtype=torch.float32)
    system = base.optical_system
    system.update()
    ghost = eng_mod.OpticalEngine(2, [operation.GhostThrough()], compile_dead_rays=True,
                                  compile_stopped_rays=True,
                                  simple_ray_inheritance={"wavelength", "slot"})
    ghost.optical_system = system
    ghost.validate_system()
    ghost.ray_trace(5)
    for set_name in ("active", "finished", "stopped"):
        s = getattr(ghost, set_name + "_rays")
        assert np.array_equal(s["slot"].cpu().numpy().astype(np.int64), want[f"{set_name}_slot"]), set_name
        for f in helpers.RAY_FIELDS_2D:
            assert helpers.rel_err(s[f].detach().cpu().numpy(), want[f"{set_name}_{f}"]) < 1e-5, (set_name, f)
    assert "dead_slot" not in want
    dead = ghost.dead_rays
    assert not bool(dead) or dead["x_start"].shape[0] == 0
