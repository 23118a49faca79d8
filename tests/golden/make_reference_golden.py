"""Generate golden vectors from the REFERENCE'S OWN SOURCE (tests/golden/ref_*.npz).

The reference (`/root/reference/tfrt`, pure Python on TensorFlow) cannot be imported as is: TensorFlow,
tfquaternion and pyvista are absent from this image.  This script puts a stand-in `tensorflow` module
on `sys.path` (oracle/tf_shim/, loaded by oracle/reference_runner.py; ~70 tf ops mapped onto the IEEE-equivalent torch float64 ops, with
`tf.GradientTape` served by torch autograd) and then imports and runs the reference's unmodified
`tfrt.geometry / engine / operation / materials / boundaries / sources / optimizer` code:

  * function level: `line_triangle_intersect`, `snells_law_3D`, `line_intersect`,
    `line_circle_intersect`, `snells_law_2D`, `angle_in_interval`, every material;
  * engine level: `OpticalSystem3D/2D` + `OpticalEngine([StandardReaction()]).ray_trace` on small
    versions of the BASELINE scenes (lens, light guide, 2D wedge, 2D arc lens), with the history sets
    in the reference's order and a `slot` field (source-ray id) carried through `simple_ray_inheritance`;
  * gradient level: the `tf.GradientTape` block of `dev/simple_3d_optimize.py:122-140` (same loss) to the
    two `ParametricTriangleBoundary.parameters` Variables.

What this pins and what it does not: every line of reference Python on the path executes as written
(broadcast orientation, where / safe-value structure, inf fill + argmin, merge and compaction order,
Snell branches, material formulas, vertices -> fields map, the tape); the arithmetic of each individual tf
op is torch's float64 implementation of the same IEEE operation, not TensorFlow's Eigen kernel.

Runs only in the build container (needs /root/reference); the .npz files are committed and travel.

    python tests/golden/make_reference_golden.py
"""
import math
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

import numpy as np  # noqa: E402
import torch  # noqa: E402

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle.reference_runner import (EPS, RAYS2, RAYS3, _polydata, f32_round,  # noqa: E402
                                     import_reference, reference_engine_2d, reference_engine_3d)


def golden_scenes():
    for p in (ROOT,):
        if p not in sys.path:
            sys.path.insert(0, p)
    from tensorflowraytrace_b200 import scenes
    return {
        "ref_lens3d": scenes.lens_optimize_config(ray_count=192, step_count=3, name="golden"),
        "ref_guide3d": scenes.config4(ray_count=150, theta_res=8, z_res=5, max_iterations=6),
        "ref_mirror_stop3d": scenes.mirror_stop_3d(),
        "ref_wedge2d": scenes.config1(ray_count=64),
        "ref_arclens2d": scenes.arc_lens_2d(ray_count=90),
        "ref_segments2d": scenes.segments_2d(),
    }


def _record_sets(out, eng, fields):
    for set_name in ("active", "finished", "stopped", "dead"):
        s = getattr(eng, set_name + "_rays")
        try:
            slot = s["slot"]
        except KeyError:
            continue
        if slot.shape[0] == 0:
            continue
        out[f"{set_name}_slot"] = slot.numpy().astype(np.int64)
        for f in fields:
            out[f"{set_name}_{f}"] = s[f].numpy().astype(np.float64)
        if set_name == "finished":
            out["finished_wavelength"] = s["wavelength"].numpy().astype(np.float64)


def run_reference(ref, name, scene):
    tf = ref.tf
    out = {}
    if scene["dimension"] == 3:
        eng, system, params = reference_engine_3d(ref, scene)
        if "goal_scale" in scene:
            # the tape block of dev/simple_3d_optimize.py:128-140, verbatim structure
            with tf.GradientTape() as tape:
                eng.optical_system.update()
                eng.clear_ray_history()
                eng.ray_trace(scene["max_iterations"])
                output = tf.stack([eng.finished_rays["y_end"], eng.finished_rays["z_end"]], axis=1)
                goal = eng.finished_rays["rank"]
                goal = goal * scene["goal_scale"]
                error = tf.math.squared_difference(output, goal)
                grads = tape.gradient(error, params)
            for i, g in enumerate(grads):
                out[f"grad_{i}"] = g.numpy()
            out["loss"] = np.array(float(error.sum()))
        else:
            eng.ray_trace(scene["max_iterations"])
        _record_sets(out, eng, RAYS3)
        # merged triangle fields as the reference built them (vertices -> fields map)
        for f in ("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2", "norm"):
            out[f"merged_{f}"] = system._merged[f].numpy().astype(np.float64)
        out["merged_catagory"] = system._merged["catagory"].numpy().astype(np.int64)
    else:
        eng, system = reference_engine_2d(ref, scene)
        eng.ray_trace(scene["max_iterations"])
        _record_sets(out, eng, RAYS2)
    return out


# ------------------------------------------------------------------------------- function level

def function_vectors(ref):
    """Random-input known answers straight out of tfrt/geometry.py and tfrt/materials.py."""
    tf, g = ref.tf, ref.geometry
    rng = np.random.default_rng(20261018)
    out = {}
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731

    # 3D: line_triangle_intersect (geometry.py:191-320): N rays x M triangles -> (M, N)
    n, m = 40, 25
    rs = rng.normal(size=(n, 3)) * 0.5 + np.array([-3.0, 0.0, 0.0])
    re = rs + rng.normal(size=(n, 3)) * 0.3 + np.array([1.0, 0.0, 0.0])
    p0 = rng.normal(size=(m, 3))
    p1 = p0 + rng.normal(size=(m, 3))
    p2 = p0 + rng.normal(size=(m, 3))
    p1[0] = p0[0] + (p2[0] - p0[0]) * 2.0            # a degenerate (zero-area) triangle
    re[1] = rs[1] + (p1[3] - p0[3])                  # a ray parallel to triangle 3's plane
    args = [T(rs[:, i]) for i in range(3)] + [T(re[:, i]) for i in range(3)]
    for p in (p0, p1, p2):
        args += [T(p[:, i]) for i in range(3)]
    res = g.line_triangle_intersect(*args, 1e-10)
    out["tri_in_rays"] = np.concatenate([rs, re], axis=1)
    out["tri_in_tris"] = np.concatenate([p0, p1, p2], axis=1)
    for key, val in zip(("x", "y", "z", "valid", "ray_u", "trig_u", "trig_v"), res):
        out[f"tri_{key}"] = val.numpy()

    # 3D: snells_law_3D (geometry.py:671-753): refraction both ways, TIR, mirror, zero indices
    k = 64
    s = rng.normal(size=(k, 3))
    e = s + rng.normal(size=(k, 3))
    nrm = rng.normal(size=(k, 3))
    n_in = rng.choice([1.0, 1.33, 1.5, 2.4, 0.0], size=k, p=[0.2, 0.2, 0.3, 0.2, 0.1])
    n_out = rng.choice([1.0, 1.33, 1.5, 0.0], size=k, p=[0.5, 0.2, 0.2, 0.1])
    res = g.snells_law_3D(*[T(s[:, i]) for i in range(3)], *[T(e[:, i]) for i in range(3)],
                          T(nrm), T(n_in), T(n_out), tf.constant(1.5, dtype=tf.float64))
    out["snell3_in"] = np.concatenate([s, e, nrm, n_in[:, None], n_out[:, None]], axis=1)
    out["snell3_out"] = np.stack([r.numpy() for r in res], axis=1)

    # 2D: line_intersect (geometry.py:27-167) -> (M, N), incl. a parallel pair
    n, m = 30, 12
    a = rng.normal(size=(n, 4))
    b = rng.normal(size=(m, 4))
    b[2, 2:] = b[2, :2] + (a[4, 2:] - a[4, :2]) * 1.7
    res = g.line_intersect(*[T(a[:, i]) for i in range(4)], *[T(b[:, i]) for i in range(4)], 1e-10)
    out["seg_in_first"], out["seg_in_second"] = a, b
    for key, val in zip(("x", "y", "valid", "u", "v"), res):
        out[f"seg_{key}"] = val.numpy()

    # 2D: line_circle_intersect (geometry.py:338-547): both roots, incl. misses and a tangent
    n, m = 30, 9
    a = rng.normal(size=(n, 4)) * 1.5
    c = np.concatenate([rng.normal(size=(m, 2)), rng.uniform(0.3, 2.0, size=(m, 1))], axis=1)
    a[0] = [-5.0, c[0, 1] + c[0, 2], 5.0, c[0, 1] + c[0, 2]]        # tangent to circle 0
    res = g.line_circle_intersect(*[T(a[:, i]) for i in range(4)], *[T(c[:, i]) for i in range(3)],
                                  1e-10)
    out["circ_in_lines"], out["circ_in_circles"] = a, c
    plus, minus = res
    for tag, d in (("plus", plus), ("minus", minus)):
        for key in ("x", "y", "valid", "u", "v"):
            out[f"circ_{tag}_{key}"] = d[key].numpy()

    # 2D: snells_law_2D (geometry.py:565-653), norms as angles
    k = 64
    s = rng.normal(size=(k, 2))
    e = s + rng.normal(size=(k, 2))
    ang = rng.uniform(-math.pi, math.pi, size=k)
    n_in = rng.choice([1.0, 1.33, 1.5, 2.4, 0.0], size=k, p=[0.2, 0.2, 0.3, 0.2, 0.1])
    n_out = rng.choice([1.0, 1.33, 1.5, 0.0], size=k, p=[0.5, 0.2, 0.2, 0.1])
    res = g.snells_law_2D(T(s[:, 0]), T(s[:, 1]), T(e[:, 0]), T(e[:, 1]), T(ang), T(n_in), T(n_out),
                          tf.constant(0.75, dtype=tf.float64))
    out["snell2_in"] = np.concatenate([s, e, ang[:, None], n_in[:, None], n_out[:, None]], axis=1)
    out["snell2_out"] = np.stack([r.numpy() for r in res], axis=1)

    # angle_in_interval (geometry.py:766-802), wrap-around windows included
    k = 200
    ang = rng.uniform(-2 * math.pi, 2 * math.pi, size=k)
    a1 = rng.uniform(-math.pi, 2 * math.pi, size=k)
    a2 = rng.uniform(-math.pi, 2 * math.pi, size=k)
    out["angle_in"] = np.stack([ang, a1, a2], axis=1)
    out["angle_out"] = g.angle_in_interval(T(ang), T(a1), T(a2)).numpy()

    # materials (materials.py:25-104), wavelength in nm
    wl = np.linspace(380.0, 780.0, 41)
    out["mat_wavelength"] = wl
    for name in ("vacuum", "acrylic", "crown_glass", "flint_glass", "fused_silica", "polycarbonate",
                 "soda_lime", "reflective"):
        out[f"mat_{name}"] = np.broadcast_to(
            np.asarray(getattr(ref.materials, name)(T(wl)), dtype=np.float64), wl.shape).copy()
    out["mat_constant_1p7"] = np.asarray(ref.materials.build_constant_material(1.7)(T(wl)),
                                         dtype=np.float64)
    return out


# ------------------------------------------------------------------------------- sources (8(f) rank 2)

def source_cases():
    """name -> (kind, constructor arguments) shared by the generator (reference classes) and the test
    (tensorflowraytrace_b200 classes).  Only deterministic (Static*) distributions.  2D Point /
    AngularSource cannot be built in the reference at HEAD (`center` setter rejects vectors,
    tfrt/sources.py:657-666, :889-898) and are therefore absent."""
    PI = math.pi
    return {
        "dist_uniform_angular": ("StaticUniformAngularDistribution", (-0.2, 0.3, 7), {}),
        "dist_lambertian_angular": ("StaticLambertianAngularDistribution", (-1.0, 0.5, 9), {}),
        "dist_beam": ("StaticUniformBeam", (-1.5, 1.5, 6), {}),
        "dist_beam_rotated": ("StaticUniformBeam", (-1.0, 3.0, 5), {"central_angle": PI / 2}),
        "dist_aperature": ("StaticUniformAperaturePoints", ((0.0, -1.0), (2.0, 1.0), 5), {}),
        "dist_square": ("StaticUniformSquare", (2.0, 3, 1.0, 4), {}),
        "dist_circle": ("StaticUniformCircle", (50,), {"radius": 2.0}),
        "dist_sphere": ("StaticUniformSphere", (PI / 8, 40), {}),
        "dist_lambertian_sphere": ("StaticLambertianSphere", (PI / 3, 40), {}),
        "src_point3d": ("PointSource", (3, (-1.0, 0.0, 0.0), (1.0, 0.0, 0.0),
                                        ("StaticUniformSphere", (PI / 8, 12), {}), [575.0, 450.0]), {}),
        "src_point3d_rotated": ("PointSource", (3, (-1.0, 0.5, 0.25), (1.0, 1.0, 0.0),
                                                ("StaticUniformSphere", (PI / 8, 12), {}), [575.0]),
                                {"ray_length": 2.5}),
        "src_angular3d_dense": ("AngularSource", (3, (-4.0, 0.0, 0.0), (1.0, 0.0, 0.0),
                                                  ("StaticUniformSphere", (PI / 16, 5), {}),
                                                  ("StaticUniformSquare", (0.2, 3), {}), [575.0, 680.0]), {}),
        "src_angular3d_undense": ("AngularSource", (3, (-4.0, 0.1, 0.0), (1.0, 0.0, 0.0),
                                                    ("StaticUniformSphere", (PI / 16, 9), {}),
                                                    ("StaticUniformSquare", (0.2, 3), {}), [575.0] * 9),
                                  {"dense": False}),
        "src_aperature2d": ("AperatureSource", (2, ("StaticUniformAperaturePoints", ((0.0, -1.0), (0.0, 1.0), 3), {}),
                                                ("StaticUniformAperaturePoints", ((2.0, -1.0), (2.0, 1.0), 4), {}),
                                                [575.0, 450.0]), {}),
    }


def build_case(dist_mod, src_mod, spec):
    """Instantiate a case from `source_cases()` with the given distributions / sources modules."""
    kind, args, kwargs = spec

    def arg(a):
        if isinstance(a, tuple) and len(a) == 3 and isinstance(a[0], str) and isinstance(a[2], dict):
            return build_case(dist_mod, src_mod, a)
        return a
    cls = getattr(src_mod, kind, None) or getattr(dist_mod, kind)
    return cls(*[arg(a) for a in args], **kwargs)


def transformation_case(dist_mod):
    """A StaticUniformSquare scaled, rotated by 90 degrees about z and translated (reference:
    BasePointTransformation, tfrt/distributions.py:2014-2120); returns the distribution."""
    base = dist_mod.StaticUniformSquare(0.5, 3, 0.25, 2)
    q = (math.cos(math.pi / 4), 0.0, 0.0, math.sin(math.pi / 4))
    keep = dist_mod.BasePointTransformation(base, rotation=q, translation=(1.0, 2.0, 3.0),
                                            scale=(1.0, 2.0, 0.5))
    base.update()
    base._keep_transformation_alive = keep
    return base


def source_vectors(ref):
    import tfrt.distributions as rd
    out = {}
    out["transformed_square__points"] = np.asarray(transformation_case(rd).points, dtype=np.float64)
    for name, spec in source_cases().items():
        obj = build_case(rd, ref.sources, spec)
        if name.startswith("dist_"):
            for attr in ("angles", "points", "ranks"):
                if hasattr(obj, attr) and getattr(obj, attr) is not None:
                    out[f"{name}__{attr}"] = np.asarray(getattr(obj, attr), dtype=np.float64)
        else:
            for key in sorted(obj.keys()):
                out[f"{name}__{key}"] = np.asarray(obj[key], dtype=np.float64)
    return out


# ------------------------------------------------------------------------------- optimiser (8(f) rank 1)

OPT_STEPS = ((1.0, 0.0), (1.0, 0.6), (0.5, 0.6), (0.5, 0.3))     # (lr_scale, momentum) per step
OPT_KW = dict(learning_rate=0.004, individual_lr=[1.0, 0.5])


def optimizer_scene():
    sys.path.insert(0, ROOT)
    from tensorflowraytrace_b200 import scenes
    return scenes.lens_optimize_config(ray_count=160, step_count=3, seed=77, name="golden_opt")


def optimizer_accumulators(scene):
    """A synthetic dense accumulator for the first surface (the reference only requires a (V, V)
    matrix, tfrt/optimizer.py:251-256), none for the second."""
    v = scene["surfaces"][0]["parameters"].shape[0]
    rng = np.random.default_rng(5)
    acc = (rng.uniform(size=(v, v)) < 0.15).astype(np.float64) + np.eye(v)
    return [acc, None]


def optimizer_vectors(ref):
    """`tfrt.optimizer.SGD_Optimizer.single_step` (Keras SGD(nesterov=True) at its default 0.01 step,
    tfrt/optimizer.py:103, :284-320) driving the reference engine: parameter trajectories."""
    import tfrt.optimizer as ropt
    tf = ref.tf
    scene = optimizer_scene()
    eng, system, params = reference_engine_3d(ref, scene)
    goal_scale = scene["goal_scale"]

    def error_function(engine):
        output = tf.stack([engine.finished_rays["y_end"], engine.finished_rays["z_end"]], axis=1)
        return tf.math.squared_difference(output, engine.finished_rays["rank"] * goal_scale)

    opt = ropt.SGD_Optimizer(eng, params, error_function, scene["max_iterations"], **OPT_KW)
    accs = [None if a is None else torch.from_numpy(a) for a in optimizer_accumulators(scene)]
    out = {"initial_0": params[0].numpy().copy(), "initial_1": params[1].numpy().copy()}
    grads, err = opt.process_gradient(accs)
    out["processed_0"], out["processed_1"] = grads[0].numpy(), grads[1].numpy()
    out["mean_error_0"] = np.array(float(err))
    errors = []
    for k, (lr_scale, momentum) in enumerate(OPT_STEPS):
        errors.append(float(opt.single_step(accs, lr_scale=lr_scale, momentum=momentum)))
        out[f"step{k}_0"], out[f"step{k}_1"] = params[0].numpy().copy(), params[1].numpy().copy()
    out["errors"] = np.array(errors)
    return out


# ------------------------------------------------------------------------------- multi-surface flow

MULTI_STEPS = ((1.0, 0.0), (1.0, 0.7), (0.5, 0.7))       # (lr_scale, momentum) per step
MULTI_KW = dict(learning_rate=0.002)
MULTI_RINGS, MULTI_SIDE = 4, 8


def multisurface_inputs():
    """Plain-data description shared by the generator (reference classes) and the GPU test (ours):
    hexagonal zero-point disk in the y-z plane, initial parameters of the two surfaces."""
    sys.path.insert(0, ROOT)
    from tensorflowraytrace_b200 import mesh_tools
    pts, faces = mesh_tools.hexagonal_mesh(1.0, MULTI_RINGS)
    verts = np.stack([np.zeros(len(pts)), pts[:, 0], pts[:, 1]], axis=1)
    r2 = pts[:, 0] ** 2 + pts[:, 1] ** 2
    return verts, faces, [-0.1 * (1.0 - r2), 0.25 + 0.1 * (1.0 - r2)]


def multisurface_vectors(ref):
    """The optimisation flow of dev/simple_3d_optimize.py:22-140 on the reference's own classes:
    `mesh_tools.get_closest_point / mesh_parametrization_tools / mesh_smoothing_tool`,
    `ParametricMultiTriangleBoundary` + `ThicknessConstraint`s + `vertex_update_map`, an undense
    `AngularSource` with a rank field, `SGD_Optimizer.single_step` with accumulators, `smooth`."""
    import tfrt.distributions as rd
    import tfrt.mesh_tools as rmt
    import tfrt.optimizer as ropt
    tf = ref.tf
    verts, faces, init = multisurface_inputs()
    zero = _polydata(ref, verts, faces)
    top = rmt.get_closest_point(zero, (0, 0, 0))
    vum, acc = rmt.mesh_parametrization_tools(zero, top)
    smo = rmt.mesh_smoothing_tool(zero, [1.0, 0.02])
    out = {"top_parent": np.array(top), "vertex_update_map": np.asarray(vum).astype(np.bool_),
           "accumulator": np.asarray(acc, dtype=np.float64), "smoother": np.asarray(smo, dtype=np.float64)}
    base_points = rd.StaticUniformSquare(0.2, MULTI_SIDE)
    angles = rd.StaticUniformSphere(math.pi / 16.0, MULTI_SIDE * MULTI_SIDE)
    source = ref.sources.AngularSource(
        3, (-4.0, 0.0, 0.0), (1.0, 0.0, 0.0), angles, base_points, [575.0] * (MULTI_SIDE * MULTI_SIDE),
        dense=False, extra_fields={"rank": ("base_point", base_points, "ranks")})
    mat = {"mat_in": tf.constant(1, dtype=tf.int64), "mat_out": tf.constant(0, dtype=tf.int64)}
    lens = ref.boundaries.ParametricMultiTriangleBoundary(
        zero, ref.boundaries.FromVectorVG((1, 0, 0)),
        [ref.boundaries.ThicknessConstraint(0.0, "min"), ref.boundaries.ThicknessConstraint(0.2, "min")],
        [True, False], initial_parameters=[torch.from_numpy(f32_round(p)) for p in init],
        material_list=[mat, mat], vertex_update_map=vum)
    tv = np.array([[8.0, -50.0, -50.0], [8.0, 50.0, -50.0], [8.0, 50.0, 50.0], [8.0, -50.0, 50.0]])
    target = ref.boundaries.ManualTriangleBoundary(mesh=_polydata(ref, tv, np.array([[0, 1, 2], [0, 2, 3]])))
    system = ref.engine.OpticalSystem3D(**EPS)
    system.optical = lens.surfaces
    system.targets = [target]
    system.sources = [source]
    system.materials = [{"n": ref.materials.vacuum}, {"n": ref.materials.acrylic}]
    system.update()
    eng = ref.engine.OpticalEngine(3, [ref.operation.StandardReaction()],
                                   simple_ray_inheritance={"wavelength", "rank"})
    eng.optical_system = system
    eng.validate_system()
    for i, p in enumerate(lens.parameters):                     # constraints applied by update()
        out[f"constrained_{i}"] = p.numpy().copy()
    for f in ("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2", "norm"):
        out[f"merged_{f}"] = system._merged[f].numpy().astype(np.float64)

    def error_function(engine):
        output = tf.stack([engine.finished_rays["y_end"], engine.finished_rays["z_end"]], axis=1)
        return tf.math.squared_difference(output, engine.finished_rays["rank"] * (-0.4))

    opt = ropt.SGD_Optimizer(eng, lens.parameters, error_function, 3, **MULTI_KW)
    accs = [torch.from_numpy(out["accumulator"])] * 2
    smoother = torch.from_numpy(out["smoother"])
    errors = []
    for k, (lr_scale, momentum) in enumerate(MULTI_STEPS):
        errors.append(float(opt.single_step(accs, lr_scale=lr_scale, momentum=momentum)))
        for p in lens.parameters:
            opt.smooth(p, smoother)
        for i, p in enumerate(lens.parameters):
            out[f"step{k}_{i}"] = p.numpy().copy()
    system.update()                                             # re-applies the constraints
    for i, p in enumerate(lens.parameters):
        out[f"final_{i}"] = p.numpy().copy()
    out["errors"] = np.array(errors)
    return out


# ------------------------------------------------------------------------------- two sources + OldestAncestor

TWO_SPLIT = 70


def two_source_scene():
    sys.path.insert(0, ROOT)
    from tensorflowraytrace_b200 import scenes
    return scenes.lens_optimize_config(ray_count=180, step_count=3, seed=5, name="golden_two")


def two_source_vectors(ref):
    """Two `ManualSource`s in one system (amalgamation order, tfrt/engine.py:203-216) traced with
    `[StandardReaction(), OldestAncestor()]` (tfrt/operation.py:166-196): every history set carries
    the running source index of the ray it descends from."""
    scene = two_source_scene()
    eng, system, params = reference_engine_3d(ref, scene)           # boundaries, materials
    srcs = []
    for sl in (slice(0, TWO_SPLIT), slice(TWO_SPLIT, None)):
        src = ref.sources.ManualSource(3)
        for k, v in scene["sources"].items():
            src[k] = torch.from_numpy(np.ascontiguousarray(f32_round(v[sl])))
        srcs.append(src)
    system.sources = srcs
    eng2 = ref.engine.OpticalEngine(3, [ref.operation.StandardReaction(), ref.operation.OldestAncestor()],
                                    compile_dead_rays=True, simple_ray_inheritance={"wavelength", "rank"})
    eng2.optical_system = system
    eng2.annotate()
    system.update()
    eng2.validate_system()
    eng2.ray_trace(scene["max_iterations"])
    out = {}
    for set_name in ("active", "finished", "dead"):
        s = getattr(eng2, set_name + "_rays")
        try:
            anc = s["oldest_ancestor"]
        except KeyError:
            continue
        if anc.shape[0] == 0:
            continue
        out[f"{set_name}_oldest_ancestor"] = anc.numpy().astype(np.int64)
        for f in RAYS3:
            out[f"{set_name}_{f}"] = s[f].numpy().astype(np.float64)
    return out


# ------------------------------------------------------------------------------- a12 variants

def boundary_inputs():
    sys.path.insert(0, ROOT)
    from tensorflowraytrace_b200 import mesh_tools
    pts, faces = mesh_tools.hexagonal_mesh(1.0, 3)
    verts = np.stack([np.zeros(len(pts)), pts[:, 0], pts[:, 1]], axis=1)
    masters = [i for i in range(len(verts)) if verts[i, 2] >= -1e-9]      # mirror symmetry in z
    init = 0.1 * (1.0 - pts[:, 0] ** 2 - pts[:, 1] ** 2)
    return verts, faces, masters, init


def attach_mirror_slaves(v, m, free):
    v = np.asarray(v.detach().cpu() if hasattr(v, "detach") else v)
    return {s for s in free if abs(v[s, 1] - v[m, 1]) < 1e-6 and abs(v[s, 2] + v[m, 2]) < 1e-6}


GUIDE_ARGS = ((0.0, 0.0, 0.0), (4.0, 0.0, 0.0), 0.5)
# caps off: with caps the reference prunes the cap vertices out of `_vertices` but keeps the faces
# that index them (tfrt/boundaries.py:1608, :1703-1717), so its own gather goes out of range
GUIDE_KW = dict(theta_res=10, z_res=7, initial_taper=(0.05, 0.2), start_cap=False, end_cap=False)


def boundary_weights(n_faces, seed):
    rng = np.random.default_rng(seed)
    return rng.normal(size=(n_faces, 12))


def _boundary_record(ref, out, name, b):
    """Fields of a reference boundary + tape gradient of a fixed random functional of them."""
    tf = ref.tf
    n = int(b["xp"].shape[0])
    w = torch.from_numpy(boundary_weights(n, n))
    with tf.GradientTape() as tape:
        b.update()
        cols = [b[k] for k in ("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2")]
        loss = sum((c * w[:, i]).sum() for i, c in enumerate(cols)) + (b["norm"] * w[:, 9:12]).sum()
        grad = tape.gradient(loss, b.parameters)
    out[f"{name}_parameters"] = b.parameters.numpy().copy()
    out[f"{name}_vertices"] = b._vertices.numpy().copy()
    for i, k in enumerate(("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2")):
        out[f"{name}_{k}"] = cols[i].numpy().copy()
    out[f"{name}_norm"] = b["norm"].numpy().copy()
    out[f"{name}_faces"] = np.asarray(b._faces)[:, 1:4].astype(np.int64)
    out[f"{name}_grad"] = grad.numpy().copy()
    out[f"{name}_loss"] = np.array(float(loss.detach()))


def boundary_vectors(ref):
    """`MasterSlaveParametricTriangleBoundary` (gather, tfrt/boundaries.py:1116-1224) and
    `ParametricCylindricalGuide` (min-pinning constraint, repeat / pad, :1416-1611), full and
    rotationally symmetric: parameters -> vertices -> fields -> norm, and the tape gradient."""
    B = ref.boundaries
    out = {}
    verts, faces, masters, init = boundary_inputs()
    ms = B.MasterSlaveParametricTriangleBoundary(
        masters, attach_mirror_slaves, _polydata(ref, verts, faces), B.FromVectorVG((1, 0, 0)),
        initial_parameters=torch.from_numpy(f32_round(init)), material_dict={})
    out["ms_gather"] = np.asarray(ms._gather).astype(np.int64)
    _boundary_record(ref, out, "ms", ms)
    # (material_dict cannot be passed to the guide: the reference forwards **kwargs to
    # mesh_tools.cylindrical_mesh, tfrt/boundaries.py:1525-1534)
    for name, sym in (("guide", False), ("guide_sym", True)):
        g = B.ParametricCylindricalGuide(*GUIDE_ARGS, rotationally_symmetric=sym, **GUIDE_KW)
        with torch.no_grad():                       # a non-trivial shape: bump the parameters
            bump = 0.03 * torch.sin(torch.arange(g.parameters.shape[0], dtype=torch.float64) * 0.7)
            g.parameters.add_(torch.from_numpy(f32_round(bump.numpy())))
        _boundary_record(ref, out, name, g)
        out[f"{name}_accumulator"] = np.asarray(g._accumulator, dtype=np.float64)
        out[f"{name}_vertex_update_map"] = np.asarray(g._vertex_update_map).astype(np.bool_)
    return out


# ------------------------------------------------------------------------------- 2D parametric lens

PARAM2D_POINTS, PARAM2D_RAYS = 9, 120


def param2d_inputs():
    """Plain data of the 2D optimisation scene: a parametric front face (points sliding from x = 0
    towards x = -0.3), a flat back face, rays from x = -3, per-ray goal -0.4 * start height."""
    y = np.linspace(-1.0, 1.0, PARAM2D_POINTS)
    rng = np.random.default_rng(21)
    y0 = rng.uniform(-0.5, 0.5, PARAM2D_RAYS)
    ang = rng.uniform(-0.12, 0.12, PARAM2D_RAYS)
    rays = {"x_start": np.full(PARAM2D_RAYS, -3.0), "y_start": y0,
            "x_end": -3.0 + np.cos(ang), "y_end": y0 + np.sin(ang),
            "wavelength": np.asarray([680.0, 575.0, 450.0])[np.arange(PARAM2D_RAYS) % 3], "rank": y0}
    return {"zero": ((0.0, -1.0), (0.0, 1.0)), "one": ((-0.3, -1.0), (-0.3, 1.0)),
            "init": 0.8 * (1.0 - y ** 2), "back": np.array([[0.4, 1.0, 0.4, -1.0]]),
            "target": np.array([[6.0, -6.0, 6.0, 6.0]]), "rays": rays, "goal_scale": -0.4}


def param2d_vectors(ref):
    """`ParametricSegmentBoundary` (tfrt/boundaries.py:528-627) in a 2D system, traced by the
    reference engine, with the GradientTape gradient of an imaging loss to its parameters: pins the
    2D backward path (segment geometry, angle-form Snell) end to end."""
    import tfrt.distributions as rd
    tf = ref.tf
    d = param2d_inputs()
    zero = rd.StaticUniformAperaturePoints(*d["zero"], PARAM2D_POINTS)
    one = rd.StaticUniformAperaturePoints(*d["one"], PARAM2D_POINTS)
    # One repair, stated: SegmentBoundaryBase.update_materials passes `self` twice
    # (tfrt/boundaries.py:485-489) and raises TypeError as soon as the fields exist, i.e. in every
    # ParametricSegmentBoundary constructor.  It is replaced by a no-op for this fixture; the
    # materials are set with annotation_helper, as the reference's scripts do.
    ref.boundaries.SegmentBoundaryBase.update_materials = lambda self: None
    front = ref.boundaries.ParametricSegmentBoundary(
        zero, one, initial_parameters=torch.from_numpy(f32_round(d["init"])))
    back = ref.boundaries.ManualSegmentBoundary()
    back.feed_segments(f32_round(d["back"]))
    for b in (front, back):
        ref.engine.annotation_helper(b, "mat_in", 1, "x_start", dtype=tf.int64)
        ref.engine.annotation_helper(b, "mat_out", 0, "x_start", dtype=tf.int64)
    target = ref.boundaries.ManualSegmentBoundary()
    target.feed_segments(f32_round(d["target"]))
    src = ref.sources.ManualSource(2)
    for k, v in d["rays"].items():
        src[k] = torch.from_numpy(np.ascontiguousarray(f32_round(v)))
    src["slot"] = torch.arange(PARAM2D_RAYS, dtype=torch.float64)
    system = ref.engine.OpticalSystem2D(**EPS)
    system.optical_segments = [front, back]
    system.target_segments = [target]
    system.sources = [src]
    system.materials = [{"n": ref.materials.vacuum}, {"n": ref.materials.acrylic}]
    system.update()
    eng = ref.engine.OpticalEngine(2, [ref.operation.StandardReaction()], compile_dead_rays=True,
                                   simple_ray_inheritance={"wavelength", "slot", "rank"})
    eng.optical_system = system
    eng.validate_system()
    with tf.GradientTape() as tape:
        system.update()
        eng.clear_ray_history()
        eng.ray_trace(3)
        fin = eng.finished_rays
        error = tf.math.squared_difference(fin["y_end"], fin["rank"] * d["goal_scale"])
        grad = tape.gradient(error, front.parameters)
    out = {"grad": grad.numpy().copy(), "loss": np.array(float(error.sum().detach())),
           "parameters": front.parameters.numpy().copy()}
    for k in ("x_start", "y_start", "x_end", "y_end"):
        out[f"front_{k}"] = front[k].numpy().copy()
    _record_sets(out, eng, RAYS2)
    return out


# ------------------------------------------------------------------------------- mesh generators

def mesh_cases():
    return {
        "hex_1": ("hexagonal_mesh", dict(radius=1.0, step_count=1)),
        "hex_2": ("hexagonal_mesh", dict(radius=0.37, step_count=2)),
        "hex_7": ("hexagonal_mesh", dict(radius=1.0, step_count=7)),
        "cyl_capped": ("cylindrical_mesh", dict(start=(0, 0, 0), end=(4, 0, 0), radius=0.5, theta_res=8, z_res=5)),
        "cyl_open": ("cylindrical_mesh", dict(start=(1, 2, 3), end=(1.5, -2, 4), radius=0.3, theta_res=10,
                                              z_res=7, start_cap=False, end_cap=False)),
        "cyl_one_cap_twist": ("cylindrical_mesh", dict(start=(0, 0, 0), end=(0, 0, 2), radius=1.0, theta_res=6,
                                                       z_res=4, start_cap=True, end_cap=False, use_twist=True)),
    }


def mesh_vectors(ref):
    """Vertex and face arrays of the reference's own mesh generators (tfrt/mesh_tools.py:713-952)."""
    import tfrt.mesh_tools as rmt
    out = {}
    for name, (fn, kw) in mesh_cases().items():
        pd = getattr(rmt, fn)(**kw)
        out[f"{name}_points"] = np.asarray(pd.points, dtype=np.float64)
        out[f"{name}_faces"] = np.asarray(pd.faces).reshape(-1, 4)[:, 1:].astype(np.int64)
    return out


# ------------------------------------------------------------------------------- PrecompiledSource file

def precompiled_file(ref):
    """A file written by the reference's own `PrecompiledSource.save` (tfrt/sources.py:1206-1215)
    from a 3D `PointSource`: the data format either side of the path (pickle of dimension,
    standard_domains, fields)."""
    import tfrt.distributions as rd
    src = ref.sources.PointSource(3, (-1.0, 0.0, 0.0), (1.0, 0.0, 0.0),
                                  rd.StaticUniformSphere(math.pi / 8, 12), [575.0, 450.0])
    pre = ref.sources.PrecompiledSource(src, do_downsample=False)
    path = os.path.join(HERE, "ref_precompiled_source.pkl")
    pre.save(path)
    return path


# ------------------------------------------------------------------------------- projection result

def projection_scene():
    """Rays that reach every category in ONE pass: the mirror / stop scene's rays moved behind the
    lens (x = 1) so that they meet the stop, the mirror, nothing, or -- aimed sideways -- the target."""
    sys.path.insert(0, ROOT)
    from tensorflowraytrace_b200 import scenes
    scene = scenes.mirror_stop_3d(ray_count=240, seed=3)
    src = scene["sources"]
    n = src["x_start"].shape[0]
    shift = 5.0
    for k in ("x_start", "x_end"):
        src[k] = src[k] + shift
    down = np.arange(n) % 4 == 0                       # a quarter of the rays go straight to the target
    src["x_end"] = np.where(down, src["x_start"], src["x_end"])
    src["y_end"] = np.where(down, src["y_start"] - 1.0, src["y_end"])
    src["z_end"] = np.where(down, src["z_start"], src["z_end"])
    up = np.arange(n) % 5 == 1                         # and a fifth leave upwards: they hit nothing
    src["x_end"] = np.where(up, src["x_start"], src["x_end"])
    src["y_end"] = np.where(up, src["y_start"] + 0.5, src["y_end"])
    return scene


def _flatten(prefix, d, out):
    for k, v in d.items():
        if isinstance(v, dict):
            _flatten(f"{prefix}{k}.", v, out)
        else:
            out[f"{prefix}{k}"] = np.asarray(v.numpy() if hasattr(v, "numpy") else v)


def projection_vectors(ref):
    """`OpticalEngine.process_projection_3D` (tfrt/engine.py:1988-2191) with
    compile_technical_intersections, all four compile flags and dead_ray_length: the full nested
    result dict of one pass (rays by category + gathered optical / stop / target boundary fields)."""
    scene = projection_scene()
    eng, system, _ = reference_engine_3d(ref, scene, compile_technical_intersections=True,
                                         dead_ray_length=5.0)
    rays = {k: v for k, v in system._amalgamated_sources.items()}
    result = eng.process_projection(dict(rays))
    out = {}
    _flatten("", result, out)
    return out


# ------------------------------------------------------------------------------- 2D projection result

def projection_scene_2d():
    """Rays placed behind the prism of `segments_2d` so that ONE pass meets every category: the
    target, the stop, optical faces (prism, mirror) and nothing."""
    sys.path.insert(0, ROOT)
    from tensorflowraytrace_b200 import scenes
    scene = scenes.segments_2d()
    n = 200
    rng = np.random.default_rng(8)
    y0 = rng.uniform(-0.5, 2.5, n)
    ang = rng.uniform(-0.6, 0.9, n)
    ang = np.where(np.arange(n) % 4 == 1, math.pi + rng.uniform(-0.5, 0.5, n), ang)   # backwards: prism
    ang = np.where(np.arange(n) % 5 == 2, math.pi / 2 + rng.uniform(-0.1, 0.1, n), ang)  # upwards: nothing
    wl = np.asarray(scenes.RAINBOW_6)[np.arange(n) % 6]
    scene["sources"] = scenes.rays_2d(np.stack([np.full(n, 3.6), y0], 1), ang, wl)
    return scene


def projection_vectors_2d(ref):
    """`OpticalEngine.process_projection_2D` (tfrt/engine.py:1544-1986), segments only, with
    compile_technical_intersections, all compile flags and dead_ray_length."""
    scene = projection_scene_2d()
    eng, system = reference_engine_2d(ref, scene, compile_technical_intersections=True,
                                      dead_ray_length=3.0)
    result = eng.process_projection(dict(system._amalgamated_sources))
    out = {}
    _flatten("", result, out)
    return out


# ------------------------------------------------------------------------------- GhostThrough (hook path)

def ghost_vectors(ref):
    """The operation-hook path (tfrt/engine.py:2193-2302 `single_pass`, tfrt/operation.py:311-333):
    the segments-only 2D scene traced with `[GhostThrough()]` instead of StandardReaction -- rays
    pass straight through every optical segment.  (2D only: the reference's 3D branch copies
    `z_start` instead of `z_end`, operation.py:330.)"""
    scene = golden_scenes()["ref_segments2d"]
    eng, system = reference_engine_2d(ref, scene)
    ghost = ref.engine.OpticalEngine(2, [ref.operation.GhostThrough()], compile_dead_rays=True,
                                     compile_stopped_rays=True,
                                     simple_ray_inheritance={"wavelength", "slot"})
    ghost.optical_system = system
    ghost.validate_system()
    ghost.ray_trace(5)
    out = {}
    _record_sets(out, ghost, RAYS2)
    return out


def main():
    ref = import_reference()
    data = function_vectors(ref)
    path = os.path.join(HERE, "ref_functions.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays, {os.path.getsize(path)} bytes")
    data = source_vectors(ref)
    path = os.path.join(HERE, "ref_sources.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays, {os.path.getsize(path)} bytes")
    print(precompiled_file(ref), os.path.getsize(os.path.join(HERE, "ref_precompiled_source.pkl")), "bytes")
    data = projection_vectors(ref)
    path = os.path.join(HERE, "ref_projection.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays:", {k: v.shape for k, v in data.items() if k.endswith("x_start")})
    data = ghost_vectors(ref)
    path = os.path.join(HERE, "ref_ghost2d.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays:", {k: v.shape for k, v in data.items() if k.endswith("slot")})
    data = projection_vectors_2d(ref)
    path = os.path.join(HERE, "ref_projection2d.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays:", {k: v.shape for k, v in data.items() if k.endswith("x_start")})
    data = mesh_vectors(ref)
    path = os.path.join(HERE, "ref_meshes.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays, {os.path.getsize(path)} bytes")
    data = param2d_vectors(ref)
    path = os.path.join(HERE, "ref_param2d.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays, loss {data['loss']}, |grad| max {np.abs(data['grad']).max()}, finished {data['finished_slot'].shape}")
    data = boundary_vectors(ref)
    path = os.path.join(HERE, "ref_boundaries.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays, {os.path.getsize(path)} bytes")
    data = two_source_vectors(ref)
    path = os.path.join(HERE, "ref_two_sources.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays, sets { {k: v.shape for k, v in data.items() if k.endswith('ancestor')} }")
    data = multisurface_vectors(ref)
    path = os.path.join(HERE, "ref_multisurface.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays, errors {data['errors']}")
    data = optimizer_vectors(ref)
    path = os.path.join(HERE, "ref_optimizer.npz")
    np.savez_compressed(path, **data)
    print(f"{path}: {len(data)} arrays, errors {data['errors']}")
    for name, scene in golden_scenes().items():
        data = run_reference(ref, name, scene)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **data)
        sizes = {k: v.shape for k, v in data.items() if k.endswith("_slot")}
        print(f"{path}: {len(data)} arrays, {os.path.getsize(path)} bytes, sets {sizes}")


if __name__ == "__main__":
    main()
