"""Run the REFERENCE'S OWN SOURCE on a stand-in TensorFlow.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference (`tfrt`, pure Python on TensorFlow) cannot be imported as is: TensorFlow, tfquaternion
and pyvista are absent from this image.  `oracle/tf_shim/` holds a stand-in `tensorflow` module (~70 tf
ops mapped onto the IEEE-equivalent torch float64 CPU ops, `tf.GradientTape` served by torch autograd)
plus stubs for `tfquaternion`, `pyvista`, `imageio`.  With it on `sys.path` the unmodified reference
modules import and run.  Two users:

  * tests/golden/make_reference_golden.py (build container, reference at /root/reference): generates
    the committed `tests/golden/ref_*.npz` vectors that pin the oracle and the CUDA path;
  * bench.py `--impl reference` / `cpu_baseline` (GPU box, reference pip-installed under
    `baseline/_ref/`, git-ignored): times the reference's own `OpticalEngine.ray_trace` (+ GradientTape)
    beside the oracle port.

Where the reference is looked for: $TFRT_REFERENCE, /root/reference, <repo>/baseline/_ref.
Nothing in the product imports this module.
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# fp32-appropriate epsilons shared with the CUDA parity tests (tests/helpers.py EPS)
EPS = dict(intersect_epsilion=1e-10, size_epsilion=1e-5, ray_start_epsilion=1e-4)
RAYS3 = ("x_start", "y_start", "z_start", "x_end", "y_end", "z_end")
RAYS2 = ("x_start", "y_start", "x_end", "y_end")


def find_reference(installed_only=False):
    """Directory that contains the `tfrt` package of the reference, or None.  installed_only: look
    only at the pip-installed copy under baseline/_ref (bench.py must not read /root/reference)."""
    cands = (os.path.join(ROOT, "baseline", "_ref"),) if installed_only else (
        os.environ.get("TFRT_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref"))
    for cand in cands:
        if cand and os.path.isfile(os.path.join(cand, "tfrt", "engine.py")):
            return cand
    return None


def f32_round(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def import_reference(installed_only=False):
    """Returns the reference's modules, imported from `find_reference()` with the tf stand-in."""
    reference = find_reference(installed_only)
    if reference is None:
        raise RuntimeError("reference_runner: no tfrt package found (TFRT_REFERENCE, /root/reference, "
                           "baseline/_ref)")
    for p in (reference, os.path.join(HERE, "tf_shim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import tensorflow as tf  # the stand-in
    assert tf.__version__ == "2.shim", "a real tensorflow is importable: use it instead of the shim"
    import tfrt.boundaries as boundaries
    import tfrt.engine as engine
    import tfrt.geometry as geometry
    import tfrt.materials as materials
    import tfrt.operation as operation
    import tfrt.sources as sources
    import pyvista as pv
    assert os.path.abspath(engine.__file__).startswith(os.path.abspath(reference))
    ref_dir = reference

    class Ref:
        pass
    ref = Ref()
    ref.tf, ref.pv, ref.path = tf, pv, ref_dir
    ref.boundaries, ref.engine, ref.geometry = boundaries, engine, geometry
    ref.materials, ref.operation, ref.sources = materials, operation, sources
    return ref


# ------------------------------------------------------------------------------- engine level

def _polydata(ref, vertices, faces):
    faces = np.asarray(faces, dtype=np.int64)
    packed = np.concatenate([np.full((faces.shape[0], 1), 3, dtype=np.int64), faces], axis=1)
    return ref.pv.PolyData(f32_round(vertices), packed.reshape(-1))


def _materials(ref, names):
    return [{"n": getattr(ref.materials, n)} for n in names]


def _manual_source(ref, scene, dimension, ray_slice=None):
    src = ref.sources.ManualSource(dimension)
    n = None
    for k, v in scene["sources"].items():
        v = v if ray_slice is None else v[ray_slice]
        src[k] = torch.from_numpy(np.ascontiguousarray(f32_round(v)))
        n = v.shape[0]
    src["slot"] = torch.arange(n, dtype=torch.float64)
    return src


def reference_engine_3d(ref, scene, ray_slice=None, **engine_kwargs):
    """tfrt objects for a 3D scene: parametric surfaces as `ParametricTriangleBoundary` along
    `FromVectorVG` (pattern of dev/simple_3d_optimize.py:69-81; flip already applied to the scene's
    faces), static surfaces (zero vectors) and targets as `ManualTriangleBoundary`."""
    tf = ref.tf
    optical, params = [], []
    for s in scene["surfaces"]:
        mat = {"mat_in": tf.constant(s["mat_in"], dtype=tf.int64),
               "mat_out": tf.constant(s["mat_out"], dtype=tf.int64)}
        if np.any(s["vectors"] != 0.0):
            vec = f32_round(s["vectors"])
            assert np.allclose(vec, vec[0]) and abs(np.linalg.norm(vec[0]) - 1.0) < 1e-12
            surf = ref.boundaries.ParametricTriangleBoundary(
                _polydata(ref, s["zero_points"], s["faces"]),
                ref.boundaries.FromVectorVG(tuple(vec[0])),
                initial_parameters=torch.from_numpy(f32_round(s["parameters"])),
                material_dict=mat)
            params.append(surf.parameters)
        else:
            verts = s["zero_points"] + s["parameters"][:, None] * s["vectors"]
            surf = ref.boundaries.ManualTriangleBoundary(
                mesh=_polydata(ref, verts, s["faces"]), material_dict=mat)
        optical.append(surf)
    targets = [ref.boundaries.ManualTriangleBoundary(mesh=_polydata(ref, m["vertices"], m["faces"]))
               for m in scene["targets"]]
    system = ref.engine.OpticalSystem3D(**EPS)
    system.optical = optical
    if scene["stops"]:
        system.stops = [ref.boundaries.ManualTriangleBoundary(
            mesh=_polydata(ref, m["vertices"], m["faces"])) for m in scene["stops"]]
    system.targets = targets
    system.sources = [_manual_source(ref, scene, 3, ray_slice)]
    system.materials = _materials(ref, scene["materials"])
    system.update()
    inherit = {k for k in scene["sources"] if k not in RAYS3} | {"slot"}
    kw = dict(compile_dead_rays=True, compile_stopped_rays=True)
    kw.update(engine_kwargs)
    eng = ref.engine.OpticalEngine(3, [ref.operation.StandardReaction()],
                                   simple_ray_inheritance=inherit, **kw)
    eng.optical_system = system
    eng.validate_system()
    return eng, system, params


def reference_engine_2d(ref, scene, ray_slice=None, **engine_kwargs):
    tf = ref.tf

    def segs(items, with_mat):
        out = []
        for it in items:
            b = ref.boundaries.ManualSegmentBoundary()
            b.feed_segments(f32_round(it["segments"]))
            if with_mat:
                ref.engine.annotation_helper(b, "mat_in", it["mat_in"], "x_start", dtype=tf.int64)
                ref.engine.annotation_helper(b, "mat_out", it["mat_out"], "x_start", dtype=tf.int64)
            out.append(b)
        return out

    def arcs(items, with_mat):
        out = []
        for it in items:
            b = ref.boundaries.ManualArcBoundary()
            a = f32_round(it["arcs"])
            for i, name in enumerate(("x_center", "y_center", "angle_start", "angle_end", "radius")):
                b[name] = torch.from_numpy(np.ascontiguousarray(a[:, i]))
            if with_mat:
                ref.engine.annotation_helper(b, "mat_in", it["mat_in"], "x_center", dtype=tf.int64)
                ref.engine.annotation_helper(b, "mat_out", it["mat_out"], "x_center", dtype=tf.int64)
            out.append(b)
        return out

    system = ref.engine.OpticalSystem2D(**EPS)
    if scene["optical_segments"]:
        system.optical_segments = segs(scene["optical_segments"], True)
    if scene["optical_arcs"]:
        system.optical_arcs = arcs(scene["optical_arcs"], True)
    if scene["stop_segments"]:
        system.stop_segments = segs(scene["stop_segments"], False)
    if scene["stop_arcs"]:
        system.stop_arcs = arcs(scene["stop_arcs"], False)
    if scene["target_segments"]:
        system.target_segments = segs(scene["target_segments"], False)
    if scene["target_arcs"]:
        system.target_arcs = arcs(scene["target_arcs"], False)
    system.sources = [_manual_source(ref, scene, 2, ray_slice)]
    system.materials = _materials(ref, scene["materials"])
    system.update()
    kw = dict(compile_dead_rays=True, compile_stopped_rays=True)
    kw.update(engine_kwargs)
    eng = ref.engine.OpticalEngine(2, [ref.operation.StandardReaction()],
                                   simple_ray_inheritance={"wavelength", "slot"}, **kw)
    eng.optical_system = system
    eng.validate_system()
    return eng, system




# ------------------------------------------------------------------------------- timing (bench.py)

def timed_reference_step(ref, scene, n_rays, backward):
    """One fwd(+bwd) step of the reference's own code on the first n_rays rays of a 3D scene:
    `system.update(); engine.ray_trace(K)` and, for optimisation configs, the GradientTape block of
    dev/simple_3d_optimize.py:128-140.  Returns (seconds, tests, live counts per iteration)."""
    import time
    tf = ref.tf
    eng, system, params = reference_engine_3d(ref, scene, ray_slice=slice(0, n_rays),
                                              compile_dead_rays=False, compile_stopped_rays=False)
    K = scene["max_iterations"]
    n_tri = int(system._merged["xp"].shape[0])
    t0 = time.perf_counter()
    with tf.GradientTape() as tape:
        eng.optical_system.update()
        eng.clear_ray_history()
        eng.ray_trace(K)
        if backward:
            output = tf.stack([eng.finished_rays["y_end"], eng.finished_rays["z_end"]], axis=1)
            goal = eng.finished_rays["rank"] * scene["goal_scale"]
            error = tf.math.squared_difference(output, goal)
            tape.gradient(error, params)
    dt = time.perf_counter() - t0
    live = [n_rays] + [int(a["x_start"].shape[0]) for a in eng._active_rays]
    live = [c for c in live[:K] if c > 0]
    return dt, float(sum(live)) * n_tri, live
