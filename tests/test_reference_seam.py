"""The drop-in boundary, exercised literally (SURVEY.md 8(b) seam b1): the REFERENCE'S
OWN engine -- unmodified `tfrt` from baseline/_ref (pip-installed by `__graft_entry__.build()`, it
travels with the repo; never /root/reference), running on the stand-in tensorflow of oracle/tf_shim --
with `OpticalSystem3D._intersection` replaced by this package's CUDA seam.  Everything else (projection,
classification, compaction, StandardReaction, Snell, materials, vertices -> fields, the GradientTape) stays
reference code.  The patched engine must reproduce the unpatched reference's fixtures
(tests/golden/ref_*.npz): same history sets, same loss, same parameter gradients.

Skipped when baseline/_ref is absent."""

import os
import sys

import numpy as np
import pytest
import torch

import helpers

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_reference_golden as mrg  # noqa: E402

pytestmark = pytest.mark.gpu


def _reference():
    from oracle import reference_runner as rr
    if rr.find_reference(installed_only=True) is None:
        pytest.skip("baseline/_ref (pip-installed reference) is not present")
    return rr, rr.import_reference(installed_only=True)


def _cuda_seam():
    """Replacement body of the reference's `_intersection` staticmethod: float64 CPU tensors in,
    this package's kernel on the GPU (fp32), float64 CPU tensors out -- attached to the autograd
    graph, so the reference's tape still reaches the boundary parameters."""
    from tensorflowraytrace_b200 import engine as ours
    calls = {"n": 0}

    def seam(rx1, ry1, rz1, rx2, ry2, rz2, xp, yp, zp, x1, y1, z1, x2, y2, z2,
             intersect_epsilion, size_epsilion, ray_start_epsilion):
        calls["n"] += 1
        dev = torch.device("cuda:0")
        args = [t.to(device=dev, dtype=torch.float32) for t in
                (rx1, ry1, rz1, rx2, ry2, rz2, xp, yp, zp, x1, y1, z1, x2, y2, z2)]
        out = ours.OpticalSystem3D._intersection(*args, float(intersect_epsilion), float(size_epsilion),
                                                 float(ray_start_epsilion))
        x, y, z, valid, ray_u, trig_u, trig_v, gather_ray, gather_trig = out
        f64 = [t.to(device="cpu", dtype=torch.float64) for t in (x, y, z)]
        uvw = [t.to(device="cpu", dtype=torch.float64) for t in (ray_u, trig_u, trig_v)]
        return (*f64, valid.cpu(), *uvw, gather_ray.cpu().long(), gather_trig.cpu().long())

    return seam, calls


@pytest.mark.parametrize("name", ["ref_lens3d", "ref_mirror_stop3d", "ref_guide3d"])
def test_reference_engine_with_the_cuda_seam_reproduces_the_reference(name):
    rr, ref = _reference()
    tf = ref.tf
    want = dict(np.load(os.path.join(HERE, "golden", f"{name}.npz")))
    scene = mrg.golden_scenes()[name]
    seam, calls = _cuda_seam()
    original = ref.engine.OpticalSystem3D.__dict__["_intersection"]
    ref.engine.OpticalSystem3D._intersection = staticmethod(seam)
    try:
        eng, system, params = rr.reference_engine_3d(ref, scene)
        with tf.GradientTape() as tape:
            eng.optical_system.update()
            eng.clear_ray_history()
            eng.ray_trace(scene["max_iterations"])
            grads = None
            if "goal_scale" in scene:
                output = tf.stack([eng.finished_rays["y_end"], eng.finished_rays["z_end"]], axis=1)
                error = tf.math.squared_difference(output, eng.finished_rays["rank"] * scene["goal_scale"])
                grads = tape.gradient(error, params)
    finally:
        ref.engine.OpticalSystem3D._intersection = original
    assert calls["n"] >= 2                                         # the seam really carried the trace
    checked = 0
    for set_name in ("active", "finished", "stopped", "dead"):
        if f"{set_name}_slot" not in want:
            continue
        s = getattr(eng, set_name + "_rays")
        assert np.array_equal(s["slot"].numpy().astype(np.int64), want[f"{set_name}_slot"]), set_name
        for f in helpers.RAY_FIELDS:
            assert helpers.rel_err(s[f].numpy(), want[f"{set_name}_{f}"],
                                   scale=_coordinate_scale(want, set_name, helpers.RAY_FIELDS)) < 1e-5, (set_name, f)
        checked += 1
    assert checked >= 2
    if grads is not None:
        assert abs(float(error.sum().detach()) - float(want["loss"])) < 1e-5 * float(want["loss"])
        for i, g in enumerate(grads):
            assert helpers.rel_err(g.numpy(), want[f"grad_{i}"]) < 1e-4, i


def _coordinate_scale(want, set_name, fields):
    """Largest |coordinate| of a ray set: what an fp32 field's absolute error is relative to."""
    return max(float(np.max(np.abs(want[f"{set_name}_{f}"]))) if want[f"{set_name}_{f}"].size else 0.0
               for f in fields if f.endswith(("_start", "_end")))


def _cuda_intersect_2d():
    """Replacement of the reference's `OpticalSystem2D.intersect(self, rays)` (seam b2,
    tfrt/engine.py:523-616): the merged segment / arc fields of the REFERENCE system go to the GPU,
    this package's 2D seam finds and evaluates the hits, the result dicts come back as float64."""
    from types import SimpleNamespace
    from tensorflowraytrace_b200 import engine as ours, kernels
    dev = torch.device("cuda:0")
    calls = {"n": 0}

    def to_dev(d):
        return {k: (v.to(device=dev, dtype=torch.float32) if v.is_floating_point() else v.to(dev))
                for k, v in d.items()} if bool(d) else {}

    def back(d):
        out = {}
        for k, v in d.items():
            v = v.cpu()
            out[k] = v.to(torch.float64) if v.is_floating_point() else \
                (v.bool() if k == "valid" else v.long())
        return out

    def intersect(self, rays):
        calls["n"] += 1
        seg, arc = to_dev(self._merged_segments), to_dev(self._merged_arcs)

        def merged_device_arrays():
            s = torch.stack([seg[k] for k in kernels.SEG_FIELDS]).contiguous() if seg else None
            a = torch.stack([arc[k] for k in kernels.ARC_FIELDS]).contiguous() if arc else None
            return s, a

        adapter = SimpleNamespace(
            _merged_segments=seg, _merged_arcs=arc, merged_device_arrays=merged_device_arrays,
            intersect_epsilion=float(self.intersect_epsilion), size_epsilion=float(self.size_epsilion),
            ray_start_epsilion=float(self.ray_start_epsilion))
        drays = {k: rays[k].to(device=dev, dtype=torch.float32) for k in kernels.RAY_FIELDS_2D}
        s_res, a_res = ours.OpticalSystem2D.intersect(adapter, drays)
        return back(s_res), back(a_res)

    return intersect, calls


@pytest.mark.parametrize("name", ["ref_wedge2d", "ref_arclens2d", "ref_segments2d"])
def test_reference_2d_engine_with_the_cuda_seam_reproduces_the_reference(name):
    rr, ref = _reference()
    want = dict(np.load(os.path.join(HERE, "golden", f"{name}.npz")))
    scene = mrg.golden_scenes()[name]
    seam, calls = _cuda_intersect_2d()
    original = ref.engine.OpticalSystem2D.intersect
    ref.engine.OpticalSystem2D.intersect = seam
    try:
        eng, system = rr.reference_engine_2d(ref, scene)
        eng.ray_trace(scene["max_iterations"])
    finally:
        ref.engine.OpticalSystem2D.intersect = original
    assert calls["n"] >= 2
    checked = 0
    for set_name in ("active", "finished", "stopped", "dead"):
        if f"{set_name}_slot" not in want:
            continue
        s = getattr(eng, set_name + "_rays")
        assert np.array_equal(s["slot"].numpy().astype(np.int64), want[f"{set_name}_slot"]), set_name
        for f in helpers.RAY_FIELDS_2D:
            assert helpers.rel_err(s[f].numpy(), want[f"{set_name}_{f}"],
                                   scale=_coordinate_scale(want, set_name, helpers.RAY_FIELDS_2D)) < 1e-5, (set_name, f)
        checked += 1
    assert checked >= 2


@pytest.mark.parametrize("name", ["ref_wedge2d", "ref_arclens2d", "ref_segments2d"])
def test_reference_2d_engine_with_per_geometry_cuda_seams(name):
    """Seam b1 in 2D: the reference's `_segment_intersection` / `_arc_intersection` staticmethods
    (tfrt/engine.py:672-866) replaced by this package's methods of the same names; the reference's
    own `intersect` (norms, `_seg_or_arc`), projection and reaction code runs around them."""
    from tensorflowraytrace_b200 import engine as ours
    rr, ref = _reference()
    want = dict(np.load(os.path.join(HERE, "golden", f"{name}.npz")))
    scene = mrg.golden_scenes()[name]
    dev = torch.device("cuda:0")
    calls = {"n": 0}

    def wrap(fn):
        def seam(*args):
            calls["n"] += 1
            tensors = [a.to(device=dev, dtype=torch.float32) for a in args[:-3]]
            out = fn(*tensors, *[float(a) for a in args[-3:]])
            x, y, valid, ray_u, geo_u, gather_ray, gather_geo = out
            f64 = lambda t: t.to(device="cpu", dtype=torch.float64)      # noqa: E731
            return (f64(x), f64(y), valid.cpu().bool(), f64(ray_u), f64(geo_u),
                    gather_ray.cpu().long(), gather_geo.cpu().long())
        return staticmethod(seam)

    cls = ref.engine.OpticalSystem2D
    saved = (cls.__dict__["_segment_intersection"], cls.__dict__["_arc_intersection"])
    cls._segment_intersection = wrap(ours.OpticalSystem2D._segment_intersection)
    cls._arc_intersection = wrap(ours.OpticalSystem2D._arc_intersection)
    try:
        eng, system = rr.reference_engine_2d(ref, scene)
        eng.ray_trace(scene["max_iterations"])
    finally:
        cls._segment_intersection, cls._arc_intersection = saved
    assert calls["n"] >= 2
    for set_name in ("active", "finished", "stopped", "dead"):
        if f"{set_name}_slot" not in want:
            continue
        s = getattr(eng, set_name + "_rays")
        assert np.array_equal(s["slot"].numpy().astype(np.int64), want[f"{set_name}_slot"]), set_name
        for f in helpers.RAY_FIELDS_2D:
            assert helpers.rel_err(s[f].numpy(), want[f"{set_name}_{f}"],
                                   scale=_coordinate_scale(want, set_name, helpers.RAY_FIELDS_2D)) < 1e-5, (set_name, f)
