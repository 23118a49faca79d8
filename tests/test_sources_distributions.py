"""Sources and distributions (tfrt/sources.py, tfrt/distributions.py) against NumPy restatements
of the reference's formulas: closed forms, dense-domain ordering (tf.meshgrid 'xy' + flatten),
rotations, extra fields.  Runs on whatever device fields live on (CPU here, the GPU on a B200)."""

import math

import numpy as np
import pytest
import torch

from tensorflowraytrace_b200 import distributions as D
from tensorflowraytrace_b200 import scenes, sources

PI = math.pi


def npy(t):
    return t.detach().cpu().numpy().astype(np.float64)


def test_config1_source_is_the_reference_point_source():
    src = sources.PointSource(2, (-2.0, 0.0), 0.0,
                              D.StaticUniformAngularDistribution(-0.2, 0.2, 1000), [scenes.YELLOW])
    want = scenes.config1()["sources"]
    for k in ("x_start", "y_start", "x_end", "y_end", "wavelength"):
        assert np.allclose(npy(src[k]), want[k], atol=2e-7), k
    ranks = D.StaticUniformAngularDistribution(-0.1, 0.2, 5).ranks
    assert np.allclose(npy(ranks), np.linspace(-0.1, 0.2, 5) / 0.2, atol=1e-7)


def test_lambertian_angles_and_limits():
    d = D.StaticLambertianAngularDistribution(-PI / 3, PI / 4, 11)
    r = np.linspace(math.sin(-PI / 3), math.sin(PI / 4), 11)
    assert np.allclose(npy(d.ranks), r, atol=1e-6) and np.allclose(npy(d.angles), np.arcsin(r), atol=1e-6)
    with pytest.raises(ValueError):
        D.StaticLambertianAngularDistribution(-2.0, 1.0, 5)
    with pytest.raises(ValueError):
        D.StaticUniformAngularDistribution(0.3, 0.1, 5)
    rnd = D.RandomUniformAngularDistribution(-0.3, 0.5, 4000)
    a = npy(rnd.angles)
    assert a.min() >= -0.3 and a.max() <= 0.5 and abs(a.mean() - 0.1) < 0.02
    before = a.copy()
    rnd.update()
    assert not np.array_equal(before, npy(rnd.angles))      # re-sampled on update


def test_beam_aperture_square_circle_formulas():
    b = D.StaticUniformBeam(-1.5, 1.5, 7)
    assert np.allclose(npy(b.points), np.stack([np.zeros(7), np.linspace(-1.5, 1.5, 7)], 1), atol=1e-6)
    assert np.allclose(npy(b.ranks), np.linspace(-1, 1, 7), atol=1e-6)
    b = D.StaticUniformBeam(-1.0, 3.0, 5, central_angle=PI / 2)   # beam along +y: points along -x..+x
    rank = np.linspace(-1 / 3, 1.0, 5)
    end = -3.0 * np.array([math.cos(0.0), math.sin(0.0)])
    assert np.allclose(npy(b.points), end[None, :] * rank[:, None], atol=1e-6)
    a = D.StaticUniformAperaturePoints((0.0, -1.0), (2.0, 1.0), 5)
    assert np.allclose(npy(a.points), np.stack([np.linspace(0, 2, 5), np.linspace(-1, 1, 5)], 1), atol=1e-6)
    s = D.StaticUniformSquare(2.0, 3, 1.0, 2)
    gx, gy = np.meshgrid(np.linspace(-2, 2, 3), np.linspace(-1, 1, 2))
    assert np.allclose(npy(s.points), np.stack([gx.reshape(-1), gy.reshape(-1)], 1), atol=1e-6)
    assert np.allclose(npy(s.ranks), npy(s.points) / 2.0)
    c = D.StaticUniformCircle(500, radius=2.0)
    idx = np.arange(500) + 0.5
    r, th = np.sqrt(idx / 500), PI * (1 + 5 ** 0.5) * idx
    assert np.allclose(npy(c.points), 2.0 * np.stack([r * np.cos(th), r * np.sin(th)], 1), atol=2e-5)
    rc = D.RandomUniformCircle(3000, radius=0.5, theta_start=0.0, theta_end=PI)
    p = npy(rc.points)
    assert np.all(np.hypot(p[:, 0], p[:, 1]) <= 0.5 + 1e-6) and np.all(p[:, 1] >= -1e-6)


def test_spheres_match_the_reference_formulas():
    s = D.StaticUniformSphere(PI / 8, 65536)
    assert np.allclose(npy(s.points), scenes.static_uniform_sphere(PI / 8, 65536), atol=3e-6)
    lam = D.StaticLambertianSphere(PI / 3, 1000)
    phi = np.arccos(np.sqrt(np.linspace(1.0, math.cos(PI / 3) ** 2, 1000)))
    assert np.allclose(npy(lam.ranks[:, 0]), phi, atol=1e-5)
    for cls in (D.RandomUniformSphere, D.RandomLambertianSphere):
        v = npy(cls(PI / 6, 5000).points)
        assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-5)
        assert np.all(v[:, 0] >= math.cos(PI / 6) - 1e-5)
    with pytest.raises(ValueError):
        D.StaticUniformSphere(2.0, 10)


def test_dense_domain_order_is_tf_meshgrid_xy():
    ang = D.ManualAngularDistribution([0.0, 0.1, 0.2])
    src = sources.PointSource(2, (0.0, 0.0), 0.0, ang, [500.0, 600.0])
    # domains (angle, wavelength): angle fastest
    assert np.allclose(npy(src["wavelength"]), [500, 500, 500, 600, 600, 600])
    assert np.allclose(npy(src["x_end"]), np.tile(np.cos([0.0, 0.1, 0.2]), 2), atol=1e-6)
    base = D.ManualBasePointDistribution([[0.0, 1.0], [0.0, 2.0]], ranks=[[0.1], [0.2]])
    asrc = sources.AngularSource(2, (0.0, 0.0), 0.0, ang, base, [500.0, 600.0],
                                 extra_fields={"rank": ("base_point", base, "ranks")})
    ga, gb, gw = np.meshgrid(np.arange(3), np.arange(2), np.arange(2))      # 'xy' indexing
    ga, gb, gw = ga.reshape(-1), gb.reshape(-1), gw.reshape(-1)
    assert np.allclose(npy(asrc["y_start"]), np.array([1.0, 2.0])[gb])
    assert np.allclose(npy(asrc["wavelength"]), np.array([500.0, 600.0])[gw])
    assert np.allclose(npy(asrc["x_end"]), np.cos(np.array([0.0, 0.1, 0.2])[ga]), atol=1e-6)
    assert np.allclose(npy(asrc["rank"])[:, 0], np.array([0.1, 0.2])[gb])
    assert asrc["x_start"].shape[0] == 12


def test_undense_sources_pair_elementwise_and_check_sizes():
    n = 64
    sq = D.RandomUniformSquare(0.2, 8)
    sph = D.RandomUniformSphere(PI / 16, n)
    src = sources.AngularSource(3, (-4.0, 0.0, 0.0), (1.0, 0.0, 0.0), sph, sq, [scenes.YELLOW],
                                dense=False, extra_fields={"rank": ("base_point", sq, "ranks")})
    assert src["x_start"].shape == (n,) and src["rank"].shape == (n, 2)
    assert np.allclose(npy(src["x_start"]), -4.0)
    assert np.allclose(npy(src["y_start"]), npy(sq.points)[:, 0], atol=1e-7)
    d = np.stack([npy(src[a + "_end"]) - npy(src[a + "_start"]) for a in "xyz"], 1)
    assert np.allclose(d, npy(sph.points), atol=1e-6)
    first = npy(src["y_start"]).copy()
    src.update()                                          # random distributions are re-sampled
    assert not np.array_equal(first, npy(src["y_start"]))
    with pytest.raises(ValueError):
        sources.AngularSource(3, (0.0, 0.0, 0.0), (1.0, 0.0, 0.0), D.RandomUniformSphere(0.1, 10),
                              D.RandomUniformSquare(0.2, 3), [500.0], dense=False)


def test_3d_rotation_and_start_on_center():
    sph = D.StaticUniformSphere(0.3, 200)
    for target in ((0.0, 1.0, 0.0), (0.0, 0.0, -2.0), (1.0, 1.0, 1.0), (-1.0, 0.0, 0.0)):
        src = sources.PointSource(3, (1.0, 2.0, 3.0), target, sph, [500.0], ray_length=2.0)
        d = np.stack([npy(src[a + "_end"]) - npy(src[a + "_start"]) for a in "xyz"], 1) / 2.0
        t = np.asarray(target) / np.linalg.norm(target)
        assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
        # rigid rotation taking +x to the target: polar angles about the new axis are preserved
        assert np.allclose(d @ t, npy(sph.points)[:, 0], atol=1e-5)
    back = sources.PointSource(3, (0.0, 0.0, 0.0), (1.0, 0.0, 0.0), sph, [500.0], start_on_center=False)
    assert np.allclose(npy(back["x_end"]), 0.0) and np.allclose(npy(back["x_start"]), npy(sph.points)[:, 0], atol=1e-6)
    quat = sources.PointSource(3, (0.0, 0.0, 0.0), (math.cos(PI / 4), 0.0, 0.0, math.sin(PI / 4)), sph,
                               [500.0], angle_type="quaternion")      # quarter turn about z: x -> y
    assert np.allclose(npy(quat["y_end"]), npy(sph.points)[:, 0], atol=1e-6)
    with pytest.raises(ValueError):
        sources.PointSource(3, (0.0, 0.0), (1.0, 0.0, 0.0), sph, [500.0])


def test_aperature_source():
    a = D.StaticUniformAperaturePoints((0.0, -1.0), (0.0, 1.0), 4)
    b = D.StaticUniformAperaturePoints((5.0, -2.0), (5.0, 2.0), 3)
    src = sources.AperatureSource(2, a, b, [450.0])
    assert src["x_start"].shape[0] == 12
    gs, ge = np.meshgrid(np.arange(4), np.arange(3))
    assert np.allclose(npy(src["y_start"]), np.linspace(-1, 1, 4)[gs.reshape(-1)], atol=1e-6)
    assert np.allclose(npy(src["y_end"]), np.linspace(-2, 2, 3)[ge.reshape(-1)], atol=1e-6)


def test_parametric_segment_boundaries_and_constraints():
    """2D parametric curves: points = zero + p (one - zero), consecutive points make segments,
    flip_norm reverses them; a multi-boundary keeps its thickness constraint on update."""
    from tensorflowraytrace_b200 import boundaries as B
    zero = D.StaticUniformAperaturePoints((0.0, -1.0), (0.0, 1.0), 5)
    one = D.StaticUniformAperaturePoints((1.0, -1.0), (1.0, 1.0), 5)
    init = [0.0, 0.1, 0.2, 0.1, 0.0]
    seg = B.ParametricSegmentBoundary(zero, one, initial_parameters=init)
    assert seg["x_start"].shape == (4,) and seg.dimension == 2
    assert np.allclose(npy(seg["x_start"]), init[:-1], atol=1e-7)
    assert np.allclose(npy(seg["x_end"]), init[1:], atol=1e-7)
    assert np.allclose(npy(seg["y_start"]), np.linspace(-1, 1, 5)[:-1], atol=1e-7)
    flipped = B.ParametricSegmentBoundary(zero, one, flip_norm=True, initial_parameters=init)
    assert np.allclose(npy(flipped["x_start"]), init[1:], atol=1e-7)
    (g,) = torch.autograd.grad(seg["x_end"].sum() + seg["x_start"].sum(), [seg.parameters])
    assert np.allclose(npy(g), [1, 2, 2, 2, 1])
    multi = B.ParametricMultiSegmentBoundary(
        zero, one, [B.ThicknessConstraint(0.0, "min"), B.ThicknessConstraint(0.3, "min")],
        [True, False], initial_parameters=[init, 0.0],
        material_list=[{"mat_in": np.int64(1), "mat_out": np.int64(0)}] * 2)
    p0, p1 = multi.parameters
    assert abs(float((0.0 - p0).max().detach())) < 1e-7           # first surface touches zero
    assert abs(float((p0 - p1).max().detach()) + 0.3) < 1e-6      # min gap 0.3 to the second
    assert multi["x_start"].shape == (8,) and len(multi.surfaces) == 2
    with torch.no_grad():
        p1.add_(0.5)
    multi.update()
    assert abs(float((p0 - p1).max().detach()) + 0.3) < 1e-6
    with pytest.raises(ValueError):
        B.ParametricMultiSegmentBoundary(zero, one, [B.NoConstraint()], [True, False])


# ------------------------------------------------------------------------------------------------
# PrecompiledSource (tfrt/sources.py:1099-1358)

def test_precompiled_source_reads_the_reference_pickle_format(tmp_path):
    """tests/golden/ref_precompiled_source.pkl was written by the reference's own
    PrecompiledSource.save (make_reference_golden.precompiled_file) from a 3D PointSource."""
    import os
    import pickle
    here = os.path.dirname(os.path.abspath(__file__))
    path = os.path.join(here, "golden", "ref_precompiled_source.pkl")
    raw = pickle.load(open(path, "rb"))
    pre = sources.PrecompiledSource(path, do_downsample=False)
    assert pre.dimension == 3 and pre.standard_domains == raw["standard_domains"]
    assert pre.sampling_domain_size == 24 and set(pre.keys()) == set(raw["fields"])
    for k, v in raw["fields"].items():
        assert np.allclose(npy(pre[k]), v, atol=1e-6), k
    # the same rays as this package's PointSource with the same arguments
    src = sources.PointSource(3, (-1.0, 0.0, 0.0), (1.0, 0.0, 0.0), D.StaticUniformSphere(PI / 8, 12),
                              [575.0, 450.0])
    for k in raw["fields"]:
        assert np.allclose(npy(src[k]), raw["fields"][k], atol=2e-6), k
    # round trip through our save(), in the reference's layout
    out = str(tmp_path / "again.pkl")
    pre.save(out)
    again = pickle.load(open(out, "rb"))
    assert set(again) == {"dimension", "standard_domains", "fields"} and again["dimension"] == 3
    assert all(isinstance(v, np.ndarray) for v in again["fields"].values())
    assert np.allclose(again["fields"]["y_end"], raw["fields"]["y_end"], atol=1e-6)


def test_precompiled_source_sampling_perturbation_and_samples():
    src = sources.PointSource(3, (0.0, 0.0, 0.0), (1.0, 0.0, 0.0), D.StaticUniformSphere(PI / 6, 50),
                              [500.0])
    pre = sources.PrecompiledSource(src, sample_count=400)
    assert pre["x_start"].shape[0] == 400 and pre.sampling_domain_size == 50
    full = np.stack([npy(src[k]) for k in ("x_end", "y_end", "z_end")], 1)
    got = np.stack([npy(pre[k]) for k in ("x_end", "y_end", "z_end")], 1)
    # every drawn ray is one of the stored rays (sampling with replacement keeps rows together)
    d = np.abs(got[:, None, :] - full[None, :, :]).max(axis=2).min(axis=1)
    assert d.max() < 1e-6
    before = got.copy()
    pre.update()
    assert not np.array_equal(before, np.stack([npy(pre[k]) for k in ("x_end", "y_end", "z_end")], 1))
    pre.do_downsample = False
    pre.end_perturbation = (0.0, 0.1, 0.0)
    pre.update()
    assert pre["x_end"].shape[0] == 50
    assert np.allclose(npy(pre["x_end"]), full[:, 0], atol=1e-6) and np.allclose(npy(pre["z_end"]), full[:, 2], atol=1e-6)
    dy = npy(pre["y_end"]) - full[:, 1]
    assert 0.03 < dy.std() < 0.25
    with pytest.raises(ValueError):
        pre.start_perturbation = (1.0, 2.0)
    with pytest.raises(ValueError):
        pre.sample_count = 0
    empty = sources.PrecompiledSource(2, do_downsample=False)
    assert not bool(empty) and empty.sampling_domain_size == 0
    a = {"x_start": np.zeros(3), "y_start": np.zeros(3), "x_end": np.ones(3), "y_end": np.arange(3.0)}
    b = {"x_start": np.zeros(2), "y_start": np.ones(2), "x_end": np.ones(2), "y_end": np.arange(2.0)}
    empty.from_samples([a, b])
    assert empty["y_end"].shape[0] == 5 and np.allclose(npy(empty["y_start"]), [0, 0, 0, 1, 1])
    empty.clear()
    assert empty.sampling_domain_size == 0
