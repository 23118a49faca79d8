"""Rays born on the device: the counter-based sampling kernels of
csrc/sources.cu against their NumPy restatement (oracle/sampling.py), shard invariance (rank r of N
reproduces slice r of the single-rank stream bit for bit), the sources / distributions on the GPU
against the arrays the reference's own classes produced (tests/golden/ref_sources.npz), and a
sharded AngularSource tracing exactly the rays the unsharded one traces."""

import math

import numpy as np
import pytest
import torch

import helpers
import test_reference_golden as trg
import test_sources_distributions as tsd
from oracle import sampling
from tensorflowraytrace_b200 import boundaries, distributions as D, kernels, sources
from tensorflowraytrace_b200.parallel import shard_range

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda")


def test_field_device_is_the_gpu():
    assert boundaries.default_device().type == "cuda"


def test_uniform_sampler_equals_the_numpy_philox_bit_for_bit():
    got = kernels.sample_uniform(77, 3, 5, 1, 1000, 4096, -2.0, 3.0, DEV).cpu().numpy()
    want = sampling.uniform(77, 3, 5, 1, 1000, 4096, -2.0, 3.0)
    assert np.array_equal(got, want)
    # beyond 2^32 samples the high counter word takes over
    big = (1 << 33) + 12345
    got = kernels.sample_uniform(1, 2, 3, 0, big, 257, 0.0, 1.0, DEV).cpu().numpy()
    assert np.array_equal(got, sampling.uniform(1, 2, 3, 0, big, 257, 0.0, 1.0))
    assert 0.0 <= got.min() and got.max() <= 1.0


@pytest.mark.parametrize("mode", ["random_uniform", "random_lambertian", "static_uniform",
                                  "static_lambertian"])
def test_sphere_sampler_matches_numpy_and_is_shard_invariant(mode):
    n, a = 100000, math.pi / 7
    pts, rk = kernels.sample_sphere(9, 4, 2, mode, 0, n, n, a, 1.5, DEV)
    wp, wr = sampling.sphere(9, 4, 2, kernels.SPHERE_MODES[mode], 0, n, n, a, 1.5)
    assert np.abs(pts.cpu().numpy() - wp).max() < 2e-7 * 1.5
    assert np.abs(rk.cpu().numpy() - wr).max() < 1e-6
    norm = torch.linalg.norm(pts, dim=1)
    assert float((norm - 1.5).abs().max()) < 1e-6
    assert float(rk[:, 0].max()) <= a + 1e-6 and float(pts[:, 0].min()) >= 1.5 * math.cos(a) - 1e-6
    # rank r of 3 generates exactly slice r
    for r in range(3):
        lo, hi = shard_range(n, r, 3)
        p2, r2 = kernels.sample_sphere(9, 4, 2, mode, lo, hi - lo, n, a, 1.5, DEV)
        assert torch.equal(p2, pts[lo:hi]) and torch.equal(r2, rk[lo:hi])
    if mode.startswith("random"):   # a new epoch is a new draw; the distribution is as designed
        p3, _ = kernels.sample_sphere(9, 4, 3, mode, 0, n, n, a, 1.5, DEV)
        assert not torch.equal(p3, pts)
        c = (pts[:, 0] / 1.5).double()
        ca = math.cos(a)
        mean = float(c.mean())
        want_mean = (1 + ca) / 2 if mode == "random_uniform" else (2.0 / 3.0) * (1 - ca ** 3) / (1 - ca ** 2)
        assert abs(mean - want_mean) < 2e-3


def test_square_sampler_matches_numpy_and_is_shard_invariant():
    n = 65537
    pts, rk = kernels.sample_square(5, 1, 1, 0, n, 0.2, 0.5, DEV)
    wp, wr = sampling.square(5, 1, 1, 0, n, 0.2, 0.5)
    assert np.array_equal(pts.cpu().numpy(), wp) and np.array_equal(rk.cpu().numpy(), wr)
    assert float(pts[:, 0].abs().max()) <= 0.2 and float(pts[:, 1].abs().max()) <= 0.5
    assert abs(float(pts[:, 1].mean())) < 5e-3 and abs(float(pts[:, 1].std()) - 0.5 / 3 ** 0.5) < 5e-3
    lo, hi = shard_range(n, 1, 2)
    p2, r2 = kernels.sample_square(5, 1, 1, lo, hi - lo, 0.2, 0.5, DEV)
    assert torch.equal(p2, pts[lo:hi]) and torch.equal(r2, rk[lo:hi])


# ---- the CPU-side tests of sources / distributions, now on the GPU
@pytest.mark.parametrize("fn", [
    tsd.test_config1_source_is_the_reference_point_source,
    tsd.test_lambertian_angles_and_limits,
    tsd.test_beam_aperture_square_circle_formulas,
    tsd.test_spheres_match_the_reference_formulas,
    tsd.test_dense_domain_order_is_tf_meshgrid_xy,
    tsd.test_undense_sources_pair_elementwise_and_check_sizes,
    tsd.test_3d_rotation_and_start_on_center,
    tsd.test_aperature_source,
    tsd.test_precompiled_source_sampling_perturbation_and_samples,
], ids=lambda f: f.__name__)
def test_sources_and_distributions_on_the_gpu(fn):
    fn()


@pytest.mark.parametrize("name", sorted(trg.SOURCE_CASES))
def test_sources_match_the_reference_classes_on_the_gpu(name):
    """tests/golden/ref_sources.npz: arrays produced by the reference's own PointSource /
    AngularSource / AperatureSource and distribution classes."""
    trg.test_sources_and_distributions_match_reference(name)
    obj = trg.mrg.build_case(D, sources, trg.SOURCE_CASES[name])
    some = obj.points if name.startswith("dist_") and getattr(obj, "points", None) is not None \
        else (obj.angles if name.startswith("dist_") else obj["x_start"])
    assert some.is_cuda


def _angular_source(n, shard=None):
    D.seed_all(2024)
    base = D.RandomUniformSquare(0.2, n, 0.2, 1)
    ang = D.RandomUniformSphere(math.pi / 16, n)
    src = sources.AngularSource(3, (-4.0, 0.0, 0.0), (1.0, 0.0, 0.0), ang, base, [575.0],
                                dense=False, extra_fields={"rank": ("base_point", base, "ranks")})
    if shard is not None:
        assert src.set_ray_shard(*shard)
        src.update()
    return src


def test_sharded_angular_source_is_the_slice_of_the_unsharded_one():
    """The bench's source (undense AngularSource, random square x random sphere): two ranks
    generate, on their own GPUs, exactly the two halves of the rays one rank generates -- same
    update number, bit for bit, extra fields included."""
    n = 50001
    whole = _angular_source(n)
    whole.update()                                  # epoch 2 of both distributions
    fields = ("x_start", "y_start", "z_start", "x_end", "y_end", "z_end", "wavelength", "rank")
    assert whole["x_start"].shape[0] == n and whole["x_start"].is_cuda
    for r in range(2):
        part = _angular_source(n, shard=(r, 2))     # constructor = epoch 1, update = epoch 2
        lo, hi = shard_range(n, r, 2)
        for f in fields:
            assert torch.equal(part[f], whole[f][lo:hi]), (r, f)
    # and a fresh update draws fresh rays
    before = whole["y_end"].clone()
    whole.update()
    assert not torch.equal(before, whole["y_end"])
    d = torch.stack([whole[a + "_end"] - whole[a + "_start"] for a in "xyz"], 1)
    assert float((torch.linalg.norm(d, dim=1) - 1.0).abs().max()) < 1e-6


def test_system_uses_the_native_shard_of_a_single_source():
    from tensorflowraytrace_b200 import engine as eng
    n = 4096
    full = eng.OpticalSystem3D(**helpers.EPS)
    full.sources = [_angular_source(n)]
    full.update()
    sh = eng.OpticalSystem3D(ray_shard=(1, 2), **helpers.EPS)
    sh.sources = [_angular_source(n)]
    assert sh._sources_native_shard
    sh.update()
    lo, hi = shard_range(n, 1, 2)
    # same number of updates on both (constructor + one): same epoch
    assert sh.sources["x_start"].shape[0] == hi - lo
    assert torch.equal(sh.sources["y_start"], full.sources["y_start"][lo:hi])
    assert torch.equal(sh.sources["rank"], full.sources["rank"][lo:hi])


@pytest.mark.parametrize("kind", ["angular_2d_base", "angular_backward", "point"])
def test_fused_directed_source_equals_the_general_path(kind, monkeypatch):
    """3D PointSource / AngularSource on the device write their rays with one kernel
    (tfrt_rays_from_directions) straight into the (7,N) block the trace reads; the same source
    built through the general torch path (rotation, offset, ends as separate fp32 operations)
    must agree to fp32 rounding, and the system must hand the block on without a copy."""
    from tensorflowraytrace_b200 import engine as eng
    n = 20000

    def build():
        D.seed_all(11)
        ang = D.RandomLambertianSphere(math.pi / 5, n)
        if kind == "point":
            return sources.PointSource(3, (0.5, -1.0, 2.0), (0.3, -0.2, 0.9), ang, [450.0, 650.0],
                                       dense=True)
        base = D.RandomUniformSquare(0.4, n, 0.3, 1)
        return sources.AngularSource(3, (-4.0, 0.25, 1.0), (1.0, 0.5, -0.25), ang, base, [575.0],
                                     dense=False, start_on_base=(kind != "angular_backward"),
                                     ray_length=2.5)

    fused = build()
    assert fused._ray_block is not None and fused._ray_block.shape[0] == 7
    monkeypatch.setattr(sources._DirectedSource, "_publish_fused", lambda *a, **k: False)
    plain = build()
    assert plain._ray_block is None
    for f in ("x_start", "y_start", "z_start", "x_end", "y_end", "z_end", "wavelength"):
        a, b = fused[f], plain[f]
        assert a.shape == b.shape and float((a - b).abs().max()) < 2e-6, f
    monkeypatch.undo()
    system = eng.OpticalSystem3D(**helpers.EPS)
    system.sources = [fused]
    system.update()
    rows = system.source_rows()
    assert rows.data_ptr() == fused._ray_block.data_ptr() and rows.shape == (7, fused["x_start"].shape[0])
    assert torch.equal(rows[3], fused["x_end"]) and torch.equal(rows[6], fused["wavelength"])
