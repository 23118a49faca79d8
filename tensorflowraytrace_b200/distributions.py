"""
Angular and base-point distributions: the sampling objects the reference's sources consume
(tfrt/distributions.py).  Everything is generated on the field device (the GPU when there is
one) with torch ops, so a source that is re-sampled at every `system.update()` -- as the
reference's optimisation scripts do (tfrt/optimizer.py:217) -- never crosses the PCIe bus.

Mirrored classes (same constructor arguments, `update()`, `angles` / `points` / `ranks`):
  AngularDistributionBase family, 2D angles            tfrt/distributions.py:27-556
  BeamPointBase / AperaturePointBase families, 2D      tfrt/distributions.py:746-1235
  SquareBase / CircleBase families (y-z plane in 3D)   tfrt/distributions.py:1238-1598
  SphereBase family (unit vectors about +x)            tfrt/distributions.py:1601-1850
The CDF / image / precompiled distributions (tfrt/distributions.py:2123-3195) are out of scope
(host-side pre-processing that never runs inside the trace loop).

Random distributions on a CUDA device are counter-based (csrc/sources.cu): sample i of an object
is Philox4x32-10 keyed by (seed, object stream) at counter (global index i, variable, update
number), so the draws do not depend on the launch shape, and a distribution told to hold only
one rank's shard (`set_shard(rank, world_size)`) produces exactly that slice of the single-rank
stream -- N ranks together trace the rays one rank would.  `seed_all(seed)` seeds the objects
created from then on.  (On a CPU device -- host-side use only -- draws come from a torch
generator.)
"""

import itertools
import math

import numpy as np
import torch

from .boundaries import as_float_field, default_device

PI = math.pi
GOLDEN_ANGLE_FACTOR = PI * (1.0 + 5.0 ** 0.5)
COUNTER = itertools.count(0)

_SEED = {"base": 0x5EED, "next": 0}


def seed_all(seed, rank=0):
    """Seed the distributions created from now on.  Every rank of a sharded job passes the SAME
    seed (the streams are indexed by global sample number; `rank` is accepted for backward
    compatibility and ignored by the counter-based generator)."""
    _SEED["base"] = int(seed)
    _SEED["next"] = 0
    _SEED["rank"] = int(rank)


def _new_generator(device, stream):
    g = torch.Generator(device=device)
    g.manual_seed(_SEED["base"] * 1_000_003 + _SEED.get("rank", 0) * 7919 + stream)
    return g


def _f(x, device=None):
    return as_float_field(x, device)


class _Sampler:
    """Shared plumbing: device, counter-based random stream, optional shard of the samples."""

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        inner = cls.__dict__.get("update")
        if inner is not None:
            def update(self, _inner=inner):
                self._begin_update()
                return _inner(self)
            update.__doc__ = inner.__doc__
            cls.update = update

    def _init_sampler(self, generator=None, device=None):
        self._device = torch.device(device) if device is not None else default_device()
        self._generator = generator
        self._seed = (generator.initial_seed() if generator is not None else _SEED["base"]) & 0xFFFFFFFF
        self._stream = _SEED["next"] & 0xFFFFFFFF
        _SEED["next"] += 1
        self._epoch, self._draw, self._depth = 0, 0, 0
        self._shard = None

    def _begin_update(self):
        """A new update = a new epoch of this object's stream (nested update wrappers of one
        class hierarchy count once)."""
        self._epoch += 1
        self._draw = 0

    # ---- sharding: hold only the samples of one rank
    supports_shard = False

    def set_shard(self, rank, world_size):
        """Keep only rank `rank`'s contiguous block of the samples from the next update() on.
        Returns False when this distribution cannot generate a slice on its own."""
        if not self.supports_shard:
            return False
        self._shard = None if world_size <= 1 else (int(rank), int(world_size))
        return True

    def _local_range(self, n):
        if self._shard is None:
            return 0, int(n)
        from .parallel import shard_range
        return shard_range(int(n), *self._shard)

    def _uniform(self, n, lo=0.0, hi=1.0):
        """Uniform samples of this object's next variable, for the local range of `n` samples."""
        first, stop = self._local_range(n)
        which = self._draw
        self._draw += 1
        if self._device.type == "cuda":
            from . import kernels
            return kernels.sample_uniform(self._seed, self._stream, self._epoch, which, first,
                                          stop - first, lo, hi, self._device)
        if self._generator is None:
            self._generator = _new_generator(self._device, self._stream)
        u = torch.rand(int(n), generator=self._generator, device=self._device)[first:stop]
        return lo + (hi - lo) * u

    def _linspace(self, lo, hi, n):
        return torch.linspace(float(lo), float(hi), int(n), device=self._device)


# =============================================================================================
# 2D angular distributions

class AngularDistributionBase(_Sampler):
    """Angles (rad) and ranks of a 2D fan (tfrt/distributions.py:27-163).  rank = angle divided
    by the larger limit magnitude unless the subclass defines its own."""

    def __init__(self, min_angle, max_angle, sample_count, name=None, generator=None, device=None):
        self._init_sampler(generator, device)
        self.min_angle = min_angle
        self.max_angle = max_angle
        self.sample_count = sample_count
        self._name = name or f"{self.__class__.__name__}-{next(COUNTER)}"
        self._angles = self._ranks = None
        self.update()

    def angle_limit_validation(self, lower_limit, upper_limit):
        if not (self.min_angle <= self.max_angle):
            raise ValueError(f"{self._name}: min_angle must be <= max_angle.")
        if self.min_angle < lower_limit:
            raise ValueError(f"{self._name}: min_angle must be >= {lower_limit}.")
        if self.max_angle > upper_limit:
            raise ValueError(f"{self._name}: max_angle must be <= {upper_limit}.")
        if self.sample_count <= 0:
            raise ValueError(f"{self._name}: sample_count must be > 0.")

    def _update_ranks(self, angles):
        return angles / max(abs(self.min_angle), abs(self.max_angle))

    name = property(lambda self: self._name)
    angles = property(lambda self: self._angles)
    ranks = property(lambda self: self._ranks)


class ManualAngularDistribution(AngularDistributionBase):
    """User-provided angles (and optional ranks) (tfrt/distributions.py:166-237)."""

    def __init__(self, angles, ranks=None, name=None, device=None, **kwargs):
        self._init_sampler(None, device)
        self._name = name or f"{self.__class__.__name__}-{next(COUNTER)}"
        self._angles = _f(angles, self._device)
        self._ranks = None if ranks is None else _f(ranks, self._device)

    def update(self):
        pass


class StaticUniformAngularDistribution(AngularDistributionBase):
    """linspace(min_angle, max_angle, sample_count) (tfrt/distributions.py:240-314)."""

    def update(self):
        self.angle_limit_validation(-PI, PI)
        self._angles = self._linspace(self.min_angle, self.max_angle, self.sample_count)
        self._ranks = self._update_ranks(self._angles)


class RandomUniformAngularDistribution(AngularDistributionBase):
    """Uniform random angles, re-sampled on every update (tfrt/distributions.py:317-391)."""

    def update(self):
        self.angle_limit_validation(-PI, PI)
        self._angles = self._uniform(self.sample_count, self.min_angle, self.max_angle)
        self._ranks = self._update_ranks(self._angles)


class StaticLambertianAngularDistribution(AngularDistributionBase):
    """rank = linspace(sin min, sin max), angle = asin(rank) (tfrt/distributions.py:394-470)."""

    def update(self):
        self.angle_limit_validation(-PI / 2.0, PI / 2.0)
        self._ranks = self._linspace(math.sin(self.min_angle), math.sin(self.max_angle),
                                     self.sample_count)
        self._angles = torch.asin(self._ranks)


class RandomLambertianAngularDistribution(AngularDistributionBase):
    """rank ~ U(sin min, sin max), angle = asin(rank) (tfrt/distributions.py:473-556)."""

    def update(self):
        self.angle_limit_validation(-PI / 2.0, PI / 2.0)
        self._ranks = self._uniform(self.sample_count, math.sin(self.min_angle),
                                    math.sin(self.max_angle))
        self._angles = torch.asin(self._ranks)


# =============================================================================================
# 2D base points

class BasePointDistributionBase(_Sampler):
    """Points, ranks and the post-update hook list of the reference's base point distributions
    (tfrt/distributions.py:559-626; RecursivelyUpdatable.post_update_handles, tfrt/update.py:52-61):
    every subclass' update() is followed by the functions in `post_update_handles`."""

    points = property(lambda self: self._points)
    ranks = property(lambda self: self._ranks)
    name = property(lambda self: self._name)

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        inner = cls.__dict__.get("update")
        if inner is not None:
            def update(self, _inner=inner):
                _inner(self)
                for handle in list(self.post_update_handles):
                    handle()
            update.__doc__ = inner.__doc__
            cls.update = update

    @property
    def post_update_handles(self):
        if "_post_update_handles" not in self.__dict__:
            self._post_update_handles = []
        return self._post_update_handles


class ManualBasePointDistribution(BasePointDistributionBase):
    """User-provided points (n, 2|3) and optional ranks (tfrt/distributions.py:629-743)."""

    def __init__(self, points, ranks=None, name=None, device=None, **kwargs):
        self._init_sampler(None, device)
        self._name = name or f"{self.__class__.__name__}-{next(COUNTER)}"
        self._points = _f(points, self._device)
        self._ranks = None if ranks is None else _f(ranks, self._device)

    def update(self):
        pass


class BeamPointBase(BasePointDistributionBase):
    """Points across the width of a beam through the origin, perpendicular to `central_angle`
    (tfrt/distributions.py:746-885): rank runs over [beam_start, beam_end] / max(|start|, |end|)
    and point = endpoint * rank, endpoint = sign(beam_start) * scale * (cos, sin)(angle - pi/2)."""

    def __init__(self, beam_start, beam_end, sample_count, central_angle=0.0, name=None,
                 generator=None, device=None):
        self._init_sampler(generator, device)
        self.beam_start, self.beam_end = float(beam_start), float(beam_end)
        self.sample_count = int(sample_count)
        self.central_angle = float(central_angle)
        self._name = name or f"{self.__class__.__name__}-{next(COUNTER)}"
        self.update()

    def update(self):
        if self.beam_start > self.beam_end:
            raise ValueError("BeamPointBase: beam_start must be < beam_end.")
        if self.sample_count <= 0:
            raise ValueError("BeamPointBase: sample_count must be > 0.")
        scale = max(abs(self.beam_start), abs(self.beam_end))
        start_rank, end_rank = self.beam_start / scale, self.beam_end / scale
        r = self.beam_start / abs(start_rank)
        endpoint = torch.tensor([r * math.cos(self.central_angle - PI / 2.0),
                                 r * math.sin(self.central_angle - PI / 2.0)],
                                dtype=torch.float32, device=self._device)
        self._ranks = self._make_ranks(start_rank, end_rank)
        self._points = endpoint.reshape(1, 2) * self._ranks.reshape(-1, 1)


class StaticUniformBeam(BeamPointBase):
    def _make_ranks(self, lo, hi):
        return self._linspace(lo, hi, self.sample_count)


class RandomUniformBeam(BeamPointBase):
    def _make_ranks(self, lo, hi):
        return self._uniform(self.sample_count, lo, hi)


class AperaturePointBase(BasePointDistributionBase):
    """Points on the segment start_point -> end_point, rank in [0, 1]
    (tfrt/distributions.py:1019-1122)."""

    def __init__(self, start_point, end_point, sample_count, name=None, generator=None,
                 device=None):
        self._init_sampler(generator, device)
        self.start_point = _f(start_point, self._device).reshape(1, 2)
        self.end_point = _f(end_point, self._device).reshape(1, 2)
        self.sample_count = int(sample_count)
        self._name = name or f"{self.__class__.__name__}-{next(COUNTER)}"
        self.update()

    def update(self):
        if self.sample_count <= 0:
            raise ValueError("AperaturePointBase: sample_count must be > 0.")
        self._ranks = self._make_ranks().reshape(-1, 1)
        self._points = self.start_point + self._ranks * (self.end_point - self.start_point)


class StaticUniformAperaturePoints(AperaturePointBase):
    def _make_ranks(self):
        return self._linspace(0.0, 1.0, self.sample_count)


class RandomUniformAperaturePoints(AperaturePointBase):
    def _make_ranks(self):
        return self._uniform(self.sample_count)


# =============================================================================================
# squares and circles (2D point sets; the y-z plane of a 3D source)

class SquareBase(BasePointDistributionBase):
    """x_res * y_res points in [-x_size, x_size] x [-y_size, y_size]; rank = point / max size
    (tfrt/distributions.py:1238-1358)."""

    def __init__(self, x_size, x_res, y_size=None, y_res=None, name=None, generator=None,
                 device=None):
        self._init_sampler(generator, device)
        self.x_size, self.x_res = float(x_size), int(x_res)
        self.y_size, self.y_res = float(y_size or x_size), int(y_res or x_res)
        if self.x_size <= 0 or self.y_size <= 0:
            raise ValueError("SquareDistribution: x_size and y_size must be > 0.")
        if self.x_res <= 0 or self.y_res <= 0:
            raise ValueError("SquareDistribution: x_res and y_res must be > 0.")
        self._name = name or f"{self.__class__.__name__}-{next(COUNTER)}"
        self.update()

    def update(self):
        self._make_points()
        self._ranks = self._points / max(self.x_size, self.y_size)


class StaticUniformSquare(SquareBase):
    """Regular grid, x fastest (tf.meshgrid 'xy' then flatten, tfrt/distributions.py:1361-1372)."""

    def _make_points(self):
        x = self._linspace(-self.x_size, self.x_size, self.x_res)
        y = self._linspace(-self.y_size, self.y_size, self.y_res)
        gy, gx = torch.meshgrid(y, x, indexing="ij")
        self._points = torch.stack([gx.reshape(-1), gy.reshape(-1)], dim=1)


class RandomUniformSquare(SquareBase):
    """x_res * y_res uniform points (tfrt/distributions.py:1379-1392); one fused sampling kernel
    on a CUDA device."""

    supports_shard = True

    def update(self):
        n = self.x_res * self.y_res
        if self._device.type == "cuda":
            from . import kernels
            first, stop = self._local_range(n)
            self._points, self._ranks = kernels.sample_square(
                self._seed, self._stream, self._epoch, first, stop - first, self.x_size,
                self.y_size, self._device)
            return
        x = self._uniform(n, -self.x_size, self.x_size)
        y = self._uniform(n, -self.y_size, self.y_size)
        self._points = torch.stack([x, y], dim=1)
        self._ranks = self._points / max(self.x_size, self.y_size)


class ThetaMod:
    """Restricts the azimuth to [theta_start, theta_end) by wrapping
    (tfrt/distributions.py:1396-1447)."""

    def _init_theta(self, theta_start, theta_end):
        self.theta_start, self.theta_end = float(theta_start), float(theta_end)
        if not self.theta_start < self.theta_end:
            raise ValueError("ThetaMod: theta_start must be < theta_end.")

    def _theta_mod(self, theta):
        if self.theta_start == 0.0 and self.theta_end == 2 * PI:
            return theta
        return torch.remainder(theta, self.theta_end - self.theta_start) + self.theta_start


class CircleBase(ThetaMod, BasePointDistributionBase):
    """Points in a disk; ranks are the unit-disk cartesian coordinates
    (tfrt/distributions.py:1450-1567)."""

    def __init__(self, sample_count, radius=1.0, theta_start=0, theta_end=2 * PI, name=None,
                 generator=None, device=None):
        self._init_sampler(generator, device)
        self._init_theta(theta_start, theta_end)
        self.sample_count, self.radius = int(sample_count), float(radius)
        if self.sample_count <= 0 or self.radius <= 0:
            raise ValueError("CircleBase: sample_count and radius must be > 0.")
        self._name = name or f"{self.__class__.__name__}-{next(COUNTER)}"
        self.update()

    def _finish(self):
        unit = torch.stack([self._r * torch.cos(self._theta), self._r * torch.sin(self._theta)], 1)
        self._points = self.radius * unit
        self._ranks = unit

    @property
    def polar_points(self):
        return torch.stack([self.radius * self._r, torch.remainder(self._theta, 2 * PI)], dim=1)

    @property
    def polar_ranks(self):
        return torch.stack([self._r, torch.remainder(self._theta, 2 * PI)], dim=1)


class StaticUniformCircle(CircleBase):
    """Golden-spiral (sunflower) disk (tfrt/distributions.py:1570-1583)."""

    def update(self):
        idx = torch.arange(self.sample_count, dtype=torch.float64, device=self._device) + 0.5
        self._r = torch.sqrt(idx / self.sample_count).float()
        theta = self._theta_mod(GOLDEN_ANGLE_FACTOR * idx)           # fp64: idx reaches millions
        self._theta = torch.remainder(theta, 2 * PI).float()
        self._finish()


class RandomUniformCircle(CircleBase):
    def update(self):
        self._r = torch.sqrt(self._uniform(self.sample_count))
        self._theta = self._theta_mod(2 * PI * self._uniform(self.sample_count))
        self._finish()


# =============================================================================================
# spheres: unit vectors in a cone about +x (the `angles` of a 3D source)

class SphereBase(ThetaMod, BasePointDistributionBase):
    """Vectors radius * (cos phi, sin phi cos theta, sin phi sin theta) with polar angle
    phi <= angular_size about +x; ranks = (phi, theta mod 2 pi)
    (tfrt/distributions.py:1601-1723)."""

    def __init__(self, angular_size, sample_count, radius=1, theta_start=0, theta_end=2 * PI,
                 name=None, generator=None, device=None):
        self._init_sampler(generator, device)
        self._init_theta(theta_start, theta_end)
        self.angular_size = float(angular_size)
        self.sample_count, self.radius = int(sample_count), float(radius)
        if not (0 < self.angular_size <= PI / 2.0 + 1e-12):
            raise ValueError("SphereBase: angular_size must be in (0, pi/2].")
        if self.sample_count <= 0 or self.radius <= 0:
            raise ValueError("SphereBase: sample_count and radius must be > 0.")
        self._name = name or f"{self.__class__.__name__}-{next(COUNTER)}"
        self.update()

    def _finish(self, cos_phi, theta):
        cos_phi = cos_phi.clamp(-1.0, 1.0)
        self._phi = torch.acos(cos_phi).float()
        self._theta = self._theta_mod(theta).float()
        sin_phi = torch.sin(self._phi)
        self._points = self.radius * torch.stack(
            [torch.cos(self._phi), sin_phi * torch.cos(self._theta),
             sin_phi * torch.sin(self._theta)], dim=1)
        self._ranks = torch.stack([self._phi, torch.remainder(self._theta, 2 * PI)], dim=1)

    def set_shard(self, rank, world_size):
        # a slice is generated by the fused device sampler only (default theta window)
        if not (self._device.type == "cuda" and self.theta_start == 0.0
                and self.theta_end == 2 * PI):
            return False
        return super().set_shard(rank, world_size)

    def _fused(self, mode):
        """One sampling kernel for points and ranks (csrc/sources.cu) on a CUDA device with the
        default theta window; returns False when the torch path has to be taken."""
        if not (self._device.type == "cuda" and self.theta_start == 0.0
                and self.theta_end == 2 * PI):
            return False
        from . import kernels
        first, stop = self._local_range(self.sample_count)
        self._points, self._ranks = kernels.sample_sphere(
            self._seed, self._stream, self._epoch, mode, first, stop - first, self.sample_count,
            self.angular_size, self.radius, self._device)
        self._phi, self._theta = self._ranks[:, 0], self._ranks[:, 1]
        return True

    def _golden_theta(self):
        idx = torch.arange(self.sample_count, dtype=torch.float64, device=self._device) + 0.5
        theta = GOLDEN_ANGLE_FACTOR * idx
        # reduce in fp64 before the fp32 trigonometry (indices reach millions)
        if self.theta_start == 0.0 and self.theta_end == 2 * PI:
            theta = torch.remainder(theta, 2 * PI)
        return theta


class StaticUniformSphere(SphereBase):
    """Golden spiral, uniform in solid angle (tfrt/distributions.py:1726-1748)."""

    supports_shard = True

    def update(self):
        if self._fused("static_uniform"):
            return
        cos_phi = torch.linspace(1.0, math.cos(self.angular_size), self.sample_count,
                                 dtype=torch.float64, device=self._device)
        self._finish(cos_phi, self._golden_theta())


class RandomUniformSphere(SphereBase):
    supports_shard = True

    def update(self):
        if self._fused("random_uniform"):
            return
        cos_phi = self._uniform(self.sample_count, math.cos(self.angular_size), 1.0)
        self._finish(cos_phi, GOLDEN_ANGLE_FACTOR * self._uniform(self.sample_count))


class StaticLambertianSphere(SphereBase):
    """cos^2(phi) uniform: Lambertian about +x (tfrt/distributions.py:1778-1811)."""

    supports_shard = True

    def update(self):
        if self._fused("static_lambertian"):
            return
        c2 = torch.linspace(1.0, math.cos(self.angular_size) ** 2, self.sample_count,
                            dtype=torch.float64, device=self._device)
        self._finish(torch.sqrt(c2), self._golden_theta())


class RandomLambertianSphere(SphereBase):
    supports_shard = True

    def update(self):
        if self._fused("random_lambertian"):
            return
        c2 = self._uniform(self.sample_count, math.cos(self.angular_size) ** 2, 1.0)
        self._finish(torch.sqrt(c2), GOLDEN_ANGLE_FACTOR * self._uniform(self.sample_count))


# =============================================================================================

class BasePointTransformation:
    """Scale, then rotate, then translate the points of a base point distribution after each of
    its updates (tfrt/distributions.py:2014-2120): lets relative base points serve as absolute
    positions, e.g. for an AperatureSource.  2D points are first lifted into the y-z plane
    (x = 0).  rotation: a unit quaternion (w, x, y, z); translation / scale: broadcastable to
    (n, 3).  The transformation is a post-update hook of `base`: it takes effect at the base's
    next update() and stops when this object is deleted."""

    def __init__(self, base, rotation=None, translation=None, scale=None):
        self._base = base
        self.rotation = rotation
        self.translation = translation
        self.scale = scale
        self._base.post_update_handles.append(self._apply_transformation)

    def __del__(self):
        try:
            self._base.post_update_handles.remove(self._apply_transformation)
        except (ValueError, AttributeError):
            pass

    def _apply_transformation(self):
        from .sources import rotate_by_quaternion
        pts = self._base._points
        if pts.shape[1] == 2:
            pts = torch.nn.functional.pad(pts, (1, 0))
        if self._scale is not None:
            pts = pts * self._scale.to(pts.device)
        if self._rotation is not None:
            pts = rotate_by_quaternion(self._rotation.cpu().numpy(), pts)
        if self._translation is not None:
            pts = pts + self._translation.to(pts.device)
        self._base._points = pts

    @property
    def rotation(self):
        return self._rotation

    @rotation.setter
    def rotation(self, val):
        if val is not None:
            val = torch.as_tensor(np.asarray(val, dtype=np.float64), dtype=torch.float32)
            if tuple(val.shape) != (4,):
                raise ValueError("BasePointTransformation: rotation must be a quaternion.")
        self._rotation = val

    @property
    def translation(self):
        return self._translation

    @translation.setter
    def translation(self, val):
        self._translation = None if val is None else torch.as_tensor(
            np.asarray(val, dtype=np.float64), dtype=torch.float32)

    @property
    def scale(self):
        return self._scale

    @scale.setter
    def scale(self, val):
        self._scale = None if val is None else torch.as_tensor(
            np.asarray(val, dtype=np.float64), dtype=torch.float32)
