"""
Sources: ray-set field dicts feeding the trace hot path (tfrt/sources.py), generated on the GPU.

The hot path consumes only the *field layout* of a source: 1-D, equal-length fp32 tensors
`x_start, y_start, (z_start,) x_end, y_end, (z_end,) wavelength` plus any extra per-ray fields
(tfrt/sources.py:72-359; tfrt/engine.py:2242-2257).  The classes below keep the reference's
constructors and its dense / undense "domain" semantics, but build the fields with torch ops on
the field device, so a source re-sampled at every `system.update()` stays in HBM.

  SourceBase       domains, dense expansion, extra fields        tfrt/sources.py:72-359
  ManualSource     filled by indexing                            tfrt/sources.py:363-382
  RotationBase     central_angle handling (2D angle, 3D vector or quaternion)   :386-460
  PointSource      rays from one point                           tfrt/sources.py:464-674
  AngularSource    base points x angles                          tfrt/sources.py:678-914
  AperatureSource  start points -> end points                    tfrt/sources.py:918-1095

Dense sources take the cross product of their domains in the reference's order: the domain
index arrays are `tf.meshgrid(*ranges)` (default 'xy' indexing) flattened row-major
(tfrt/sources.py:283-292), i.e. for domains (d0, d1, d2, ...) the flattened order has d1
slowest, then d0, then d2 ... fastest.  Undense sources pair their variables element-wise.
  PrecompiledSource  ray sets stored in / loaded from the reference's pickle format   :1099-1358
"""

import itertools
import math
import pickle

import numpy as np
import torch

from .boundaries import as_field, as_float_field, default_device
from .update import RecursivelyUpdatable

COUNTER = itertools.count(0)


class SourceBase(RecursivelyUpdatable):
    """Base class of sources (tfrt/sources.py:72-359).

    extra_fields : dict field -> (domain, value) or (domain, object, attribute).  The value (a
        tensor, array, scalar or zero-argument callable) is expanded like the source's own
        variables of that domain and published as a ray field (e.g. ranks).
    dense : cross every domain with every other (True) or pair them element-wise (False).
    """

    def __init__(self, extra_fields=None, standard_domains=frozenset(), dense=True,
                 always_resize=False, name=None, **kwargs):
        self.extra_fields = {} if extra_fields is None else extra_fields
        self._standard_domains = set(standard_domains) | {"whole", "wavelength"}
        self.dense = dense
        self._fields = {}
        self._domain_gathers = {}
        self._domain_sizes = {}
        self._needs_resize = True
        self.always_resize = always_resize
        self._name = name or f"{type(self).__name__}-{next(COUNTER)}"
        super().__init__(**kwargs)

    def _set_dimension(self, dimension):
        if dimension not in {2, 3}:
            raise ValueError("Source: dimension must be 2 or 3")
        self._dimension = dimension

    @staticmethod
    def validate_extra_fields(extra_fields):
        for field, args in extra_fields.items():
            if type(field) is not str:
                raise ValueError("Source extra fields: keys must be strings.")
            if len(args) not in (2, 3):
                raise ValueError(
                    "Source extra fields: every entry must be either a 2-tuple of "
                    "(domain, value), or a 3-tuple of (domain, object, attribute).")

    # ------------------------------------------------------------------ domain machinery
    @staticmethod
    def _resolve(items):
        if len(items) == 2:
            domain, value = items
        else:
            domain, obj, attr = items
            try:
                value = obj[attr]
            except (TypeError, KeyError, IndexError):
                value = getattr(obj, attr)
        if callable(value):
            value = value()
        return domain, value

    def resize(self):
        self._needs_resize = True

    def _resize(self):
        """Work out the size of every domain and, for dense sources, the gather index arrays
        (tfrt/sources.py:206-315)."""
        self._needs_resize = False
        sizes = {}
        entries = list(self._internal_vars().values()) + \
            [self._resolve(items) for items in self._extra_fields.values()]
        for domain, value in entries:
            t = torch.as_tensor(value) if not torch.is_tensor(value) else value
            sizes.setdefault(domain, []).append(t.shape[0] if t.dim() > 0 else 1)
        self._domain_sizes = {}
        for domain, found in sizes.items():
            distinct = set(found) - {1}
            if len(distinct) > 1:
                raise ValueError("Source resize: found incompatible shapes in the same domain.")
            self._domain_sizes[domain] = distinct.pop() if distinct else 1
        if self.dense:
            domains = [d for d in self._domain_sizes if d != "whole"]
            dev = default_device()
            ranges = [torch.arange(self._domain_sizes[d], device=dev) for d in domains]
            grids = torch.meshgrid(*ranges, indexing="xy") if len(ranges) > 1 else ranges
            self._domain_gathers = {d: g.reshape(-1) for d, g in zip(domains, grids)}
            self._domain_sizes["whole"] = int(np.prod([self._domain_sizes[d] for d in domains]))
        else:
            self._domain_gathers = {}
            whole = 1
            for size in self._domain_sizes.values():
                if size == 1:
                    continue
                if whole == 1:
                    whole = size
                if whole != size:
                    raise ValueError("Source resize: found incompatibly sized variables with "
                                     "an undense source.")
            self._domain_sizes["whole"] = whole

    def _expand(self, domain, value):
        t = as_field(value)
        if self.dense and domain != "whole":
            if t.dim() == 0:
                t = t.reshape(1)
            return t[self._domain_gathers[domain].to(t.device)]
        if t.dim() < 2:
            return torch.broadcast_to(t, (self._domain_sizes["whole"],)).clone() \
                if t.dim() == 0 or t.shape[0] == 1 else t
        return t

    def make_vars(self, internal_vars):
        return {name: self._expand(domain, value) for name, (domain, value) in internal_vars.items()}

    def publish_extra_fields(self):
        for field, items in self._extra_fields.items():
            domain, value = self._resolve(items)
            self[field] = self._expand(domain, value)

    def _update(self):
        if self._needs_resize or self.always_resize:
            self._resize()
        self._internal_update(self.make_vars(self._internal_vars()))
        self.publish_extra_fields()

    def _internal_vars(self):
        return {}

    def _internal_update(self, expanded_vars):
        pass

    def snapshot(self, do_update=True):
        if do_update:
            self.update()
        return {field: data.clone() for field, data in self.items()}

    # ------------------------------------------------------------------ multi-GPU ray sharding
    def set_ray_shard(self, rank, world_size):
        """Ask the source to generate only rank `rank`'s contiguous block of its rays (from the
        next update() on) instead of all of them.  Returns True when it can: its distributions
        are then sampled for that slice only -- counter-based, so the ranks of a job together
        produce exactly the rays a single rank would (csrc/sources.cu).  False: the caller
        slices the full ray set itself (OpticalSystemBase._amalgamate_sources)."""
        return False

    def _extra_domains_block_shard(self):
        """Extra fields must follow a distribution (a 3-tuple naming an object's attribute, e.g.
        its ranks) or be scalars: a fixed full-length array could not be sliced here."""
        for items in self._extra_fields.values():
            if len(items) == 2:
                v = items[1]
                if callable(v):
                    return True
                t = torch.as_tensor(v) if not torch.is_tensor(v) else v
                if t.dim() > 0 and t.shape[0] > 1:
                    return True
        return False

    def _shardable_wavelengths(self):
        w = getattr(self, "_wavelengths", None)
        return w is None or w.dim() == 0 or w.shape[0] <= 1

    # ------------------------------------------------------------------ dict protocol
    @property
    def extra_fields(self):
        return self._extra_fields

    @extra_fields.setter
    def extra_fields(self, val):
        self.validate_extra_fields(val)
        self._extra_fields = val

    @property
    def standard_domains(self):
        return self._standard_domains

    @property
    def name(self):
        return self._name

    @property
    def dimension(self):
        return self._dimension

    def __getitem__(self, key):
        return self._fields[key]

    def __setitem__(self, key, item):
        self._fields[key] = as_field(item)

    def __bool__(self):
        return bool(self._fields)

    @property
    def keys(self):
        return self._fields.keys

    @property
    def items(self):
        return self._fields.items

    def _generate_update_handles(self):
        return []


class ManualSource(SourceBase):
    """A do-nothing source that is filled by indexing (tfrt/sources.py:363-382).  Register
    functions in `update_handles` / `post_update_handles` to regenerate it on update()."""

    def __init__(self, dimension, **kwargs):
        self._set_dimension(dimension)
        super().__init__(**kwargs)

    def _update(self):
        pass

    def feed(self, fields):
        """Set several fields at once from a dict of arrays."""
        for k, v in fields.items():
            self[k] = v
        return self


# =============================================================================================
# rotations (tfrt/sources.py:386-460; the reference delegates to tfquaternion)

def quaternion_from_x_to(v):
    """Unit quaternion (w, x, y, z) of the shortest rotation taking +x onto `v`."""
    v = np.asarray(v, dtype=np.float64)
    v = v / np.linalg.norm(v)
    w = 1.0 + v[0]
    if w < 1e-12:                                   # antiparallel: half a turn about z
        return np.array([0.0, 0.0, 0.0, 1.0])
    q = np.array([w, 0.0, -v[2], v[1]])            # (1 + x.v, x cross v)
    return q / np.linalg.norm(q)


def rotate_by_quaternion(q, vectors):
    """Rotate (n,3) vectors by the unit quaternion q = (w, x, y, z): v' = v + 2 r x (r x v + w v)."""
    q = torch.as_tensor(np.asarray(q, dtype=np.float64), dtype=vectors.dtype, device=vectors.device)
    w, r = q[0], q[1:].reshape(1, 3)
    t = torch.linalg.cross(r.expand_as(vectors), vectors, dim=1) + w * vectors
    return vectors + 2.0 * torch.linalg.cross(r.expand_as(vectors), t, dim=1)


class RotationBase:
    """central_angle handling shared by PointSource and AngularSource.  2D: a scalar angle.
    3D: angle_type "vector" (the direction the source faces; the shortest rotation from +x) or
    "quaternion" ((w, x, y, z) given directly)."""

    def __init__(self, central_angle, angle_type):
        if angle_type not in ("vector", "quaternion"):
            raise ValueError("Source: angle_type must be 'vector' or 'quaternion'.")
        self._angle_type = angle_type
        self.central_angle = central_angle

    @property
    def central_angle(self):
        return self._central_angle

    @central_angle.setter
    def central_angle(self, val):
        val = np.asarray(val, dtype=np.float64)
        if self.dimension == 2:
            if val.shape != ():
                raise ValueError("Source: central_angle must be scalar.")
            self._central_angle = float(val)
        elif self._angle_type == "vector":
            if val.shape != (3,):
                raise ValueError("Source: central_angle must be size (3,).")
            self._central_angle = quaternion_from_x_to(val)
        else:
            if val.shape != (4,):
                raise ValueError("Source: central_angle must be size (4,).")
            self._central_angle = val / np.linalg.norm(val)

    def _rotate_angles(self, angles):
        if self._dimension == 2:
            return angles + self._central_angle
        return rotate_by_quaternion(self._central_angle, angles)

    def _rotate_points(self, points):
        if self._dimension == 2:
            c, s = math.cos(self._central_angle), math.sin(self._central_angle)
            return torch.stack([c * points[:, 0] - s * points[:, 1],
                                s * points[:, 0] + c * points[:, 1]], dim=1)
        if points.shape[1] == 2:     # 2D base points live in the y-z plane
            points = torch.cat([torch.zeros_like(points[:, :1]), points], dim=1)
        return rotate_by_quaternion(self._central_angle, points)


def _center(val, dimension, who):
    t = np.asarray(val, dtype=np.float64)
    if t.shape != (dimension,):
        raise ValueError(f"{who}: center must be size ({dimension},).")
    return as_float_field(t)


class _DirectedSource(SourceBase, RotationBase):
    """Common tail of PointSource / AngularSource: write the ray fields from start points and
    unit directions (or 2D angles)."""

    def _publish_fused(self, dirs, base, forward, expanded_vars):
        """3D, device-resident, nothing to differentiate: rotation, offset and both ray ends in one
        kernel (csrc/sources.cu), written straight into the (7,N) block the trace kernels read;
        the ray fields are its rows (OpticalSystemBase.source_rows hands the block on as it is).
        Returns False when the general path has to run."""
        wl = expanded_vars.get("wavelengths")
        tensors = [t for t in (dirs, base, wl) if t is not None]
        if self._dimension != 3 or not all(t.is_cuda and not t.requires_grad for t in tensors):
            return False
        n = dirs.shape[0]
        if dirs.shape != (n, 3) or (base is not None and (base.shape[0] != n or base.shape[1] not in (2, 3))) \
                or (wl is not None and wl.shape != (n,)):
            return False
        from . import kernels
        block = kernels.rays_from_directions(dirs, base, self._central_angle, self._center_host,
                                             self.ray_length, forward, wl)
        for i, name in enumerate(("x_start", "y_start", "z_start", "x_end", "y_end", "z_end")):
            self._fields[name] = block[i]
        if wl is not None:
            self._fields["wavelength"] = block[6]
        self._ray_block = block
        return True

    def _publish(self, start, angles, forward, expanded_vars):
        self._ray_block = None
        if self._dimension == 2:
            end = start + self.ray_length * torch.stack([torch.cos(angles), torch.sin(angles)], 1)
        else:
            end = start + self.ray_length * angles
        first, second = (start, end) if forward else (end, start)
        for i, ax in enumerate(("x", "y", "z")[: self._dimension]):
            self._fields[f"{ax}_start"] = first[:, i].contiguous()
            self._fields[f"{ax}_end"] = second[:, i].contiguous()
        if "wavelengths" in expanded_vars:
            self._fields["wavelength"] = expanded_vars["wavelengths"]

    def _angle_var(self):
        dist = self._angular_distribution
        return dist.angles if getattr(dist, "angles", None) is not None else dist.points


class PointSource(_DirectedSource):
    """Rays that all start at `center` (tfrt/sources.py:464-674).  In 2D the angular
    distribution provides angles (rad, relative to central_angle); in 3D unit vectors about +x
    (a sphere distribution's `points`) that are rotated onto `central_angle`.
    end = start + ray_length * direction (:598-603); start_on_center=False swaps the ends."""

    def __init__(self, dimension, center, central_angle, angular_distribution, wavelengths,
                 start_on_center=True, ray_length=1.0, angle_type="vector", **kwargs):
        self._set_dimension(dimension)
        RotationBase.__init__(self, central_angle, angle_type)
        self.center = center
        self._angular_distribution = angular_distribution
        self._wavelengths = None if wavelengths is None else as_float_field(wavelengths)
        self.start_on_center = start_on_center
        self.ray_length = ray_length
        SourceBase.__init__(self, standard_domains={"angle"}, **kwargs)

    def _internal_vars(self):
        out = {"angles": ("angle", self._angle_var())}
        if self._wavelengths is not None:
            out["wavelengths"] = ("wavelength", self._wavelengths)
        return out

    def _internal_update(self, expanded_vars):
        if self._publish_fused(expanded_vars["angles"], None, self.start_on_center, expanded_vars):
            return
        angles = self._rotate_angles(expanded_vars["angles"])
        start = torch.broadcast_to(self._center.to(angles.device).reshape(1, -1),
                                   (angles.shape[0], self._dimension))
        self._publish(start, angles, self.start_on_center, expanded_vars)

    def _generate_update_handles(self):
        return [self._angular_distribution.update]

    def set_ray_shard(self, rank, world_size):
        # rays = angles (x a single wavelength): shard the angular distribution
        if not self._shardable_wavelengths() or self._extra_domains_block_shard():
            return False
        if not getattr(self._angular_distribution, "set_shard", lambda *a: False)(rank, world_size):
            return False
        self.resize()
        return True

    @property
    def center(self):
        return self._center

    @center.setter
    def center(self, val):
        self._center = _center(val, self.dimension, "PointSource")
        self._center_host = np.asarray(val, dtype=np.float64).reshape(-1)

    @property
    def angular_distribution(self):
        return self._angular_distribution


class AngularSource(_DirectedSource):
    """Rays that start on base points (rotated by central_angle, then offset by `center`) and
    leave along the angular distribution (tfrt/sources.py:678-914).  dense=True crosses base
    points, angles and wavelengths; dense=False pairs them element-wise."""

    def __init__(self, dimension, center, central_angle, angular_distribution,
                 base_point_distribution, wavelengths, start_on_base=True, ray_length=1.0,
                 angle_type="vector", **kwargs):
        self._set_dimension(dimension)
        RotationBase.__init__(self, central_angle, angle_type)
        self.center = center
        self._angular_distribution = angular_distribution
        self._base_point_distribution = base_point_distribution
        self._wavelengths = None if wavelengths is None else as_float_field(wavelengths)
        self.start_on_base = start_on_base
        self.ray_length = ray_length
        SourceBase.__init__(self, standard_domains={"base_point", "angle"}, **kwargs)

    def _internal_vars(self):
        out = {"angles": ("angle", self._angle_var()),
               "base_points": ("base_point", self._base_point_distribution.points)}
        if self._wavelengths is not None:
            out["wavelengths"] = ("wavelength", self._wavelengths)
        return out

    def _internal_update(self, expanded_vars):
        if self._publish_fused(expanded_vars["angles"], expanded_vars["base_points"],
                               self.start_on_base, expanded_vars):
            return
        angles = self._rotate_angles(expanded_vars["angles"])
        base = self._rotate_points(expanded_vars["base_points"])
        start = self._center.to(base.device).reshape(1, -1) + base
        self._publish(start, angles, self.start_on_base, expanded_vars)

    def _generate_update_handles(self):
        return [self._angular_distribution.update, self._base_point_distribution.update]

    def set_ray_shard(self, rank, world_size):
        # paired (undense) base points and angles: shard both distributions alike
        if self.dense or not self._shardable_wavelengths() or self._extra_domains_block_shard():
            return False
        dists = (self._angular_distribution, self._base_point_distribution)
        if not all(getattr(d, "supports_shard", False) for d in dists):
            return False
        if not all(d.set_shard(rank, world_size) for d in dists):
            for d in dists:
                d.set_shard(0, 1)
            return False
        self.resize()
        return True

    @property
    def center(self):
        return self._center

    @center.setter
    def center(self, val):
        self._center = _center(val, self.dimension, "AngularSource")
        self._center_host = np.asarray(val, dtype=np.float64).reshape(-1)

    @property
    def angular_distribution(self):
        return self._angular_distribution

    @property
    def base_point_distribution(self):
        return self._base_point_distribution


class AperatureSource(SourceBase):
    """Rays from the points of one distribution to the points of another
    (tfrt/sources.py:918-1095)."""

    def __init__(self, dimension, start_point_distribution, end_point_distribution, wavelengths,
                 **kwargs):
        self._set_dimension(dimension)
        self._start_point_distribution = start_point_distribution
        self._end_point_distribution = end_point_distribution
        self._wavelengths = None if wavelengths is None else as_float_field(wavelengths)
        super().__init__(standard_domains={"start_point", "end_point"}, **kwargs)

    def _internal_vars(self):
        out = {"start_points": ("start_point", self._start_point_distribution.points),
               "end_points": ("end_point", self._end_point_distribution.points)}
        if self._wavelengths is not None:
            out["wavelengths"] = ("wavelength", self._wavelengths)
        return out

    def _internal_update(self, expanded_vars):
        start, end = expanded_vars["start_points"], expanded_vars["end_points"]
        for i, ax in enumerate(("x", "y", "z")[: self._dimension]):
            self._fields[f"{ax}_start"] = start[:, i].contiguous()
            self._fields[f"{ax}_end"] = end[:, i].contiguous()
        if "wavelengths" in expanded_vars:
            self._fields["wavelength"] = expanded_vars["wavelengths"]

    def _generate_update_handles(self):
        return [self._start_point_distribution.update, self._end_point_distribution.update]

    @property
    def start_point_distribution(self):
        return self._start_point_distribution

    @property
    def end_point_distribution(self):
        return self._end_point_distribution


# =============================================================================================

class PrecompiledSource(RecursivelyUpdatable):
    """A ray set that was generated once (a random source behind a static optic, a costly extra
    field, the output of a previous trace) and is re-used, optionally drawing a random subset and
    jittering the end points at every update (tfrt/sources.py:1099-1358).

    arg : str  -> file written by `save` here or by the reference's `PrecompiledSource.save`
                  (pickle of {"dimension", "standard_domains", "fields": {name: ndarray}} --
                  the reference's format; unpickling executes code, so load only files you trust);
          int  -> dimension of a new, empty source (fill it with `from_samples`);
          else -> an updated source object whose fields are copied.
    The stored rays live on the field device; `update()` samples `sample_count` of them with
    replacement (`do_downsample`) and adds N(0, sigma) jitter per coordinate when
    `start_perturbation` / `end_perturbation` are given, all on the device."""

    def __init__(self, arg, sample_count=100, do_downsample=True, start_perturbation=None,
                 end_perturbation=None, **kwargs):
        if isinstance(arg, str):
            with open(arg, "rb") as f:
                data = pickle.load(f)
            self._dimension = int(data["dimension"])
            self._standard_domains = set(data["standard_domains"])
            full = data["fields"]
        elif isinstance(arg, (int, np.integer)):
            self._dimension, self._standard_domains, full = int(arg), set(), {}
        else:
            self._dimension = int(arg.dimension)
            self._standard_domains = set(getattr(arg, "_standard_domains", ()))
            full = dict(arg.items())
        if self._dimension not in (2, 3):
            raise ValueError("PrecompiledSource: dimension must be 2 or 3")
        self._full_fields = {k: as_field(v) for k, v in full.items()}
        self._fields = {}
        self.start_perturbation = start_perturbation
        self.end_perturbation = end_perturbation
        self.sample_count = sample_count
        self.do_downsample = do_downsample
        super().__init__(**kwargs)

    # ---- storage
    def save(self, filename):
        """Write the CURRENT (possibly down-sampled) fields in the reference's pickle layout."""
        out = {"dimension": self._dimension, "standard_domains": set(self._standard_domains),
               "fields": {k: v.detach().cpu().numpy() for k, v in self._fields.items()}}
        with open(filename, "wb") as f:
            pickle.dump(out, f, pickle.HIGHEST_PROTOCOL)

    def from_samples(self, samples):
        """Fill the store by concatenating the fields of several field dicts (e.g. the
        `finished_rays` of previous traces)."""
        lists = {}
        for sample in samples:
            for field, data in sample.items():
                lists.setdefault(field, []).append(as_field(data))
        self._full_fields = {k: torch.cat(v, dim=0) for k, v in lists.items()}
        self.update()

    def clear(self):
        self._full_fields = {}
        self._fields = self._full_fields

    # ---- update
    def _generate_update_handles(self):
        return []

    def _update(self):
        n = self.sampling_domain_size
        if self.do_downsample and n > 0:
            dev = self._full_fields["x_start"].device
            idx = torch.randint(0, n, (self._sample_count,), device=dev)
            self._fields = {k: v.index_select(0, idx) for k, v in self._full_fields.items()}
        else:
            self._fields = dict(self._full_fields)
        if not self._fields:
            return
        dims = ("x", "y") if self._dimension == 2 else ("x", "y", "z")
        for which, sigma in (("start", self._start_perturbation), ("end", self._end_perturbation)):
            if sigma is None:
                continue
            for d, s in zip(dims, sigma.tolist()):
                key = f"{d}_{which}"
                self._fields[key] = self._fields[key] + float(s) * torch.randn_like(self._fields[key])

    # ---- field access (the data contract of a source)
    def __getitem__(self, key):
        return self._fields[key]

    def __setitem__(self, key, item):
        self._fields[key] = as_field(item)

    def __bool__(self):
        return bool(self._fields)

    @property
    def keys(self):
        return self._fields.keys

    @property
    def items(self):
        return self._fields.items

    @property
    def dimension(self):
        return self._dimension

    @property
    def standard_domains(self):
        return self._standard_domains

    @property
    def sampling_domain_size(self):
        f = self._full_fields.get("x_start")
        return 0 if f is None else int(f.shape[0])

    @property
    def sample_count(self):
        return self._sample_count

    @sample_count.setter
    def sample_count(self, val):
        if int(val) != val:
            raise TypeError("PrecompiledSource: sample_count must be of integer type.")
        if val <= 0:
            raise ValueError("PrecompiledSource: sample_count must be > 0.")
        self._sample_count = int(val)

    def _perturbation(self, val, which):
        if val is None:
            return None
        try:
            return torch.broadcast_to(torch.as_tensor(val, dtype=torch.float32),
                                      (self._dimension,)).clone()
        except RuntimeError as e:
            raise ValueError(f"PrecompiledSource: {which}_perturbation must be None, scalar, or "
                             "must have one entry per dimension.") from e

    @property
    def start_perturbation(self):
        return self._start_perturbation

    @start_perturbation.setter
    def start_perturbation(self, val):
        self._start_perturbation = self._perturbation(val, "start")

    @property
    def end_perturbation(self):
        return self._end_perturbation

    @end_perturbation.setter
    def end_perturbation(self, val):
        self._end_perturbation = self._perturbation(val, "end")
