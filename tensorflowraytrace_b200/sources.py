"""
Sources: ray-set field dicts feeding the trace hot path (tfrt/sources.py).

The hot path only consumes the *field layout* of a source: 1-D, equal-length tensors
`x_start, y_start, (z_start,) x_end, y_end, (z_end,) wavelength` plus any extra per-ray
fields that ray operations declare or `simple_ray_inheritance` names (tfrt/sources.py:72-359;
tfrt/engine.py:2242-2257).  Ray generation itself (distributions, quaternion rotations) is
outside the hot path ("next" rank 2); the generic container and the two
closed-form sources the BASELINE configs use are provided here.

  SourceBase      common container + update protocol     tfrt/sources.py:72-359
  ManualSource    filled by indexing                      tfrt/sources.py:363-382
  PointSource     rays from one point along given unit vectors   tfrt/sources.py:464-674
                  (end = start + ray_length * direction, :598-603)
  AngularSource   base points x directions, dense or undense     tfrt/sources.py:678-914
"""

import itertools

import numpy as np
import torch

from .boundaries import as_field
from .update import RecursivelyUpdatable

COUNTER = itertools.count(0)


class SourceBase(RecursivelyUpdatable):
    def __init__(self, dimension, extra_fields=None, name=None, **kwargs):
        if dimension not in {2, 3}:
            raise ValueError("Source: dimension must be 2 or 3")
        self._dimension = dimension
        self._fields = {}
        self.extra_fields = {} if extra_fields is None else extra_fields
        self._name = name or f"{type(self).__name__}-{next(COUNTER)}"
        super().__init__(**kwargs)

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

    def snapshot(self, do_update=True):
        if do_update:
            self.update()
        return dict(self._fields)

    def _publish_extra_fields(self, count):
        for field, value in self.extra_fields.items():
            value = value() if callable(value) else value
            t = as_field(value)
            if t.dim() < 1:
                t = torch.broadcast_to(t, (count,)).clone()
            self[field] = t

    def _generate_update_handles(self):
        return []


class ManualSource(SourceBase):
    """A do-nothing source that is filled by indexing (tfrt/sources.py:363-382).  Register
    functions in `update_handles` / `post_update_handles` to regenerate it on update()."""

    def __init__(self, dimension, **kwargs):
        super().__init__(dimension, **kwargs)

    def _update(self):
        pass

    def feed(self, fields):
        """Set several fields at once from a dict of arrays."""
        for k, v in fields.items():
            self[k] = v
        return self


class PointSource(SourceBase):
    """Rays leaving `center` along unit vectors `directions` (N,dim), one wavelength per ray or
    a list of wavelengths crossed with the directions when dense (tfrt/sources.py:464-674).
    end = start + ray_length * direction (tfrt/sources.py:598-603)."""

    def __init__(self, dimension, center, directions, wavelengths, ray_length=1.0, dense=True,
                 **kwargs):
        self.center = center
        self.directions = directions
        self.wavelengths = wavelengths
        self.ray_length = ray_length
        self.dense = dense
        super().__init__(dimension, **kwargs)

    def _update(self):
        d = as_field(self.directions() if callable(self.directions) else self.directions)
        d = d.reshape(-1, self.dimension)
        wl = as_field(np.atleast_1d(np.asarray(self.wavelengths, dtype=np.float64))).to(d.device)
        if self.dense:   # every direction x every wavelength, direction index fastest
            n_d, n_w = d.shape[0], wl.shape[0]
            d = d.repeat(n_w, 1)
            wl = wl.repeat_interleave(n_d)
        else:
            wl = torch.broadcast_to(wl, (d.shape[0],)).clone()
        c = as_field(np.asarray(self.center, dtype=np.float64)).to(d.device).reshape(1, -1)
        start = torch.broadcast_to(c, d.shape)
        end = start + self.ray_length * d
        names = ("x", "y", "z")[: self.dimension]
        for i, ax in enumerate(names):
            self._fields[f"{ax}_start"] = start[:, i].contiguous()
            self._fields[f"{ax}_end"] = end[:, i].contiguous()
        self._fields["wavelength"] = wl
        self._publish_extra_fields(d.shape[0])


class AngularSource(SourceBase):
    """Base points x directions (tfrt/sources.py:678-914).  dense=True crosses every base point
    with every direction (and wavelength); dense=False pairs them element-wise.  The optional
    `ranks` (N, dim-1) of the base points is published as the field `rank`, the quantity the
    reference's imaging losses use (dev/simple_3d_optimize.py:136-137)."""

    def __init__(self, dimension, base_points, directions, wavelengths, ray_length=1.0,
                 dense=False, ranks=None, **kwargs):
        self.base_points = base_points
        self.directions = directions
        self.wavelengths = wavelengths
        self.ray_length = ray_length
        self.dense = dense
        self.ranks = ranks
        super().__init__(dimension, **kwargs)

    def _update(self):
        get = lambda v: as_field(v() if callable(v) else v)  # noqa: E731
        b = get(self.base_points).reshape(-1, self.dimension)
        d = get(self.directions).reshape(-1, self.dimension)
        wl = as_field(np.atleast_1d(np.asarray(self.wavelengths, dtype=np.float64))).to(b.device)
        rank = None if self.ranks is None else get(self.ranks)
        if self.dense:
            n_b, n_d, n_w = b.shape[0], d.shape[0], wl.shape[0]
            b = b.repeat(n_d * n_w, 1)
            if rank is not None:
                rank = rank.repeat(n_d * n_w, *([1] * (rank.dim() - 1)))
            d = d.repeat_interleave(n_b, dim=0).repeat(n_w, 1)
            wl = wl.repeat_interleave(n_b * n_d)
        else:
            if b.shape[0] != d.shape[0]:
                raise ValueError(
                    "Source resize: found incompatibly sized variables with an undense source.")
            wl = torch.broadcast_to(wl, (b.shape[0],)).clone()
        end = b + self.ray_length * d
        names = ("x", "y", "z")[: self.dimension]
        for i, ax in enumerate(names):
            self._fields[f"{ax}_start"] = b[:, i].contiguous()
            self._fields[f"{ax}_end"] = end[:, i].contiguous()
        self._fields["wavelength"] = wl
        if rank is not None:
            self._fields["rank"] = rank
        self._publish_extra_fields(b.shape[0])
