"""
OpticalSystem3D / OpticalEngine: the tracing engine of tfrt/engine.py on B200.

`OpticalEngine.ray_trace` is the hot path.  When the operation list is made of
`StandardReaction` (+ optionally `OldestAncestor`) the whole iteration loop -- intersection of
every live ray with every triangle, nearest hit, classification, material look-up, Snell /
reflection / TIR, order-preserving compaction, history -- runs fused on the GPU through
libtfrt_b200.so (one autograd node, `autograd.TraceFunction`).  Any other operation list runs
the same loop one projection at a time (`single_pass`): the O(rays x triangles) search is still
the CUDA kernel (the `_intersection` seam, tfrt/engine.py:1102-1166), the O(rays) bookkeeping is
torch on the GPU.  There is no CPU path.

Mirrored reference interfaces (same names, arguments, result field names and ordering):
  amalgamate, ReadOnlySet           tfrt/engine.py:27-76
  OpticalSystemBase / 3D            tfrt/engine.py:146-250, 870-1166
  OpticalEngine                     tfrt/engine.py:1216-1403, 1988-2330
"""

from abc import abstractmethod

import torch

from . import geometry, kernels
from . import operation as op
import weakref

from .autograd import TraceFunction, TracePlan, WorkspaceLease, run_trace
from .materials import descriptor_of
from .update import RecursivelyUpdatable

OPTICAL = 0
STOP = 1
TARGET = 2

SEGMENT_GEO_SIG = {"x_start", "y_start", "x_end", "y_end"}
ARC_GEO_SIG = {"x_center", "y_center", "angle_start", "angle_end", "radius"}
SOURCE_3D_SIG = {"x_start", "y_start", "z_start", "x_end", "y_end", "z_end"}
TRIANGLE_GEO_SIG = {"xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2", "norm"}

TRI_FIELDS = kernels.TRI_FIELDS
RAY_FIELDS_3D = kernels.RAY_FIELDS


class ReadOnlySet:
    """Read-only view of a ray or boundary field dict (tfrt/engine.py:27-46)."""

    def __init__(self, fields):
        self._fields = fields

    def __getitem__(self, key):
        try:
            return self._fields[key]
        except KeyError as e:
            raise KeyError(f"key {key} not in the signature of this set.") from e

    def __bool__(self):
        return bool(self._fields)

    @property
    def keys(self):
        return self._fields.keys


def _field_dict(item):
    return item._fields if hasattr(item, "_fields") else item


def amalgamate(stuff, signature=None):
    """Concatenate field dicts along axis 0 over the common (or given) signature
    (tfrt/engine.py:50-76).  Empty items are skipped."""
    processed = [_field_dict(item) for item in stuff if bool(item)]
    if not processed:
        return {}
    if not signature:
        signature = None
        for item in processed:
            keys = set(item.keys())
            signature = keys if signature is None else (signature & keys)
    if len(processed) == 1:
        return {field: processed[0][field] for field in signature}
    return {field: torch.cat([item[field] for item in processed], 0) for field in signature}


def annotation_helper(parent, field, value, valid_shape_field, dtype=None):
    """Keep `parent[field]` annotated on every update (tfrt/engine.py:103-142): `value` is
    broadcast to the shape of `parent[valid_shape_field]` (or, if callable, called with that
    shape and `dtype`) by a post-update handle, e.g. to apply materials to a boundary."""
    cache = {}

    def f():
        ref = parent[valid_shape_field]
        if callable(value):
            parent[field] = value(ref.shape, dtype)
            return
        key = (tuple(ref.shape), ref.device)
        if cache.get("key") != key:      # constant annotation: rebuilt only when the shape changes
            t = torch.as_tensor(value, device=ref.device)
            if dtype is not None:
                t = t.to(dtype)
            elif t.is_floating_point():
                t = t.to(torch.float32)
            cache["key"], cache["tensor"] = key, torch.broadcast_to(t, ref.shape).clone()
        parent[field] = cache["tensor"]
    parent.post_update_handles.append(f)
    f()


# =========================================================================================

class OpticalSystemBase(RecursivelyUpdatable):
    """Holds sources, boundaries and materials (tfrt/engine.py:146-250).

    intersect_epsilion / size_epsilion / ray_start_epsilion have the reference's meaning
    (tfrt/engine.py:160-179) and are applied exactly, in fp64, to the hit the fp32 search
    proposes: the search itself runs with scale-aware allowances on top of them
    (`search_margin_ulps` units in the last place of the triangle coordinates on the edge
    inflation, 4x that on the ray-start exclusion), and a proposal the exact rule rejects is
    searched again in fp64.  So any values work, including the reference's
    1e-10 defaults and scenes in millimetres or far from the origin; the defaults here (1e-10
    on |det|, 1e-5 edge inflation, 1e-4 ray-start exclusion) keep the exact re-searches rare.

    ray_shard : None or (rank, world_size).  When set, `update()` keeps only this rank's
    contiguous block of the amalgamated source rays (multi-GPU ray sharding, parallel.py).

    accel : None (default: every ray is tested against every boundary, the brute-force search
    the reference performs) or "grid" (3D): a uniform grid over the triangles is rebuilt on the
    device at every update() and the search walks it -- same hit indices, same gradients, far
    fewer tests (csrc/grid3d.cu); `grid_max_cells` bounds its size.
    """

    def __init__(self, manual_update_management=False, intersect_epsilion=1e-10,
                 size_epsilion=1e-5, ray_start_epsilion=1e-4, ray_shard=None,
                 search_margin_ulps=kernels.DEFAULT_MARGIN_ULPS, accel=None,
                 grid_max_cells=kernels.DEFAULT_GRID_MAX_CELLS, **kwargs):
        if accel not in (None, "grid"):
            raise ValueError("OpticalSystem: accel must be None or 'grid'.")
        self.accel = accel
        self.grid_max_cells = int(grid_max_cells)
        self._sources = []
        self._read_only_sources = None
        self.source_handles = []
        self._amalgamated_sources = {}
        self.manual_update_management = manual_update_management
        self.materials = []
        self.intersect_epsilion = intersect_epsilion
        self.size_epsilion = size_epsilion
        self.ray_start_epsilion = ray_start_epsilion
        self.search_margin_ulps = float(search_margin_ulps)
        self._sources_native_shard = False
        self._ray_shard = None
        self.ray_shard = ray_shard
        self.projection_results = {}
        super().__init__(**kwargs)

    @abstractmethod
    def refresh_update_handles(self):
        raise NotImplementedError

    @property
    @abstractmethod
    def dimension(self):
        raise NotImplementedError

    @property
    def sources(self):
        if not bool(self._read_only_sources):
            self._read_only_sources = ReadOnlySet(self._amalgamated_sources)
        return self._read_only_sources

    @sources.setter
    def sources(self, new):
        self.source_handles = []
        for each in new:
            assert each.dimension == self.dimension
            self.source_handles.append(each.update)
        self._sources = new
        self._apply_ray_shard()
        self.refresh_update_handles()

    @property
    def ray_shard(self):
        return self._ray_shard

    @ray_shard.setter
    def ray_shard(self, val):
        self._ray_shard = None if val is None else (int(val[0]), int(val[1]))
        self._apply_ray_shard()

    def _apply_ray_shard(self):
        """A single source that can generate its own shard does (rays born on this rank's GPU
        only); otherwise update() slices the amalgamated rays."""
        self._sources_native_shard = False
        srcs = getattr(self, "_sources", None) or []
        if self._ray_shard is not None and len(srcs) == 1 and hasattr(srcs[0], "set_ray_shard"):
            self._sources_native_shard = bool(srcs[0].set_ray_shard(*self._ray_shard))

    @property
    def materials(self):
        return self._materials

    @materials.setter
    def materials(self, val):
        assert type(val) is list
        self._materials = val
        self._material_table = None

    def _amalgamate_sources(self):
        if bool(self._sources):
            merged = amalgamate(self._sources)
            if self.ray_shard is not None and merged and not self._sources_native_shard:
                from .parallel import shard_range
                n = next(iter(merged.values())).shape[0]
                lo, hi = shard_range(n, *self.ray_shard)
                merged = {k: v[lo:hi] for k, v in merged.items()}
            self._amalgamated_sources = merged
        self._source_rows = None

    def source_rows(self):
        """The amalgamated source rays as one fp32 block [x_start .. z_end, wavelength] -- (7,N)
        in 3D, (5,N) in 2D, the layout the trace kernels read -- built once per update()."""
        rows = getattr(self, "_source_rows", None)
        if rows is None and len(self._sources) == 1 and self.dimension == 3:
            # a directed source writes its rays straight into this layout (sources._publish_fused)
            blk = getattr(self._sources[0], "_ray_block", None)
            f = self._amalgamated_sources
            if blk is not None and f.get("wavelength") is not None and \
                    f["x_start"].data_ptr() == blk.data_ptr() and \
                    f["x_start"].shape[0] == blk.shape[1] and \
                    f["wavelength"].data_ptr() == blk[6].data_ptr():
                rows = self._source_rows = blk
        if rows is None:
            f = self._amalgamated_sources
            n = f["x_start"].shape[0]
            cols = [f[k] for k in (RAY_FIELDS_3D if self.dimension == 3 else kernels.RAY_FIELDS_2D)]
            cols.append(f["wavelength"] if "wavelength" in f
                        else torch.zeros(n, dtype=torch.float32, device=cols[0].device))
            rows = torch.stack([c.to(torch.float32) for c in cols]).contiguous()
            self._source_rows = rows
        return rows

    def material_table(self, device):
        """Device table of kernel material descriptors for `self.materials` (cached)."""
        if self._material_table is None or self._material_table.device != device:
            self._material_table = kernels.material_table(
                [descriptor_of(m) for m in self._materials], device)
        return self._material_table

    def materials_have_descriptors(self):
        """Every entry of `materials` is a closed-form tensorflowraytrace_b200.materials.Material
        (evaluated per ray inside the reaction kernel)."""
        from .materials import Material
        return all(isinstance(m["n"] if isinstance(m, dict) else m, Material)
                   for m in self._materials)

    def material_index_table(self, wavelength):
        """Refractive indices of ARBITRARY material callables for the fused reaction kernel
        (tfrt/operation.py:265-267 evaluates any `mat["n"](wavelength)`): every callable is
        evaluated once, on the device, over the distinct wavelengths of the source set.  Returns
        (n_table (n_materials, n_lambda) f64, lambda_index (N,) int32)."""
        distinct, inverse = torch.unique(wavelength, return_inverse=True)
        rows = []
        for m in self._materials:
            fn = m["n"] if isinstance(m, dict) else m
            n = torch.as_tensor(fn(distinct), device=distinct.device)
            rows.append(torch.broadcast_to(n.to(torch.float64), distinct.shape))
        return torch.stack(rows).contiguous(), inverse.to(torch.int32)

    def _category_tensor(self, cat, like):
        """Constant int64 category labels, cached per (category, count, device)."""
        cache = self.__dict__.setdefault("_category_cache", {})
        key = (cat, like.shape[0], like.device)
        if key not in cache:
            cache[key] = torch.full((like.shape[0],), cat, dtype=torch.int64, device=like.device)
        return cache[key]

    @abstractmethod
    def intersect(self, rays):
        raise NotImplementedError


class OpticalSystem3D(OpticalSystemBase):
    """Triangle-boundary optical system (tfrt/engine.py:870-1166)."""

    def __init__(self, **kwargs):
        self._optical = []
        self.optical_handles = []
        self._amalgamated_optical = {}
        self._stop = []
        self.stop_handles = []
        self._amalgamated_stop = {}
        self._target = []
        self.target_handles = []
        self._amalgamated_target = {}
        self._materials = []
        self._merged = {}
        self._norm_is_geometric = False
        self._optical_count = self._stop_count = self._target_count = 0
        self._dev_cache = None
        self.clear_read_only()
        super().__init__(**kwargs)

    def clear_read_only(self):
        self._read_only_sources = None
        self._read_only_optical = None
        self._read_only_stop = None
        self._read_only_target = None

    def refresh_update_handles(self):
        if not self.manual_update_management:
            self.update_handles = self._generate_update_handles()

    def _generate_update_handles(self):
        return (self.source_handles + self.optical_handles + self.stop_handles
                + self.target_handles)

    @property
    def dimension(self):
        return 3

    def _update(self):
        self._amalgamate_sources()
        if not self._merge_blocks():
            if bool(self._optical):
                self._amalgamated_optical = amalgamate(self._optical)
            if bool(self._stop):
                self._amalgamated_stop = amalgamate(self._stop)
            if bool(self._target):
                self._amalgamated_target = amalgamate(self._target)
            self._merge_boundaries()
        self.clear_read_only()

    def _cached_cat(self, key, parts):
        """torch.cat of per-boundary fields that rarely change (materials): reused while the
        very same tensors are handed in."""
        cache = self.__dict__.setdefault("_cat_cache", {})
        ids = tuple((id(t), t._version) for t in parts)
        hit = cache.get(key)
        if hit is not None and hit[0] == ids:
            return hit[2]
        out = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
        cache[key] = (ids, list(parts), out)     # the parts are kept alive: ids stay unique
        return out

    def _merge_blocks(self):
        """Fast form of amalgamate + _merge_boundaries (tfrt/engine.py:906-918, 971-1018) when
        every boundary keeps its geometry as one (9,F) block (all triangle boundaries of this
        package do): ONE concatenation builds the merged geometry in the layout the kernels
        read, and the per-field views the reference's API exposes are rows / slices of it.
        Returns False when some boundary is not block-based (the field-by-field path runs)."""
        groups = ((OPTICAL, "optical", self._optical), (STOP, "stop", self._stop),
                  (TARGET, "target", self._target))
        blocks, norms, counts = [], [], []
        for _, _, items in groups:
            n = 0
            for b in items:
                if not bool(b):
                    continue
                blk = getattr(b, "_tri_block", None)
                f = getattr(b, "_fields", None)
                if blk is None or f is None or f.get("xp") is None or \
                        f["xp"].data_ptr() != blk.data_ptr() or f["xp"].shape[0] != blk.shape[1] \
                        or blk.dtype != torch.float32 or f.get("norm") is None:
                    return False
                blocks.append(blk)
                norms.append(f["norm"])
                n += blk.shape[1]
            counts.append(n)
        if not blocks:
            return False
        tri = blocks[0] if len(blocks) == 1 else torch.cat(blocks, 1)
        norm = norms[0] if len(norms) == 1 else torch.cat(norms, 0)
        self._optical_count, self._stop_count, self._target_count = counts
        lo = 0
        for (cat, name, items), n in zip(groups, counts):
            d = {}
            if n:
                whole = n == tri.shape[1]
                d.update(zip(TRI_FIELDS, (tri if whole else tri[:, lo:lo + n]).unbind(0)))
                d["norm"] = norm if whole else norm[lo:lo + n]
                live = [b for b in items if bool(b)]
                extra = set(live[0].keys()) - set(TRI_FIELDS) - {"norm"}
                for b in live[1:]:
                    extra &= set(b.keys())
                for k in extra:
                    d[k] = self._cached_cat((name, k), [b[k] for b in live])
                d["catagory"] = self._category_tensor(cat, d["xp"])
            setattr(self, "_amalgamated_" + name, d)
            lo += n
        merged = dict(zip(TRI_FIELDS, tri.unbind(0)))
        merged["norm"] = norm
        merged["catagory"] = self._cached_cat(
            ("merged", "catagory"),
            [d["catagory"] for d in (self._amalgamated_optical, self._amalgamated_stop,
                                     self._amalgamated_target) if bool(d)])
        self._merged = merged
        self._norm_is_geometric = all(
            getattr(b, "has_geometric_norm", lambda: False)()
            for b in list(self._optical) + list(self._stop) + list(self._target) if bool(b))
        self._merged_tri_block = tri
        self._dev_cache = None
        return True

    def _boundary_property(name, store, handles, cache):  # noqa: N805  (class-body helper)
        def getter(self):
            if not bool(getattr(self, cache)):
                setattr(self, cache, ReadOnlySet(getattr(self, "_amalgamated_" + store)))
            return getattr(self, cache)

        def setter(self, new):
            hs = []
            for each in new:
                assert each.dimension == self.dimension
                hs.append(each.update)
            setattr(self, handles, hs)
            setattr(self, "_" + store, new)
            self.refresh_update_handles()

        return property(getter, setter)

    optical = _boundary_property("optical", "optical", "optical_handles", "_read_only_optical")
    stops = _boundary_property("stops", "stop", "stop_handles", "_read_only_stop")
    targets = _boundary_property("targets", "target", "target_handles", "_read_only_target")
    del _boundary_property

    def _merge_boundaries(self):
        """Label and merge optical -> stop -> target (tfrt/engine.py:971-1018)."""
        counts, parts = [], []
        for cat, d in ((OPTICAL, self._amalgamated_optical), (STOP, self._amalgamated_stop),
                       (TARGET, self._amalgamated_target)):
            if bool(d):
                d["catagory"] = self._category_tensor(cat, d["xp"])
                counts.append(d["xp"].shape[0])
            else:
                counts.append(0)
            parts.append(d)
        self._optical_count, self._stop_count, self._target_count = counts
        self._merged = amalgamate(parts, TRIANGLE_GEO_SIG | {"catagory"})
        # every `norm` is the geometric face normal of its triangle (not user-overridden): the
        # fused trace then evaluates normals in fp64 from the corners instead of the fp32 field
        self._norm_is_geometric = all(
            getattr(b, "has_geometric_norm", lambda: False)()
            for b in list(self._optical) + list(self._stop) + list(self._target) if bool(b))
        self._merged_tri_block = None
        self._dev_cache = None

    # ------------------------------------------------------------------ kernel-side views
    def merged_device_arrays(self):
        """(tri (9,T) f32, norm (T,3) f32, category (T,) int32, packed) of the merged set.  tri
        and norm stay attached to the autograd graph of the boundary parameters; the packed
        plane records for the intersection kernel are rebuilt once per update()."""
        if self._dev_cache is None:
            m = self._merged
            blk = getattr(self, "_merged_tri_block", None)
            if blk is not None and m["xp"].data_ptr() == blk.data_ptr():
                tri = blk if blk.is_contiguous() else blk.contiguous()
            else:
                tri = torch.stack([m[k] for k in TRI_FIELDS]).to(torch.float32).contiguous()
            norm = m["norm"].to(torch.float32).contiguous()
            ckey = (self._optical_count, self._stop_count, self._target_count, tri.device)
            c32 = self.__dict__.setdefault("_category32_cache", {})
            if ckey not in c32:      # labels are a function of the three counts only
                c32[ckey] = m["catagory"].to(torch.int32).contiguous()
            category = c32[ckey]
            packed = kernels.pack_triangles(tri.detach(), self.size_epsilion,
                                            self.search_margin_ulps)
            self._grid = None
            if self.accel == "grid":
                self._grid = kernels.TriangleGrid(tri.detach(), packed, self.size_epsilion,
                                                  self.search_margin_ulps, self.grid_max_cells)
            self._dev_cache = (tri, norm, category, packed)
        return self._dev_cache

    def grad_triangle_count(self):
        """How many leading triangles of the merged set can receive cotangents: all of them when
        a target boundary is differentiable, optical + stop when a stop is, else the optical
        ones (skipping the others removes the atomics of every finished ray into the few target
        triangles).  Decided from the boundaries' own fields: the merged block requires grad as
        a whole as soon as one boundary does."""
        def needs_grad(items):
            return any(bool(b) and any(torch.is_tensor(v) and v.requires_grad
                                       for v in (b[k] for k in b.keys()))
                       for b in items)
        if needs_grad(self._target):
            return 0
        if needs_grad(self._stop):
            return self._optical_count + self._stop_count
        return max(self._optical_count, 1)

    @property
    def grid(self):
        """The acceleration grid of the current merged triangles (None in brute-force mode)."""
        self.merged_device_arrays()
        return getattr(self, "_grid", None)

    def intersect(self, rays):
        """Nearest valid intersection of each ray with the merged triangles
        (tfrt/engine.py:1020-1078): dict with x, y, z, valid, ray_u, trig_u, trig_v,
        gather_ray, gather_trig, norm."""
        result = {}
        if bool(self._merged):
            (result["x"], result["y"], result["z"], result["valid"], result["ray_u"],
             result["trig_u"], result["trig_v"], result["gather_ray"],
             result["gather_trig"]) = self._intersection(
                *[rays[k] for k in RAY_FIELDS_3D], *[self._merged[k] for k in TRI_FIELDS],
                self.intersect_epsilion, self.size_epsilion, self.ray_start_epsilion,
                _packed=self.merged_device_arrays()[3], _margin_ulps=self.search_margin_ulps,
                _grid=self.grid)
            result["norm"] = self._merged["norm"][result["gather_trig"]]
        return result

    @staticmethod
    def _intersection(rx1, ry1, rz1, rx2, ry2, rz2, xp, yp, zp, x1, y1, z1, x2, y2, z2,
                      intersect_epsilion, size_epsilion, ray_start_epsilion, _packed=None,
                      _margin_ulps=kernels.DEFAULT_MARGIN_ULPS, _grid=None):
        """The intersection seam (tfrt/engine.py:1102-1166): N rays x M triangles -> per-ray
        nearest hit.  The search runs in the CUDA kernel; the gathered values are then
        re-evaluated 1-to-1 in torch so they carry the reference's gradient (through ray_u
        only).  Outputs for rays without a valid hit are unspecified, as in the reference."""
        rays = torch.stack([rx1, ry1, rz1, rx2, ry2, rz2]).detach()
        tri = torch.stack([xp, yp, zp, x1, y1, z1, x2, y2, z2]).detach()
        if _packed is None:
            _packed = kernels.pack_triangles(tri, size_epsilion, _margin_ulps)
        raw = kernels.intersect3d(rays, tri, _packed, intersect_epsilion, size_epsilion,
                                  ray_start_epsilion, _margin_ulps, grid=_grid)
        g = raw["gather_trig"]
        # the 1-to-1 evaluation runs in fp64 on the fp32 fields, like the kernels' per-ray
        # arithmetic, and is stored back as fp32
        f64, f32 = torch.float64, torch.float32
        hit, _, ray_u, trig_u, trig_v = geometry.ray_triangle_hit(
            torch.stack([rx1, ry1, rz1], 1).to(f64), torch.stack([rx2, ry2, rz2], 1).to(f64),
            torch.stack([xp[g], yp[g], zp[g]], 1).to(f64),
            torch.stack([x1[g], y1[g], z1[g]], 1).to(f64),
            torch.stack([x2[g], y2[g], z2[g]], 1).to(f64), intersect_epsilion)
        hit = hit.to(f32)
        return (hit[:, 0], hit[:, 1], hit[:, 2], raw["valid"], ray_u.to(f32), trig_u.to(f32),
                trig_v.to(f32), raw["gather_ray"], g)


class OpticalSystem2D(OpticalSystemBase):
    """Segment + arc optical system (tfrt/engine.py:253-866): six boundary lists
    (optical / stop / target x segments / arcs), merged optical -> stop -> target per geometry
    (tfrt/engine.py:425-521)."""

    _KINDS = ("optical_segments", "optical_arcs", "stop_segments", "stop_arcs",
              "target_segments", "target_arcs")

    def __init__(self, **kwargs):
        for kind in self._KINDS:
            setattr(self, "_" + kind, [])
            setattr(self, kind + "_handles", [])
            setattr(self, "_amalgamated_" + kind, {})
        self._materials = []
        self._merged_segments = {}
        self._merged_arcs = {}
        self._optical_seg_count = self._stop_seg_count = self._target_seg_count = 0
        self._optical_arc_count = self._stop_arc_count = self._target_arc_count = 0
        self.clear_read_only()
        super().__init__(**kwargs)

    def clear_read_only(self):
        self._read_only_sources = None
        for kind in self._KINDS:
            setattr(self, "_read_only_" + kind, None)

    def refresh_update_handles(self):
        if not self.manual_update_management:
            self.update_handles = self._generate_update_handles()

    def _generate_update_handles(self):
        out = list(self.source_handles)
        for kind in self._KINDS:
            out += getattr(self, kind + "_handles")
        return out

    @property
    def dimension(self):
        return 2

    def _update(self):
        self._amalgamate_sources()
        for kind in self._KINDS:
            items = getattr(self, "_" + kind)
            if bool(items):
                setattr(self, "_amalgamated_" + kind, amalgamate(items))
        self._merge_boundaries()
        self.clear_read_only()

    def _kind_property(kind):  # noqa: N805  (class-body helper)
        def getter(self):
            if not bool(getattr(self, "_read_only_" + kind)):
                setattr(self, "_read_only_" + kind,
                        ReadOnlySet(getattr(self, "_amalgamated_" + kind)))
            return getattr(self, "_read_only_" + kind)

        def setter(self, new):
            hs = []
            for each in new:
                assert each.dimension == self.dimension
                hs.append(each.update)
            setattr(self, kind + "_handles", hs)
            setattr(self, "_" + kind, new)
            self.refresh_update_handles()

        return property(getter, setter)

    optical_segments = _kind_property("optical_segments")
    optical_arcs = _kind_property("optical_arcs")
    stop_segments = _kind_property("stop_segments")
    stop_arcs = _kind_property("stop_arcs")
    target_segments = _kind_property("target_segments")
    target_arcs = _kind_property("target_arcs")
    del _kind_property

    def _merge_boundaries(self):
        """Label and merge optical -> stop -> target, per geometry (tfrt/engine.py:425-521)."""
        def merge(suffix, key, sig):
            counts, parts = [], []
            for cat, name in ((OPTICAL, "optical"), (STOP, "stop"), (TARGET, "target")):
                d = getattr(self, f"_amalgamated_{name}_{suffix}")
                if bool(d):
                    d["catagory"] = self._category_tensor(cat, d[key])
                    counts.append(d[key].shape[0])
                else:
                    counts.append(0)
                parts.append(d)
            # static boundaries hand in the very same tensors at every update: the merged set
            # (and the (4,S) / (5,A) device blocks built from it) is then reused as it is
            full = sig | {"catagory"}
            ids = tuple((id(d[f]), d[f]._version) for d in parts if bool(d) for f in full)
            cache = self.__dict__.setdefault("_merge_cache", {})
            hit = cache.get(suffix)
            if hit is not None and hit[0] == ids:
                return counts, hit[2]
            merged = amalgamate(parts, full)
            cache[suffix] = (ids, [dict(d) for d in parts], merged)   # inputs kept alive
            return counts, merged

        (self._optical_seg_count, self._stop_seg_count, self._target_seg_count), \
            self._merged_segments = merge("segments", "x_start", SEGMENT_GEO_SIG)
        (self._optical_arc_count, self._stop_arc_count, self._target_arc_count), \
            self._merged_arcs = merge("arcs", "x_center", ARC_GEO_SIG)

    def merged_device_arrays(self):
        """(seg (4,S) f32 or None, arc (5,A) f32 or None), attached to the autograd graph of
        the boundary parameters."""
        def block(merged, fields, name):
            if not bool(merged):
                return None
            cols = [merged[k] for k in fields]
            ids = tuple((id(c), c._version) for c in cols)
            cache = self.__dict__.setdefault("_block_cache", {})
            hit = cache.get(name)
            if hit is not None and hit[0] == ids:
                return hit[2]
            out = torch.stack(cols).to(torch.float32).contiguous()
            if not out.requires_grad:      # a block that carries a graph is rebuilt every time
                cache[name] = (ids, cols, out)
            return out

        seg = block(self._merged_segments, kernels.SEG_FIELDS, "seg")
        arc = block(self._merged_arcs, kernels.ARC_FIELDS, "arc")
        return seg, arc

    def intersect(self, rays):
        """Nearest valid intersection of each ray with the merged segments and arcs
        (tfrt/engine.py:523-616).  Returns (segment_result, arc_result), each a dict with x, y,
        valid, ray_u, segment_u | arc_u, gather_ray, gather_segment | gather_arc, norm ({} when
        the system has no boundary of that geometry).  The search runs in the CUDA kernel; the
        gathered values are re-evaluated 1-to-1 in torch so they carry the reference's
        gradient."""
        seg, arc = self.merged_device_arrays()
        if seg is None and arc is None:
            return {}, {}
        ray_rows = torch.stack([rays[k] for k in kernels.RAY_FIELDS_2D])
        raw_seg, raw_arc = kernels.intersect2d(
            ray_rows.detach(), None if seg is None else seg.detach(),
            None if arc is None else arc.detach(), self.intersect_epsilion, self.size_epsilion,
            self.ray_start_epsilion)
        xs, ys, xe, ye = (rays[k] for k in kernels.RAY_FIELDS_2D)
        if raw_seg:
            g = raw_seg["gather_segment"]
            m = self._merged_segments
            x, y, u, v = geometry.line_hit_2d(xs, ys, xe, ye, m["x_start"][g], m["y_start"][g],
                                              m["x_end"][g], m["y_end"][g],
                                              self.intersect_epsilion)
            norm = torch.atan2(m["y_end"] - m["y_start"], m["x_end"] - m["x_start"]) \
                + geometry.PI / 2.0
            raw_seg.update(x=x, y=y, ray_u=u, segment_u=v, norm=norm[g])
        if raw_arc:
            g = raw_arc["gather_arc"]
            m = self._merged_arcs
            x, y, u, v = geometry.circle_hit_2d(xs, ys, xe, ye, m["x_center"][g],
                                                m["y_center"][g], m["radius"][g],
                                                raw_arc["ray_u"].detach())
            arc_norm = torch.where(m["radius"][g] < 0, v + geometry.PI, v)
            arc_norm = torch.remainder(arc_norm + geometry.PI, 2 * geometry.PI) - geometry.PI
            raw_arc.update(x=x, y=y, ray_u=u, arc_u=v, norm=arc_norm)
        return raw_seg, raw_arc


    # ---- the reference's per-geometry seams (tfrt/engine.py:626-866), same names and signatures
    @staticmethod
    def _seg_or_arc(seg_u, arc_u, seg_valid, arc_valid):
        """Where a ray hits both a segment and an arc keep the nearer (tfrt/engine.py:626-657)."""
        has_both = seg_valid & arc_valid
        seg_less = seg_u < arc_u
        return (torch.where(has_both, seg_less, seg_valid),
                torch.where(has_both, ~seg_less, arc_valid))

    @staticmethod
    def _get_arc_norm(radius, arc_u, gather_ray):
        """Arc normal angle: the hit's polar angle, turned by pi for negative radii, wrapped to
        (-pi, pi] (tfrt/engine.py:666-670; `gather_ray` indexes the arcs, as in the reference)."""
        radius = radius[gather_ray]
        arc_norm = torch.where(radius < 0, arc_u + geometry.PI, arc_u)
        return torch.remainder(arc_norm + geometry.PI, 2 * geometry.PI) - geometry.PI

    @staticmethod
    def _one_geometry(rays4, seg=None, arc=None, eps=(1e-10, 1e-5, 1e-4)):
        from types import SimpleNamespace
        names = kernels.SEG_FIELDS if seg is not None else kernels.ARC_FIELDS
        merged = dict(zip(names, seg if seg is not None else arc))
        stacked = torch.stack(list(seg if seg is not None else arc)).to(torch.float32).contiguous()
        adapter = SimpleNamespace(
            _merged_segments=merged if seg is not None else {},
            _merged_arcs=merged if arc is not None else {},
            merged_device_arrays=lambda: (stacked, None) if seg is not None else (None, stacked),
            intersect_epsilion=float(eps[0]), size_epsilion=float(eps[1]),
            ray_start_epsilion=float(eps[2]))
        rays = dict(zip(kernels.RAY_FIELDS_2D, rays4))
        seg_res, arc_res = OpticalSystem2D.intersect(adapter, rays)
        return seg_res if seg is not None else arc_res

    @staticmethod
    def _segment_intersection(rx1, ry1, rx2, ry2, sx1, sy1, sx2, sy2, intersect_epsilion,
                              size_epsilion, ray_start_epsilion):
        """N rays x M segments -> per-ray nearest valid hit (tfrt/engine.py:672-749): x, y, valid,
        ray_u, seg_u, gather_ray, gather_segment; values of rays without a valid hit are
        unspecified, as in the reference."""
        r = OpticalSystem2D._one_geometry(
            (rx1, ry1, rx2, ry2), seg=(sx1, sy1, sx2, sy2),
            eps=(intersect_epsilion, size_epsilion, ray_start_epsilion))
        return (r["x"], r["y"], r["valid"], r["ray_u"], r["segment_u"], r["gather_ray"],
                r["gather_segment"])

    @staticmethod
    def _arc_intersection(rx1, ry1, rx2, ry2, xc, yc, a1, a2, r, intersect_epsilion,
                          size_epsilion, ray_start_epsilion):
        """N rays x M arcs -> per-ray nearest valid hit (tfrt/engine.py:751-866): x, y, valid,
        ray_u, arc_u (the hit's polar angle), gather_ray, gather_arc."""
        res = OpticalSystem2D._one_geometry(
            (rx1, ry1, rx2, ry2), arc=(xc, yc, a1, a2, r),
            eps=(intersect_epsilion, size_epsilion, ray_start_epsilion))
        return (res["x"], res["y"], res["valid"], res["ray_u"], res["arc_u"], res["gather_ray"],
                res["gather_arc"])


# =========================================================================================

class OpticalEngine:
    """The ray tracer (tfrt/engine.py:1170-2330).  Constructor arguments as in the reference
    (tfrt/engine.py:1216-1230)."""

    def __init__(self, dimension, operations, optical_system=None,
                 compile_technical_intersections=False, compile_stopped_rays=False,
                 compile_dead_rays=False, compile_finished_rays=True, compile_active_rays=True,
                 dead_ray_length=None, compile_geometry_specific_result=False,
                 new_ray_length=1.0, simple_ray_inheritance={"wavelength"}):
        if dimension == 2:
            self.process_projection = self.process_projection_2D
        elif dimension == 3:
            self.process_projection = self.process_projection_3D
        else:
            raise ValueError(f"RayEngine: dimension must be 2 or 3, but was given {dimension}.")
        self._dimension = dimension
        self._check_exclusions(operations)
        self._operations = operations
        self._optical_system = optical_system
        self.compile_technical_intersections = compile_technical_intersections
        self.compile_stopped_rays = compile_stopped_rays
        self.compile_dead_rays = compile_dead_rays
        self.compile_finished_rays = compile_finished_rays
        self.compile_active_rays = compile_active_rays
        self.dead_ray_length = dead_ray_length
        self.compile_geometry_specific_result = compile_geometry_specific_result
        self.new_ray_length = new_ray_length
        self.clear_ray_history()
        self.last_projection_result = {}
        self.last_trace = None     # kernels.Trace3D of the last fused trace (tape, live counts)
        self._history_slots = {}
        self._history_iterations = {}
        self._pending = None       # fused 3D trace enqueued, histories not materialised yet
        self._workspace = None     # kernels.TraceWorkspace kept from trace to trace
        self._lease_ref = None     # weak reference to the lease of the workspace's last user

        self.input_signature = set()
        self.output_signature = set()
        self.optical_signature = set()
        self.stop_signature = set()
        self.target_signature = set()
        self.material_signature = set()
        self.simple_ray_inheritance = set(simple_ray_inheritance)
        for o in operations:
            self.input_signature |= o.input_signature
            self.output_signature |= o.output_signature
            self.optical_signature |= o.optical_signature
            self.stop_signature |= o.stop_signature
            self.target_signature |= o.target_signature
            self.material_signature |= o.material_signature
            self.simple_ray_inheritance |= o.simple_ray_inheritance

    def add_inheritable_field(self, fields):
        if type(fields) is str:
            fields = {fields}
        self.simple_ray_inheritance = self.simple_ray_inheritance | set(fields)

    def _check_exclusions(self, operations):
        exclusions, used = set(), set()
        for o in operations:
            used.add(o.__class__)
            exclusions |= o.exclusions
        matches = used & exclusions
        if bool(matches):
            raise RuntimeError(f"RayEngine: discovered exclusive operations: {matches}")
        self.operations = operations

    def update(self):
        if self.optical_system is not None:
            self.optical_system.update()

    def annotate(self, op_list=None):
        if self.optical_system is not None:
            for o in (self._operations if op_list is None else op_list):
                o.annotate(self)
        else:
            print("No optical system found, so annotating nothing.")

    @property
    def dimension(self):
        return self._dimension

    @property
    def new_ray_length(self):
        return self._new_ray_length

    @new_ray_length.setter
    def new_ray_length(self, val):
        self._new_ray_length = float(val)

    @property
    def optical_system(self):
        return self._optical_system

    @optical_system.setter
    def optical_system(self, val):
        if val.dimension != self.dimension:
            raise ValueError(
                f"OpticalEngine: attempted to set an optical system with dimension "
                f"{val.dimension}, but this engine is set to dimension {self.dimension}")
        self._optical_system = val

    # ------------------------------------------------------------------ histories
    # The fused 3D trace is enqueued by ray_trace() without waiting for the GPU; the sizes of the
    # history sets are data dependent, so they are materialised (one wait for the counters, one
    # small kernel per set, the autograd node) when a history property is first read.
    @property
    def active_rays(self):
        self._materialise_pending()
        return ReadOnlySet(amalgamate(self._active_rays))

    @property
    def finished_rays(self):
        self._materialise_pending()
        return ReadOnlySet(amalgamate(self._finished_rays))

    @property
    def dead_rays(self):
        self._materialise_pending()
        return ReadOnlySet(amalgamate(self._dead_rays))

    @property
    def stopped_rays(self):
        self._materialise_pending()
        return ReadOnlySet(amalgamate(self._stopped_rays))

    @property
    def all_rays(self):
        self._materialise_pending()
        return ReadOnlySet(amalgamate(
            self._active_rays + self._finished_rays + self._dead_rays + self._stopped_rays))

    @property
    def unfinished_rays(self):
        return self._unfinished_rays

    @property
    def last_history_slots(self):
        """Source-ray index of every row of each compiled history set of the last fused trace."""
        self._materialise_pending()
        return self._history_slots

    @last_history_slots.setter
    def last_history_slots(self, val):
        self._history_slots = val

    @property
    def last_history_iterations(self):
        """Iteration at which every row of each compiled history set was produced."""
        self._materialise_pending()
        return self._history_iterations

    @last_history_iterations.setter
    def last_history_iterations(self, val):
        self._history_iterations = val

    def clear_ray_history(self):
        self._pending = None
        self._active_rays = []
        self._finished_rays = []
        self._stopped_rays = []
        self._dead_rays = []
        self._unfinished_rays = {}

    # ------------------------------------------------------------------ validation
    def validate_system(self):
        """Signature checks of materials, sources and boundaries (tfrt/engine.py:1416-1522)."""
        system = self.optical_system
        if system is None:
            print("No optical system found, so validating nothing.")
            return
        for material in system.materials:
            sig = set(material.keys())
            if not (sig >= self.material_signature):
                raise RuntimeError(
                    f"Optical engine failed materials signature check.  System signature is "
                    f"{sig} but needed {self.material_signature}")

        def check(fields, required, what):
            if bool(fields):
                sig = set(fields.keys())
                if not (sig >= required):
                    raise RuntimeError(
                        f"Optical engine failed {what} signature check.  System signature is "
                        f"{sig}, but needed {required}.")

        if self.dimension == 2:
            check(system._amalgamated_sources, SEGMENT_GEO_SIG | self.input_signature, "sources")
            for kind, sig in (("optical", self.optical_signature), ("stop", self.stop_signature),
                              ("target", self.target_signature)):
                check(getattr(system, f"_amalgamated_{kind}_segments"), SEGMENT_GEO_SIG | sig,
                      f"{kind} segments")
                check(getattr(system, f"_amalgamated_{kind}_arcs"), ARC_GEO_SIG | sig,
                      f"{kind} arcs")
        else:
            check(system._amalgamated_sources, SOURCE_3D_SIG | self.input_signature, "sources")
            check(system._amalgamated_optical, TRIANGLE_GEO_SIG | self.optical_signature,
                  "optical segments")
            check(system._amalgamated_stop, TRIANGLE_GEO_SIG | self.stop_signature,
                  "stop segments")
            check(system._amalgamated_target, TRIANGLE_GEO_SIG | self.target_signature,
                  "target segments")

    def validate_output(self):
        if self.dimension == 2:
            required = SEGMENT_GEO_SIG | self.output_signature
            for rays in (self.active_rays, self.finished_rays, self.stopped_rays, self.dead_rays):
                if bool(rays):
                    sig = set(rays.keys())
                    if not (sig >= required):
                        raise RuntimeError(
                            f"Optical engine failed output signature check.  System signature "
                            f"is {sig}, but needed {required}.")

    # ------------------------------------------------------------------ the trace loop
    def _fusable(self):
        if not (self.dimension in (2, 3) and not self.compile_technical_intersections
                and not self.compile_geometry_specific_result
                and any(type(o) is op.StandardReaction for o in self._operations)
                and all(type(o) in (op.StandardReaction, op.OldestAncestor) and o.active
                        for o in self._operations)):
            return False
        # The 3D reaction kernel takes any {"n": callable} (closed-form materials are evaluated
        # per ray, other callables through a table over the distinct wavelengths); the 2D kernel
        # only knows the closed forms, other callables take the projection-at-a-time path.
        if self.dimension == 2 and self.optical_system is not None:
            reaction = next(o for o in self._operations if type(o) is op.StandardReaction)
            if reaction.refractive_index_type == "index" and \
                    not self.optical_system.materials_have_descriptors():
                return False
        return True

    def ray_trace(self, max_iterations=25):
        """Trace the optical system (tfrt/engine.py:2311-2330): iterate projection + reaction
        until `max_iterations` or until no active rays remain."""
        if not bool(self.optical_system):
            return
        self.clear_ray_history()
        if self._fusable():
            if self.dimension == 2:
                self._ray_trace_fused_2d(int(max_iterations))
            else:
                self._ray_trace_fused(int(max_iterations))
            return
        starting_rays = dict(self.optical_system._amalgamated_sources)
        for _ in range(max_iterations):
            result = self.single_pass(starting_rays)
            if bool(result):
                starting_rays = result
            else:
                break
        self._unfinished_rays = starting_rays

    def _ray_trace_fused(self, max_iterations):
        """Enqueue the whole trace loop on the GPU (asynchronous); see _materialise_pending."""
        system = self.optical_system
        src_fields = system._amalgamated_sources
        if not bool(src_fields) or not bool(system._merged):
            return
        reaction = next(o for o in self._operations if type(o) is op.StandardReaction)
        tri, norm, category, packed = system.merged_device_arrays()
        dev = tri.device
        src = system.source_rows()

        optical = system._amalgamated_optical
        kw = {}
        if reaction.refractive_index_type == "index":
            none = torch.zeros(0, dtype=torch.int32, device=dev)
            kw.update(mat_in=optical.get("mat_in", none), mat_out=optical.get("mat_out", none),
                      n_materials=len(system.materials))
            if system.materials_have_descriptors():
                kw.update(materials=system.material_table(dev))
            else:   # arbitrary callables: tabulated over the distinct wavelengths
                n_table, lambda_index = system.material_index_table(src[6])
                kw.update(n_table=n_table, lambda_index=lambda_index)
        else:
            none = torch.zeros(0, dtype=torch.float32, device=dev)
            kw.update(n_in=optical.get("n_in", none), n_out=optical.get("n_out", none))
        compile_sets = set()
        if self.compile_active_rays:
            compile_sets.add("active")
        if self.compile_finished_rays:
            compile_sets.add("finished")
        if self.compile_stopped_rays:
            compile_sets.add("stopped")
        if self.compile_dead_rays:
            compile_sets.add("dead")
        plan = TracePlan(
            category=category, n_optical=system._optical_count, max_iterations=max_iterations,
            intersect_eps=system.intersect_epsilion, size_eps=system.size_epsilion,
            ray_start_eps=system.ray_start_epsilion, new_ray_length=self.new_ray_length,
            margin_ulps=system.search_margin_ulps, grid=system.grid,
            dead_ray_length=self.dead_ray_length if bool(self.dead_ray_length) else None,
            compile=compile_sets, **kw)
        plan.packed = packed
        # stop / target boundaries that do not require grad need no cotangents (merged order is
        # optical -> stop -> target, so they are the tail of the triangle list)
        plan.n_grad_tri = system.grad_triangle_count()

        # Reuse the previous trace's workspace (tape + scratch) unless an autograd graph can
        # still read it: its lease is then still alive, and this trace gets a fresh one.
        ws = self._workspace
        if ws is not None and self._lease_ref is not None and self._lease_ref() is not None:
            ws = None
        norm_in = None if system._norm_is_geometric else norm
        tr = run_trace(tri, norm_in, src, plan, workspace=ws)
        plan.lease = WorkspaceLease(tr.ws)
        self._workspace, self._lease_ref = tr.ws, weakref.ref(plan.lease)
        self.last_trace = tr
        self._pending = (tri, norm_in, src, plan, src_fields)
        self.last_projection_result = {}

    def _materialise_pending(self):
        """Build the history sets of the pending fused trace in the reference's order (waits for
        the trace's counters) and attach them to the autograd graph of the boundary parameters."""
        if self._pending is None:
            return
        if self.dimension == 2:
            pending, self._pending = self._pending, None
            self._materialise_pending_2d(*pending)
            return
        tri, norm_in, src, plan, src_fields = self._pending
        self._pending = None
        active, stopped, finished, dead = TraceFunction.apply(tri, norm_in, src, plan)
        inherit = [f for f in src_fields if f not in RAY_FIELDS_3D]

        def as_set(name, rays):
            # a compiled set exists even when no ray fell into it: the reference appends one
            # (possibly empty) field dict per iteration (tfrt/engine.py:2028-2111)
            out = {k: rays[i] for i, k in enumerate(RAY_FIELDS_3D)}
            slot = plan.slots[name].long()
            for f in inherit:
                out[f] = src_fields[f][slot]
            return out

        for name, rays, store in (("active", active, self._active_rays),
                                  ("finished", finished, self._finished_rays),
                                  ("stopped", stopped, self._stopped_rays),
                                  ("dead", dead, self._dead_rays)):
            if name in plan.compile:
                store.append(as_set(name, rays))
        # source-ray index / iteration of every row of each compiled history set
        self._history_slots = dict(plan.slots)
        self._history_iterations = dict(plan.iters)
        plan.lease = None     # from here on only the autograd node (if any) holds the lease

    def _ray_trace_fused_2d(self, max_iterations):
        """Enqueue the whole 2D trace as one kernel launch (K4, asynchronous); the history sets
        are built when first read (_materialise_pending_2d)."""
        system = self.optical_system
        src_fields = system._amalgamated_sources
        seg, arc = system.merged_device_arrays()
        if not bool(src_fields) or (seg is None and arc is None):
            return
        reaction = next(o for o in self._operations if type(o) is op.StandardReaction)
        dev = (seg if seg is not None else arc).device
        src = system.source_rows()
        index_mode = reaction.refractive_index_type == "index"
        keys = ("mat_in", "mat_out") if index_mode else ("n_in", "n_out")

        def mats(d):
            return (d[keys[0]], d[keys[1]]) if bool(d) else None

        kw = dict(seg_mat=mats(system._amalgamated_optical_segments),
                  arc_mat=mats(system._amalgamated_optical_arcs),
                  intersect_eps=system.intersect_epsilion, size_eps=system.size_epsilion,
                  ray_start_eps=system.ray_start_epsilion, new_ray_length=self.new_ray_length)
        if index_mode:
            kw.update(materials=system.material_table(dev), n_materials=len(system.materials))
        counts = dict(seg_optical=system._optical_seg_count, seg_stop=system._stop_seg_count,
                      arc_optical=system._optical_arc_count, arc_stop=system._stop_arc_count)
        # Reuse the previous trace's tape unless an autograd graph can still read it (its lease
        # is then still alive), as the 3D path does with its workspace.
        old = self._workspace
        if old is not None and self._lease_ref is not None and self._lease_ref() is not None:
            old = None
        tr = kernels.Trace2D(src.detach(), None if seg is None else seg.detach(),
                             None if arc is None else arc.detach(), counts, max_iterations,
                             reuse=old, **kw)
        tr.forward()
        lease = WorkspaceLease(tr)
        plan = {"trace": tr, "lease": lease}
        self._workspace, self._lease_ref = tr, weakref.ref(lease)
        self.last_trace = tr
        self._pending = (seg, arc, src, plan, src_fields, max_iterations)
        self.last_projection_result = {}

    def _materialise_pending_2d(self, seg, arc, src, plan, src_fields, max_iterations):
        """Histories of the pending 2D trace, gathered from the per-(iteration, source ray) tape
        in the reference's order: per iteration the rays that hit segments, then those that hit
        arcs, each order-preserving (tfrt/engine.py:1640-1660, :1797-1815); the next pass sees
        them in the same order (:1950-1956).  The gathers are plain torch indexing of the tape,
        so the cotangents of the sets scatter back into the tape's cotangent."""
        from .autograd import Trace2DFunction
        tr = plan["trace"]
        dev = src.device
        n = tr.N
        tape = Trace2DFunction.apply(seg, arc, src, plan)        # (K, 4, N)
        plan["lease"] = None     # from here on only the autograd node (if any) holds the lease

        both = seg is not None and arc is not None
        code = {"active": 0, "stopped": 1, "finished": 2, "dead": 3}
        sets = {}
        if n * max_iterations <= (1 << 13) or (both and n * max_iterations <= (1 << 20)):
            # small traces are latency bound: ONE device->host copy of the status / hit tape,
            # the ordering on the host (NumPy), one host->device copy of the gather indices
            import numpy as np
            packed = torch.cat([tr.tape_status[:, :n].to(torch.int32), tr.tape_hit[:, :n]]).cpu().numpy()
            st_h, arc_h = packed[:max_iterations], packed[max_iterations:] < -1
            if not both:
                for name, c in code.items():
                    sets[name] = np.nonzero(st_h == c)
            else:
                acc = {name: ([], []) for name in code}
                order = np.arange(n)
                for k in range(max_iterations):
                    if order.size == 0:
                        break
                    st, ka = st_h[k, order], arc_h[k, order]
                    for name, c in code.items():
                        parts = [order[st == c]] if name == "dead" else \
                            [order[(st == c) & ~ka], order[(st == c) & ka]]
                        for part in parts:
                            acc[name][0].append(np.full(part.shape, k, dtype=np.int64))
                            acc[name][1].append(part)
                    order = np.concatenate([order[(st == 0) & ~ka], order[(st == 0) & ka]])
                for name in code:
                    sets[name] = (np.concatenate(acc[name][0]) if acc[name][0] else np.zeros(0, np.int64),
                                  np.concatenate(acc[name][1]) if acc[name][1] else np.zeros(0, np.int64))
            names = [nm for nm in code]
            sizes = [sets[nm][0].shape[0] for nm in names]
            flat = np.concatenate([np.concatenate([sets[nm][0], sets[nm][1]]) for nm in names]) \
                if sum(sizes) else np.zeros(0, np.int64)
            flat_dev = torch.from_numpy(flat.astype(np.int64)).to(dev)
            off = 0
            for nm, sz in zip(names, sizes):
                sets[nm] = (flat_dev[off:off + sz], flat_dev[off + sz:off + 2 * sz])
                off += 2 * sz
        else:
            status = tr.tape_status[:, :n]
            is_arc = tr.tape_hit[:, :n] < -1
            if not both:
                # one geometry: the live list stays in source order, so every set is the
                # row-major (iteration, slot) list of its status code.  One stable sort groups
                # the (iteration, slot) positions by code and one count gives the set sizes: a
                # single wait for the device instead of one per set
                flat = status.reshape(-1)
                _, where = torch.sort(flat, stable=True)
                sizes = torch.bincount(flat, minlength=4)[:4].tolist()
                off = 0
                for name, c in code.items():
                    idx = where[off:off + sizes[c]]
                    sets[name] = (torch.div(idx, n, rounding_mode="floor"), idx % n)
                    off += sizes[c]
            else:
                acc = {name: ([], []) for name in code}
                order = torch.arange(n, device=dev)
                for k in range(max_iterations):
                    if order.numel() == 0:
                        break
                    st, ka = status[k, order], is_arc[k, order]
                    for name, c in code.items():
                        parts = [order[st == c]] if name == "dead" else \
                            [order[(st == c) & ~ka], order[(st == c) & ka]]
                        for part in parts:
                            acc[name][0].append(torch.full_like(part, k))
                            acc[name][1].append(part)
                    order = torch.cat([order[(st == 0) & ~ka], order[(st == 0) & ka]])
                empty = torch.zeros(0, dtype=torch.long, device=dev)
                for name in code:
                    sets[name] = (torch.cat(acc[name][0]) if acc[name][0] else empty,
                                  torch.cat(acc[name][1]) if acc[name][1] else empty)

        inherit = [f for f in src_fields if f not in kernels.RAY_FIELDS_2D]
        self._history_slots, self._history_iterations = {}, {}
        wanted = [(name, store) for name, flag, store in (
            ("active", self.compile_active_rays, self._active_rays),
            ("finished", self.compile_finished_rays, self._finished_rays),
            ("stopped", self.compile_stopped_rays, self._stopped_rays),
            ("dead", self.compile_dead_rays, self._dead_rays)) if flag]
        # one gather for all compiled sets (latency matters here), then per-set views
        sizes = [int(sets[name][0].shape[0]) for name, _ in wanted]
        it_all = torch.cat([sets[name][0] for name, _ in wanted]) if wanted else None
        slot_all = torch.cat([sets[name][1] for name, _ in wanted]) if wanted else None
        rays_all = tape[it_all, :, slot_all] if wanted else None          # (n_total, 4)
        inherited = {f: src_fields[f][slot_all] for f in inherit} if wanted else {}
        off = 0
        for (name, store), size in zip(wanted, sizes):
            rays = rays_all[off:off + size]
            out = {k: rays[:, i] for i, k in enumerate(kernels.RAY_FIELDS_2D)}
            if name == "dead" and bool(self.dead_ray_length):
                for c in "xy":
                    out[c + "_end"] = out[c + "_start"] + self.dead_ray_length * (
                        out[c + "_end"] - out[c + "_start"])
            for f in inherit:
                out[f] = inherited[f][off:off + size]
            store.append(out)
            self._history_slots[name] = slot_all[off:off + size]
            self._history_iterations[name] = it_all[off:off + size]
            off += size
        self.last_projection_result = {}

    # ------------------------------------------------------------------ unfused single pass
    def process_projection_2D(self, input_rays):
        """Project rays onto the 2D system and split them by what they hit
        (tfrt/engine.py:1544-1986).  Result layout as in the reference (`rays`, `optical`,
        optionally `stop` / `target`); with both geometries present the combined sets are
        segment hits followed by arc hits, rays and boundary data alike (the reference
        concatenates the boundary data arc-first, tfrt/engine.py:1958-1965: a latent bug that
        mis-pairs rays and boundaries, not reproduced -- DESIGN.md section 9)."""
        system = self.optical_system
        fields = list(input_rays.keys())
        seg_proj, arc_proj = system.intersect(input_rays)
        has_segments, has_arcs = bool(seg_proj), bool(arc_proj)
        result = {}
        if self.compile_dead_rays:
            valid_any = None
            for proj in (seg_proj, arc_proj):
                if proj:
                    valid_any = proj["valid"] if valid_any is None else (valid_any | proj["valid"])
            is_dead = torch.logical_not(valid_any)
            dead = {f: input_rays[f][is_dead] for f in fields}
            if bool(self.dead_ray_length):
                for c in "xy":
                    dead[c + "_end"] = dead[c + "_start"] + self.dead_ray_length * (
                        dead[c + "_end"] - dead[c + "_start"])
            self._dead_rays.append(dead)

        input_rays = dict(input_rays)
        geo = {}
        for name, proj, merged, gkey, counts in (
                ("segment", seg_proj, system._merged_segments, "gather_segment",
                 (system._optical_seg_count, system._stop_seg_count)),
                ("arc", arc_proj, system._merged_arcs, "gather_arc",
                 (system._optical_arc_count, system._stop_arc_count))):
            if not proj:
                continue
            suffix = "segments" if name == "segment" else "arcs"
            valid = proj["valid"]
            for c in "xy":
                input_rays[c + "_end"] = torch.where(valid, proj[c], input_rays[c + "_end"])
            btype = merged["catagory"][proj[gkey]]
            entry = {"rays": {}}
            masks = {"active": valid & (btype == OPTICAL), "finished": valid & (btype == TARGET),
                     "stopped": valid & (btype == STOP)}
            entry["rays"]["active"] = {f: input_rays[f][masks["active"]] for f in fields}
            if self.compile_active_rays:
                self._active_rays.append(entry["rays"]["active"])
            if self.compile_finished_rays:
                entry["rays"]["finished"] = {f: input_rays[f][masks["finished"]] for f in fields}
                self._finished_rays.append(entry["rays"]["finished"])
            if self.compile_stopped_rays:
                entry["rays"]["stopped"] = {f: input_rays[f][masks["stopped"]] for f in fields}
                self._stopped_rays.append(entry["rays"]["stopped"])
            g = proj[gkey][masks["active"]]
            optical = {f: v[g] for f, v in getattr(system, f"_amalgamated_optical_{suffix}").items()}
            optical["norm"] = proj["norm"][masks["active"]]
            entry["optical"] = optical
            if self.compile_technical_intersections and self.compile_stopped_rays:
                g = (proj[gkey] - counts[0])[masks["stopped"]]
                entry["stop"] = {f: v[g] for f, v in
                                 getattr(system, f"_amalgamated_stop_{suffix}").items()}
                entry["stop"]["norm"] = proj["norm"][masks["stopped"]]
            if self.compile_technical_intersections and self.compile_finished_rays:
                g = (proj[gkey] - counts[0] - counts[1])[masks["finished"]]
                entry["target"] = {f: v[g] for f, v in
                                   getattr(system, f"_amalgamated_target_{suffix}").items()}
                entry["target"]["norm"] = proj["norm"][masks["finished"]]
            if self.compile_dead_rays:
                entry["rays"]["dead"] = dead
            geo[name] = entry
        if self.compile_geometry_specific_result:
            result.update(geo)
        if has_segments != has_arcs:
            only = geo["segment" if has_segments else "arc"]
            result["rays"] = dict(only["rays"])
            for k in ("optical", "stop", "target"):
                if k in only:
                    result[k] = only[k]
        else:
            a, b = geo["segment"], geo["arc"]
            result["rays"] = {t: amalgamate([a["rays"][t], b["rays"][t]])
                              for t in a["rays"] if t != "dead"}
            for k in ("optical", "stop", "target"):
                if k in a:
                    common = set(a[k].keys()) & set(b[k].keys())
                    result[k] = {f: torch.cat([a[k][f], b[k][f]], 0) for f in common}
        if self.compile_dead_rays:
            result["rays"]["dead"] = dead
        return result

    def process_projection_3D(self, input_rays):
        """Project rays onto the system and split them by what they hit
        (tfrt/engine.py:1988-2191).  Order-preserving masks, same result layout."""
        system = self.optical_system
        result = {"rays": {}}
        fields = list(input_rays.keys())
        inter = system.intersect(input_rays)
        valid = inter["valid"]

        if self.compile_dead_rays:
            is_dead = torch.logical_not(valid)
            dead = {f: input_rays[f][is_dead] for f in fields}
            if bool(self.dead_ray_length):
                for c in "xyz":
                    dead[c + "_end"] = dead[c + "_start"] + self.dead_ray_length * (
                        dead[c + "_end"] - dead[c + "_start"])
            self._dead_rays.append(dead)
            result["rays"]["dead"] = dead

        input_rays = dict(input_rays)
        for c in "xyz":
            input_rays[c + "_end"] = torch.where(valid, inter[c], input_rays[c + "_end"])
        boundary_type = system._merged["catagory"][inter["gather_trig"]]

        is_active = valid & (boundary_type == OPTICAL)
        active = {f: input_rays[f][is_active] for f in fields}
        if self.compile_active_rays:
            self._active_rays.append(active)
        result["rays"]["active"] = active

        is_finished = valid & (boundary_type == TARGET)
        if self.compile_finished_rays:
            finished = {f: input_rays[f][is_finished] for f in fields}
            self._finished_rays.append(finished)
            result["rays"]["finished"] = finished

        is_stopped = valid & (boundary_type == STOP)
        if self.compile_stopped_rays:
            stopped = {f: input_rays[f][is_stopped] for f in fields}
            self._stopped_rays.append(stopped)
            result["rays"]["stopped"] = stopped

        gather_optical = inter["gather_trig"][is_active]
        optical = {f: v[gather_optical] for f, v in system._amalgamated_optical.items()}
        if gather_optical.shape[0] > 0:
            optical["norm"] = inter["norm"][is_active]
        result["optical"] = optical

        if self.compile_technical_intersections and self.compile_stopped_rays:
            g = (inter["gather_trig"] - system._optical_count)[is_stopped]
            stop = {f: v[g] for f, v in system._amalgamated_stop.items()}
            if g.shape[0] > 0:
                stop["norm"] = inter["norm"][is_stopped]
            result["stop"] = stop
        if self.compile_technical_intersections and self.compile_finished_rays:
            g = (inter["gather_trig"] - system._optical_count - system._stop_count)[is_finished]
            target = {f: v[g] for f, v in system._amalgamated_target.items()}
            if g.shape[0] > 0:
                target["norm"] = inter["norm"][is_finished]
            result["target"] = target
        return result

    def _react_projection(self, reaction, proj_result):
        """StandardReaction.main on one projection result (tfrt/operation.py:255-307)."""
        rays = proj_result["rays"]["active"]
        optical = proj_result["optical"]
        if reaction.refractive_index_type == "index":
            wavelength = rays["wavelength"]
            if self.dimension == 3:
                wavelength = wavelength.to(torch.float64)
            n_stack = torch.stack(
                [torch.broadcast_to(torch.as_tensor(
                    (m["n"] if isinstance(m, dict) else m)(wavelength),
                    device=wavelength.device).to(wavelength.dtype), wavelength.shape)
                 for m in self.optical_system.materials])
            idx = torch.arange(wavelength.shape[0], device=wavelength.device)
            n_in = n_stack[optical["mat_in"].long(), idx]
            n_out = n_stack[optical["mat_out"].long(), idx]
        else:
            n_in, n_out = optical["n_in"], optical["n_out"]
        new_rays = {}
        if self.dimension == 2:
            (new_rays["x_start"], new_rays["y_start"], new_rays["x_end"],
             new_rays["y_end"]) = geometry.snells_law_2D(
                rays["x_start"], rays["y_start"], rays["x_end"], rays["y_end"],
                optical["norm"], n_in, n_out, self.new_ray_length)
        else:
            # fp64 arithmetic on the fp32 fields (what the fused reaction kernel does), fp32 out
            f64 = torch.float64
            out = geometry.snells_law_3D(
                *[rays[k].to(f64) for k in RAY_FIELDS_3D], optical["norm"].to(f64),
                n_in.to(f64), n_out.to(f64), self.new_ray_length)
            for k, v in zip(RAY_FIELDS_3D, out):
                new_rays[k] = v.to(torch.float32)
        valid = torch.ones_like(rays["x_end"], dtype=torch.bool)
        return {"active": {"rays": new_rays, "valid": valid}}

    def single_pass(self, input_rays):
        """One projection + reaction step (tfrt/engine.py:2193-2302)."""
        if not bool(self.optical_system):
            return {}
        self.last_projection_result = self.process_projection(input_rays)
        for branch in (self.last_projection_result,
                       self.last_projection_result.get("arc", {}),
                       self.last_projection_result.get("seg", {})):
            domain = branch.get("rays")
            if domain:
                for t in list(domain.keys()):
                    if domain[t]["x_start"].shape[0] == 0:
                        domain.pop(t)

        for o in self._operations:
            o.preprocess(self, self.last_projection_result)
        new_ray_dict = {}
        for o in self._operations:
            op_result = o.main(self, self.last_projection_result)
            if bool(op_result):
                new_ray_dict[o] = op_result

        for op_entry in new_ray_dict.values():
            for entry_type, entry in op_entry.items():
                base, _, geo = entry_type.partition("_")
                parent = (self.last_projection_result[geo]["rays"][base] if geo
                          else self.last_projection_result["rays"][base])
                for sig in self.simple_ray_inheritance:
                    entry["rays"][sig] = parent[sig]

        for o in self._operations:
            o.postprocess(self, self.last_projection_result, new_ray_dict)

        new_ray_list = []
        for op_entry in new_ray_dict.values():
            for geo_entry in op_entry.values():
                valid = geo_entry["valid"]
                new_ray_list.append({k: f[valid] for k, f in geo_entry["rays"].items()})
        return amalgamate(new_ray_list)
