"""
Boundaries: the field-dict objects of tfrt/boundaries.py that feed the trace hot path.

Every boundary is a `RecursivelyUpdatable` dict-like object whose fields are 1-D torch tensors of
equal length (one entry per segment / arc / triangle), fp32 on the GPU.  The triangle boundaries
run the vertices -> fields map (and its VJP) in the K3 CUDA kernels (tfrt/boundaries.py:890-923).

Reference classes mirrored here (same names / constructor arguments / attributes):
  BoundaryBase                 tfrt/boundaries.py:387-438
  ManualSegmentBoundary        tfrt/boundaries.py:493-524   (2D)
  ManualArcBoundary            tfrt/boundaries.py:458-473   (2D)
  TriangleBoundaryBase         tfrt/boundaries.py:830-938
  ManualTriangleBoundary       tfrt/boundaries.py:942-963
  ParametricTriangleBoundary   tfrt/boundaries.py:967-1112
  VectorGeneratorBase, FromVectorVG, FromPointVG, FromAxisVG, SecondSurfaceVG   :238-383

pyvista is not a dependency: a mesh is anything with `.points` (V,3) and `.faces` (pyvista's
packed [3,i,j,k, 3,i,j,k, ...] array or an (F,3) / (F,4) array) and `.copy()`; `TriangleMesh`
below is the minimal stand-in.
"""

from abc import ABC, abstractmethod

import numpy as np
import torch

from .autograd import FacesFunction
from .update import RecursivelyUpdatable

_DEFAULT_DEVICE = None


def default_device():
    """Device new field tensors are created on: the current CUDA device when there is one."""
    if _DEFAULT_DEVICE is not None:
        return _DEFAULT_DEVICE
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() \
        else torch.device("cpu")


def set_default_device(device):
    global _DEFAULT_DEVICE
    _DEFAULT_DEVICE = None if device is None else torch.device(device)


def as_field(value, device=None):
    """Convert to a torch tensor on the field device; floating data becomes fp32 (the kernels'
    storage type), integer / bool data keeps its type."""
    device = device or default_device()
    if torch.is_tensor(value):
        t = value
    else:
        t = torch.as_tensor(np.asarray(value))
    if t.is_floating_point() and t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.to(device)


def as_float_field(value, device=None):
    """as_field for geometric data: integer input (e.g. the tuple (1, 0, 0)) becomes fp32."""
    t = as_field(value, device)
    return t if t.is_floating_point() else t.to(torch.float32)


class TriangleMesh:
    """Minimal mesh container (stand-in for pyvista.PolyData): points (V,3), faces (F,3)."""

    def __init__(self, points, faces):
        self.points = np.array(points, dtype=np.float64).reshape(-1, 3)
        self.faces = unpack_faces(faces)

    def copy(self):
        return TriangleMesh(self.points.copy(), self.faces.copy())


def unpack_faces(faces):
    """pyvista-packed [3,i,j,k,...] / (F,4) / (F,3) -> (F,3) int64 numpy
    (tfrt/mesh_tools.py:1152-1157)."""
    f = np.asarray(faces.cpu() if torch.is_tensor(faces) else faces).astype(np.int64)
    if f.ndim == 1:
        if f.size % 4 != 0:
            raise ValueError("TriangleBoundary: mesh must consist entirely of triangles.")
        f = f.reshape(-1, 4)
    if f.shape[1] == 4:
        if f.size and not np.all(f[:, 0] == 3):
            raise ValueError("TriangleBoundary: mesh must consist entirely of triangles.")
        f = f[:, 1:]
    return np.ascontiguousarray(f)


# =========================================================================================
# vector generators (tfrt/boundaries.py:238-383)

class VectorGeneratorBase(ABC):
    @abstractmethod
    def generate(self, zero):
        raise NotImplementedError

    @staticmethod
    def normalize(val):
        return val / torch.linalg.norm(val, dim=1, keepdim=True)


class SecondSurfaceVG(VectorGeneratorBase):
    """Vectors point from the zero points to the matching vertices of a second surface."""

    def __init__(self, surface):
        self.surface = surface

    def generate(self, zero):
        return self.normalize(self.points.to(zero.device) - zero)

    @property
    def surface(self):
        return self._surface

    @surface.setter
    def surface(self, val):
        if hasattr(val, "points"):
            self.points = val.points
            self._surface = val
        else:
            self.points = val
            self._surface = None

    @property
    def points(self):
        return self._points

    @points.setter
    def points(self, val):
        self._points = as_float_field(val)


class FromPointVG(VectorGeneratorBase):
    """Vectors project away from a single 3-D point."""

    def __init__(self, point):
        self.point = point

    def generate(self, zero):
        return self.normalize(zero - self.point.to(zero.device))

    @property
    def point(self):
        return self._point

    @point.setter
    def point(self, val):
        self._point = as_float_field(val)


class FromVectorVG(VectorGeneratorBase):
    """Copies one 3-vector (or an (N,3) list of vectors) to every vertex."""

    def __init__(self, vector):
        self.vector = vector

    def generate(self, zero):
        return self.normalize(torch.broadcast_to(self.vector.to(zero.device), zero.shape))

    @property
    def vector(self):
        return self._vector

    @vector.setter
    def vector(self, val):
        self._vector = as_float_field(val)


class FromAxisVG(VectorGeneratorBase):
    """Vectors project perpendicular to an axis, through the points.  The axis is given by a
    point plus a kwarg `point` (second point) or `direction`."""

    def __init__(self, first, **kwargs):
        self.axis_point = as_float_field(first)
        if "point" in kwargs:
            axis_vector = as_float_field(kwargs["point"]) - self.axis_point
        elif "direction" in kwargs:
            axis_vector = as_float_field(kwargs["direction"])
        else:
            raise ValueError(
                "FromAxisVG: Must provide a kwarg 'point' or 'direction' to define the axis.")
        self.axis_vector = self.normalize(axis_vector.reshape(1, 3))

    def generate(self, zero):
        axis_point = self.axis_point.to(zero.device)
        axis_vector = torch.broadcast_to(self.axis_vector.to(zero.device), zero.shape)
        d = torch.sum((zero - axis_point) * axis_vector, dim=1)
        closest_point = axis_point + axis_vector * d.reshape(-1, 1)
        return self.normalize(zero - closest_point)


# =========================================================================================
# constraints between parametric surfaces (tfrt/boundaries.py:17-235).  A constraint's make()
# returns a handler that is appended to a surface's update_handles; it shifts or clips the
# surface's `parameters` in place and is not differentiated (the reference uses assign_add
# inside the tape).

class Constraint(ABC):
    """Base class.  `parent` selects what the constraint is relative to: "prev" (previous
    surface of a multi-boundary, or zero for the first), "zero", "literal" (target and parent
    passed to make() are the surfaces themselves) or an integer surface index."""

    def __init__(self, parent="prev"):
        if type(parent) is int:
            if parent < 0:
                raise ValueError("Constraint: integer parent must be >= 0")
        elif type(parent) is str:
            if parent not in {"zero", "prev", "literal"}:
                raise ValueError("Constraint: string parent must be 'zero' or 'prev'.")
        self._parent = parent

    @property
    def parent(self):
        return self._parent

    def interpret_params(self, target, parent):
        """(target parameters, parent parameters) for the mode chosen at construction."""
        if self._parent == "literal":
            return target.parameters, parent.parameters
        if self._parent == "zero":
            return target.parameters, torch.zeros_like(target.parameters)
        if self._parent == "prev":
            mine = parent[target].parameters
            if target == 0:
                return mine, torch.zeros_like(mine)
            return mine, parent[target - 1].parameters
        return parent[target].parameters, parent[self._parent].parameters

    @abstractmethod
    def make(self, target, parent):
        raise NotImplementedError


class NoConstraint(Constraint):
    def make(self, target, parent):
        return lambda: None


class PointConstraint(Constraint):
    """Fixes the parameter-space distance between one vertex of the target surface and one
    vertex of the parent surface (tfrt/boundaries.py:124-158)."""

    def __init__(self, distance, target_vertex, parent_vertex=None, **kwargs):
        super().__init__(**kwargs)
        self.distance = distance
        self.target_vertex = target_vertex
        self.parent_vertex = parent_vertex or target_vertex

    def make(self, target, parent):
        def handler(target=target, parent=parent):
            t, p = self.interpret_params(target, parent)
            with torch.no_grad():
                t.add_(p[self.parent_vertex] - t[self.target_vertex] + self.distance)
        return handler


class ThicknessConstraint(Constraint):
    """Fixes the minimum ("min") or maximum ("max") parameter-space distance between the target
    surface and its parent (tfrt/boundaries.py:162-215)."""

    def __init__(self, distance, mode, **kwargs):
        super().__init__(**kwargs)
        self.distance = distance
        self.mode = mode

    @property
    def mode(self):
        return self._mode

    @mode.setter
    def mode(self, val):
        if val not in {"min", "max"}:
            raise ValueError("ThicknessConstraint: mode must be either 'min' or 'max'.")
        self._mode = val
        self._reduce = torch.max if val == "min" else torch.min

    def make(self, target, parent):
        def handler(target=target, parent=parent):
            t, p = self.interpret_params(target, parent)
            with torch.no_grad():
                t.add_(self._reduce(p - t) + self.distance)
        return handler


class ClipConstraint:
    """Clips a surface's parameters to [lower, upper] (tfrt/boundaries.py:219-235)."""

    def __init__(self, lower, upper):
        self.lower = lower
        self.upper = upper
        self.parent = "zero"

    def make(self, target, _):
        def handler(target=target):
            with torch.no_grad():
                target.parameters.clamp_(self.lower, self.upper)
        return handler


# =========================================================================================

class BoundaryBase(RecursivelyUpdatable):
    """Base class of boundaries (tfrt/boundaries.py:387-438).  `material_dict` entries become
    fields, scalars broadcast to one value per boundary element."""

    def __init__(self, name=None, material_dict=None, **kwargs):
        self._name = name
        self._fields = {}
        self.material_dict = {} if material_dict is None else material_dict
        super().__init__(**kwargs)
        self.update_materials()

    @property
    def name(self):
        return self._name

    @property
    @abstractmethod
    def dimension(self):
        raise NotImplementedError

    @abstractmethod
    def update_materials(self):
        raise NotImplementedError

    def _update_materials(self, count, device):
        for field, value in self.material_dict.items():
            t = torch.as_tensor(value)
            if t.dim() < 1:
                t = torch.broadcast_to(t, (count,)).clone()
            self[field] = as_field(t, device)

    @property
    def keys(self):
        return self._fields.keys

    def __getitem__(self, key):
        return self._fields[key]

    def __setitem__(self, key, item):
        self._fields[key] = item

    def __bool__(self):
        return bool(self._fields)


# ----------------------------------------------------------------------------------------- 2D

class ManualSegmentBoundary(BoundaryBase):
    """Static 2D segments: fields x_start, y_start, x_end, y_end set by indexing
    (tfrt/boundaries.py:493-524)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    @property
    def dimension(self):
        return 2

    def feed_segments(self, segments):
        """segments: (n, 4) rows (x_start, y_start, x_end, y_end)."""
        s = as_field(segments).reshape(-1, 4)
        self["x_start"], self["y_start"] = s[:, 0].contiguous(), s[:, 1].contiguous()
        self["x_end"], self["y_end"] = s[:, 2].contiguous(), s[:, 3].contiguous()
        self.update_materials()

    def update_materials(self):
        if "x_start" in self._fields:
            self._update_materials(self["x_start"].shape[0], self["x_start"].device)

    def _generate_update_handles(self):
        return []

    def _update(self):
        pass


class ManualArcBoundary(BoundaryBase):
    """Static 2D arcs: fields x_center, y_center, angle_start, angle_end, radius
    (tfrt/boundaries.py:458-473)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    @property
    def dimension(self):
        return 2

    def feed_arcs(self, arcs):
        """arcs: (n, 5) rows (x_center, y_center, angle_start, angle_end, radius)."""
        a = as_field(arcs).reshape(-1, 5)
        for i, k in enumerate(("x_center", "y_center", "angle_start", "angle_end", "radius")):
            self[k] = a[:, i].contiguous()
        self.update_materials()

    def update_materials(self):
        if "x_center" in self._fields:
            self._update_materials(self["x_center"].shape[0], self["x_center"].device)

    def _generate_update_handles(self):
        return []

    def _update(self):
        pass


# ----------------------------------------------------------------------------------------- 3D

class TriangleBoundaryBase(BoundaryBase):
    """Base class of triangle boundaries (tfrt/boundaries.py:830-938)."""

    def __init__(self, file_name=None, mesh=None, vertex_update_map=None, **kwargs):
        if file_name:
            self._mesh = _read_mesh(file_name)
        elif mesh is not None:
            self._mesh = mesh
        else:
            self._mesh = None
        self._vertices = None
        self._faces = None
        self.vertex_update_map = vertex_update_map
        super().__init__(**kwargs)

    @property
    def dimension(self):
        return 3

    def update_materials(self):
        if "xp" in self._fields:
            self._update_materials(self["xp"].shape[0], self["xp"].device)

    @property
    def mesh(self):
        return self._mesh

    @property
    def vertices(self):
        return self._vertices

    @property
    def faces(self):
        return self._faces

    @property
    def vertex_update_map(self):
        return self._vertex_update_map

    @vertex_update_map.setter
    def vertex_update_map(self, val):
        self._vertex_update_map = val
        self._update_map_dev = None

    def save(self, filename, **kwargs):
        if self._mesh is not None and hasattr(self._mesh, "save"):
            self._mesh.save(filename, **kwargs)

    def update_vertices_from_mesh(self):
        """(Re-)read vertices and faces from the mesh (tfrt/boundaries.py:871-888).  The device
        copies are kept when the mesh arrays are bit-identical to the ones uploaded last time
        (a static boundary is re-read at every system.update(); comparing a few KB on the host
        is cheaper than a host-to-device copy and a kernel)."""
        if self._mesh is None:
            return False
        points = np.ascontiguousarray(np.asarray(self._mesh.points))
        faces = np.asarray(self._mesh.faces)
        cached = getattr(self, "_mesh_upload", None)
        if (cached is not None and self._vertices is not None and self._faces is not None
                and cached[0].shape == points.shape and cached[1].shape == faces.shape
                and np.array_equal(cached[0], points) and np.array_equal(cached[1], faces)):
            return False
        self._vertices = as_field(points)
        self._faces = torch.from_numpy(unpack_faces(faces)).to(
            dtype=torch.int32, device=self._vertices.device)
        self._mesh_upload = (points.copy(), faces.copy())
        return True

    def update_fields_from_vertices(self):
        """vertices -> xp..z2 + unit norm through the K3 kernel; differentiable w.r.t. the
        vertices except where vertex_update_map is False (tfrt/boundaries.py:890-923)."""
        if self._faces is None or self._vertices is None:
            return
        if self._vertex_update_map is not None and self._update_map_dev is None:
            m = np.asarray(self._vertex_update_map).astype(np.uint8).reshape(-1, 3)
            self._update_map_dev = torch.from_numpy(np.ascontiguousarray(m)).to(
                self._vertices.device)
        tri, norm = FacesFunction.apply(self._vertices, self._faces, self._update_map_dev)
        for i, k in enumerate(("xp", "yp", "zp", "x1", "y1", "z1", "x2", "y2", "z2")):
            self[k] = tri[i]
        self["norm"] = norm
        self._geometric_norm = norm   # the fused trace may re-derive it from the corners in fp64
        self._tri_block = tri         # (9,F): the geometry fields are its rows (merged as a block)

    def has_geometric_norm(self):
        """True while the `norm` field is the one computed from the vertices (not overridden)."""
        return self._fields.get("norm") is getattr(self, "_geometric_norm", None) is not None

    def update_from_mesh(self):
        if self._mesh is not None:
            changed = self.update_vertices_from_mesh()
            if changed or "xp" not in self._fields or getattr(self, "_tri_block", None) is None:
                self.update_fields_from_vertices()

    def update_mesh_from_vertices(self):
        if self._vertices is not None and self._mesh is not None:
            self._mesh.points = self._vertices.detach().cpu().numpy().astype(np.float64)


def _read_mesh(file_name):
    try:
        import pyvista as pv
    except ImportError as e:
        raise RuntimeError(
            "TriangleBoundary: reading mesh files needs pyvista, which is not installed; pass "
            "mesh=TriangleMesh(points, faces) instead.") from e
    return pv.read(file_name)


class ManualTriangleBoundary(TriangleBoundaryBase):
    """Static triangle boundary built from a mesh (tfrt/boundaries.py:942-963)."""

    def _generate_update_handles(self):
        return []

    def _update(self):
        self.update_from_mesh()


class ParametricTriangleBoundary(TriangleBoundaryBase):
    """A single parametric surface: vertices = zero_points + parameters * vectors
    (tfrt/boundaries.py:967-1112).  `parameters` is a leaf tensor with requires_grad=True (the
    tf.Variable of the reference) that an optimizer can target."""

    def __init__(self, zero_points, vector_generator, flip_norm=False, initial_parameters=0.0,
                 validate_shape=True, parameters=None, auto_update_mesh=False,
                 vertex_update_map=None, **kwargs):
        if isinstance(zero_points, str):
            zero_points = _read_mesh(zero_points)
        else:
            zero_points = zero_points.copy()
        if flip_norm:
            zero_points = self._flip_norm(zero_points)
            if vertex_update_map is not None:
                vertex_update_map = np.take(vertex_update_map, [2, 1, 0], axis=1)
        self._zero_points_mesh = zero_points
        self._zero_points = as_field(zero_points.points)
        self.vertex_update_map = vertex_update_map

        self._mesh = zero_points.copy()
        self._vertices = None
        self._faces = None
        TriangleBoundaryBase.update_vertices_from_mesh(self)

        self.vector_generator = vector_generator
        self.auto_update_mesh = auto_update_mesh
        self.reparametrize(self._zero_points)

        if parameters is None:
            init = torch.broadcast_to(
                torch.as_tensor(initial_parameters, dtype=torch.float32),
                (self._zero_points.shape[0],)).clone().to(self._zero_points.device)
            self.parameters = init.requires_grad_(True)
        else:
            self.parameters = parameters

        BoundaryBase.__init__(self, **kwargs)
        if not self.auto_update_mesh:
            self.update_mesh_from_vertices()

    def _generate_update_handles(self):
        return []

    def _update(self):
        self._vertices = self._update_internal(self._zero_points, self._vectors, self.parameters)
        if self.auto_update_mesh:
            self.update_mesh_from_vertices()
        self.update_fields_from_vertices()

    @staticmethod
    def _update_internal(zero, vectors, parameter):
        return zero + parameter.reshape(-1, 1) * vectors

    def reparametrize(self, zero_points):
        self._vectors = self.vector_generator.generate(self._zero_points)

    @property
    def zero_points(self):
        return self._zero_points_mesh

    @property
    def vectors(self):
        return self._vectors

    def _flip_norm(self, mesh):
        """Reverse the winding of every face (tfrt/boundaries.py:1059-1065).  The faces are
        written back in the layout the mesh object uses: (F,3) for TriangleMesh, pyvista's
        packed [3, i, j, k, ...] for anything else (a pyvista.PolyData would misread (F,3))."""
        flipped = np.ascontiguousarray(unpack_faces(mesh.faces)[:, ::-1])
        if isinstance(mesh, TriangleMesh):
            mesh.faces = flipped
        else:
            mesh.faces = np.hstack(
                [np.full((flipped.shape[0], 1), 3, dtype=flipped.dtype), flipped]).ravel()
        return mesh

    def update_vertices_from_mesh(self):
        raise RuntimeError(
            "ParametricTriangleBoundary: update_vertices_from_mesh is disabled for parametric "
            "boundaries.")

    def update_from_mesh(self):
        raise RuntimeError(
            "ParametricTriangleBoundary: update_from_mesh is disabled for parametric boundaries.")


class ParametricMultiTriangleBoundary(TriangleBoundaryBase):
    """Several parametric triangle surfaces over one set of zero points, tied together by
    constraints (tfrt/boundaries.py:1233-1412).  `surfaces` is the list to hand to
    OpticalSystem3D.optical; `parameters` the list of their parameter tensors."""

    def __init__(self, zero_points, vector_generator, constraints, flip_norm,
                 initial_parameters=0.0, validate_shape=True, parameters=None, material_list=(),
                 **kwargs):
        if isinstance(zero_points, str):
            zero_points = _read_mesh(zero_points)
        self.zero_points = zero_points
        self.vector_generator = vector_generator
        try:
            self._surface_count = len(constraints)
        except TypeError as e:
            raise ValueError("ParametricMultiTriangleBoundary: constraints must be iterable.") from e

        def per_surface(value, name, scalar_ok=True):
            try:
                if len(value) != self._surface_count:
                    raise ValueError(f"ParametricMultiTriangleBoundary: constraints and {name} "
                                     "must have the same size.")
                return list(value)
            except TypeError as e:
                if scalar_ok:
                    return [value] * self._surface_count
                raise ValueError(f"ParametricMultiTriangleBoundary: {name} must be iterable.") from e

        flip_norm = per_surface(flip_norm, "flip_norm", scalar_ok=False)
        initial_parameters = per_surface(initial_parameters, "initial_parameters")
        if parameters is not None:
            parameters = per_surface(parameters, "parameters", scalar_ok=False)
        material_list = list(material_list) if len(material_list) else [{}] * self._surface_count
        if len(material_list) != self._surface_count:
            raise ValueError("ParametricMultiTriangleBoundary: constraints and material_list "
                             "must have the same size.")
        vmap = kwargs.pop("vertex_update_map", None)
        self._surfaces = [
            ParametricTriangleBoundary(
                self.zero_points, self.vector_generator, flip_norm=flip_norm[i],
                initial_parameters=initial_parameters[i],
                parameters=None if parameters is None else parameters[i],
                material_dict=material_list[i], vertex_update_map=vmap, **kwargs)
            for i in range(self._surface_count)]
        self.constraints = list(constraints)
        for i, (surface, constraint) in enumerate(zip(self._surfaces, self.constraints)):
            if constraint.parent != "zero":
                surface.update_handles.append(constraint.make(i, self._surfaces))
            else:
                surface.update_handles.append(constraint.make(surface, None))
        self._name = kwargs.get("name")
        self._fields = {}
        self.material_dict = {}
        RecursivelyUpdatable.__init__(self)

    def _update(self):
        from .engine import amalgamate
        self._fields = amalgamate(self._surfaces)

    def _generate_update_handles(self):
        return [surface.update for surface in self._surfaces]

    def update_materials(self):
        pass

    @property
    def surface_count(self):
        return self._surface_count

    @property
    def surfaces(self):
        return self._surfaces

    @property
    def parameters(self):
        return [surface.parameters for surface in self._surfaces]


# ----------------------------------------------------------------------------------------- 2D

class ParametricSegmentBoundary(BoundaryBase):
    """A curve approximated by segments whose vertices slide along the lines from the points
    of `zero_distribution` to those of `one_distribution`:
    points = zero + parameters * (one - zero) (tfrt/boundaries.py:528-627).  The distributions
    are objects with `.points` (n,2) (and optionally `.update`) or plain arrays."""

    def __init__(self, zero_distribution, one_distribution, flip_norm=False,
                 initial_parameters=0.0, validate_shape=True, parameters=None, **kwargs):
        self._zero_distribution = zero_distribution
        self._one_distribution = one_distribution
        self.flip_norm = flip_norm
        zero = self._points(zero_distribution)
        if parameters is None:
            init = torch.broadcast_to(torch.as_tensor(initial_parameters, dtype=torch.float32),
                                      (zero.shape[0],)).clone().to(zero.device)
            self.parameters = init.requires_grad_(True)
        else:
            self.parameters = parameters
        super().__init__(**kwargs)

    @staticmethod
    def _points(dist):
        return as_field(getattr(dist, "points", dist)).reshape(-1, 2)

    @property
    def dimension(self):
        return 2

    def update_materials(self):
        if "x_start" in self._fields:
            self._update_materials(self["x_start"].shape[0], self["x_start"].device)

    def _generate_update_handles(self):
        return [d.update for d in (self._zero_distribution, self._one_distribution)
                if hasattr(d, "update")]

    def _update(self):
        zero, one = self._points(self._zero_distribution), self._points(self._one_distribution)
        (self["x_start"], self["y_start"], self["x_end"], self["y_end"]) = self._update_internal(
            zero, one, self.parameters, self.flip_norm)
        self.update_materials()

    @staticmethod
    def _update_internal(zero, one, parameter, flip_norm):
        points = zero + parameter.reshape(-1, 1) * (one - zero)
        if flip_norm:
            return points[1:, 0], points[1:, 1], points[:-1, 0], points[:-1, 1]
        return points[:-1, 0], points[:-1, 1], points[1:, 0], points[1:, 1]

    @property
    def zero_distribution(self):
        return self._zero_distribution

    @property
    def one_distribution(self):
        return self._one_distribution


class MasterSlaveParametricTriangleBoundary(ParametricTriangleBoundary):
    """Parametric triangle boundary in which several vertices share one parameter
    (tfrt/boundaries.py:1116-1229): `filter_masters` (callable on the (V,3) vertices returning a
    boolean mask or index list, or the index list itself) names the vertices that own a
    parameter; `attach_slaves(vertices, master, available_slaves)` returns the slave vertices
    that follow each master.  parameters has one entry per master; the expansion to vertices is a
    gather whose transpose (scatter-add) autograd applies to the vertex gradients."""

    def __init__(self, filter_masters, attach_slaves, *args, **kwargs):
        self._gather = None
        super().__init__(*args, **kwargs)
        verts = self._vertices.detach().cpu().numpy()
        masters = filter_masters(verts) if callable(filter_masters) else filter_masters
        masters = np.asarray(masters)
        if masters.dtype == np.bool_:
            masters = np.nonzero(masters)[0]
        masters = [int(m) for m in masters]
        master_index = {m: i for i, m in enumerate(masters)}
        unclaimed = set(range(verts.shape[0])) - set(masters)
        owner = dict(master_index)
        for m in masters:
            slaves = set(int(s) for s in attach_slaves(verts, m, unclaimed))
            unclaimed -= slaves
            for s in slaves:
                owner[s] = master_index[m]
        missing = [v for v in range(verts.shape[0]) if v not in owner]
        if missing:
            raise ValueError(f"MasterSlaveParametricTriangleBoundary: vertices {missing[:8]} "
                             "were attached to no master.")
        dev = self._zero_points.device
        self._gather = torch.tensor([owner[v] for v in range(verts.shape[0])], dtype=torch.long,
                                    device=dev)
        self.masters = masters
        self.parameters = self.parameters.detach()[torch.tensor(masters, device=dev)].clone() \
            .requires_grad_(True)
        self._update()

    def _update(self):
        if self._gather is None:       # base-class constructor, before the gather exists
            return super()._update()
        self._vertices = self._update_internal(self._zero_points, self._vectors,
                                               self.parameters[self._gather])
        if self.auto_update_mesh:
            self.update_mesh_from_vertices()
        self.update_fields_from_vertices()


class ParametricCylindricalGuide(TriangleBoundaryBase):
    """A closed cylinder-like light guide whose radius varies along (and optionally around) its
    axis (tfrt/boundaries.py:1416-1717): vertices = cylinder of `minimum_radius` + parameters *
    radial unit vectors, parameters >= 0 with min(parameters) pinned to 0 on every update.
    rotationally_symmetric=True: one parameter per axial layer (z_res), else one per side vertex
    (z_res * theta_res).  Cap vertices (on the axis) never move.  The class builds its own mesh
    (mesh_tools.cylindrical_mesh), vertex_update_map and accumulator
    (mesh_tools.mesh_parametrization_tools)."""

    def __init__(self, start, end, minimum_radius, theta_res=6, z_res=8, start_cap=True,
                 end_cap=True, rotationally_symmetric=False, initial_parameters=0.0,
                 initial_taper=None, auto_update_mesh=False, use_vertex_update_map=True,
                 **kwargs):
        from . import mesh_tools as mt
        pts, faces = mt.cylindrical_mesh(start, end, radius=minimum_radius, theta_res=theta_res,
                                         z_res=z_res, start_cap=start_cap, end_cap=end_cap)
        self._mesh = TriangleMesh(pts, faces)
        self._zero_points_mesh = self._mesh.copy()
        self._zero_points = as_float_field(pts)
        self._vertices = self._zero_points
        self._faces = torch.from_numpy(faces).to(dtype=torch.int32, device=self._zero_points.device)
        self._n_side = theta_res * z_res
        # cap vertices first / last, as in the reference's mesh (tfrt/mesh_tools.py:905-945)
        self._cap_pad = (int(bool(start_cap)), int(bool(end_cap)))
        self._theta_res, self._z_res = theta_res, z_res
        self._guide_map, self._accumulator = mt.mesh_parametrization_tools(
            self._mesh, mt.get_closest_point(self._mesh, start),
            active_vertices=range(self._cap_pad[0], self._cap_pad[0] + self._n_side))
        self._vertex_update_map = None
        self._update_map_dev = None
        self.use_vertex_update_map = use_vertex_update_map
        self.vector_generator = FromAxisVG(start, point=end)
        self.auto_update_mesh = auto_update_mesh
        self._rotationally_symmetric = rotationally_symmetric
        self.reparametrize(self._zero_points)
        dev = self._zero_points.device
        size = z_res if rotationally_symmetric else self._n_side
        if initial_taper is not None:
            try:
                t0, t1 = float(initial_taper[0]), float(initial_taper[1])
            except (IndexError, TypeError) as e:
                raise ValueError(
                    "ParametricCylindricalGuide: initial_taper must be None or a 2-tuple.") from e
            init = torch.linspace(t0, t1, z_res, device=dev)
            if not rotationally_symmetric:
                init = init.repeat_interleave(theta_res)
        else:
            init = torch.broadcast_to(torch.as_tensor(initial_parameters, dtype=torch.float32),
                                      (size,)).clone().to(dev)
        self.parameters = init.to(torch.float32).requires_grad_(True)
        BoundaryBase.__init__(self, **kwargs)
        if not self.auto_update_mesh:
            self.update_mesh_from_vertices()

    def _generate_update_handles(self):
        return []

    def _constraint(self):
        with torch.no_grad():
            self.parameters.sub_(self.parameters.min())

    def _update(self):
        self._constraint()
        p = self.parameters
        if self._rotationally_symmetric:
            p = p.repeat_interleave(self._theta_res)
        p = torch.nn.functional.pad(p, self._cap_pad)            # cap vertices stay on the axis
        self._vertices = self._zero_points + p.reshape(-1, 1) * self._vectors
        if self.auto_update_mesh:
            self.update_mesh_from_vertices()
        self.update_fields_from_vertices()

    def reparametrize(self, zero_points):
        v = self.vector_generator.generate(self._zero_points)
        # the cap vertices lie on the axis: no radial direction (the generator yields 0/0)
        self._vectors = torch.where(torch.isfinite(v), v, torch.zeros_like(v))

    @property
    def zero_points(self):
        return self._zero_points_mesh

    @property
    def vectors(self):
        return self._vectors

    @property
    def use_vertex_update_map(self):
        return self._use_vertex_update_map

    @use_vertex_update_map.setter
    def use_vertex_update_map(self, val):
        self._use_vertex_update_map = bool(val)
        self.vertex_update_map = self._guide_map if val else None

    @property
    def accumulator(self):
        return self._accumulator

    @property
    def rotationally_symmetric(self):
        return self._rotationally_symmetric

    def update_vertices_from_mesh(self):
        raise RuntimeError("ParametricCylindricalGuide: update_vertices_from_mesh is disabled "
                           "for parametric boundaries.")

    def update_from_mesh(self):
        raise RuntimeError("ParametricCylindricalGuide: update_from_mesh is disabled for "
                           "parametric boundaries.")


class ParametricMultiSegmentBoundary(BoundaryBase):
    """Several parametric segment curves over one pair of base-point distributions, tied
    together by constraints (tfrt/boundaries.py:631-826); the 2D sibling of
    ParametricMultiTriangleBoundary.  `surfaces` is the list for OpticalSystem2D.optical_segments."""

    def __init__(self, zero_distribution, one_distribution, constraints, flip_norm,
                 initial_parameters=0.0, validate_shape=True, parameters=None, material_list=(),
                 **kwargs):
        self._zero_distribution = zero_distribution
        self._one_distribution = one_distribution
        try:
            self._surface_count = len(constraints)
        except TypeError as e:
            raise ValueError("ParametricMultiSegmentBoundary: constraints must be iterable.") from e

        def per_surface(value, name, scalar_ok=True):
            try:
                if len(value) != self._surface_count:
                    raise ValueError(f"ParametricMultiSegmentBoundary: constraints and {name} "
                                     "must have the same size.")
                return list(value)
            except TypeError as e:
                if scalar_ok:
                    return [value] * self._surface_count
                raise ValueError(f"ParametricMultiSegmentBoundary: {name} must be iterable.") from e

        self.flip_norm = per_surface(flip_norm, "flip_norm", scalar_ok=False)
        initial_parameters = per_surface(initial_parameters, "initial_parameters")
        if parameters is not None:
            parameters = per_surface(parameters, "parameters", scalar_ok=False)
        material_list = list(material_list) if len(material_list) else [{}] * self._surface_count
        if len(material_list) != self._surface_count:
            raise ValueError("ParametricMultiSegmentBoundary: constraints and material_list "
                             "must have the same size.")
        self.surfaces = [
            ParametricSegmentBoundary(
                zero_distribution, one_distribution, flip_norm=self.flip_norm[i],
                initial_parameters=initial_parameters[i],
                parameters=None if parameters is None else parameters[i],
                material_dict=material_list[i])
            for i in range(self._surface_count)]
        self.constraints = list(constraints)
        for i, (surface, constraint) in enumerate(zip(self.surfaces, self.constraints)):
            if constraint.parent != "zero":
                surface.update_handles.append(constraint.make(i, self.surfaces))
            else:
                surface.update_handles.append(constraint.make(surface, None))
        super().__init__(**kwargs)

    @property
    def dimension(self):
        return 2

    def update_materials(self):
        pass

    def _update(self):
        from .engine import amalgamate
        self._fields = amalgamate(self.surfaces)

    def _generate_update_handles(self):
        handles = [d.update for d in (self._zero_distribution, self._one_distribution)
                   if hasattr(d, "update")]
        return handles + [surface.update for surface in self.surfaces]

    @property
    def surface_count(self):
        return self._surface_count

    @property
    def parameters(self):
        return [surface.parameters for surface in self.surfaces]
