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
        self._points = as_field(val)


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
        self._point = as_field(val)


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
        self._vector = as_field(val)


class FromAxisVG(VectorGeneratorBase):
    """Vectors project perpendicular to an axis, through the points.  The axis is given by a
    point plus a kwarg `point` (second point) or `direction`."""

    def __init__(self, first, **kwargs):
        self.axis_point = as_field(first)
        if "point" in kwargs:
            axis_vector = as_field(kwargs["point"]) - self.axis_point
        elif "direction" in kwargs:
            axis_vector = as_field(kwargs["direction"])
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
        if self._mesh is not None:
            self._vertices = as_field(self._mesh.points)
            self._faces = torch.from_numpy(unpack_faces(self._mesh.faces)).to(
                dtype=torch.int32, device=self._vertices.device)

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

    def has_geometric_norm(self):
        """True while the `norm` field is the one computed from the vertices (not overridden)."""
        return self._fields.get("norm") is getattr(self, "_geometric_norm", None) is not None

    def update_from_mesh(self):
        if self._mesh is not None:
            self.update_vertices_from_mesh()
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
        mesh.faces = np.ascontiguousarray(unpack_faces(mesh.faces)[:, ::-1])
        return mesh

    def update_vertices_from_mesh(self):
        raise RuntimeError(
            "ParametricTriangleBoundary: update_vertices_from_mesh is disabled for parametric "
            "boundaries.")

    def update_from_mesh(self):
        raise RuntimeError(
            "ParametricTriangleBoundary: update_from_mesh is disabled for parametric boundaries.")
