"""Stand-in for `pyvista` (test infrastructure only; see tensorflow/__init__.py in this directory):
just the PolyData container the reference's triangle boundaries read (`.points (V,3)`, `.faces` in the
packed `[3, i, j, k, 3, ...]` layout, `.copy()`), `tfrt/boundaries.py:840-940, 1010-1030`."""
import numpy as np


class PolyData:
    def __init__(self, points=None, faces=None):
        if isinstance(points, PolyData):
            faces = points.faces.copy()
            points = points.points.copy()
        self.points = None if points is None else np.array(points, dtype=np.float64)
        self.faces = None if faces is None else np.array(faces, dtype=np.int64).reshape(-1)

    def copy(self):
        return PolyData(self.points.copy(), self.faces.copy())

    def save(self, *a, **k):
        raise NotImplementedError


def read(name):
    raise NotImplementedError("pyvista stub: no mesh I/O")
