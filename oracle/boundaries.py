"""
Oracle restatement of the vertex -> triangle-field map of tfrt/boundaries.py (float64 torch).
TEST INFRASTRUCTURE ONLY.
"""

import torch


def parametric_vertices(zero, vectors, parameter):
    """ParametricTriangleBoundary._update_internal, tfrt/boundaries.py:1081-1083."""
    return zero + parameter.reshape(-1, 1) * vectors


def fields_from_vertices(vertices, faces, vertex_update_map=None):
    """TriangleBoundaryBase.update_fields_from_vertices, tfrt/boundaries.py:890-923.

    vertices : (V, 3) float64.  faces : (F, 3) or pyvista-packed (F, 4) [3, i, j, k] int.
    vertex_update_map : optional (F, 3) bool; False entries receive no gradient
    (tf.stop_gradient, boundaries.py:900-913).
    Returns the field dict xp..z2 (F,) and norm (F, 3).
    """
    faces = torch.as_tensor(faces, dtype=torch.int64)
    if faces.shape[1] == 4:
        faces = faces[:, 1:]
    first_points = vertices[faces[:, 0]]
    second_points = vertices[faces[:, 1]]
    third_points = vertices[faces[:, 2]]
    if vertex_update_map is not None:
        m = torch.as_tensor(vertex_update_map, dtype=torch.bool)
        first_points = torch.where(m[:, 0:1], first_points, first_points.detach())
        second_points = torch.where(m[:, 1:2], second_points, second_points.detach())
        third_points = torch.where(m[:, 2:3], third_points, third_points.detach())
    out = {}
    out["xp"], out["yp"], out["zp"] = torch.unbind(first_points, dim=1)
    out["x1"], out["y1"], out["z1"] = torch.unbind(second_points, dim=1)
    out["x2"], out["y2"], out["z2"] = torch.unbind(third_points, dim=1)
    cross = torch.linalg.cross(second_points - first_points, third_points - second_points, dim=1)
    # tf.linalg.normalize(..., axis=1)[0] == x / ||x||
    out["norm"] = cross / torch.linalg.norm(cross, dim=1, keepdim=True)
    return out
