"""
Triangle-mesh generators (NumPy, float64) for the synthetic scenes.

These play the role of the pure-NumPy parts of the reference's tfrt/mesh_tools.py generators
(`hexagonal_mesh` tfrt/mesh_tools.py:713-795: 6*S^2 faces, 1+3S(S+1) vertices;
`cylindrical_mesh` :800-952: 2*theta*(z-1) side faces + theta faces per cap), without pyvista:
a mesh is a plain `(points (V,3) float64, faces (F,3) int64)` pair.  They are written from the
lattice description, not from the reference's layer-weaving loops, so vertex/face *ordering*
differs from the reference; the scenes are synthetic inputs shared bit-identically by the
oracle and the CUDA path, so ordering only has to be deterministic.
"""

import numpy as np


def hexagonal_mesh(radius=1.0, step_count=10):
    """Regular hexagon of equilateral triangles in the z=0 plane, normals +z.

    Lattice points p = i*a1 + j*a2 (axial coordinates, |i|,|j|,|i+j| <= S); every unit lattice
    cell contributes an "up" and a "down" triangle when all corners lie in the hexagon.
    """
    S = int(step_count)
    h = radius / S
    ii, jj = np.meshgrid(np.arange(-S, S + 1), np.arange(-S, S + 1), indexing="ij")
    inside = (np.abs(ii) <= S) & (np.abs(jj) <= S) & (np.abs(ii + jj) <= S)
    index = -np.ones(inside.shape, dtype=np.int64)
    index[inside] = np.arange(inside.sum())
    pi_, pj_ = ii[inside].astype(np.float64), jj[inside].astype(np.float64)
    points = np.stack([h * (pi_ + 0.5 * pj_), h * (np.sqrt(3.0) / 2.0) * pj_,
                       np.zeros_like(pi_)], axis=1)

    def idx(di, dj):
        out = -np.ones(inside.shape, dtype=np.int64)
        src = index[max(di, 0):index.shape[0] + min(di, 0), max(dj, 0):index.shape[1] + min(dj, 0)]
        out[max(-di, 0):index.shape[0] + min(-di, 0), max(-dj, 0):index.shape[1] + min(-dj, 0)] = src
        return out

    a, b, c, d = index, idx(1, 0), idx(0, 1), idx(1, 1)
    up = np.stack([a, b, c], axis=-1).reshape(-1, 3)       # (i,j) (i+1,j) (i,j+1): CCW
    down = np.stack([b, d, c], axis=-1).reshape(-1, 3)     # (i+1,j) (i+1,j+1) (i,j+1): CCW
    faces = np.concatenate([up, down], axis=0)
    faces = faces[(faces >= 0).all(axis=1)]
    return points, faces.astype(np.int64)


def circular_cap_sag(points, curvature_radius):
    """Sag of a spherical cap of the given curvature radius over in-plane points (x, y)."""
    r2 = points[:, 0] ** 2 + points[:, 1] ** 2
    return curvature_radius - np.sqrt(np.maximum(curvature_radius ** 2 - r2, 0.0))


def cylindrical_mesh(start, end, radius=1.0, theta_res=6, z_res=8, start_cap=True, end_cap=True):
    """Closed (optionally capped) cylinder between two axis points, normals pointing outward."""
    start = np.asarray(start, dtype=np.float64).reshape(3)
    end = np.asarray(end, dtype=np.float64).reshape(3)
    axis = end - start
    u = np.cross(axis, (1.0, 0.0, 0.0))
    if np.linalg.norm(u) < 1e-6:
        u = np.cross(axis, (0.0, 1.0, 0.0))
    if np.linalg.norm(u) < 1e-6:
        raise ValueError("cylindrical_mesh: could not find vectors perpendicular to axis.")
    u = u * radius / np.linalg.norm(u)
    v = np.cross(axis, u)
    v = v * radius / np.linalg.norm(v)

    theta = np.linspace(0.0, 2.0 * np.pi, theta_res + 1)[:-1]
    z = np.linspace(0.0, 1.0, z_res)
    ring = (np.cos(theta)[:, None] * u[None, :] + np.sin(theta)[:, None] * v[None, :])
    side = start[None, None, :] + z[:, None, None] * axis[None, None, :] + ring[None, :, :]
    points = [side.reshape(-1, 3)]
    n_side = z_res * theta_res

    t = np.arange(theta_res)
    tn = (t + 1) % theta_res
    faces = []
    for k in range(1, z_res):
        lo, hi = (k - 1) * theta_res, k * theta_res
        faces.append(np.stack([lo + tn, hi + t, lo + t], axis=1))
        faces.append(np.stack([hi + t, lo + tn, hi + tn], axis=1))
    next_vertex = n_side
    if start_cap:
        points.append(start[None, :])
        faces.append(np.stack([t, np.full_like(t, next_vertex), tn], axis=1))
        next_vertex += 1
    if end_cap:
        points.append(end[None, :])
        top = (z_res - 1) * theta_res
        faces.append(np.stack([top + tn, np.full_like(t, next_vertex), top + t], axis=1))
        next_vertex += 1
    points = np.concatenate(points, axis=0)
    faces = np.concatenate(faces, axis=0).astype(np.int64)

    # orient every face outward (away from the axis mid-point) so "outside" is unambiguous
    p0, p1, p2 = points[faces[:, 0]], points[faces[:, 1]], points[faces[:, 2]]
    nrm = np.cross(p1 - p0, p2 - p1)
    centroid = (p0 + p1 + p2) / 3.0
    rel = centroid - (start + 0.5 * axis)[None, :]
    # remove the axial component for side faces so the test is radial there
    flip = np.einsum("ij,ij->i", nrm, rel) < 0
    faces[flip] = faces[flip][:, ::-1]
    return points, faces


def plane_mesh(center, half_size, axis=0):
    """Two-triangle square plane perpendicular to `axis`, normal along +axis."""
    c = np.asarray(center, dtype=np.float64)
    a, b = [k for k in range(3) if k != axis]
    corners = np.tile(c, (4, 1))
    for n, (sa, sb) in enumerate(((-1, -1), (1, -1), (1, 1), (-1, 1))):
        corners[n, a] += sa * half_size
        corners[n, b] += sb * half_size
    faces = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int64)
    p = corners
    nrm = np.cross(p[1] - p[0], p[2] - p[1])
    if nrm[axis] < 0:
        faces = faces[:, ::-1].copy()
    return corners, faces
