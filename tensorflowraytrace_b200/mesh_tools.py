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
    """Regular hexagon of equilateral triangles in the z = 0 plane (corner distance `radius`,
    `step_count` rings: 1 + 3 S (S + 1) vertices, 6 S^2 faces), with the reference's exact vertex and
    face order (tfrt/mesh_tools.py:713-795): vertex 0 is the centre, ring r holds 6 r vertices
    starting at the corner on +x and running counter-clockwise; the faces of a ring are woven
    edge by edge, alternating between a triangle with two outer vertices and one with two inner
    ones.  Returns (points (V, 3), faces (F, 3) int64)."""
    S = int(step_count)
    radii = np.linspace(0.0, radius, S + 1)
    corner = np.stack([np.cos(np.pi / 3.0 * np.arange(7)), np.sin(np.pi / 3.0 * np.arange(7)),
                       np.zeros(7)], axis=1)
    points = [np.zeros((1, 3))]
    faces = []
    inner_base, inner_count = 0, 1          # the previous ring (the centre counts as a ring of one)
    outer_base = 1
    for r in range(1, S + 1):
        # ring vertices: r per hexagon edge, the edge's end corner belongs to the next edge
        edges = [np.linspace(radii[r] * corner[e], radii[r] * corner[e + 1], r + 1)[:-1]
                 for e in range(6)]
        points.append(np.concatenate(edges, axis=0))
        outer_count = 6 * r
        po = pi = 0                         # positions on the outer / inner ring (cyclic)
        first = outer_base + po; po += 1
        second = inner_base + pi % inner_count; pi += 1
        kept = second
        for _ in range(6):
            take_outer = True
            for _ in range(2 * r - 1):
                if take_outer:
                    third = outer_base + po % outer_count; po += 1
                    faces.append((third, second, first))
                else:
                    third = inner_base + pi % inner_count; pi += 1
                    faces.append((first, second, third))
                first, kept, second = second, second, third
                take_outer = not take_outer
            first, second = third, kept
        inner_base, inner_count = outer_base, outer_count
        outer_base += outer_count
    return np.concatenate(points, axis=0), np.array(faces, dtype=np.int64).reshape(-1, 3)


def circular_cap_sag(points, curvature_radius):
    """Sag of a spherical cap of the given curvature radius over in-plane points (x, y)."""
    r2 = points[:, 0] ** 2 + points[:, 1] ** 2
    return curvature_radius - np.sqrt(np.maximum(curvature_radius ** 2 - r2, 0.0))


def cylindrical_mesh(start, end, radius=1.0, theta_res=6, z_res=8, start_cap=True, end_cap=True,
                     use_twist=False, epsilion=1e-6):
    """Cylinder between two axis points with optional caps, with the reference's exact vertex and
    face order (tfrt/mesh_tools.py:800-952): vertices = [start point if start_cap] + ring 0 ..
    ring z_res-1 + [end point if end_cap]; faces = start-cap fan, then per ring and per theta the
    two triangles of the quad, then the end-cap fan; windings as the reference writes them (the
    normals (p2-p1) x (p3-p2) point outward)."""
    start = np.asarray(start, dtype=np.float64).reshape(3)
    end = np.asarray(end, dtype=np.float64).reshape(3)
    axis = end - start
    u = np.cross(axis, (1.0, 0.0, 0.0))
    if np.linalg.norm(u) < epsilion:
        u = np.cross(axis, (0.0, 1.0, 0.0))
    if np.linalg.norm(u) < epsilion:
        raise ValueError("cylindrical_mesh: could not find vectors perpendicular to axis.  Try "
                         "decreasing epsilion?")
    u = u * radius / np.linalg.norm(u)
    v = np.cross(axis, u)
    v = v * radius / np.linalg.norm(v)

    theta, z = np.meshgrid(np.linspace(0.0, 2.0 * np.pi, theta_res + 1)[:-1],
                           np.linspace(0.0, 1.0, z_res))
    if use_twist:                       # half a triangle of twist per layer
        theta = theta + (np.pi / theta_res * np.arange(z_res)).reshape(-1, 1)
    side = (start + z[..., None] * axis + np.cos(theta)[..., None] * u
            + np.sin(theta)[..., None] * v)                                  # (z_res, theta_res, 3)

    off = 1 if start_cap else 0
    points = ([start[None, :]] if start_cap else []) + [side.reshape(-1, 3)] + \
             ([end[None, :]] if end_cap else [])
    points = np.concatenate(points, axis=0)

    t = np.arange(theta_res)
    tn = (t + 1) % theta_res
    faces = []
    if start_cap:
        faces.append(np.stack([t + 1, np.zeros_like(t), tn + 1], axis=1))
    for k in range(1, z_res):
        lo, hi = (k - 1) * theta_res + off, k * theta_res + off
        quad = np.stack([np.stack([lo + tn, hi + t, lo + t], axis=1),
                         np.stack([hi + t, lo + tn, hi + tn], axis=1)], axis=1)   # (theta, 2, 3)
        faces.append(quad.reshape(-1, 3))
    if end_cap:
        last = points.shape[0] - 1
        top = (z_res - 1) * theta_res + off
        faces.append(np.stack([tn + top, np.full_like(t, last), t + top], axis=1))
    faces = np.concatenate(faces, axis=0).astype(np.int64)
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


# =============================================================================================
# Graph tools for parametric meshes (tfrt/mesh_tools.py:221-331, 345-421): which corners a face
# may move (`vertex_update_map`), the gradient accumulator and the parameter smoother.  The
# reference builds dense V x V float64 matrices with O(V^2) Python set scans (5 GB at config 5's
# 25 117 vertices); here everything is vectorised on a CSR vertex adjacency and the matrices are
# scipy.sparse CSR (SGD_Optimizer and `smooth` accept them), identical in value.

def _faces_of(mesh_or_faces):
    faces = getattr(mesh_or_faces, "faces", mesh_or_faces)
    f = np.asarray(faces).astype(np.int64)
    if f.ndim == 1:
        f = f.reshape(-1, 4)
    if f.shape[1] == 4:
        f = f[:, 1:]
    return f


def _points_of(mesh_or_points):
    return np.asarray(getattr(mesh_or_points, "points", mesh_or_points), dtype=np.float64)


def vertex_adjacency(faces, n_vertices):
    """Symmetric boolean CSR adjacency: two vertices are neighbours when they share a face
    (the reference's neighbors_from_faces, tfrt/mesh_tools.py:119-129)."""
    import scipy.sparse as sp
    i = np.concatenate([faces[:, 0], faces[:, 1], faces[:, 2], faces[:, 1], faces[:, 2], faces[:, 0]])
    j = np.concatenate([faces[:, 1], faces[:, 2], faces[:, 0], faces[:, 0], faces[:, 1], faces[:, 2]])
    a = sp.csr_matrix((np.ones(i.shape[0], dtype=np.int8), (i, j)), shape=(n_vertices, n_vertices))
    a.data[:] = 1
    return a


def get_closest_point(mesh, point):
    """Index of the mesh vertex nearest to `point` (tfrt/mesh_tools.py `get_closest_point`)."""
    pts = _points_of(mesh)
    return int(np.argmin(np.sum((pts - np.asarray(point, dtype=np.float64)) ** 2, axis=1)))


def vertex_rings(mesh, top_parent):
    """Breadth-first ring index of every vertex from `top_parent` over the face adjacency
    (-1 for vertices that cannot be reached): the wave front the reference advances as
    `active_edge` (tfrt/mesh_tools.py:241-272)."""
    faces = _faces_of(mesh)
    n = _points_of(mesh).shape[0]
    adj = vertex_adjacency(faces, n)
    ring = np.full(n, -1, dtype=np.int64)
    frontier = np.array([top_parent], dtype=np.int64)
    ring[frontier] = 0
    k = 0
    while frontier.size:
        k += 1
        reach = np.unique(adj[frontier].indices)
        frontier = reach[ring[reach] < 0]
        ring[frontier] = k
    return ring, adj


def mesh_parametrization_tools(mesh, top_parent, active_vertices=None, dense=False):
    """vertex_update_map (F,3) bool and the gradient accumulator (V,V)
    (tfrt/mesh_tools.py:221-331).

    A face is visited when the wave front from `top_parent` first touches it and may move only
    its vertices that lie ahead of the front (ring > the face's lowest ring); faces whose three
    vertices share one ring would move nothing and are allowed all three ("orphaned faces",
    tfrt/mesh_tools.py:471-480).  accumulator[i, j] = 1 when j == i or j is an ancestor of i,
    a vertex's parents being its neighbours one ring closer to `top_parent`; left-multiplied
    onto a gradient it moves a vertex together with everything downstream of it.
    Returns a scipy.sparse CSR accumulator unless dense=True."""
    import scipy.sparse as sp
    faces = _faces_of(mesh)
    n = _points_of(mesh).shape[0]
    ring, adj = vertex_rings(mesh, top_parent)
    missed = ring < 0
    fring = np.where(ring[faces] < 0, np.iinfo(np.int64).max, ring[faces])
    update_map = fring > fring.min(axis=1, keepdims=True)
    update_map[~update_map.any(axis=1)] = True

    # parents: neighbours one ring closer; unreachable vertices take their reachable neighbours
    coo = adj.tocoo()
    is_parent = (ring[coo.col] == ring[coo.row] - 1) & ~missed[coo.row] & ~missed[coo.col]
    is_parent |= missed[coo.row] & ~missed[coo.col]
    parents = sp.csr_matrix((np.ones(int(is_parent.sum()), dtype=np.float64),
                             (coo.row[is_parent], coo.col[is_parent])), shape=(n, n))
    # ancestors ring by ring: rows of ring k are final once rings < k are
    order = [np.nonzero(ring == k)[0] for k in range(1, int(ring.max()) + 1)] + [np.nonzero(missed)[0]]
    closure = sp.csr_matrix((n, n), dtype=np.float64)
    eye = sp.identity(n, dtype=np.float64, format="csr")
    for rows in order:
        if rows.size == 0:
            continue
        block = parents[rows] @ (eye + closure)
        block.data[:] = 1.0
        sel = sp.csr_matrix((np.ones(rows.size), (rows, np.arange(rows.size))), shape=(n, rows.size))
        closure = closure + sel @ block
    accumulator = (closure + eye).tocsr()
    if active_vertices is not None:
        keep = np.array(sorted(set(int(v) for v in active_vertices)), dtype=np.int64)
        accumulator = accumulator[keep][:, keep]
    return update_map, (accumulator.toarray() if dense else accumulator)


def gaussian_weights(sigma, count):
    """Un-normalised gaussian weights per neighbour order (tfrt/mesh_tools.py:335-341)."""
    x = np.arange(count) / sigma
    return np.exp(-0.5 * x ** 2)


def mesh_smoothing_tool(mesh, weights, active_vertices=None, dense=False):
    """Smoother (V,V): left-multiplied onto the parameters, every vertex keeps weights[0]/sum of
    itself and spreads weights[k]/sum evenly over its k-th order neighbours
    (tfrt/mesh_tools.py:345-421)."""
    import scipy.sparse as sp
    faces = _faces_of(mesh)
    n = _points_of(mesh).shape[0]
    adj = vertex_adjacency(faces, n).astype(np.float64)
    w = np.asarray(weights, dtype=np.float64)
    w = w / w.sum()
    eye = sp.identity(n, dtype=np.float64, format="csr")
    taken, current = eye.copy(), eye.copy()
    smoother = eye * w[0]
    for order in range(1, w.shape[0]):
        reach = current @ adj
        reach.data[:] = 1.0
        new = reach - reach.multiply(taken)          # not yet taken
        new.eliminate_zeros()
        counts = np.asarray(new.sum(axis=1)).ravel()
        scale = sp.diags(np.where(counts > 0, w[order] / np.maximum(counts, 1.0), 0.0))
        smoother = smoother + scale @ new
        taken = taken + new
        current = new
    smoother = smoother.tocsr()
    if active_vertices is not None:
        keep = np.array(sorted(set(int(v) for v in active_vertices)), dtype=np.int64)
        smoother = smoother[keep][:, keep]
    return smoother.toarray() if dense else smoother


def get_flat_initial(mesh, axis=0):
    """Flatten one axis of `mesh.points` in place and return the removed coordinates, to be used
    as initial parameters (tfrt/mesh_tools.py:425-456)."""
    if axis not in (0, 1, 2):
        raise ValueError("get_flat_initial: axis must be in {0, 1, 2}.")
    initial = mesh.points[:, axis].copy()
    mesh.points[:, axis] = 0.0
    return initial


def to_torch_sparse(matrix, device=None, dtype=None):
    """scipy.sparse / dense matrix -> torch sparse CSR-compatible COO tensor for
    SGD_Optimizer accumulators and smoothers."""
    import scipy.sparse as sp
    import torch
    dtype = dtype or torch.float32
    if sp.issparse(matrix):
        coo = matrix.tocoo()
        idx = torch.from_numpy(np.stack([coo.row, coo.col]).astype(np.int64))
        t = torch.sparse_coo_tensor(idx, torch.from_numpy(coo.data).to(dtype), coo.shape).coalesce()
    else:
        t = torch.as_tensor(np.asarray(matrix), dtype=dtype)
    return t.to(device) if device is not None else t
