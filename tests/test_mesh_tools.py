"""mesh_tools graph helpers vs a literal set-based restatement of the reference's algorithms
(tfrt/mesh_tools.py:221-331, 345-421, 460-505).  CPU only."""

import numpy as np
import pytest

from tensorflowraytrace_b200 import mesh_tools as mt


class _Mesh:
    def __init__(self, points, faces):
        self.points, self.faces = points, faces


def _neighbours(p, face_sets):
    out = set()
    for f in face_sets:
        if p in f:
            out |= f
    out.discard(p)
    return out


def _reference_parametrization(faces, n, top):
    face_sets = [set(f) for f in faces]
    movable = [set()] * len(face_sets)
    to_visit, active, last = set(range(len(face_sets))), {top}, set()
    avail, parents, anc, missed = set(range(n)), [set()] * n, [set()] * n, set(range(n))
    while to_visit:
        nxt, just = set(), set()
        avail -= active
        for f in to_visit:
            for v in active:
                if v in face_sets[f]:
                    mv = face_sets[f] & avail
                    nxt |= mv
                    movable[f] = mv
                    just.add(f)
        for v in active:
            missed.remove(v)
            parents[v] = _neighbours(v, face_sets) & last
            anc[v] = parents[v].copy()
            for p in parents[v]:
                anc[v] |= anc[p]
        to_visit -= just
        last, active = active, nxt
    for v in missed:
        parents[v] = _neighbours(v, face_sets) - missed
        anc[v] = parents[v].copy()
        for p in parents[v]:
            anc[v] |= anc[p]
    umap = np.array([[v in movable[f] for v in faces[f]] for f in range(len(faces))])
    umap[~umap.any(1)] = True
    acc = np.array([[1.0 if j in anc[i] else 0.0 for j in range(n)] for i in range(n)]) + np.eye(n)
    return umap, acc


@pytest.mark.parametrize("rings,top_at", [(1, (0, 0, 0)), (4, (0, 0, 0)), (5, (0.7, 0.2, 0))])
def test_parametrization_tools_match_the_reference_algorithm(rings, top_at):
    pts, faces = mt.hexagonal_mesh(1.0, rings)
    mesh = _Mesh(pts, faces)
    top = mt.get_closest_point(mesh, top_at)
    umap, acc = mt.mesh_parametrization_tools(mesh, top, dense=True)
    want_map, want_acc = _reference_parametrization(faces, pts.shape[0], top)
    assert np.array_equal(umap, want_map)
    assert np.array_equal(acc, want_acc)
    keep = list(range(0, pts.shape[0], 2))
    _, sub = mt.mesh_parametrization_tools(mesh, top, active_vertices=keep, dense=True)
    assert np.array_equal(sub, want_acc[np.ix_(keep, keep)])


def test_smoothing_tool_matches_the_reference_algorithm():
    pts, faces = mt.cylindrical_mesh((0, 0, 0), (2, 0, 0), 0.5, 8, 5)
    mesh = _Mesh(pts, faces)
    weights = mt.gaussian_weights(1.2, 4)
    got = mt.mesh_smoothing_tool(mesh, weights, dense=True)
    face_sets = [set(f) for f in faces]
    n = pts.shape[0]
    w = weights / weights.sum()
    want = np.zeros((n, n))
    for p in range(n):
        nth, taken, cur = [{p}], {p}, {p}
        for _ in range(1, len(w)):
            new = set()
            for e in cur:
                new |= _neighbours(e, face_sets)
            new -= taken
            nth.append(new)
            cur = new
            taken |= new
        for o in range(len(w)):
            if nth[o]:
                want[p, list(nth[o])] = w[o] / len(nth[o])
    assert np.allclose(got, want)
    assert np.allclose(got.sum(axis=1)[:3], 1.0)


def test_flat_initial_and_closest_point():
    pts, faces = mt.hexagonal_mesh(1.0, 2)
    pts = pts.copy()
    pts[:, 2] = 0.3 * pts[:, 0]
    mesh = _Mesh(pts, faces)
    init = mt.get_flat_initial(mesh, axis=2)
    assert np.all(mesh.points[:, 2] == 0.0) and np.allclose(init, 0.3 * pts[:, 0])
    assert mt.get_closest_point(mesh, (0, 0, 0)) == int(np.argmin((mesh.points ** 2).sum(1)))
    with pytest.raises(ValueError):
        mt.get_flat_initial(mesh, axis=5)
