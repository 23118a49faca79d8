"""The search kernel's arithmetic, restated in NumPy (CPU): the projected plane records written by
`k_pack_triangles` (csrc/pack.cu) and the single-condition inside test of `k_intersect3d`
(csrc/intersect3d.cu):

    det = D[k] + m1 D[k1] + m2 D[k2],  t = (dd - O[k] - m1 O[k1] - m2 O[k2]) / det,
    P1, P2 = (O + t D)[k1, k2],  S = s.P + s0 = 2(u+v) - C,  E = e.P + e0 = u - v,
    inside  <=>  |E| + |S - |E||  <=  C          (C = 1 + 3 size_eps)

must agree with the reference's formulation (Cramer solve + three inequalities,
tfrt/geometry.py:275-320, tfrt/engine.py:1138-1141) evaluated by the float64 oracle: same t, same
inside / outside decision for every pair that is not within rounding of an edge.  This pins the
algebra (dominant-axis elimination, the two min-folds) on the CPU; the CUDA code itself is checked by
the GPU parity suite."""

import numpy as np
import torch

from oracle import geometry as og

EPS = 1e-5


def pack(tri, size_eps, dtype):
    """(T, 9) corners -> records (k, m1, m2, dd, s1, s2, s0, e1, e2, e0, nk), coefficients rounded
    to `dtype` as the kernel stores them (computed in float64 like pack.cu)."""
    A, B, Cc = tri[:, 0:3], tri[:, 3:6], tri[:, 6:9]
    e1, e2 = B - A, Cc - A
    N = np.cross(e1, e2)
    inv = 1.0 / np.einsum("ij,ij->i", N, N)
    av = np.cross(e2, N) * inv[:, None]
    bv = np.cross(N, e1) * inv[:, None]
    ad = -np.einsum("ij,ij->i", av, A) + size_eps
    bd = -np.einsum("ij,ij->i", bv, A) + size_eps
    k = np.argmax(np.abs(N), axis=1)
    k1, k2 = (k + 1) % 3, (k + 2) % 3
    r = np.arange(tri.shape[0])
    m1, m2 = N[r, k1] / N[r, k], N[r, k2] / N[r, k]
    dd = A[r, k] + m1 * A[r, k1] + m2 * A[r, k2]
    u1, u2, u0 = av[r, k1] - av[r, k] * m1, av[r, k2] - av[r, k] * m2, ad + av[r, k] * dd
    v1, v2, v0 = bv[r, k1] - bv[r, k] * m1, bv[r, k2] - bv[r, k] * m2, bd + bv[r, k] * dd
    C = 1.0 + 3.0 * size_eps
    rec = dict(m1=m1, m2=m2, dd=dd, s1=2 * (u1 + v1), s2=2 * (u2 + v2), s0=2 * (u0 + v0) - C,
               e1=u1 - v1, e2=u2 - v2, e0=u0 - v0, nk=N[r, k])
    rec = {key: val.astype(dtype) for key, val in rec.items()}
    rec["k"] = k
    return rec


def kernel_test(O, D, rec, size_eps, dtype):
    """All rays x all triangles with the kernel's operation order in `dtype`: (t, inside, det)."""
    O, D = O.astype(dtype), D.astype(dtype)
    k = rec["k"]
    k1, k2 = (k + 1) % 3, (k + 2) % 3
    Ok, Ok1, Ok2 = O[:, k].T, O[:, k1].T, O[:, k2].T              # (T, N)
    Dk, Dk1, Dk2 = D[:, k].T, D[:, k1].T, D[:, k2].T
    col = lambda name: rec[name][:, None]                          # noqa: E731
    det = col("m2") * Dk2 + (col("m1") * Dk1 + Dk)
    tq = (col("dd") - Ok) - col("m1") * Ok1 - col("m2") * Ok2
    with np.errstate(divide="ignore", invalid="ignore"):
        t = tq * (dtype(1.0) / det)
    p1, p2 = Ok1 + t * Dk1, Ok2 + t * Dk2
    S = col("s2") * p2 + (col("s1") * p1 + col("s0"))
    E = col("e2") * p2 + (col("e1") * p1 + col("e0"))
    x = np.abs(E) + np.abs(S - np.abs(E))
    return t, x <= dtype(1.0 + 3.0 * size_eps), det * col("nk")


def scene(seed, n_rays=300, n_tri=120):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n_tri, 3))
    tri = np.concatenate([A, A + rng.normal(size=(n_tri, 3)) * 0.6,
                          A + rng.normal(size=(n_tri, 3)) * 0.6], axis=1)
    tri[:8, [0, 3, 6]] = tri[:8, [0]]                 # axis-aligned planes (m1 = m2 = 0)
    tri[8:16, [1, 4, 7]] = tri[8:16, [1]]
    tri[16:24, [2, 5, 8]] = tri[16:24, [2]]
    O = rng.normal(size=(n_rays, 3)) * 2.0
    D = rng.normal(size=(n_rays, 3))
    return tri, O, D


def reference_test(tri, O, D):
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))      # noqa: E731
    E = O + D
    args = [T(O[:, i]) for i in range(3)] + [T(E[:, i]) for i in range(3)] + [T(tri[:, i]) for i in range(9)]
    x, y, z, valid, ray_u, trig_u, trig_v = [v.numpy() for v in og.line_triangle_intersect(*args, 1e-10)]
    return ray_u, trig_u, trig_v, valid


def test_projected_form_equals_the_reference_formulation_in_float64():
    for seed in range(4):
        tri, O, D = scene(seed)
        ray_u, u, v, valid = reference_test(tri, O, D)
        t, inside, den = kernel_test(O, D, pack(tri, EPS, np.float64), EPS, np.float64)
        assert valid.all()
        assert np.allclose(t, ray_u, rtol=1e-9, atol=1e-9)
        want = (u >= -EPS) & (v >= -EPS) & (u + v <= 1 + EPS)
        margin = np.minimum(np.minimum(u + EPS, v + EPS), 1 + EPS - u - v)
        clear = np.abs(margin) > 1e-9
        assert np.array_equal(inside[clear], want[clear])
        assert want.sum() > 50                                     # the scenes do contain hits
        # det * n[k] is the reference's denominator up to sign (|den| >= intersect_epsilion)
        e1, e2 = tri[:, 3:6] - tri[:, 0:3], tri[:, 6:9] - tri[:, 0:3]
        n_dot_d = np.cross(e1, e2) @ D.T
        assert np.allclose(np.abs(den), np.abs(n_dot_d), rtol=1e-9, atol=1e-12)


def test_projected_form_in_float32_decides_like_the_reference_away_from_edges():
    worst_t, checked = 0.0, 0
    for seed in range(4):
        tri, O, D = scene(100 + seed)
        ray_u, u, v, _ = reference_test(tri.astype(np.float32).astype(np.float64),
                                        O.astype(np.float32).astype(np.float64),
                                        D.astype(np.float32).astype(np.float64))
        t, inside, _ = kernel_test(O, D, pack(tri.astype(np.float32).astype(np.float64), EPS, np.float32),
                                   EPS, np.float32)
        want = (u >= -EPS) & (v >= -EPS) & (u + v <= 1 + EPS)
        margin = np.minimum(np.minimum(u + EPS, v + EPS), 1 + EPS - u - v)
        # fp32 rounding of the barycentrics grows with |P| / triangle size: stay 1e-3 clear of edges
        scale = np.maximum(np.abs(u), np.abs(v)) + 1.0
        clear = (np.abs(margin) > 2e-3 * scale) & (np.abs(ray_u) < 1e3)
        assert np.array_equal(inside[clear], want[clear])
        hit = want & clear & (ray_u > 1e-3)
        # t' cancels against the ray origin (|O| ~ 2 here): absolute error ~ 1e-6, so scale by max(|t|, 1)
        worst_t = max(worst_t, float(np.max(np.abs(t[hit] - ray_u[hit]) / np.maximum(np.abs(ray_u[hit]), 1.0))))
        checked += int(hit.sum())
    assert checked > 150 and worst_t < 5e-6, (checked, worst_t)
