"""Diagnostic: rays aimed at vertices / edges of the hex91 lens surface; report rays whose chosen
triangle does not touch the aimed point, with fp64 barycentrics of the chosen and the best triangle."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tensorflowraytrace_b200 import scenes, engine as eng

step_count = int(sys.argv[1]) if len(sys.argv) > 1 else 91
rng = np.random.default_rng(step_count)
surf = scenes.lens_surfaces(step_count)[0]
verts = scenes.surface_vertices(surf)
faces = surf["faces"]
r = np.hypot(verts[:, 1], verts[:, 2])
inner = np.nonzero(r < 0.9)[0]
pick = rng.choice(inner, size=min(3000, inner.size), replace=False)
targets = [verts[pick]]
f = faces[rng.choice(faces.shape[0], size=min(3000, faces.shape[0]), replace=False)]
f = f[(r[f] < 0.9).all(axis=1)]
for a, b in ((0, 1), (1, 2), (2, 0)):
    lam = rng.uniform(0.05, 0.95, size=(f.shape[0], 1))
    targets.append(lam * verts[f[:, a]] + (1 - lam) * verts[f[:, b]])
aim = np.concatenate(targets)
start = aim + np.stack([np.full(aim.shape[0], -3.0), rng.uniform(-0.3, 0.3, aim.shape[0]),
                        rng.uniform(-0.3, 0.3, aim.shape[0])], axis=1)
dev = torch.device("cuda:0")
for perm in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
    V = verts[:, perm].astype(np.float32)
    S, A = start[:, perm].astype(np.float32), aim[:, perm].astype(np.float32)
    E = (S + (A - S) * np.float32(0.25)).astype(np.float32)
    tri = [torch.tensor(np.ascontiguousarray(V[faces[:, c], k]), device=dev) for c in range(3) for k in range(3)]
    rays = [torch.tensor(np.ascontiguousarray(S[:, k]), device=dev) for k in range(3)] + \
           [torch.tensor(np.ascontiguousarray(E[:, k]), device=dev) for k in range(3)]
    out = eng.OpticalSystem3D._intersection(*rays, *tri, 1e-10, 1e-5, 1e-4)
    valid, g = out[3].cpu().numpy(), out[8].cpu().numpy()
    corners = V[faces[g]].astype(np.float64)
    lo, hi = corners.min(axis=1) - 2e-5, corners.max(axis=1) + 2e-5
    bad = np.nonzero(~((A >= lo) & (A <= hi)).all(axis=1))[0]
    print("perm", perm, "invalid", int((~valid).sum()), "not touching", bad.size)
    for i in bad[:6]:
        # fp64 barycentrics of the hit on the chosen triangle
        P0, P1, P2 = corners[i]
        s, e = S[i].astype(np.float64), E[i].astype(np.float64)
        d = e - s
        e1, e2 = P1 - P0, P2 - P0
        n = np.cross(e1, e2)
        t = np.dot(n, P0 - s) / np.dot(n, d)
        P = s + t * d
        nn = np.dot(n, n)
        u = np.dot(np.cross(e2, n), P - P0) / nn
        v = np.dot(np.cross(n, e1), P - P0) / nn
        # brute-force fp64 over all triangles
        C = V[faces].astype(np.float64)
        E1, E2 = C[:, 1] - C[:, 0], C[:, 2] - C[:, 0]
        N = np.cross(E1, E2)
        T = np.einsum("ij,ij->i", N, C[:, 0] - s) / (N @ d)
        PP = s + T[:, None] * d
        NN = np.einsum("ij,ij->i", N, N)
        U = np.einsum("ij,ij->i", np.cross(E2, N), PP - C[:, 0]) / NN
        W = np.einsum("ij,ij->i", np.cross(N, E1), PP - C[:, 0]) / NN
        ok = (U >= -1e-5) & (W >= -1e-5) & (U + W <= 1 + 1e-5) & (T >= 1e-4)
        best = np.where(ok, T, np.inf).argmin()
        print(f"  ray {i}: chosen tri {g[i]} t={t:.7f} u={u:.3e} v={v:.3e} w={1-u-v:.3e} |P-A|={np.abs(P-A[i]).max():.2e}; "
              f"fp64 best tri {best} t={T[best]:.7f} (u={U[best]:.3e}, v={W[best]:.3e}); candidates {np.nonzero(ok)[0].tolist()}")
