"""Stand-in for the un-vendored `tfquaternion` dependency (test infrastructure only; see
tensorflow/__init__.py in this directory).  Only the two functions the reference calls
(`tfrt/sources.py:428-458`, `tfrt/distributions.py:2078`).  Quaternions are (w, x, y, z)."""
import torch


def get_rotation_quaternion_from_u_to_v(u, v, dtype=torch.float64, epsilon=1e-12):
    u = torch.as_tensor(u, dtype=dtype)
    v = torch.as_tensor(v, dtype=dtype)
    u = u / torch.linalg.norm(u)
    v = v / torch.linalg.norm(v)
    w = 1.0 + torch.dot(u, v)
    if float(w) < epsilon:       # antiparallel: any perpendicular axis, half turn
        axis = torch.linalg.cross(u, torch.tensor([0.0, 0.0, 1.0], dtype=dtype))
        if float(torch.linalg.norm(axis)) < epsilon:
            axis = torch.linalg.cross(u, torch.tensor([0.0, 1.0, 0.0], dtype=dtype))
        q = torch.cat([torch.zeros(1, dtype=dtype), axis])
    else:
        q = torch.cat([w.reshape(1), torch.linalg.cross(u, v)])
    return q / torch.linalg.norm(q)


def rotate_vector_by_quaternion(q, v, q_ndims=None, v_ndims=None):
    q = torch.as_tensor(q, dtype=torch.float64)
    v = torch.as_tensor(v, dtype=torch.float64)
    w, r = q[..., :1], q[..., 1:]
    r = r.expand_as(v) if r.dim() < v.dim() or r.shape != v.shape else r
    t = 2.0 * torch.linalg.cross(r, v, dim=-1)
    return v + w * t + torch.linalg.cross(r, t, dim=-1)
