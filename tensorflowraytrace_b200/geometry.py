"""
Per-ray (1-to-1) differentiable geometry in torch, used by the *unfused* engine path
(`OpticalEngine.single_pass` with arbitrary RayOperations): once the CUDA intersection kernel
has chosen the nearest boundary per ray, the O(rays) evaluation of the hit and of the reaction
is expressed in torch ops so that autograd reproduces the reference's gradient.

Reference functions covered (same argument meaning, vector implementation):
  raw_line_triangle_intersect   tfrt/geometry.py:275-320   -> ray_triangle_hit
  snells_law_3D                 tfrt/geometry.py:671-753
  snells_law_2D                 tfrt/geometry.py:565-653   (vector form, 2D mirror rule)
The O(rays x boundaries) searches themselves live in csrc/ (intersect3d.cu, trace2d.cu).
"""

import torch


def _dot(a, b):
    return torch.sum(a * b, dim=-1, keepdim=True)


def l2_normalize(x, eps=1e-12):
    """x * rsqrt(max(|x|^2, eps)) (tf.math.l2_normalize)."""
    return x * torch.rsqrt(torch.clamp(_dot(x, x), min=eps))


def ray_triangle_hit(start, end, p0, p1, p2, epsilion):
    """Moller-Trumbore solve for paired rays and triangles, all (N,3).  Returns
    (hit (N,3), valid (N,), ray_u, trig_u, trig_v (N,)) with the reference's conventions:
    ray_u in units of |end - start|, trig_u / trig_v the weights of p1 / p2, a unit safe
    denominator where |det| < epsilion (tfrt/geometry.py:300-318)."""
    d = end - start
    e1, e2 = p1 - p0, p2 - p0
    pvec = torch.linalg.cross(d, e2, dim=-1)
    det = _dot(e1, pvec)
    valid = torch.abs(det) >= epsilion
    inv = 1.0 / torch.where(valid, det, torch.ones_like(det))
    tvec = start - p0
    qvec = torch.linalg.cross(tvec, e1, dim=-1)
    trig_u = _dot(tvec, pvec) * inv
    trig_v = _dot(d, qvec) * inv
    ray_u = _dot(e2, qvec) * inv
    hit = start + ray_u * d
    return hit, valid[:, 0], ray_u[:, 0], trig_u[:, 0], trig_v[:, 0]


def _react(u, n, n_in, n_out, mirror_rule):
    nu = _dot(n, u)
    internal = nu > 0
    n_in = n_in.reshape(-1, 1)
    n_out = n_out.reshape(-1, 1)
    in_ok, out_ok = n_in != 0, n_out != 0
    n_in_s = torch.where(in_ok, n_in, torch.ones_like(n_in))
    n_out_s = torch.where(out_ok, n_out, torch.ones_like(n_out))
    n1 = torch.where(out_ok, n_in_s / n_out_s, torch.zeros_like(n_in))
    n2 = torch.where(in_ok, n_out_s / n_in_s, torch.zeros_like(n_in))
    eta = torch.where(internal, n1, n2)
    nu_eta = eta * nu
    radicand = 1 - eta * eta + nu_eta * nu_eta
    tir = radicand < 0
    safe = torch.where(tir, torch.ones_like(radicand), radicand)
    refract = (torch.sign(nu) * torch.sqrt(safe) - nu_eta) * n + eta * u
    reflect = u - 2 * nu * n
    if mirror_rule == "3d":      # tfrt/geometry.py:747-749
        do_reflect = tir | (n_in == 0)
    else:                        # tfrt/geometry.py:640-646
        do_reflect = tir | (eta == 0)
    return torch.where(do_reflect, reflect, refract)


def snells_law_3D(x_start, y_start, z_start, x_end, y_end, z_end, norm, n_in, n_out,
                  new_ray_length):
    """Refraction / reflection / TIR in vector form (tfrt/geometry.py:671-753).  `norm` is
    (N,3); returns the six fields of the new rays."""
    u = l2_normalize(torch.stack([x_end - x_start, y_end - y_start, z_end - z_start], dim=1))
    n = l2_normalize(norm)
    w = _react(u, n, n_in, n_out, "3d")
    end = torch.stack([x_end, y_end, z_end], dim=1) + new_ray_length * w
    return x_end, y_end, z_end, end[:, 0], end[:, 1], end[:, 2]


def snells_law_2D(x_start, y_start, x_end, y_end, norm, n_in, n_out, new_ray_length):
    """2D reaction; `norm` is the surface-normal *angle* as in the reference
    (tfrt/geometry.py:565-653).  Evaluated in vector form, which equals the reference's angle
    form but keeps the 2D mirror rule (reflect iff TIR or eta == 0)."""
    u = l2_normalize(torch.stack([x_end - x_start, y_end - y_start], dim=1))
    n = torch.stack([torch.cos(norm), torch.sin(norm)], dim=1)
    w = _react(u, n, n_in, n_out, "2d")
    end = torch.stack([x_end, y_end], dim=1) + new_ray_length * w
    return x_end, y_end, end[:, 0], end[:, 1]


# ---------------------------------------------------------------------------------------------
# 2D, paired (1-to-1) and differentiable

import math  # noqa: E402

PI = math.pi


def line_hit_2d(xs, ys, xe, ye, sx1, sy1, sx2, sy2, epsilion):
    """Ray / segment-line solve for paired rays and segments (tfrt/geometry.py:136-167).
    Returns (x, y, ray_u, segment_u)."""
    x1, y1, x2, y2 = xe - xs, ye - ys, sx2 - sx1, sy2 - sy1
    den = x1 * y2 - y1 * x2
    valid = torch.abs(den) >= epsilion
    inv = 1.0 / torch.where(valid, den, torch.ones_like(den))
    one = torch.ones_like(den)
    u = torch.where(valid, (x2 * (ys - sy1) - y2 * (xs - sx1)) * inv, one)
    v = torch.where(valid, (y1 * (sx1 - xs) - x1 * (sy1 - ys)) * inv, one)
    return xs + u * x1, ys + u * y1, u, v


def circle_hit_2d(xs, ys, xe, ye, xc, yc, r, u_hint):
    """Ray / circle solve for paired rays and arcs (tfrt/geometry.py:464-547), returning the
    root the search kernel selected (`u_hint`, its ray parameter).  Returns (x, y, ray_u,
    arc_u) where arc_u is the polar angle of the hit about the centre."""
    inv_r = 1.0 / r
    xr, yr = (xs - xc) * inv_r, (ys - yc) * inv_r
    xd, yd = (xe - xs) * inv_r, (ye - ys) * inv_r
    a = xd * xd + yd * yd
    b = 2.0 * xr * xd + 2.0 * yr * yd
    c = xr * xr + yr * yr - 1.0
    rad = b * b - 4.0 * a * c
    bad = (rad < 0) | (a == 0)
    sq = torch.sqrt(torch.where(bad, torch.ones_like(rad), rad))
    inv = 1.0 / torch.where(bad, torch.ones_like(a), 2.0 * a)
    um, up = (-b - sq) * inv, (-b + sq) * inv
    u = torch.where((um - u_hint).abs() <= (up - u_hint).abs(), um, up)
    u = torch.where(bad, torch.ones_like(u), u)
    x, y = xs + (xe - xs) * u, ys + (ye - ys) * u
    return x, y, u, torch.atan2(y - yc, x - xc)
