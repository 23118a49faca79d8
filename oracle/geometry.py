"""
Oracle restatement of tfrt/geometry.py (float64 torch CPU).  TEST INFRASTRUCTURE ONLY.

Each function follows the reference op-for-op (same safe-value / `where` structure, same
operation order) so that results agree with the TensorFlow float64 reference to rounding.
"""

import math

import torch

PI = math.pi


def _f64(x):
    return torch.as_tensor(x, dtype=torch.float64)


# --------------------------------------------------------------------------------------
# 3D ray x triangle


def raw_line_triangle_intersect(
    rx1, ry1, rz1, rx2, ry2, rz2, xp, yp, zp, x1, y1, z1, x2, y2, z2, epsilion
):
    """tfrt/geometry.py:275-320.  Cramer solve of ray/triangle; any co-broadcastable shape."""
    a = rx1 - rx2
    b = x1 - xp
    c = x2 - xp
    d = ry1 - ry2
    f = y1 - yp
    g = y2 - yp
    h = rz1 - rz2
    k = z1 - zp
    l = z2 - zp

    q = rx1 - xp
    r = ry1 - yp
    s = rz1 - zp

    denominator = a * g * k + b * d * l + c * f * h - a * f * l - b * g * h - c * d * k
    ray_u_numerator = b * l * r + c * f * s + g * k * q - b * g * s - c * k * r - f * l * q
    trig_u_numerator = a * g * s + c * h * r + d * l * q - a * l * r - c * d * s - g * h * q
    trig_v_numerator = a * k * r + b * d * s + f * h * q - a * f * s - b * h * r - d * k * q

    # safe division (geometry.py:305-311)
    valid = torch.abs(denominator) >= epsilion
    safe_value = torch.ones_like(denominator)
    safe_denominator = torch.where(valid, denominator, safe_value)
    ray_u = ray_u_numerator / safe_denominator
    trig_u = trig_u_numerator / safe_denominator
    trig_v = trig_v_numerator / safe_denominator

    # geometry.py:316-318
    x = rx1 - ray_u * a
    y = ry1 - ray_u * d
    z = rz1 - ray_u * h
    return x, y, z, valid, ray_u, trig_u, trig_v


def line_triangle_intersect(
    rx1, ry1, rz1, rx2, ry2, rz2, xp, yp, zp, x1, y1, z1, x2, y2, z2, epsilion
):
    """tfrt/geometry.py:191-251.  N rays x M triangles -> (M, N) tensors (tf.meshgrid
    default 'xy' indexing puts the second argument, the triangles, on axis 0)."""
    ray = [t.reshape(1, -1) for t in (rx1, ry1, rz1, rx2, ry2, rz2)]
    tri = [t.reshape(-1, 1) for t in (xp, yp, zp, x1, y1, z1, x2, y2, z2)]
    return raw_line_triangle_intersect(*ray, *tri, epsilion)


def l2_normalize(x, dim):
    """tf.math.l2_normalize: x * rsqrt(max(sum(x**2), 1e-12))."""
    sq = torch.sum(x * x, dim=dim, keepdim=True)
    return x * torch.rsqrt(torch.clamp(sq, min=1e-12))


def snells_law_3D(x_start, y_start, z_start, x_end, y_end, z_end, norm, n_in, n_out,
                  new_ray_length):
    """tfrt/geometry.py:671-753.  Vector-form refraction / reflection / TIR."""
    u = torch.stack([x_end - x_start, y_end - y_start, z_end - z_start], dim=1)
    u = l2_normalize(u, 1)

    n = l2_normalize(norm, 1)
    nu = torch.sum(n * u, dim=1, keepdim=True)

    internal_mask = nu > 0
    one = torch.ones_like(n_in)
    zero = torch.zeros_like(n_in)

    n_in_is_safe = n_in != 0.0
    n_in_safe = torch.where(n_in_is_safe, n_in, one)
    n_out_is_safe = n_out != 0.0
    n_out_safe = torch.where(n_out_is_safe, n_out, one)

    n1 = torch.where(n_out_is_safe, n_in_safe / n_out_safe, zero).reshape(-1, 1)
    n2 = torch.where(n_in_is_safe, n_out_safe / n_in_safe, zero).reshape(-1, 1)
    eta = torch.where(internal_mask, n1, n2)
    nu_eta = eta * nu

    radicand = 1 - eta * eta + nu_eta * nu_eta
    do_tir = radicand < 0
    safe_radicand = torch.where(do_tir, torch.ones_like(radicand), radicand)
    refract = (torch.sign(nu) * torch.sqrt(safe_radicand) - nu_eta) * n + eta * u

    reflect = -2 * nu * n + u

    reflective_surface = (n_in == 0).reshape(-1, 1)
    do_reflect = torch.logical_or(do_tir, reflective_surface)
    new_vector = torch.where(do_reflect, reflect, refract)

    new_endpoints = torch.stack([x_end, y_end, z_end], dim=1) + new_ray_length * new_vector
    x, y, z = torch.unbind(new_endpoints, dim=1)
    return x_end, y_end, z_end, x, y, z


# --------------------------------------------------------------------------------------
# 2D line x line, line x circle


def raw_line_intersect(x1s, y1s, x1e, y1e, x2s, y2s, x2e, y2e, epsilion):
    """tfrt/geometry.py:96-167."""
    x1 = x1e - x1s
    y1 = y1e - y1s
    x2 = x2e - x2s
    y2 = y2e - y2s
    denominator = x1 * y2 - y1 * x2

    valid_intersection = torch.abs(denominator) >= epsilion
    safe_value = torch.ones_like(denominator)
    safe_denominator = torch.where(valid_intersection, denominator, safe_value)
    safe_denominator = 1.0 / safe_denominator

    u = torch.where(
        valid_intersection,
        (x2 * (y1s - y2s) - y2 * (x1s - x2s)) * safe_denominator,
        safe_value,
    )
    v = torch.where(
        valid_intersection,
        (y1 * (x2s - x1s) - x1 * (y2s - y1s)) * safe_denominator,
        safe_value,
    )
    x = x1s + u * x1
    y = y1s + u * y1
    return x, y, valid_intersection, u, v


def line_intersect(x1s, y1s, x1e, y1e, x2s, y2s, x2e, y2e, epsilion):
    """tfrt/geometry.py:27-94.  N first lines x M second lines -> (M, N)."""
    first = [t.reshape(1, -1) for t in (x1s, y1s, x1e, y1e)]
    second = [t.reshape(-1, 1) for t in (x2s, y2s, x2e, y2e)]
    return raw_line_intersect(*first, *second, epsilion)


def raw_line_circle_intersect(xs, ys, xe, ye, xc, yc, r, epsilion):
    """tfrt/geometry.py:420-547.  Returns (plus, minus) dicts with x, y, valid, u, v."""
    inverse_r = 1.0 / r
    xr = (xs - xc) * inverse_r
    yr = (ys - yc) * inverse_r
    xd = (xe - xs) * inverse_r
    yd = (ye - ys) * inverse_r

    a = xd * xd + yd * yd
    b = 2.0 * xr * xd + 2.0 * yr * yd
    c = xr * xr + yr * yr - 1.0
    rad = b * b - 4.0 * a * c

    # tangency snap (geometry.py:492-496)
    rad = torch.where(torch.abs(rad) < epsilion, torch.zeros_like(rad), rad)

    safe_value = torch.ones_like(rad)
    radLess = rad < 0
    uminus_valid = uplus_valid = torch.logical_not(radLess)
    safe_rad = torch.where(radLess, safe_value, rad)
    safe_rad = torch.sqrt(safe_rad)
    uminus = torch.where(radLess, safe_value, (-b - safe_rad))
    uplus = torch.where(radLess, safe_value, (-b + safe_rad))

    azero = torch.abs(a) < epsilion
    safe_denominator = torch.where(azero, safe_value, 2 * a)
    safe_denominator = 1.0 / safe_denominator
    uminus_valid = torch.logical_and(uminus_valid, torch.logical_not(azero))
    uminus = torch.where(azero, safe_value, uminus * safe_denominator)
    uplus_valid = torch.logical_and(uplus_valid, torch.logical_not(azero))
    uplus = torch.where(azero, safe_value, uplus * safe_denominator)

    xminus = xs + (xe - xs) * uminus
    xplus = xs + (xe - xs) * uplus
    yminus = ys + (ye - ys) * uminus
    yplus = ys + (ye - ys) * uplus
    vminus = torch.atan2(yminus - yc, xminus - xc)
    vplus = torch.atan2(yplus - yc, xplus - xc)

    return (
        {"x": xplus, "y": yplus, "valid": uplus_valid, "u": uplus, "v": vplus},
        {"x": xminus, "y": yminus, "valid": uminus_valid, "u": uminus, "v": vminus},
    )


def line_circle_intersect(xs, ys, xe, ye, xc, yc, r, epsilion):
    """tfrt/geometry.py:338-418.  N lines x M circles -> (M, N)."""
    line = [t.reshape(1, -1) for t in (xs, ys, xe, ye)]
    circ = [t.reshape(-1, 1) for t in (xc, yc, r)]
    return raw_line_circle_intersect(*line, *circ, epsilion)


def angle_in_interval(angle, start, end):
    """tfrt/geometry.py:766-802.  Closed interval with wrap; inputs in [-pi, pi]."""
    reduced_angle = angle - start
    reduced_angle = torch.where(reduced_angle < 0.0, reduced_angle + 2 * PI, reduced_angle)
    reduced_end = end - start
    reduced_end = torch.where(reduced_end < 0.0, reduced_end + 2 * PI, reduced_end)
    return reduced_angle <= reduced_end


def tf_mod(x, m):
    """tf.math.mod: floored modulo (result has the sign of the divisor)."""
    return torch.remainder(x, m)


def snells_law_2D(x_start, y_start, x_end, y_end, norm, n_in, n_out, new_ray_length):
    """tfrt/geometry.py:565-653.  Angle-form Snell; norm is an absolute angle."""
    norm = tf_mod(norm, 2 * PI)
    ray_angle = torch.atan2(y_start - y_end, x_start - x_end)
    ray_angle = tf_mod(ray_angle, 2 * PI)
    theta1 = norm - ray_angle
    theta1 = torch.where(theta1 > PI, theta1 - (2 * PI), theta1)
    theta1 = torch.where(theta1 < -PI, theta1 + (2 * PI), theta1)

    internal_mask = torch.abs(theta1) >= PI / 2

    one = torch.ones_like(theta1)
    zero = torch.zeros_like(theta1)

    n_in = torch.broadcast_to(_f64(n_in), theta1.shape)
    n_in_is_safe = n_in != 0.0
    n_in_safe = torch.where(n_in_is_safe, n_in, one)

    n_out = torch.broadcast_to(_f64(n_out), theta1.shape)
    n_out_is_safe = n_out != 0.0
    n_out_safe = torch.where(n_out_is_safe, n_out, one)

    n1 = torch.where(n_out_is_safe, n_in_safe / n_out_safe, zero)
    n2 = torch.where(n_in_is_safe, n_out_safe / n_in_safe, zero)
    n = torch.where(internal_mask, n1, n2)

    norm = torch.where(internal_mask, norm, norm + PI)
    theta1 = torch.where(internal_mask, theta1 + PI, theta1)

    theta2 = n * torch.sin(theta1)
    use_refract = torch.logical_and(torch.abs(theta2) <= 1.0, n != 0.0)
    # the reference evaluates asin on every element (NaN where |theta2| > 1) and selects with
    # tf.where; clamp the unselected branch so autograd never sees NaN.
    safe_theta2 = torch.where(use_refract, theta2, torch.zeros_like(theta2))
    new_angle = torch.where(use_refract, norm - torch.asin(safe_theta2), norm + theta1 + PI)

    x_start = x_end
    y_start = y_end
    x_end = x_start + new_ray_length * torch.cos(new_angle)
    y_end = y_start + new_ray_length * torch.sin(new_angle)
    return x_start, y_start, x_end, y_end
