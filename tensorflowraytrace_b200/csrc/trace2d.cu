// trace2d.cu -- K4: the 2D path (segments + arcs), forward and backward (sm_100a).
//
// Reference path restated here, per ray instead of dense (M,N) tensors:
//   geometry.raw_line_intersect            tfrt/geometry.py:96-167
//   geometry.raw_line_circle_intersect     tfrt/geometry.py:420-547
//   geometry.angle_in_interval             tfrt/geometry.py:766-802
//   OpticalSystem2D.intersect              tfrt/engine.py:523-616   (segment norm :580-586)
//   _seg_or_arc / _get_arc_norm            tfrt/engine.py:626-670
//   _segment_intersection/_arc_intersection tfrt/engine.py:687-866
//   process_projection_2D                  tfrt/engine.py:1544-1986
//   StandardReaction.main (2D)             tfrt/operation.py:255-307
//   geometry.snells_law_2D                 tfrt/geometry.py:565-653
//
// The 2D workloads of the reference are small (config 1: 1 000 rays x 3 segments x 3
// iterations; dev/performance.py:77-86 goes up to 10 000 rays x 1 001 segments) and
// launch-latency bound, so the design goal is ONE launch per trace: a ray is followed through
// every iteration (rays never split or interact) by a group of LANES threads (1, 4 or 32, chosen
// from the ray and boundary counts) that share the scan over the boundaries -- each lane takes
// every LANES-th segment / arc, then the group reduces to the nearest hit with shuffles (lowest
// index on exact ties) -- so a trace of few rays over many boundaries still fills the machine.
// The boundaries are staged once per CTA in shared memory when they fit.  All arithmetic is fp64 on the
// fp32 inputs -- the reference's own precision -- so the nearest-hit choice needs no near-tie
// allowance.  The per-(iteration, slot) tape feeds the backward kernel, which walks the
// iterations in reverse with the closed-form VJPs derived in DESIGN.md section 6.
#include "common.cuh"

namespace {

constexpr int kThreads2D = 128;
constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * kPi;

__device__ __forceinline__ double floored_mod(double x, double m) {   // tf.math.mod
    const double r = fmod(x, m);
    return (r != 0.0 && ((r < 0.0) != (m < 0.0))) ? r + m : r;
}

// geometry.angle_in_interval (tfrt/geometry.py:790-802)
__device__ __forceinline__ bool angle_in_interval(double angle, double start, double end) {
    double ra = angle - start;
    if (ra < 0.0) ra += kTwoPi;
    double re = end - start;
    if (re < 0.0) re += kTwoPi;
    return ra <= re;
}

struct SegHit {
    bool valid;
    double u, v;   // ray parameter, segment parameter
};

// raw_line_intersect (tfrt/geometry.py:136-167) + pruning (tfrt/engine.py:722-724)
__device__ __forceinline__ SegHit seg_test(double xs, double ys, double xe, double ye,
                                           double sx1, double sy1, double sx2, double sy2,
                                           double eps, double size_eps, double start_eps) {
    const double x1 = xe - xs, y1 = ye - ys, x2 = sx2 - sx1, y2 = sy2 - sy1;
    const double den = x1 * y2 - y1 * x2;
    SegHit h;
    h.valid = fabs(den) >= eps;
    const double inv = 1.0 / (h.valid ? den : 1.0);
    h.u = h.valid ? (x2 * (ys - sy1) - y2 * (xs - sx1)) * inv : 1.0;
    h.v = h.valid ? (y1 * (sx1 - xs) - x1 * (sy1 - ys)) * inv : 1.0;
    h.valid = h.valid && h.v >= -size_eps && h.v <= 1.0 + size_eps && h.u >= start_eps;
    return h;
}

struct ArcHit {
    bool valid;
    double u, v;   // ray parameter, polar angle of the hit about the centre
};

// raw_line_circle_intersect (tfrt/geometry.py:464-547) + pruning and root choice
// (tfrt/engine.py:803-845)
__device__ __forceinline__ ArcHit arc_test(double xs, double ys, double xe, double ye, double xc,
                                           double yc, double a1, double a2, double r, double eps,
                                           double start_eps) {
    const double inv_r = 1.0 / r;
    const double xr = (xs - xc) * inv_r, yr = (ys - yc) * inv_r;
    const double xd = (xe - xs) * inv_r, yd = (ye - ys) * inv_r;
    const double a = xd * xd + yd * yd;
    const double b = 2.0 * xr * xd + 2.0 * yr * yd;
    const double c = xr * xr + yr * yr - 1.0;
    double rad = b * b - 4.0 * a * c;
    if (fabs(rad) < eps) rad = 0.0;                 // tangency snap
    const bool rad_less = rad < 0.0;
    const bool azero = fabs(a) < eps;
    const double sq = sqrt(rad_less ? 1.0 : rad);
    const double inv = 1.0 / (azero ? 1.0 : 2.0 * a);
    const bool ok = !rad_less && !azero;
    const double um = azero ? 1.0 : (rad_less ? 1.0 : (-b - sq)) * inv;
    const double up = azero ? 1.0 : (rad_less ? 1.0 : (-b + sq)) * inv;
    const double xm = xs + (xe - xs) * um, ym = ys + (ye - ys) * um;
    const double xp = xs + (xe - xs) * up, yp = ys + (ye - ys) * up;
    const double vm = atan2(ym - yc, xm - xc), vp = atan2(yp - yc, xp - xc);
    const bool valid_m = ok && um >= start_eps && angle_in_interval(vm, a1, a2);
    const bool valid_p = ok && up >= start_eps && angle_in_interval(vp, a1, a2);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double mu = valid_m ? um : inf, pu = valid_p ? up : inf;
    const bool choose_minus = mu < pu;
    ArcHit h;
    h.valid = choose_minus ? valid_m : valid_p;
    h.u = choose_minus ? mu : pu;
    h.v = choose_minus ? vm : vp;
    return h;
}

struct Nearest2D {
    bool seg_valid, arc_valid;
    int seg, arc;            // argmin index (0 when nothing valid, like tf.argmin)
    double seg_u, seg_v, arc_u, arc_v;
};

struct Geo2D {
    const float *seg[4];
    const float *arc[5];
    int n_seg, n_arc;
    double eps, size_eps, start_eps;
};

// nearest segment and nearest arc.  The LANES threads of a group share the scan (lane `sub`
// takes boundaries sub, sub + LANES, ...; strict '<' on an ascending scan = lowest index on ties
// within a lane), then reduce to the lexicographic minimum of (u, index): the lowest index on
// exact ties, the behaviour of the reference's CPU argmin.  Then _seg_or_arc: with both valid
// the segment wins iff seg_u < arc_u.  Every lane of the group returns the same result.
template <int LANES>
__device__ __forceinline__ Nearest2D nearest2d(const Geo2D &G, double xs, double ys, double xe,
                                               double ye, int sub) {
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    Nearest2D n;
    n.seg_valid = n.arc_valid = false;
    n.seg = n.arc = 0;
    n.seg_u = n.arc_u = inf;
    n.seg_v = n.arc_v = 0.0;
    for (int i = sub; i < G.n_seg; i += LANES) {
        const SegHit h = seg_test(xs, ys, xe, ye, G.seg[0][i], G.seg[1][i], G.seg[2][i],
                                  G.seg[3][i], G.eps, G.size_eps, G.start_eps);
        if (h.valid && h.u < n.seg_u) {
            n.seg_valid = true; n.seg = i; n.seg_u = h.u; n.seg_v = h.v;
        }
    }
    for (int i = sub; i < G.n_arc; i += LANES) {
        const ArcHit h = arc_test(xs, ys, xe, ye, G.arc[0][i], G.arc[1][i], G.arc[2][i],
                                  G.arc[3][i], G.arc[4][i], G.eps, G.start_eps);
        if (h.valid && h.u < n.arc_u) {
            n.arc_valid = true; n.arc = i; n.arc_u = h.u; n.arc_v = h.v;
        }
    }
    if (LANES > 1) {
#pragma unroll
        for (int off = LANES / 2; off > 0; off >>= 1) {
            {
                const double ou = __shfl_xor_sync(0xffffffffu, n.seg_u, off);
                const double ov = __shfl_xor_sync(0xffffffffu, n.seg_v, off);
                const int oi = __shfl_xor_sync(0xffffffffu, n.seg, off);
                const int ok = __shfl_xor_sync(0xffffffffu, (int)n.seg_valid, off);
                if (ok && (!n.seg_valid || ou < n.seg_u || (ou == n.seg_u && oi < n.seg))) {
                    n.seg_valid = true; n.seg = oi; n.seg_u = ou; n.seg_v = ov;
                }
            }
            {
                const double ou = __shfl_xor_sync(0xffffffffu, n.arc_u, off);
                const double ov = __shfl_xor_sync(0xffffffffu, n.arc_v, off);
                const int oi = __shfl_xor_sync(0xffffffffu, n.arc, off);
                const int ok = __shfl_xor_sync(0xffffffffu, (int)n.arc_valid, off);
                if (ok && (!n.arc_valid || ou < n.arc_u || (ou == n.arc_u && oi < n.arc))) {
                    n.arc_valid = true; n.arc = oi; n.arc_u = ou; n.arc_v = ov;
                }
            }
        }
    }
    if (n.seg_valid && n.arc_valid) {
        const bool seg_less = n.seg_u < n.arc_u;
        n.seg_valid = seg_less;
        n.arc_valid = !seg_less;
    }
    return n;
}

// Copy the boundaries into shared memory (when the launcher reserved room for them) and point
// the geometry at the copies.  Called by every thread of the CTA.
__device__ __forceinline__ void stage_boundaries(Geo2D &G, float *smem, int staged) {
    if (!staged) return;
    float *dst = smem;
    for (int c = 0; c < 4; ++c) {
        if (G.n_seg) {
            for (int i = threadIdx.x; i < G.n_seg; i += blockDim.x) dst[i] = G.seg[c][i];
            G.seg[c] = dst;
            dst += G.n_seg;
        }
    }
    for (int c = 0; c < 5; ++c) {
        if (G.n_arc) {
            for (int i = threadIdx.x; i < G.n_arc; i += blockDim.x) dst[i] = G.arc[c][i];
            G.arc[c] = dst;
            dst += G.n_arc;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double seg_norm(const Geo2D &G, int i) {   // tfrt/engine.py:580-586
    return atan2((double)G.seg[3][i] - (double)G.seg[1][i],
                 (double)G.seg[2][i] - (double)G.seg[0][i]) + kPi / 2.0;
}

__device__ __forceinline__ double arc_norm(const Geo2D &G, int i, double arc_u) {   // :666-670
    const double r = G.arc[4][i];
    return floored_mod((r < 0.0 ? arc_u + kPi : arc_u) + kPi, kTwoPi) - kPi;
}

// snells_law_2D (tfrt/geometry.py:601-651).  Returns the new ray angle; the optional outputs
// feed the backward pass: d(new_angle)/d(norm) and d(new_angle)/d(ray_angle).
__device__ __forceinline__ double snell2d(double xs, double ys, double xh, double yh, double norm,
                                          double n_in, double n_out, double *d_norm,
                                          double *d_ray) {
    norm = floored_mod(norm, kTwoPi);
    const double ray_angle = floored_mod(atan2(ys - yh, xs - xh), kTwoPi);
    double theta1 = norm - ray_angle;
    if (theta1 > kPi) theta1 -= kTwoPi;
    if (theta1 < -kPi) theta1 += kTwoPi;
    const bool internal = fabs(theta1) >= kPi / 2.0;
    const bool in_safe = n_in != 0.0, out_safe = n_out != 0.0;
    const double nis = in_safe ? n_in : 1.0, nos = out_safe ? n_out : 1.0;
    const double n1 = out_safe ? nis / nos : 0.0;
    const double n2 = in_safe ? nos / nis : 0.0;
    const double n = internal ? n1 : n2;
    if (!internal) norm += kPi;
    if (internal) theta1 += kPi;
    const double theta2 = n * sin(theta1);
    const bool use_refract = fabs(theta2) <= 1.0 && n != 0.0;
    if (use_refract) {
        if (d_norm) {
            const double den = sqrt(1.0 - theta2 * theta2);
            const double c = den > 0.0 ? n * cos(theta1) / den : 0.0;
            *d_norm = 1.0 - c;
            *d_ray = c;
        }
        return norm - asin(theta2);
    }
    if (d_norm) {
        *d_norm = 2.0;
        *d_ray = -1.0;
    }
    return norm + theta1 + kPi;
}

struct Mat2D {
    int index_mode;
    const float *wavelength;
    const int32_t *seg_mat_in, *seg_mat_out, *arc_mat_in, *arc_mat_out;
    const float *seg_n_in, *seg_n_out, *arc_n_in, *arc_n_out;
    const TfrtMaterial *materials;
};

__device__ __forceinline__ void indices2d(const Mat2D &M, bool is_arc, int idx, int slot,
                                          double &n_in, double &n_out) {
    if (M.index_mode) {
        const double wl = M.wavelength[slot];
        const int mi = is_arc ? M.arc_mat_in[idx] : M.seg_mat_in[idx];
        const int mo = is_arc ? M.arc_mat_out[idx] : M.seg_mat_out[idx];
        n_in = tfrt_material_n(M.materials[mi], wl);
        n_out = tfrt_material_n(M.materials[mo], wl);
    } else {
        n_in = is_arc ? M.arc_n_in[idx] : M.seg_n_in[idx];
        n_out = is_arc ? M.arc_n_out[idx] : M.seg_n_out[idx];
    }
}

__device__ __forceinline__ int category_of(int idx, int n_optical, int n_stop) {
    return idx < n_optical ? TFRT_STATUS_ACTIVE
                           : (idx < n_optical + n_stop ? TFRT_STATUS_STOPPED : TFRT_STATUS_FINISHED);
}

struct Fwd2D {
    Geo2D G;
    Mat2D M;
    int n_rays, K;
    int n_seg_optical, n_seg_stop, n_arc_optical, n_arc_stop;
    double new_ray_length;
    const float *src[4];
    float *tape_f;
    double *tape_d;
    int32_t *tape_hit;
    uint8_t *tape_status;
    int32_t *live_counts;
};

template <int LANES>
__global__ void __launch_bounds__(kThreads2D) k_trace2d_forward(Fwd2D P, int staged) {
    extern __shared__ float s_boundaries[];
    stage_boundaries(P.G, s_boundaries, staged);
    // LANES consecutive threads follow one ray; all of them carry its state, lane 0 writes
    const int slot = (blockIdx.x * kThreads2D + threadIdx.x) / LANES;
    const int sub = threadIdx.x % LANES;
    const size_t N = (size_t)P.n_rays;
    bool live = slot < P.n_rays;
    double xs = 0, ys = 0, xe = 0, ye = 0;
    if (live) {
        xs = P.src[0][slot]; ys = P.src[1][slot]; xe = P.src[2][slot]; ye = P.src[3][slot];
    }
    for (int k = 0; k < P.K; ++k) {
        const unsigned m = __ballot_sync(0xffffffffu, live && sub == 0);
        if (m == 0) break;
        if ((threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd(&P.live_counts[k], __popc(m));
        // (a group is live or dead as a whole; dead groups of a live warp take part in the
        // shuffles of nearest2d with no boundaries to scan)
        Geo2D G = P.G;
        if (!live) { G.n_seg = 0; G.n_arc = 0; }
        const Nearest2D n = nearest2d<LANES>(G, xs, ys, xe, ye, sub);
        if (!live) continue;
        int status = TFRT_STATUS_DEAD, hit = TFRT_HIT2D_NONE;
        double px = xe, py = ye, ray_u = 0.0, bnd_u = 0.0;
        if (n.seg_valid || n.arc_valid) {
            const bool is_arc = n.arc_valid;
            const int idx = is_arc ? n.arc : n.seg;
            ray_u = is_arc ? n.arc_u : n.seg_u;
            bnd_u = is_arc ? n.arc_v : n.seg_v;
            // x = x1s + u * x1 (tfrt/geometry.py:164-165, :521-524)
            px = xs + (xe - xs) * ray_u;
            py = ys + (ye - ys) * ray_u;
            hit = is_arc ? -2 - idx : idx;
            status = is_arc ? category_of(idx, P.n_arc_optical, P.n_arc_stop)
                            : category_of(idx, P.n_seg_optical, P.n_seg_stop);
        }
        if (sub == 0) {
            float *tf = P.tape_f + (size_t)k * 4 * N;
            tf[slot] = (float)xs; tf[N + slot] = (float)ys;
            tf[2 * N + slot] = (float)px; tf[3 * N + slot] = (float)py;
            double *td = P.tape_d + (size_t)k * 2 * N;
            td[slot] = ray_u; td[N + slot] = bnd_u;
            P.tape_hit[(size_t)k * N + slot] = hit;
            P.tape_status[(size_t)k * N + slot] = (uint8_t)status;
        }
        if (status != TFRT_STATUS_ACTIVE) {
            live = false;
            continue;
        }
        const bool is_arc = hit < 0;
        const int idx = is_arc ? -2 - hit : hit;
        const double norm = is_arc ? arc_norm(P.G, idx, bnd_u) : seg_norm(P.G, idx);
        double n_in, n_out;
        indices2d(P.M, is_arc, idx, slot, n_in, n_out);
        const double ang = snell2d(xs, ys, px, py, norm, n_in, n_out, nullptr, nullptr);
        xs = px; ys = py;
        xe = px + P.new_ray_length * cos(ang);
        ye = py + P.new_ray_length * sin(ang);
    }
}

// ---------------------------------------------------------------------------------------------
// backward

struct Bwd2D {
    Fwd2D F;
    const float *g_tape;   // K*4*N
    double *g_acc;         // 4*n_seg + 3*n_arc
    float *g_src[4];
};

// add `vals[n]` into dst[0..n) once per group of lanes sharing `key` (warp-aggregated atomics)
template <int NV>
__device__ __forceinline__ void grouped_atomic_add(bool active, int key, double *dst,
                                                   double (&vals)[NV]) {
    const int lane = threadIdx.x & 31;
    const unsigned any = __ballot_sync(0xffffffffu, active);
    if (any == 0) return;
    const unsigned peers = __match_any_sync(0xffffffffu, active ? key : -1 - lane);
    const int leader = __ffs(peers) - 1;
    int maxn = __popc(peers);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) maxn = max(maxn, __shfl_xor_sync(0xffffffffu, maxn, off));
    for (int step = 1; step < maxn; ++step) {
        const unsigned src = __fns(peers, 0, step + 1);
        const int src_lane = (src == 0xffffffffu) ? lane : (int)src;
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            const double other = __shfl_sync(0xffffffffu, vals[c], src_lane);
            if (lane == leader && src != 0xffffffffu) vals[c] += other;
        }
    }
    if (active && lane == leader) {
#pragma unroll
        for (int c = 0; c < NV; ++c)
            if (vals[c] != 0.0) atomicAdd(dst + c, vals[c]);
    }
}

__global__ void __launch_bounds__(kThreads2D) k_trace2d_backward(Bwd2D B) {
    const Fwd2D &P = B.F;
    const int slot = blockIdx.x * kThreads2D + threadIdx.x;
    const size_t N = (size_t)P.n_rays;
    const bool mine = slot < P.n_rays;
    double gS[2] = {0, 0}, gE[2] = {0, 0};   // cotangent of the ray produced by iteration k
    for (int k = P.K - 1; k >= 0; --k) {
        const int status = mine ? P.tape_status[(size_t)k * N + slot] : 255;
        if (__ballot_sync(0xffffffffu, status != 255) == 0) continue;
        bool seg_contrib = false, arc_contrib = false;
        int idx = 0;
        double gseg[4] = {0, 0, 0, 0}, garc[3] = {0, 0, 0};
        if (status != 255) {
            const float *tf = P.tape_f + (size_t)k * 4 * N;
            const float *gt = B.g_tape + (size_t)k * 4 * N;
            const double S[2] = {tf[slot], tf[N + slot]};
            const double H[2] = {tf[2 * N + slot], tf[3 * N + slot]};
            const double ghS[2] = {gt[slot], gt[N + slot]};
            const double ghP[2] = {gt[2 * N + slot], gt[3 * N + slot]};
            if (status == TFRT_STATUS_DEAD) {
                gS[0] = ghS[0]; gS[1] = ghS[1];
                gE[0] = ghP[0]; gE[1] = ghP[1];
            } else {
                const int hit = P.tape_hit[(size_t)k * N + slot];
                const bool is_arc = hit < 0;
                idx = is_arc ? -2 - hit : hit;
                const double t = P.tape_d[(size_t)k * 2 * N + slot];
                const double bu = P.tape_d[(size_t)k * 2 * N + N + slot];
                double gH[2] = {ghP[0], ghP[1]};
                double gSd[2] = {ghS[0], ghS[1]};
                // boundary geometry
                double A[2] = {0, 0}, Bp[2] = {0, 0}, C[2] = {0, 0}, r = 1.0;
                if (is_arc) {
                    C[0] = __ldg(P.G.arc[0] + idx); C[1] = __ldg(P.G.arc[1] + idx);
                    r = __ldg(P.G.arc[4] + idx);
                } else {
                    A[0] = __ldg(P.G.seg[0] + idx); A[1] = __ldg(P.G.seg[1] + idx);
                    Bp[0] = __ldg(P.G.seg[2] + idx); Bp[1] = __ldg(P.G.seg[3] + idx);
                }
                if (status == TFRT_STATUS_ACTIVE) {
                    // reaction: new ray (S', E') = (H, H + L (cos phi, sin phi))
                    const double norm = is_arc ? arc_norm(P.G, idx, bu) : seg_norm(P.G, idx);
                    double n_in, n_out, d_norm, d_ray;
                    indices2d(P.M, is_arc, idx, slot, n_in, n_out);
                    const double phi = snell2d(S[0], S[1], H[0], H[1], norm, n_in, n_out, &d_norm,
                                               &d_ray);
                    const double g_phi = P.new_ray_length * (-sin(phi) * gE[0] + cos(phi) * gE[1]);
                    gH[0] += gS[0] + gE[0];
                    gH[1] += gS[1] + gE[1];
                    const double g_norm = g_phi * d_norm, g_rho = g_phi * d_ray;
                    // ray_angle rho = atan2(a_y, a_x), a = S - H
                    const double a[2] = {S[0] - H[0], S[1] - H[1]};
                    const double a2 = a[0] * a[0] + a[1] * a[1];
                    if (a2 > 0.0) {
                        const double ga[2] = {-g_rho * a[1] / a2, g_rho * a[0] / a2};
                        gSd[0] += ga[0]; gSd[1] += ga[1];
                        gH[0] -= ga[0]; gH[1] -= ga[1];
                    }
                    if (is_arc) {   // norm = atan2(q_y, q_x) (+ const), q = H - C
                        const double q[2] = {H[0] - C[0], H[1] - C[1]};
                        const double q2 = q[0] * q[0] + q[1] * q[1];
                        if (q2 > 0.0) {
                            const double gq[2] = {-g_norm * q[1] / q2, g_norm * q[0] / q2};
                            gH[0] += gq[0]; gH[1] += gq[1];
                            garc[0] -= gq[0]; garc[1] -= gq[1];
                        }
                    } else {        // norm = atan2(w_y, w_x) + pi/2, w = B - A
                        const double w[2] = {Bp[0] - A[0], Bp[1] - A[1]};
                        const double w2 = w[0] * w[0] + w[1] * w[1];
                        if (w2 > 0.0) {
                            const double gw[2] = {-g_norm * w[1] / w2, g_norm * w[0] / w2};
                            gseg[0] -= gw[0]; gseg[1] -= gw[1];
                            gseg[2] += gw[0]; gseg[3] += gw[1];
                        }
                    }
                }
                // hit-point block: H = S + t D constrained to the boundary
                const double D[2] = {H[0] - S[0], H[1] - S[1]};
                double m[2];
                if (is_arc) { m[0] = H[0] - C[0]; m[1] = H[1] - C[1]; }
                else { m[0] = -(Bp[1] - A[1]); m[1] = Bp[0] - A[0]; }
                const double mD = m[0] * D[0] + m[1] * D[1];
                const double kappa = mD != 0.0 ? (gH[0] * D[0] + gH[1] * D[1]) / mD : 0.0;
                if (is_arc) {
                    garc[0] += kappa * m[0]; garc[1] += kappa * m[1];
                    garc[2] += kappa * r;
                    arc_contrib = true;
                } else {
                    gseg[0] += (1.0 - bu) * kappa * m[0]; gseg[1] += (1.0 - bu) * kappa * m[1];
                    gseg[2] += bu * kappa * m[0]; gseg[3] += bu * kappa * m[1];
                    seg_contrib = true;
                }
                for (int c = 0; c < 2; ++c) {
                    const double proj = gH[c] - kappa * m[c];
                    gS[c] = gSd[c] + (1.0 - t) * proj;
                    gE[c] = t * proj;
                }
            }
            if (k == 0 && B.g_src[0]) {
                B.g_src[0][slot] = (float)gS[0]; B.g_src[1][slot] = (float)gS[1];
                B.g_src[2][slot] = (float)gE[0]; B.g_src[3][slot] = (float)gE[1];
            }
        }
        grouped_atomic_add<4>(seg_contrib, idx, B.g_acc + (size_t)idx * 4, gseg);
        grouped_atomic_add<3>(arc_contrib, idx,
                              B.g_acc + (size_t)P.G.n_seg * 4 + (size_t)idx * 3, garc);
    }
}

__global__ void k_finish2d(const double *g_acc, int n_seg, int n_arc, float *s0, float *s1,
                           float *s2, float *s3, float *a0, float *a1, float *a2, float *a3,
                           float *a4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_seg) {
        s0[i] = (float)g_acc[4 * i]; s1[i] = (float)g_acc[4 * i + 1];
        s2[i] = (float)g_acc[4 * i + 2]; s3[i] = (float)g_acc[4 * i + 3];
    }
    if (i < n_arc) {
        const double *g = g_acc + (size_t)n_seg * 4 + (size_t)i * 3;
        a0[i] = (float)g[0]; a1[i] = (float)g[1]; a2[i] = 0.f; a3[i] = 0.f; a4[i] = (float)g[2];
    }
}

// b1-2D seam: one projection, both geometries reported separately like the reference
template <int LANES>
__global__ void __launch_bounds__(kThreads2D)
k_intersect2d(Geo2D G, const float *rx1, const float *ry1, const float *rx2, const float *ry2,
              int n_rays, float *sx, float *sy, float *su, float *sv, float *sn,
              uint8_t *seg_valid, int32_t *gather_segment, float *ax, float *ay, float *au,
              float *av, float *an, uint8_t *arc_valid, int32_t *gather_arc, int staged) {
    extern __shared__ float s_boundaries[];
    stage_boundaries(G, s_boundaries, staged);
    const int i = (blockIdx.x * kThreads2D + threadIdx.x) / LANES;
    const int sub = threadIdx.x % LANES;
    const bool mine = i < n_rays;
    double xs = 0, ys = 0, xe = 0, ye = 0;
    if (mine) { xs = rx1[i]; ys = ry1[i]; xe = rx2[i]; ye = ry2[i]; }
    Geo2D Gs = G;
    if (!mine) { Gs.n_seg = 0; Gs.n_arc = 0; }
    const Nearest2D n = nearest2d<LANES>(Gs, xs, ys, xe, ye, sub);
    if (!mine || sub != 0) return;
    if (G.n_seg > 0) {
        // entries of rays without a valid hit are unspecified (tfrt/engine.py:528-533)
        const double u = isfinite(n.seg_u) ? n.seg_u : 0.0;
        if (sx) sx[i] = (float)(xs + (xe - xs) * u);
        if (sy) sy[i] = (float)(ys + (ye - ys) * u);
        if (su) su[i] = (float)u;
        if (sv) sv[i] = (float)n.seg_v;
        if (sn) sn[i] = (float)seg_norm(G, n.seg);
        if (seg_valid) seg_valid[i] = n.seg_valid ? 1 : 0;
        if (gather_segment) gather_segment[i] = n.seg;
    }
    if (G.n_arc > 0) {
        const double u = isfinite(n.arc_u) ? n.arc_u : 0.0;
        if (ax) ax[i] = (float)(xs + (xe - xs) * u);
        if (ay) ay[i] = (float)(ys + (ye - ys) * u);
        if (au) au[i] = (float)u;
        if (av) av[i] = (float)n.arc_v;
        if (an) an[i] = (float)arc_norm(G, n.arc, n.arc_v);
        if (arc_valid) arc_valid[i] = n.arc_valid ? 1 : 0;
        if (gather_arc) gather_arc[i] = n.arc;
    }
}

Fwd2D make_fwd(const TfrtTrace2DArgs *a) {
    Fwd2D P;
    for (int c = 0; c < 4; ++c) P.G.seg[c] = a->seg[c];
    for (int c = 0; c < 5; ++c) P.G.arc[c] = a->arc[c];
    P.G.n_seg = a->n_seg; P.G.n_arc = a->n_arc;
    P.G.eps = a->intersect_eps; P.G.size_eps = a->size_eps; P.G.start_eps = a->ray_start_eps;
    P.M.index_mode = a->index_mode; P.M.wavelength = a->src[4];
    P.M.seg_mat_in = a->seg_mat_in; P.M.seg_mat_out = a->seg_mat_out;
    P.M.arc_mat_in = a->arc_mat_in; P.M.arc_mat_out = a->arc_mat_out;
    P.M.seg_n_in = a->seg_n_in; P.M.seg_n_out = a->seg_n_out;
    P.M.arc_n_in = a->arc_n_in; P.M.arc_n_out = a->arc_n_out;
    P.M.materials = a->materials;
    P.n_rays = a->n_rays; P.K = a->max_iterations;
    P.n_seg_optical = a->n_seg_optical; P.n_seg_stop = a->n_seg_stop;
    P.n_arc_optical = a->n_arc_optical; P.n_arc_stop = a->n_arc_stop;
    P.new_ray_length = a->new_ray_length;
    for (int c = 0; c < 4; ++c) P.src[c] = a->src[c];
    P.tape_f = a->tape_f; P.tape_d = a->tape_d; P.tape_hit = a->tape_hit;
    P.tape_status = a->tape_status; P.live_counts = a->live_counts;
    return P;
}

// Threads per ray: as many as it takes to fill the machine while every lane keeps a worthwhile
// share of the boundaries.
int lanes_for(long long n_rays, long long n_boundaries) {
    long long lanes = 1;
    while (lanes < 32 && n_rays * lanes < 148LL * 1024 && n_boundaries / (2 * lanes) >= 8) lanes *= 2;
    return lanes <= 1 ? 1 : (lanes <= 4 ? 4 : 32);
}

// dynamic shared memory needed to stage the boundaries (0: leave them in global memory)
size_t staged_bytes(int n_seg, int n_arc) {
    const size_t bytes = sizeof(float) * (4 * (size_t)n_seg + 5 * (size_t)n_arc);
    return bytes <= 160 * 1024 ? bytes : 0;
}

int check2d(const TfrtTrace2DArgs *a, const char *who) {
    TFRT_REQUIRE(a, "%s: null args", who);
    TFRT_REQUIRE(a->n_rays >= 0 && a->max_iterations > 0 && a->n_seg >= 0 && a->n_arc >= 0 &&
                     a->n_seg + a->n_arc > 0,
                 "%s: n_rays=%d max_iterations=%d n_seg=%d n_arc=%d", who, a->n_rays,
                 a->max_iterations, a->n_seg, a->n_arc);
    TFRT_REQUIRE(a->n_seg_optical + a->n_seg_stop <= a->n_seg &&
                     a->n_arc_optical + a->n_arc_stop <= a->n_arc,
                 "%s: category counts exceed the boundary counts", who);
    TFRT_REQUIRE(a->tape_f && a->tape_d && a->tape_hit && a->tape_status && a->live_counts,
                 "%s: null tape pointer", who);
    if (a->n_rays == 0) return 0;   // nothing is dereferenced
    if (a->index_mode) {
        TFRT_REQUIRE(a->materials && a->src[4], "%s: index mode needs materials and wavelength", who);
        TFRT_REQUIRE((a->n_seg_optical == 0 || (a->seg_mat_in && a->seg_mat_out)) &&
                         (a->n_arc_optical == 0 || (a->arc_mat_in && a->arc_mat_out)),
                     "%s: mat_in / mat_out missing", who);
    } else {
        TFRT_REQUIRE((a->n_seg_optical == 0 || (a->seg_n_in && a->seg_n_out)) &&
                         (a->n_arc_optical == 0 || (a->arc_n_in && a->arc_n_out)),
                     "%s: n_in / n_out missing", who);
    }
    return 0;
}

}  // namespace

extern "C" int tfrt_intersect2d(const float *const ray[4], int32_t n_rays,
                                const float *const seg[4], int32_t n_seg,
                                const float *const arc[5], int32_t n_arc, float intersect_eps,
                                float size_eps, float ray_start_eps, float *const seg_out[5],
                                uint8_t *seg_valid, int32_t *gather_segment,
                                float *const arc_out[5], uint8_t *arc_valid, int32_t *gather_arc,
                                void *stream) {
    TFRT_REQUIRE(n_rays >= 0 && n_seg >= 0 && n_arc >= 0 && n_seg + n_arc > 0,
                 "tfrt_intersect2d: bad arguments");
    if (n_rays == 0) return 0;
    Geo2D G;
    for (int c = 0; c < 4; ++c) G.seg[c] = n_seg ? seg[c] : nullptr;
    for (int c = 0; c < 5; ++c) G.arc[c] = n_arc ? arc[c] : nullptr;
    G.n_seg = n_seg; G.n_arc = n_arc;
    G.eps = intersect_eps; G.size_eps = size_eps; G.start_eps = ray_start_eps;
    float *so[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float *ao[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int c = 0; c < 5; ++c) {
        if (seg_out) so[c] = seg_out[c];
        if (arc_out) ao[c] = arc_out[c];
    }
    const int lanes = lanes_for(n_rays, (long long)n_seg + n_arc);
    const size_t smem = staged_bytes(n_seg, n_arc);
    const int grid = tfrt_div_up((long long)n_rays * lanes, kThreads2D);
#define TFRT_LAUNCH2D(L)                                                                        \
    do {                                                                                        \
        if (smem > 48 * 1024)                                                                   \
            TFRT_CHECK_CUDA(cudaFuncSetAttribute(k_intersect2d<L>,                              \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                                 (int)smem));                                   \
        k_intersect2d<L><<<grid, kThreads2D, smem, (cudaStream_t)stream>>>(                     \
            G, ray[0], ray[1], ray[2], ray[3], n_rays, so[0], so[1], so[2], so[3], so[4],       \
            seg_valid, gather_segment, ao[0], ao[1], ao[2], ao[3], ao[4], arc_valid, gather_arc, \
            smem > 0 ? 1 : 0);                                                                  \
    } while (0)
    if (lanes == 1) TFRT_LAUNCH2D(1);
    else if (lanes == 4) TFRT_LAUNCH2D(4);
    else TFRT_LAUNCH2D(32);
#undef TFRT_LAUNCH2D
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfrt_trace2d_forward(const TfrtTrace2DArgs *a, void *stream) {
    if (int rc = check2d(a, "tfrt_trace2d_forward")) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = (size_t)a->n_rays, K = (size_t)a->max_iterations;
    TFRT_CHECK_CUDA(cudaMemsetAsync(a->live_counts, 0, sizeof(int32_t) * K, st));
    if (N == 0) return 0;
    TFRT_CHECK_CUDA(cudaMemsetAsync(a->tape_status, 0xff, K * N, st));
    const int lanes = lanes_for(a->n_rays, (long long)a->n_seg + a->n_arc);
    const size_t smem = staged_bytes(a->n_seg, a->n_arc);
    const int grid = tfrt_div_up((long long)a->n_rays * lanes, kThreads2D);
#define TFRT_LAUNCH2D(L)                                                                        \
    do {                                                                                        \
        if (smem > 48 * 1024)                                                                   \
            TFRT_CHECK_CUDA(cudaFuncSetAttribute(k_trace2d_forward<L>,                          \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                                 (int)smem));                                   \
        k_trace2d_forward<L><<<grid, kThreads2D, smem, st>>>(make_fwd(a), smem > 0 ? 1 : 0);    \
    } while (0)
    if (lanes == 1) TFRT_LAUNCH2D(1);
    else if (lanes == 4) TFRT_LAUNCH2D(4);
    else TFRT_LAUNCH2D(32);
#undef TFRT_LAUNCH2D
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfrt_trace2d_backward(const TfrtTrace2DArgs *a, const float *g_tape, double *g_acc,
                                     float *const g_seg[4], float *const g_arc[5],
                                     float *const g_src[4], void *stream) {
    if (int rc = check2d(a, "tfrt_trace2d_backward")) return rc;
    TFRT_REQUIRE(g_tape && g_acc, "tfrt_trace2d_backward: null g_tape / g_acc");
    TFRT_REQUIRE(a->n_seg == 0 || (g_seg && g_seg[0] && g_seg[1] && g_seg[2] && g_seg[3]),
                 "tfrt_trace2d_backward: null g_seg");
    TFRT_REQUIRE(a->n_arc == 0 || (g_arc && g_arc[0] && g_arc[1] && g_arc[2] && g_arc[3] && g_arc[4]),
                 "tfrt_trace2d_backward: null g_arc");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_acc = (size_t)a->n_seg * 4 + (size_t)a->n_arc * 3;
    TFRT_CHECK_CUDA(cudaMemsetAsync(g_acc, 0, sizeof(double) * n_acc, st));
    Bwd2D B;
    B.F = make_fwd(a);
    B.g_tape = g_tape;
    B.g_acc = g_acc;
    for (int c = 0; c < 4; ++c) B.g_src[c] = g_src ? g_src[c] : nullptr;
    if (a->n_rays > 0) {
        if (B.g_src[0])
            for (int c = 0; c < 4; ++c)
                TFRT_CHECK_CUDA(cudaMemsetAsync(B.g_src[c], 0, sizeof(float) * a->n_rays, st));
        k_trace2d_backward<<<tfrt_div_up(a->n_rays, kThreads2D), kThreads2D, 0, st>>>(B);
        tfrt_note_launch(1);
    }
    const int n_max = a->n_seg > a->n_arc ? a->n_seg : a->n_arc;
    k_finish2d<<<tfrt_div_up(n_max, 128), 128, 0, st>>>(
        g_acc, a->n_seg, a->n_arc, a->n_seg ? g_seg[0] : nullptr, a->n_seg ? g_seg[1] : nullptr,
        a->n_seg ? g_seg[2] : nullptr, a->n_seg ? g_seg[3] : nullptr,
        a->n_arc ? g_arc[0] : nullptr, a->n_arc ? g_arc[1] : nullptr,
        a->n_arc ? g_arc[2] : nullptr, a->n_arc ? g_arc[3] : nullptr,
        a->n_arc ? g_arc[4] : nullptr);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}
