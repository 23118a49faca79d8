// backward3d.cu -- K2: reverse pass of the trace loop (the tf.GradientTape of
// tfrt/optimizer.py:216-220 restricted to OpticalEngine.ray_trace).
//
// The forward tape holds, per (iteration k, slot): ray start S, projected end P (hit H or the
// unchanged end for a miss), ray_u t, barycentrics (u, v), hit triangle, status and the slot's
// position in the output list it was appended to.  argmin / masks / barycentric comparisons carry
// no gradient in the reference, so the VJP is O(rays x iterations): walk k = K-1 .. 0 and apply,
// per live ray, the closed forms (derived in DESIGN.md, checked against float64 autograd of the
// oracle in tests/):
//
//   reaction (ACTIVE only, tfrt/geometry.py:715-753), new ray (S', E') = (H, H + L w):
//       g_w = L g_E' ;  g_H = g_S' + g_E' (+ history cotangent on H)
//       refract: c = sgn(nu) sqrt(r) - eta nu, c' = sgn(nu) eta^2 nu / sqrt(r) - eta
//                g_n^ = c g_w + (g_w.n^) c' u^ ;  g_u^ = eta g_w + (g_w.n^) c' n^
//       reflect: g_u^ = g_w - 2 (g_w.n^) n^ ;    g_n^ = -2 [nu g_w + (g_w.n^) u^]
//       normalisations: g_x = (g_y^ - (g_y^.y^) y^) / |x|   for x = H - S and x = norm
//   hit point (tfrt/geometry.py:286-318), H = S + t D on the plane of (A, B, C), m = e1 x e2:
//       kappa = (g_H.D)/(m.D) ; g_A = (1-u-v) kappa m ; g_B = u kappa m ; g_C = v kappa m
//       g_S += (1-t)(g_H - kappa m) ;  g_E = t (g_H - kappa m)
//
// Every triangle cotangent of a hit is a per-triangle linear map of SIX per-ray scalars: the corner
// cotangents of the hit point are (1-u-v) kappa m, u kappa m, v kappa m with m a constant of the
// triangle, and the cotangent of its normal reaches the corners through the constant Jacobian of
// n = m/|m|.  So a hit adds six numbers -- (1-u-v) kappa, u kappa, v kappa and the 3-vector g_m
// (or g_n when the caller supplies normals) -- with warp-aggregated fp64 atomics (lanes that hit
// the same triangle are combined first) into one 64-byte row per triangle; k_finish_tri_grads
// applies the per-triangle maps once and rounds to fp32.  The triangle corners are read from a
// 48-byte record next to that row (three 16-byte loads instead of nine scattered ones), written
// by k_backward_prepare, which also zeroes the sums.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int kBwdThreads = 256;
constexpr int kRowDoubles = 16;   // 128-byte row per triangle: doubles 0-5 sums, bytes 64-111 corners
constexpr int kSums = 6;

// Zero the sums and write the corner record (A.xyz, B.xyz, C.xyz as fp32) of every triangle.
__global__ void k_backward_prepare(int n_tri, const float *t0, const float *t1, const float *t2,
                                   const float *t3, const float *t4, const float *t5,
                                   const float *t6, const float *t7, const float *t8,
                                   double *__restrict__ g_acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tri) return;
    double4 *row = reinterpret_cast<double4 *>(g_acc + (size_t)i * kRowDoubles);
    row[0] = make_double4(0, 0, 0, 0);
    row[1] = make_double4(0, 0, 0, 0);
    float4 *rec = reinterpret_cast<float4 *>(row + 2);
    rec[0] = make_float4(t0[i], t1[i], t2[i], t3[i]);
    rec[1] = make_float4(t4[i], t5[i], t6[i], t7[i]);
    rec[2] = make_float4(t8[i], 0.f, 0.f, 0.f);
}

struct BwdParams {
    int n_rays, n_iters, index_mode, n_grad_tri;
    double new_ray_length, dead_ray_length;
    const int32_t *counters;    // (K+1)*4 device counters of the forward pass
    const float *tape_f;        // K*9*N
    const int32_t *tape_tri;    // K*N
    const uint8_t *tape_status; // K*N
    const int32_t *tape_outpos; // K*N
    const float *g_hist[4][6];
    const float *norm;
    const float *src_wavelength;
    const int32_t *mat_in, *mat_out;
    const float *n_in, *n_out;
    const TfrtMaterial *materials;
    const double *n_table;      // optional tabulated indices (see TfrtTrace3DArgs)
    const int32_t *lambda_index;
    int n_lambda;
    double *g_acc;              // T rows of kRowDoubles: six sums, two spare, corner record
    float *g_src[6];            // written at k == 0 when non-null
};

// Sum `vals[12]` over the lanes of `peers` (lanes sharing one triangle) into the group's lowest
// lane: the leader walks its peers in ascending lane order (one shuffle round per peer; groups are
// tiny for incoherent rays, and the order makes the sum independent of the launch shape).
__device__ __forceinline__ void aggregate_peers(unsigned peers, int lane, double vals[kSums],
                                                bool &is_leader) {
    const int n = __popc(peers);
    int maxn = n;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) maxn = max(maxn, __shfl_xor_sync(0xffffffffu, maxn, off));
    const int leader = __ffs(peers) - 1;
    is_leader = (lane == leader);
    unsigned rest = peers & (peers - 1);                   // the peers above the leader
#pragma unroll 1
    for (int step = 1; step < maxn; ++step) {
        // lane of the peer with rank `step`, in ascending order (peeled off with ffs)
        const int src_lane = rest ? __ffs(rest) - 1 : lane;
        rest &= rest - 1;
#pragma unroll
        for (int c = 0; c < kSums; ++c) {
            const double other = __shfl_sync(0xffffffffu, vals[c], src_lane);
            if (is_leader && step < n) vals[c] += other;
        }
    }
}

// One thread per source ray walks its iterations in reverse (a ray is live for a prefix
// 0 .. k_last of the iterations); the running cotangent of the ray stays in registers, the
// tape is read slot-indexed (coalesced), and each iteration's triangle cotangents are combined
// across the lanes that hit the same triangle before the fp64 atomics.  One launch per trace.
// kMinBlocks = 3 caps the kernel at 80 registers (a few spilled words) for 24 instead of 16 resident
// warps per SM: the kernel waits on scattered L2 reads and atomics, so the extra warps pay.
template <int kMinBlocks>
__global__ void __launch_bounds__(kBwdThreads, kMinBlocks) k_backward3d(BwdParams P) {
    const int slot = blockIdx.x * kBwdThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const size_t N = (size_t)P.n_rays;
    const int K = P.n_iters;
    const bool mine = slot < P.n_rays;
    // last iteration this ray took part in (-1: never; not possible for a source ray)
    int k_last = -1;
    if (mine) {
        for (int k = K - 1; k >= 0; --k)
            if (P.tape_status[(size_t)k * N + slot] != 255) { k_last = k; break; }
    }
    int k_top = k_last;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) k_top = max(k_top, __shfl_xor_sync(0xffffffffu, k_top, off));
    // position of iteration k's first ray in the active history = rays that reacted before k
    long long active_before = 0;
    for (int j = 0; j <= k_top; ++j) active_before += P.counters[(j + 1) * 4 + 0];
    double gSn[3] = {0, 0, 0}, gEn[3] = {0, 0, 0};   // cotangent of the ray leaving iteration k

    for (int k = k_top; k >= 0; --k) {
    active_before -= P.counters[(k + 1) * 4 + 0];
    const int active_offset = (int)active_before;
    bool contributes = false;
    int tri = -1;
    double acc[kSums];
#pragma unroll
    for (int c = 0; c < kSums; ++c) acc[c] = 0.0;

    if (mine && k <= k_last) {
        const int status = P.tape_status[(size_t)k * N + slot];
        tri = P.tape_tri[(size_t)k * N + slot];
        const float *tf = P.tape_f + (size_t)k * 9 * N;
        const double S[3] = {tf[0 * N + slot], tf[1 * N + slot], tf[2 * N + slot]};
        const double Pt[3] = {tf[3 * N + slot], tf[4 * N + slot], tf[5 * N + slot]};
        const double t = tf[6 * N + slot], u = tf[7 * N + slot], v = tf[8 * N + slot];

        // cotangent of this iteration's history record (start, projected end)
        double ghS[3] = {0, 0, 0}, ghP[3] = {0, 0, 0};
        {
            const int pos = P.tape_outpos[(size_t)k * N + slot] +
                            (status == TFRT_STATUS_ACTIVE ? active_offset : 0);
            const float *const *g = P.g_hist[status];
            for (int c = 0; c < 3; ++c) {
                if (g[c]) ghS[c] = g[c][pos];
                if (g[3 + c]) ghP[c] = g[3 + c][pos];
            }
        }
        double gS[3], gE[3];
        if (status == TFRT_STATUS_DEAD) {
            // ray unchanged (optionally rescaled: end' = S + L (E - S), tfrt/engine.py:2036-2046)
            const double L = P.dead_ray_length;
            for (int c = 0; c < 3; ++c) {
                if (L >= 0.0) {
                    gS[c] = ghS[c] + (1.0 - L) * ghP[c];
                    gE[c] = L * ghP[c];
                } else {
                    gS[c] = ghS[c];
                    gE[c] = ghP[c];
                }
            }
        } else {
            const float4 *rec =
                reinterpret_cast<const float4 *>(P.g_acc + (size_t)tri * kRowDoubles + 8);
            const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2];
            const double A[3] = {r0.x, r0.y, r0.z}, B[3] = {r0.w, r1.x, r1.y};
            const double C[3] = {r1.z, r1.w, r2.x};
            double gH[3] = {ghP[0], ghP[1], ghP[2]};
            double gSd[3] = {ghS[0], ghS[1], ghS[2]};
            if (status == TFRT_STATUS_ACTIVE) {
                // (gSn, gEn): cotangent of the reacted ray, left by iteration k + 1 of this slot
                // (zero when that iteration was not executed)
                double gw[3];
                for (int c = 0; c < 3; ++c) {
                    gw[c] = P.new_ray_length * gEn[c];
                    gH[c] += gSn[c] + gEn[c];
                }
                // recompute the reaction's intermediates (same formulas as the forward kernel)
                double d[3] = {Pt[0] - S[0], Pt[1] - S[1], Pt[2] - S[2]};
                const double d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                const double dinv = rsqrt(fmax(d2, 1e-12));
                const double uh[3] = {d[0] * dinv, d[1] * dinv, d[2] * dinv};
                double nr[3], cu[3], cv[3], inv_clen = 1.0;
                if (P.norm) {
                    nr[0] = P.norm[3 * tri]; nr[1] = P.norm[3 * tri + 1]; nr[2] = P.norm[3 * tri + 2];
                } else {   // geometric normal from the corners, as in the forward kernel
                    for (int c = 0; c < 3; ++c) { cu[c] = B[c] - A[c]; cv[c] = C[c] - B[c]; }
                    const double cr[3] = {cu[1] * cv[2] - cu[2] * cv[1], cu[2] * cv[0] - cu[0] * cv[2],
                                          cu[0] * cv[1] - cu[1] * cv[0]};
                    // 1/|c| once (rsqrt) instead of a square root and six divisions
                    inv_clen = rsqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
                    for (int c = 0; c < 3; ++c) nr[c] = cr[c] * inv_clen;
                }
                const double n2 = nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2];
                const double ninv = rsqrt(fmax(n2, 1e-12));
                const double nh[3] = {nr[0] * ninv, nr[1] * ninv, nr[2] * ninv};
                const double nu = nh[0] * uh[0] + nh[1] * uh[1] + nh[2] * uh[2];
                double n_in, n_out;
                if (P.n_table) {
                    const int li = P.lambda_index[slot];
                    n_in = P.n_table[(size_t)P.mat_in[tri] * P.n_lambda + li];
                    n_out = P.n_table[(size_t)P.mat_out[tri] * P.n_lambda + li];
                } else if (P.index_mode) {
                    const double wl = P.src_wavelength[slot];
                    n_in = tfrt_material_n(P.materials[P.mat_in[tri]], wl);
                    n_out = tfrt_material_n(P.materials[P.mat_out[tri]], wl);
                } else {
                    n_in = P.n_in[tri];
                    n_out = P.n_out[tri];
                }
                const bool in_safe = n_in != 0.0, out_safe = n_out != 0.0;
                const double nis = in_safe ? n_in : 1.0, nos = out_safe ? n_out : 1.0;
                const double eta = (nu > 0.0) ? (out_safe ? nis / nos : 0.0)
                                              : (in_safe ? nos / nis : 0.0);
                const double r = 1.0 - eta * eta + eta * eta * nu * nu;
                const double gwn = gw[0] * nh[0] + gw[1] * nh[1] + gw[2] * nh[2];
                double g_uh[3], g_nh[3];
                if (r < 0.0 || n_in == 0.0) {
                    for (int c = 0; c < 3; ++c) {
                        g_uh[c] = gw[c] - 2.0 * gwn * nh[c];
                        g_nh[c] = -2.0 * (nu * gw[c] + gwn * uh[c]);
                    }
                } else {
                    const double sgn = (nu > 0.0) ? 1.0 : ((nu < 0.0) ? -1.0 : 0.0);
                    const double sq = sqrt(r);
                    const double cc = sgn * sq - eta * nu;
                    const double cp = (sq > 0.0 ? sgn * eta * eta * nu / sq : 0.0) - eta;
                    for (int c = 0; c < 3; ++c) {
                        g_nh[c] = cc * gw[c] + gwn * cp * uh[c];
                        g_uh[c] = eta * gw[c] + gwn * cp * nh[c];
                    }
                }
                // through the two normalisations (x * rsqrt(max(|x|^2, 1e-12)))
                if (d2 > 1e-12) {
                    const double dot = g_uh[0] * uh[0] + g_uh[1] * uh[1] + g_uh[2] * uh[2];
                    for (int c = 0; c < 3; ++c) {
                        const double gd = (g_uh[c] - dot * uh[c]) * dinv;
                        gH[c] += gd;
                        gSd[c] -= gd;
                    }
                } else {
                    for (int c = 0; c < 3; ++c) {
                        gH[c] += g_uh[c] * dinv;
                        gSd[c] -= g_uh[c] * dinv;
                    }
                }
                {
                    const double dot = g_nh[0] * nh[0] + g_nh[1] * nh[1] + g_nh[2] * nh[2];
                    double g_nr[3];   // cotangent of the (unit) normal fed to the reaction
                    for (int c = 0; c < 3; ++c)
                        g_nr[c] = (n2 > 1e-12) ? (g_nh[c] - dot * nh[c]) * ninv : g_nh[c] * ninv;
                    if (P.norm) {
                        for (int c = 0; c < 3; ++c) acc[3 + c] = g_nr[c];
                    } else {
                        // through n = c/|c|, c = u x v (u = B - A, v = C - B): the cotangent of
                        // the unnormalised cross product, g_c = (g_n - (g_n.n) n)/|c|; its map
                        // onto the corners is applied once per triangle by k_finish_tri_grads
                        const double dn = g_nr[0] * nr[0] + g_nr[1] * nr[1] + g_nr[2] * nr[2];
                        for (int c = 0; c < 3; ++c) acc[3 + c] = (g_nr[c] - dn * nr[c]) * inv_clen;
                    }
                }
            }
            // hit-point block
            const double e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
            const double e2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
            const double m[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2],
                                 e1[0] * e2[1] - e1[1] * e2[0]};
            // D is only needed up to scale: use H - S (t > 0 for a valid hit)
            const double D[3] = {Pt[0] - S[0], Pt[1] - S[1], Pt[2] - S[2]};
            const double mD = m[0] * D[0] + m[1] * D[1] + m[2] * D[2];
            const double gD = gH[0] * D[0] + gH[1] * D[1] + gH[2] * D[2];
            const double kappa = (mD != 0.0) ? gD / mD : 0.0;
            acc[0] = (1.0 - u - v) * kappa;
            acc[1] = u * kappa;
            acc[2] = v * kappa;
            for (int c = 0; c < 3; ++c) {
                const double km = kappa * m[c];
                const double proj = gH[c] - km;
                gS[c] = gSd[c] + (1.0 - t) * proj;
                gE[c] = t * proj;
            }
            contributes = tri < P.n_grad_tri;
        }
        for (int c = 0; c < 3; ++c) {
            gSn[c] = gS[c];
            gEn[c] = gE[c];
            if (k == 0 && P.g_src[c]) {
                P.g_src[c][slot] = (float)gS[c];
                P.g_src[3 + c][slot] = (float)gE[c];
            }
        }
    }

    // warp-aggregated fp64 atomics into the per-triangle accumulators
    const unsigned active = __ballot_sync(0xffffffffu, contributes);
    if (active != 0) {
        const int key = contributes ? tri : -1 - lane;   // non-contributors never match anybody
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        bool is_leader = false;
        aggregate_peers(peers, lane, acc, is_leader);
        if (contributes && is_leader) {
            double *dst = P.g_acc + (size_t)tri * kRowDoubles;
#pragma unroll
            for (int c = 0; c < kSums; ++c)
                if (acc[c] != 0.0) atomicAdd(dst + c, acc[c]);
        }
    }
    }   // iterations, in reverse
}

// Per-triangle maps of the six sums onto the corner cotangents (see the header):
//   g_A = s0 m - g_u ; g_B = s1 m + g_u - g_v ; g_C = s2 m + g_v,  m = (B-A) x (C-A),
//   g_u = (C-B) x G ; g_v = G x (B-A)  with G = sum of g_c (geometric normals), else G = 0 and
//   g_norm = the sum itself (caller-supplied normals).
__global__ void k_finish_tri_grads(const double *__restrict__ g_acc, int n_tri, int has_norm,
                                   float *o0, float *o1, float *o2, float *o3, float *o4,
                                   float *o5, float *o6, float *o7, float *o8, float *g_norm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tri) return;
    const double *g = g_acc + (size_t)i * kRowDoubles;
    const float4 *rec = reinterpret_cast<const float4 *>(g + 8);
    const float4 r0 = rec[0], r1 = rec[1], r2 = rec[2];
    const double A[3] = {r0.x, r0.y, r0.z}, B[3] = {r0.w, r1.x, r1.y}, C[3] = {r1.z, r1.w, r2.x};
    const double e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    const double e2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
    const double cv[3] = {C[0] - B[0], C[1] - B[1], C[2] - B[2]};
    const double m[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2],
                         e1[0] * e2[1] - e1[1] * e2[0]};
    double gu[3] = {0, 0, 0}, gv[3] = {0, 0, 0};
    if (!has_norm) {
        const double G[3] = {g[3], g[4], g[5]};
        gu[0] = cv[1] * G[2] - cv[2] * G[1]; gu[1] = cv[2] * G[0] - cv[0] * G[2];
        gu[2] = cv[0] * G[1] - cv[1] * G[0];
        gv[0] = G[1] * e1[2] - G[2] * e1[1]; gv[1] = G[2] * e1[0] - G[0] * e1[2];
        gv[2] = G[0] * e1[1] - G[1] * e1[0];
    }
    float *const oa[3] = {o0, o1, o2}, *const ob[3] = {o3, o4, o5}, *const oc[3] = {o6, o7, o8};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        oa[c][i] = (float)(g[0] * m[c] - gu[c]);
        ob[c][i] = (float)(g[1] * m[c] + gu[c] - gv[c]);
        oc[c][i] = (float)(g[2] * m[c] + gv[c]);
        g_norm[3 * i + c] = has_norm ? (float)g[3 + c] : 0.f;
    }
}

}  // namespace

extern "C" int tfrt_trace3d_backward(const TfrtTrace3DArgs *a, const int32_t *ch,
                                     const TfrtTrace3DBwdArgs *b, void *stream) {
    (void)ch;   // the iteration counts are read on the device: the call needs no host copy
    TFRT_REQUIRE(a && b, "tfrt_trace3d_backward: null argument");
    TFRT_REQUIRE(b->g_acc && b->g_norm_out, "tfrt_trace3d_backward: null workspace");
    for (int c = 0; c < 9; ++c)
        TFRT_REQUIRE(b->g_tri_out[c], "tfrt_trace3d_backward: null g_tri_out[%d]", c);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = a->n_rays, K = a->max_iterations, T = a->n_tri;
    k_backward_prepare<<<tfrt_div_up(T, 256), 256, 0, st>>>(
        T, a->tri[0], a->tri[1], a->tri[2], a->tri[3], a->tri[4], a->tri[5], a->tri[6], a->tri[7],
        a->tri[8], b->g_acc);
    tfrt_note_launch(1);
    if (N > 0) {
        for (int c = 0; c < 6; ++c)
            if (b->g_src[c]) TFRT_CHECK_CUDA(cudaMemsetAsync(b->g_src[c], 0, sizeof(float) * N, st));
        BwdParams P;
        P.n_rays = N; P.n_iters = K; P.index_mode = a->index_mode;
        P.n_grad_tri = b->n_grad_tri > 0 ? b->n_grad_tri : T;
        P.new_ray_length = a->new_ray_length;
        P.dead_ray_length = b->dead_ray_length;
        P.counters = a->counters;
        P.tape_f = a->tape_f; P.tape_tri = a->tape_tri;
        P.tape_status = a->tape_status; P.tape_outpos = a->tape_outpos;
        for (int w = 0; w < 4; ++w)
            for (int c = 0; c < 6; ++c) P.g_hist[w][c] = b->g_hist[w][c];
        P.norm = a->norm; P.src_wavelength = a->src[6];
        P.mat_in = a->mat_in; P.mat_out = a->mat_out; P.n_in = a->n_in; P.n_out = a->n_out;
        P.materials = a->materials;
        P.n_table = a->n_table; P.lambda_index = a->lambda_index; P.n_lambda = a->n_lambda;
        for (int c = 0; c < 6; ++c) P.g_src[c] = b->g_src[c];
        P.g_acc = b->g_acc;
        static const int min_blocks = [] {
            const char *e = getenv("TFRT_BWD_MINBLOCKS");     // tuning switch (2 or 3)
            return e ? atoi(e) : 3;
        }();
        if (min_blocks >= 3)
            k_backward3d<3><<<tfrt_div_up(N, kBwdThreads), kBwdThreads, 0, st>>>(P);
        else
            k_backward3d<2><<<tfrt_div_up(N, kBwdThreads), kBwdThreads, 0, st>>>(P);
        tfrt_note_launch(1);
    }
    k_finish_tri_grads<<<tfrt_div_up(T, 256), 256, 0, st>>>(
        b->g_acc, T, a->norm != nullptr, b->g_tri_out[0], b->g_tri_out[1], b->g_tri_out[2], b->g_tri_out[3],
        b->g_tri_out[4], b->g_tri_out[5], b->g_tri_out[6], b->g_tri_out[7], b->g_tri_out[8],
        b->g_norm_out);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}
