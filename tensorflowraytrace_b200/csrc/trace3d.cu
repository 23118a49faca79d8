// trace3d.cu -- K1b and the whole-loop driver: exact hit evaluation for the winning triangle,
// classification, fused material-index + Snell reaction, per-iteration tape, order-preserving
// compaction of the live set, and materialisation of the ray histories.
//
// Reference path restated here (per live ray instead of dense tensors):
//   OpticalSystem3D._intersection gathers      tfrt/engine.py:1157-1164
//   raw_line_triangle_intersect                tfrt/geometry.py:275-320
//   process_projection_3D                      tfrt/engine.py:1988-2191
//   StandardReaction.main                      tfrt/operation.py:255-307
//   snells_law_3D                              tfrt/geometry.py:671-753
//   single_pass / ray_trace                    tfrt/engine.py:2193-2330
//
// Per-ray arithmetic here is fp64 on fp32 inputs (O(rays) work, negligible next to the
// O(rays x triangles) intersection); results are stored fp32.
#include "common.cuh"

int tfrt_launch_intersect3d(const float *const ray[6], const int32_t *live,
                            const int32_t *n_live_ptr, int n_upper, const float *packed,
                            int n_tri, float intersect_eps, float ray_start_eps, float size_eps,
                            uint64_t *hit, int dir_mode, cudaStream_t stream);

namespace {

constexpr int kReactThreads = 256;

struct HitEval {
    double t, u, v;      // ray_u, trig_u, trig_v
    double h[3];         // hit point
    bool den_ok;
};

// The reference's Cramer solve, same variable names (tfrt/geometry.py:286-318), in fp64.
__device__ __forceinline__ HitEval eval_hit(const double r1[3], const double r2[3],
                                            const double p[3], const double p1[3],
                                            const double p2[3], double eps) {
    const double a = r1[0] - r2[0], b = p1[0] - p[0], c = p2[0] - p[0];
    const double d = r1[1] - r2[1], f = p1[1] - p[1], g = p2[1] - p[1];
    const double h = r1[2] - r2[2], k = p1[2] - p[2], l = p2[2] - p[2];
    const double q = r1[0] - p[0], r = r1[1] - p[1], s = r1[2] - p[2];
    const double den = a * g * k + b * d * l + c * f * h - a * f * l - b * g * h - c * d * k;
    const double nt = b * l * r + c * f * s + g * k * q - b * g * s - c * k * r - f * l * q;
    const double nu = a * g * s + c * h * r + d * l * q - a * l * r - c * d * s - g * h * q;
    const double nv = a * k * r + b * d * s + f * h * q - a * f * s - b * h * r - d * k * q;
    HitEval e;
    e.den_ok = fabs(den) >= eps;
    const double safe = e.den_ok ? den : 1.0;
    e.t = nt / safe;
    e.u = nu / safe;
    e.v = nv / safe;
    e.h[0] = r1[0] - e.t * a;
    e.h[1] = r1[1] - e.t * d;
    e.h[2] = r1[2] - e.t * h;
    return e;
}

struct ReactParams {
    int n_tri, n_optical, index_mode, n_materials;
    double intersect_eps, new_ray_length;
    const float *src_wavelength;
    const float *tri[9];
    const float *norm;
    const int32_t *category, *mat_in, *mat_out;
    const float *n_in, *n_out;
    const TfrtMaterial *materials;
    float *cur[6];
    double *cur_d;             // 6*N: the same state in fp64 (the chain between iterations)
    const uint64_t *hit;
    const int32_t *live;       // live[k]
    const int32_t *n_live_ptr; // &counters[k*4]
    float *tape_f;             // tape_f + k*9*N
    int32_t *tape_tri;         // + k*N
    uint8_t *tape_status;      // + k*N
    int32_t *block_counts;
    int n_rays;
};

// snells_law_3D for one ray (tfrt/geometry.py:715-753); writes the new direction w.
__device__ __forceinline__ void snell3d(const double S[3], const double H[3], const double nrm[3],
                                        double n_in, double n_out, double w[3]) {
    double u[3] = {H[0] - S[0], H[1] - S[1], H[2] - S[2]};
    double s = rsqrt(fmax(u[0] * u[0] + u[1] * u[1] + u[2] * u[2], 1e-12));
    u[0] *= s; u[1] *= s; u[2] *= s;
    s = rsqrt(fmax(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2], 1e-12));
    const double n[3] = {nrm[0] * s, nrm[1] * s, nrm[2] * s};
    const double nu = n[0] * u[0] + n[1] * u[1] + n[2] * u[2];
    const bool internal = nu > 0.0;
    const bool in_safe = n_in != 0.0, out_safe = n_out != 0.0;
    const double nis = in_safe ? n_in : 1.0, nos = out_safe ? n_out : 1.0;
    const double n1 = out_safe ? nis / nos : 0.0;
    const double n2 = in_safe ? nos / nis : 0.0;
    const double eta = internal ? n1 : n2;
    const double nu_eta = eta * nu;
    const double radicand = 1.0 - eta * eta + nu_eta * nu_eta;
    const bool tir = radicand < 0.0;
    if (tir || n_in == 0.0) {
        for (int c = 0; c < 3; ++c) w[c] = -2.0 * nu * n[c] + u[c];
    } else {
        const double sgn = (nu > 0.0) ? 1.0 : ((nu < 0.0) ? -1.0 : 0.0);
        const double coef = sgn * sqrt(radicand) - nu_eta;
        for (int c = 0; c < 3; ++c) w[c] = coef * n[c] + eta * u[c];
    }
}

__global__ void __launch_bounds__(kReactThreads) k_react3d(ReactParams P) {
    __shared__ int s_counts[4];
    const int n_live = *P.n_live_ptr;
    const int p = blockIdx.x * kReactThreads + threadIdx.x;
    if (blockIdx.x * kReactThreads >= n_live) return;   // whole block idle (uniform)
    if (threadIdx.x < 4) s_counts[threadIdx.x] = 0;
    __syncthreads();

    int status = -1;
    if (p < n_live) {
        const int slot = P.live[p];
        const uint64_t hit = P.hit[p];
        const uint32_t key = (uint32_t)(hit >> 32);
        // working state = (start, direction) in fp64; end = start + direction
        const size_t N = (size_t)P.n_rays;
        double S[3], E[3];
        for (int c = 0; c < 3; ++c) {
            S[c] = P.cur_d[c * N + slot];
            E[c] = S[c] + P.cur_d[(3 + c) * N + slot];
        }
        float *tf = P.tape_f;
        int tri = -1;
        double Pt[3] = {E[0], E[1], E[2]};
        double t = 0.0, u = 0.0, v = 0.0;
        status = TFRT_STATUS_DEAD;
        if (key < TFRT_NOHIT_KEY) {
            tri = (int)(uint32_t)hit;
            double A[3], B[3], C[3];
            for (int c = 0; c < 3; ++c) {
                A[c] = P.tri[c][tri];
                B[c] = P.tri[3 + c][tri];
                C[c] = P.tri[6 + c][tri];
            }
            const HitEval e = eval_hit(S, E, A, B, C, P.intersect_eps);
            t = e.t; u = e.u; v = e.v;
            Pt[0] = e.h[0]; Pt[1] = e.h[1]; Pt[2] = e.h[2];
            status = P.category[tri];   // OPTICAL=0 -> ACTIVE, STOP=1, TARGET=2
            if (status == TFRT_STATUS_ACTIVE) {
                double n_in, n_out;
                if (P.index_mode) {
                    const double wl = P.src_wavelength[slot];
                    n_in = tfrt_material_n(P.materials[P.mat_in[tri]], wl);
                    n_out = tfrt_material_n(P.materials[P.mat_out[tri]], wl);
                } else {
                    n_in = P.n_in[tri];
                    n_out = P.n_out[tri];
                }
                double nrm[3];
                if (P.norm) {
                    nrm[0] = P.norm[3 * tri]; nrm[1] = P.norm[3 * tri + 1]; nrm[2] = P.norm[3 * tri + 2];
                } else {
                    // geometric normal normalize((p2-p1) x (p3-p2)) (tfrt/boundaries.py:918-923)
                    // in fp64 from the corners: no fp32 rounding of the normal enters the chain
                    const double ux = B[0] - A[0], uy = B[1] - A[1], uz = B[2] - A[2];
                    const double vx = C[0] - B[0], vy = C[1] - B[1], vz = C[2] - B[2];
                    const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
                    const double inv = 1.0 / sqrt(cx * cx + cy * cy + cz * cz);
                    nrm[0] = cx * inv; nrm[1] = cy * inv; nrm[2] = cz * inv;
                }
                double w[3];
                snell3d(S, Pt, nrm, n_in, n_out, w);
                // new ray = (hit, hit + new_ray_length * w)   (tfrt/geometry.py:751-753), kept as
                // (hit, new_ray_length * w): the direction keeps its full fp32 precision instead
                // of being re-derived from two rounded end points
                // The fp64 copy carries the ray from one iteration to the next, so rounding does
                // not compound over bounces; the fp32 copy feeds the intersection search.
                for (int c = 0; c < 3; ++c) {
                    const double d = P.new_ray_length * w[c];
                    P.cur_d[c * N + slot] = Pt[c];
                    P.cur_d[(3 + c) * N + slot] = d;
                    P.cur[c][slot] = (float)Pt[c];
                    P.cur[3 + c][slot] = (float)d;
                }
            }
        }
        tf[0 * N + slot] = (float)S[0]; tf[1 * N + slot] = (float)S[1]; tf[2 * N + slot] = (float)S[2];
        tf[3 * N + slot] = (float)Pt[0]; tf[4 * N + slot] = (float)Pt[1]; tf[5 * N + slot] = (float)Pt[2];
        tf[6 * N + slot] = (float)t; tf[7 * N + slot] = (float)u; tf[8 * N + slot] = (float)v;
        P.tape_tri[slot] = tri;
        P.tape_status[slot] = (uint8_t)status;
    }
    // per-block category counts for the order-preserving compaction
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, status == c);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&s_counts[c], __popc(m));
    }
    __syncthreads();
    if (threadIdx.x < 4) P.block_counts[blockIdx.x * 4 + threadIdx.x] = s_counts[threadIdx.x];
}

// Exclusive scan of the per-block category counts (single block), plus the running totals:
//   counters[k+1][0] = #active  (live count of the next iteration)
//   counters[k+1][c] = counters[k][c] + #category c   (c = 1 stopped, 2 finished, 3 dead)
__global__ void __launch_bounds__(1024) k_scan_blocks(int32_t *block_counts, int n_blocks_upper,
                                                      const int32_t *counters_k,
                                                      int32_t *counters_next) {
    __shared__ int s_part[4][1024];
    const int n_live = counters_k[0];
    const int n_blocks = min(n_blocks_upper, (n_live + kReactThreads - 1) / kReactThreads);
    const int per = (n_blocks + 1023) / 1024;
    const int b0 = threadIdx.x * per, b1 = min(n_blocks, b0 + per);
    int sum[4] = {0, 0, 0, 0};
    for (int b = b0; b < b1; ++b)
        for (int c = 0; c < 4; ++c) sum[c] += block_counts[b * 4 + c];
    for (int c = 0; c < 4; ++c) s_part[c][threadIdx.x] = sum[c];
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partials
    for (int off = 1; off < 1024; off <<= 1) {
        int vals[4];
        for (int c = 0; c < 4; ++c)
            vals[c] = (threadIdx.x >= off) ? s_part[c][threadIdx.x - off] : 0;
        __syncthreads();
        for (int c = 0; c < 4; ++c) s_part[c][threadIdx.x] += vals[c];
        __syncthreads();
    }
    int run[4];
    for (int c = 0; c < 4; ++c) run[c] = s_part[c][threadIdx.x] - sum[c];
    for (int b = b0; b < b1; ++b)
        for (int c = 0; c < 4; ++c) {
            const int v = block_counts[b * 4 + c];
            block_counts[b * 4 + c] = run[c];
            run[c] += v;
        }
    if (threadIdx.x == 1023) {
        counters_next[0] = s_part[0][1023];
        for (int c = 1; c < 4; ++c) counters_next[c] = counters_k[c] + s_part[c][1023];
    }
}

// Order-preserving scatter of the live slots into: the next live list (ACTIVE) and the three
// history id lists (STOPPED / FINISHED / DEAD, appended after the earlier iterations' entries).
__global__ void __launch_bounds__(kReactThreads)
k_scatter(const int32_t *__restrict__ live, const uint8_t *__restrict__ tape_status,
          const int32_t *__restrict__ block_offsets, const int32_t *__restrict__ counters_k,
          int32_t *__restrict__ live_next, int32_t *__restrict__ stop_ids,
          int32_t *__restrict__ fin_ids, int32_t *__restrict__ dead_ids,
          int32_t *__restrict__ tape_outpos) {
    __shared__ int s_warp[4][kReactThreads / 32];
    const int n_live = counters_k[0];
    if (blockIdx.x * kReactThreads >= n_live) return;
    const int p = blockIdx.x * kReactThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int slot = -1, status = -1;
    if (p < n_live) {
        slot = live[p];
        status = tape_status[slot];
    }
    int rank_in_warp = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, status == c);
        if (status == c) rank_in_warp = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[c][warp] = __popc(m);
    }
    __syncthreads();
    if (status < 0) return;
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[status][w];
    const int pos = block_offsets[blockIdx.x * 4 + status] + before + rank_in_warp;
    if (status == TFRT_STATUS_ACTIVE) {
        live_next[pos] = slot;
        tape_outpos[slot] = pos;
    } else {
        const int gpos = counters_k[status] + pos;
        int32_t *dst = status == TFRT_STATUS_STOPPED ? stop_ids
                       : status == TFRT_STATUS_FINISHED ? fin_ids : dead_ids;
        dst[gpos] = slot;
        tape_outpos[slot] = gpos;
    }
}

__global__ void k_iota(int32_t *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

__global__ void k_copy_rays(const float *s0, const float *s1, const float *s2, const float *s3,
                            const float *s4, const float *s5, float *d0, float *d1, float *d2,
                            float *d3, float *d4, float *d5, double *cur_d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // working state = (start, direction); the fp64 difference of two floats is exact
    const double a0 = s0[i], a1 = s1[i], a2 = s2[i];
    const double b0 = (double)s3[i] - a0, b1 = (double)s4[i] - a1, b2 = (double)s5[i] - a2;
    const size_t N = (size_t)n;
    cur_d[i] = a0; cur_d[N + i] = a1; cur_d[2 * N + i] = a2;
    cur_d[3 * N + i] = b0; cur_d[4 * N + i] = b1; cur_d[5 * N + i] = b2;
    d0[i] = (float)a0; d1[i] = (float)a1; d2[i] = (float)a2;
    d3[i] = (float)b0; d4[i] = (float)b1; d5[i] = (float)b2;
}

__global__ void k_init_counters(int32_t *counters, int n_rays, int n_entries) {
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i < n_entries) counters[i] = (i == 0) ? n_rays : 0;
}

// history materialisation: one thread per output ray.  For the stopped/finished/dead sets the
// ids are iteration-major and segment k ends at counters[(k+1)*4 + which] (cumulative counts);
// for the active set one launch per iteration passes iter_fixed >= 0.
__global__ void k_materialize(const int32_t *__restrict__ ids, const int32_t *__restrict__ counters,
                              int which, int n_iters, int iter_fixed, int n_out,
                              const float *__restrict__ tape_f, size_t n_rays,
                              float dead_ray_length, float *o0, float *o1, float *o2, float *o3,
                              float *o4, float *o5, int32_t *out_slot, int32_t *out_iter) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_out) return;
    int k = iter_fixed;
    if (k < 0) {
        k = 0;
        while (k < n_iters - 1 && q >= counters[(k + 1) * 4 + which]) ++k;
    }
    const int slot = ids[q];
    const float *tf = tape_f + (size_t)k * 9 * n_rays;
    const float sx = tf[0 * n_rays + slot], sy = tf[1 * n_rays + slot], sz = tf[2 * n_rays + slot];
    float px = tf[3 * n_rays + slot], py = tf[4 * n_rays + slot], pz = tf[5 * n_rays + slot];
    if (dead_ray_length >= 0.f) {   // tfrt/engine.py:2036-2046
        px = sx + dead_ray_length * (px - sx);
        py = sy + dead_ray_length * (py - sy);
        pz = sz + dead_ray_length * (pz - sz);
    }
    o0[q] = sx; o1[q] = sy; o2[q] = sz; o3[q] = px; o4[q] = py; o5[q] = pz;
    if (out_slot) out_slot[q] = slot;
    if (out_iter) out_iter[q] = k;
}

// simple per-ray evaluation for the b1 seam (tfrt_intersect3d)
__global__ void k_intersect_outputs(const float *sx, const float *sy, const float *sz,
                                    const float *ex, const float *ey, const float *ez,
                                    const float *xp, const float *yp, const float *zp,
                                    const float *x1, const float *y1, const float *z1,
                                    const float *x2, const float *y2, const float *z2,
                                    const uint64_t *hit, int n, double eps, float *x, float *y,
                                    float *z, uint8_t *valid, float *ray_u, float *trig_u,
                                    float *trig_v, int32_t *gather_trig) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t h = hit[i];
    const bool ok = (uint32_t)(h >> 32) < TFRT_NOHIT_KEY;
    const int tri = ok ? (int)(uint32_t)h : 0;   // argmin of an all-invalid column is 0
    const double S[3] = {sx[i], sy[i], sz[i]}, E[3] = {ex[i], ey[i], ez[i]};
    const double A[3] = {xp[tri], yp[tri], zp[tri]}, B[3] = {x1[tri], y1[tri], z1[tri]},
                 C[3] = {x2[tri], y2[tri], z2[tri]};
    const HitEval e = eval_hit(S, E, A, B, C, eps);
    x[i] = (float)e.h[0]; y[i] = (float)e.h[1]; z[i] = (float)e.h[2];
    valid[i] = ok ? 1 : 0;
    ray_u[i] = (float)e.t; trig_u[i] = (float)e.u; trig_v[i] = (float)e.v;
    gather_trig[i] = tri;
}

}  // namespace

// ------------------------------------------------------------------------------------------------

extern "C" int tfrt_intersect3d(const float *const ray[6], int32_t n_rays,
                                const float *const tri[9], const float *packed, int32_t n_tri,
                                float intersect_eps, float size_eps, float ray_start_eps,
                                uint64_t *hit_ws, float *x, float *y, float *z, uint8_t *valid,
                                float *ray_u,
                                float *trig_u, float *trig_v, int32_t *gather_trig, void *stream) {
    TFRT_REQUIRE(n_rays >= 0 && n_tri > 0 && packed && hit_ws, "tfrt_intersect3d: bad arguments");
    if (n_rays == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = tfrt_launch_intersect3d(ray, nullptr, nullptr, n_rays, packed, n_tri, intersect_eps,
                                     ray_start_eps, size_eps, hit_ws, /*dir_mode=*/0, st);
    if (rc) return rc;
    k_intersect_outputs<<<tfrt_div_up(n_rays, 256), 256, 0, st>>>(
        ray[0], ray[1], ray[2], ray[3], ray[4], ray[5], tri[0], tri[1], tri[2], tri[3], tri[4],
        tri[5], tri[6], tri[7], tri[8], hit_ws, n_rays, (double)intersect_eps, x, y, z, valid,
        ray_u, trig_u, trig_v, gather_trig);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfrt_trace3d_forward(const TfrtTrace3DArgs *a, void *stream) {
    TFRT_REQUIRE(a, "tfrt_trace3d_forward: null args");
    TFRT_REQUIRE(a->n_rays >= 0 && a->n_tri > 0 && a->max_iterations > 0,
                 "tfrt_trace3d_forward: n_rays=%d n_tri=%d max_iterations=%d", a->n_rays,
                 a->n_tri, a->max_iterations);
    TFRT_REQUIRE(a->cur_d, "tfrt_trace3d_forward: null cur_d workspace");
    TFRT_REQUIRE(a->packed && a->category && a->hit && a->tape_f && a->tape_tri &&
                     a->tape_status && a->tape_outpos && a->live && a->block_counts && a->counters,
                 "tfrt_trace3d_forward: null workspace pointer");
    TFRT_REQUIRE(a->n_optical == 0 || a->n_rays == 0 ||
                     (a->index_mode ? (a->materials && a->mat_in && a->mat_out && a->src[6])
                                    : (a->n_in && a->n_out)),
                 "tfrt_trace3d_forward: material inputs missing for index_mode=%d", a->index_mode);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = a->n_rays, K = a->max_iterations;
    k_init_counters<<<tfrt_div_up((K + 1) * 4, 128), 128, 0, st>>>(a->counters, N, (K + 1) * 4);
    if (N == 0) {
        TFRT_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    const int nb = tfrt_div_up(N, 256);
    k_copy_rays<<<nb, 256, 0, st>>>(a->src[0], a->src[1], a->src[2], a->src[3], a->src[4],
                                    a->src[5], a->cur[0], a->cur[1], a->cur[2], a->cur[3],
                                    a->cur[4], a->cur[5], a->cur_d, N);
    k_iota<<<nb, 256, 0, st>>>(a->live, N);
    tfrt_note_launch(3);
    TFRT_CHECK_CUDA(cudaMemsetAsync(a->tape_status, 0xff, (size_t)K * N, st));

    const int n_blocks = tfrt_div_up(N, kReactThreads);
    for (int k = 0; k < K; ++k) {
        const int32_t *live_k = a->live + (size_t)k * N;
        int32_t *live_next = a->live + (size_t)(k + 1) * N;
        const int32_t *counters_k = a->counters + k * 4;
        const float *cur_c[6] = {a->cur[0], a->cur[1], a->cur[2], a->cur[3], a->cur[4], a->cur[5]};
        int rc = tfrt_launch_intersect3d(cur_c, live_k, counters_k, N, a->packed, a->n_tri,
                                         a->intersect_eps, a->ray_start_eps, a->size_eps, a->hit,
                                         /*dir_mode=*/1, st);
        if (rc) return rc;
        ReactParams P;
        P.n_tri = a->n_tri; P.n_optical = a->n_optical; P.index_mode = a->index_mode;
        P.n_materials = a->n_materials;
        P.intersect_eps = a->intersect_eps; P.new_ray_length = a->new_ray_length;
        P.src_wavelength = a->src[6];
        for (int c = 0; c < 9; ++c) P.tri[c] = a->tri[c];
        P.norm = a->norm; P.category = a->category; P.mat_in = a->mat_in; P.mat_out = a->mat_out;
        P.n_in = a->n_in; P.n_out = a->n_out; P.materials = a->materials;
        for (int c = 0; c < 6; ++c) P.cur[c] = a->cur[c];
        P.cur_d = a->cur_d;
        P.hit = a->hit; P.live = live_k; P.n_live_ptr = counters_k;
        P.tape_f = a->tape_f + (size_t)k * 9 * N;
        P.tape_tri = a->tape_tri + (size_t)k * N;
        P.tape_status = a->tape_status + (size_t)k * N;
        P.block_counts = a->block_counts;
        P.n_rays = N;
        k_react3d<<<n_blocks, kReactThreads, 0, st>>>(P);
        k_scan_blocks<<<1, 1024, 0, st>>>(a->block_counts, n_blocks, counters_k,
                                          a->counters + (k + 1) * 4);
        k_scatter<<<n_blocks, kReactThreads, 0, st>>>(
            live_k, P.tape_status, a->block_counts, counters_k, live_next, a->cat_ids[0],
            a->cat_ids[1], a->cat_ids[2], a->tape_outpos + (size_t)k * N);
        tfrt_note_launch(3);
    }
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// which: 0 active, 1 stopped, 2 finished, 3 dead
static void history_segments(const TfrtTrace3DArgs *a, const int32_t *ch, int which,
                             int32_t *seg_end /*K*/, int64_t *total) {
    const int K = a->max_iterations;
    int64_t run = 0;
    for (int k = 0; k < K; ++k) {
        int64_t cnt = (which == 0) ? ch[(k + 1) * 4 + 0] : (ch[(k + 1) * 4 + which] - ch[k * 4 + which]);
        run += cnt;
        seg_end[k] = (int32_t)run;
    }
    *total = run;
}

extern "C" int tfrt_history_count(const TfrtTrace3DArgs *a, const int32_t *counters_host,
                                  int which, int64_t *n_out) {
    TFRT_REQUIRE(a && counters_host && n_out && which >= 0 && which < 4,
                 "tfrt_history_count: bad arguments");
    TFRT_REQUIRE(a->max_iterations <= 4096, "tfrt_history_count: max_iterations too large");
    int32_t seg[4096];
    history_segments(a, counters_host, which, seg, n_out);
    return 0;
}

extern "C" int tfrt_materialize3d(const TfrtTrace3DArgs *a, const int32_t *counters_host,
                                  int which, float dead_ray_length, float *const out[6],
                                  int32_t *out_slot, int32_t *out_iter, void *stream) {
    TFRT_REQUIRE(a && counters_host && which >= 0 && which < 4, "tfrt_materialize3d: bad arguments");
    TFRT_REQUIRE(a->max_iterations <= 4096, "tfrt_materialize3d: max_iterations too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int K = a->max_iterations;
    const size_t N = (size_t)a->n_rays;
    int32_t seg[4096];
    int64_t total = 0;
    history_segments(a, counters_host, which, seg, &total);
    if (total == 0) return 0;
    if (which == 0) {
        // active rays of iteration k are the live list of iteration k+1: one launch per segment
        int64_t begin = 0;
        for (int k = 0; k < K; ++k) {
            const int cnt = (int)(seg[k] - begin);
            if (cnt > 0)
                k_materialize<<<tfrt_div_up(cnt, 256), 256, 0, st>>>(
                    a->live + (size_t)(k + 1) * N, nullptr, 0, K, k, cnt, a->tape_f, N, -1.f,
                    out[0] + begin, out[1] + begin, out[2] + begin, out[3] + begin,
                    out[4] + begin, out[5] + begin, out_slot ? out_slot + begin : nullptr,
                    out_iter ? out_iter + begin : nullptr);
            if (cnt > 0) tfrt_note_launch(1);
            begin = seg[k];
        }
        TFRT_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    k_materialize<<<tfrt_div_up((int)total, 256), 256, 0, st>>>(
        a->cat_ids[which - 1], a->counters, which, K, -1, (int)total, a->tape_f, N,
        which == 3 ? dead_ray_length : -1.f, out[0], out[1], out[2], out[3], out[4], out[5],
        out_slot, out_iter);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}
