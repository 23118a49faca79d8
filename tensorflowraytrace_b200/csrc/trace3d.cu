// trace3d.cu -- K1b and the whole-loop driver: exact evaluation and re-validation of the search's
// winner, classification, fused material-index + Snell reaction, per-iteration tape,
// order-preserving compaction of the live set (single pass, decoupled look-back), and
// materialisation of the ray histories.
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
//
// Launches of one trace: k_trace_init, then per iteration k_intersect3d (intersect3d.cu),
// k_validate3d and k_react3d -- 1 + 3 K launches, no memsets, no host round trip.
#include "common.cuh"

int tfrt_launch_intersect3d(const float *const ray[6], const int32_t *live,
                            const int32_t *n_live_ptr, int n_upper, const float *packed,
                            int n_tri, float intersect_eps, float ray_start_eps, float size_eps,
                            uint64_t *hit, int dir_mode, unsigned int *work_counter,
                            float start_floor_ulps, cudaStream_t stream);

int tfrt_launch_grid_intersect3d(const float *const ray[6], const int32_t *live,
                                 const int32_t *n_live_ptr, int n_upper, const float *packed,
                                 int n_tri, const void *grid_ws, int max_cells,
                                 float intersect_eps, float ray_start_eps, float size_eps,
                                 uint64_t *hit, int dir_mode, float start_floor_ulps,
                                 cudaStream_t stream);

namespace {

constexpr int kReactThreads = 256;
constexpr int kReactWarps = kReactThreads / 32;

struct HitEval {
    double t, u, v;      // ray_u, trig_u, trig_v
    double h[3];         // hit point
    double den;
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
    e.den = den;
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

// The reference's validity rule for one (ray, triangle) pair (tfrt/geometry.py:306,
// tfrt/engine.py:1138-1141), in fp64.
struct ValidRule {
    double intersect_eps, size_eps, ray_start_eps;
    __device__ __forceinline__ bool operator()(const HitEval &e) const {
        return e.den_ok && e.u >= -size_eps && e.v >= -size_eps &&
               e.u + e.v <= 1.0 + size_eps && e.t >= ray_start_eps;
    }
};

struct TriArrays {
    const float *c[9];   // xp,yp,zp,x1,y1,z1,x2,y2,z2
    __device__ __forceinline__ void load(int i, double A[3], double B[3], double C[3]) const {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            A[k] = c[k][i];
            B[k] = c[3 + k][i];
            C[k] = c[6 + k][i];
        }
    }
};

// Shared-memory scratch of the block-cooperative exact search.
struct ResearchSmem {
    int turn;
    double ray[6];
    double wt[kReactWarps];
    int wi[kReactWarps];
};

// What the exact search needs to know about the mesh.
struct SearchMesh {
    TriArrays tri;
    const float4 *packed;     // tfrt_pack_triangles output (pair records)
    int n_tri, n_tiles;
    float w_one;              // C = 1 + 3 size_eps, the search's acceptance constant
};

// Exact nearest-hit search for the (rare) rays whose fp32 winner failed the fp64 re-validation.
// The whole CTA scans all pair records with the search's own projected fp32 test as a FILTER
// (slightly widened, any sign of the ray parameter), evaluates every triangle that passes it with
// the reference's Cramer solve in fp64, applies the reference's rule, and keeps the nearest
// (lowest index on exact ties, like the reference's CPU argmin).  Candidates are the few
// triangles whose plane the ray crosses inside the (inflated) triangle, so the fp64 work is a
// handful of evaluations per ray and the scan costs about as much as the search's own pass over
// the mesh for one ray.  Must be called by every thread of the CTA; threads with `fail` set
// get their new winner (or -1) back, all others -2.
__device__ int block_research(bool fail, const double S[3], const double E[3],
                              const SearchMesh &mesh, const ValidRule &rule, ResearchSmem &sm) {
    int result = -2;
    if (__syncthreads_or(fail ? 1 : 0) == 0) return result;   // CTA-uniform
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float c_wide = mesh.w_one * 1.002f;
    for (;;) {
        if (tid == 0) sm.turn = kReactThreads;
        __syncthreads();
        if (fail) atomicMin(&sm.turn, tid);
        __syncthreads();
        const int turn = sm.turn;
        if (turn == kReactThreads) break;                     // CTA-uniform
        if (tid == turn) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { sm.ray[c] = S[c]; sm.ray[3 + c] = E[c]; }
        }
        __syncthreads();
        const double s[3] = {sm.ray[0], sm.ray[1], sm.ray[2]};
        const double e[3] = {sm.ray[3], sm.ray[4], sm.ray[5]};
        const float o[3] = {(float)s[0], (float)s[1], (float)s[2]};
        const float d[3] = {(float)(e[0] - s[0]), (float)(e[1] - s[1]), (float)(e[2] - s[2])};
        double bt = 0.0;
        int bi = -1;
        const int n_records = mesh.n_tiles * TFRT_TILE_PAIRS;
        for (int q = tid; q < n_records; q += kReactThreads) {
            const int tile = q / TFRT_TILE_PAIRS, j = q - tile * TFRT_TILE_PAIRS;
            const float4 *rec0 = mesh.packed + (size_t)tile * TFRT_TILE_PAIRS * 6;
            const int header = __float_as_int(__ldg(&rec0[5]).z);
            const int n0 = header & 0xff, n1 = n0 + ((header >> 8) & 0xff);
            const int n2 = n1 + ((header >> 16) & 0xff);
            if (j >= n2) continue;                            // unused pair slot
            const int k = j < n0 ? 0 : (j < n1 ? 1 : 2);
            // ray components in the roles (k, k1, k2), selected without indexed registers
            const float ok = k == 0 ? o[0] : (k == 1 ? o[1] : o[2]);
            const float ok1 = k == 0 ? o[1] : (k == 1 ? o[2] : o[0]);
            const float ok2 = k == 0 ? o[2] : (k == 1 ? o[0] : o[1]);
            const float dk = k == 0 ? d[0] : (k == 1 ? d[1] : d[2]);
            const float dk1 = k == 0 ? d[1] : (k == 1 ? d[2] : d[0]);
            const float dk2 = k == 0 ? d[2] : (k == 1 ? d[0] : d[1]);
            const float4 *rec = rec0 + (size_t)j * 6;
            const float4 c0 = __ldg(rec + 0), c1 = __ldg(rec + 1), c2 = __ldg(rec + 2);
            const float4 c3 = __ldg(rec + 3), c4 = __ldg(rec + 4);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float m1 = h ? c0.y : c0.x, m2 = h ? c0.w : c0.z, dd = h ? c1.y : c1.x;
                const float s1 = h ? c1.w : c1.z, s2 = h ? c2.y : c2.x, s0 = h ? c2.w : c2.z;
                const float e1 = h ? c3.y : c3.x, e2 = h ? c3.w : c3.z, e0 = h ? c4.y : c4.x;
                const float det = fmaf(m2, dk2, fmaf(m1, dk1, dk));
                const float tq = fmaf(-m2, ok2, fmaf(-m1, ok1, dd - ok));
                const float t = tq / det;
                const float p1 = fmaf(t, dk1, ok1), p2 = fmaf(t, dk2, ok2);
                const float sv = fmaf(s2, p2, fmaf(s1, p1, s0));
                const float ev = fabsf(fmaf(e2, p2, fmaf(e1, p1, e0)));
                const float x = ev + fabsf(sv - ev);
                if (!(x <= c_wide)) continue;                 // NaN fails too
                const int i = __float_as_int(h ? c4.w : c4.z);
                if (i < 0 || i >= mesh.n_tri) continue;
                double A[3], B[3], C[3];
                mesh.tri.load(i, A, B, C);
                const HitEval hv = eval_hit(s, e, A, B, C, rule.intersect_eps);
                if (rule(hv) && (bi < 0 || hv.t < bt || (hv.t == bt && i < bi))) { bt = hv.t; bi = i; }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ot = __shfl_xor_sync(0xffffffffu, bt, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (oi >= 0 && (bi < 0 || ot < bt || (ot == bt && oi < bi))) { bt = ot; bi = oi; }
        }
        if (lane == 0) { sm.wt[warp] = bt; sm.wi[warp] = bi; }
        __syncthreads();
        if (tid == turn) {
            bt = sm.wt[0]; bi = sm.wi[0];
#pragma unroll
            for (int w = 1; w < kReactWarps; ++w) {
                const double ot = sm.wt[w];
                const int oi = sm.wi[w];
                if (oi >= 0 && (bi < 0 || ot < bt || (ot == bt && oi < bi))) { bt = ot; bi = oi; }
            }
            result = bi;
            fail = false;
        }
        __syncthreads();
    }
    return result;
}

// Re-validation of the fp32 search's winner with the reference's exact rule in fp64.
// Returns true when the winner stands.  `gross` is set when it fails by more than rounding can
// explain: the fp32 hit point is uncertain by a few ulps of the coordinates involved (the
// triangle's and the ray origin's), amplified by 1 / |cos incidence| for an oblique ray; a
// proposal whose fp64 hit lies farther than 16 search margins of that from the triangle (or from
// the ray-start threshold) is a search defect, not rounding, and tests assert it never happens.
__device__ __forceinline__ bool revalidate(const HitEval &e, const ValidRule &rule,
                                           const double S[3], const double E[3],
                                           const double A[3], const double B[3],
                                           const double C[3], double margin_ulps, bool &gross) {
    gross = false;
    if (rule(e)) return true;
    const double e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    const double e2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
    const double m[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2],
                         e1[0] * e2[1] - e1[1] * e2[0]};
    const double D[3] = {E[0] - S[0], E[1] - S[1], E[2] - S[2]};
    const double mD = fabs(m[0] * D[0] + m[1] * D[1] + m[2] * D[2]);
    const double mm = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
    const double dn = sqrt(D[0] * D[0] + D[1] * D[1] + D[2] * D[2]);
    if (!(mD > 0.0) || !(dn > 0.0)) return false;         // parallel / null ray: all rounding
    const double obliq = fmax(1.0, mm * dn / mD);         // 1 / |cos incidence|
    const double scale = fmax(tfrt_tri_scale(A, e1, e2),
                              fmax(fabs(S[0]), fmax(fabs(S[1]), fabs(S[2]))));
    // distance the fp32 hit point may be off, with a wide safety factor
    const double dist = 16.0 * fmax(margin_ulps, 4.0) * 1.1920928955078125e-7 * scale * obliq;
    const double slack_b = dist * tfrt_tri_bary_gradient(e1, e2);     // in barycentric units
    const double slack_t = dist / dn;                                 // in ray-parameter units
    gross = !e.den_ok || e.u < -(rule.size_eps + slack_b) || e.v < -(rule.size_eps + slack_b) ||
            e.u + e.v > 1.0 + rule.size_eps + 2.0 * slack_b ||
            e.t < 0.5 * rule.ray_start_eps - slack_t;
    return false;
}

// ------------------------------------------------------------------------------------------------
// single-pass order-preserving compaction: decoupled look-back over CTAs.
//
// A CTA's four category counts (active, stopped, finished, dead) travel in two 64-bit words,
//   word = flag << 62 | count_hi << 31 | count_lo      (counts < 2^31)
// flag 0 = not ready, 1 = this CTA's own counts, 2 = inclusive prefix over CTAs 0..b.  CTAs take
// their index from an atomic ticket, so every predecessor of a CTA has started and will publish:
// the look-back cannot deadlock whatever order the hardware schedules CTAs in.

constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62;
constexpr unsigned long long kValueMask = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive prefix (packed pair of counts) of CTA `bid` for state word `w`; one warp.
__device__ unsigned long long lookback(const unsigned long long *state, int w, int bid,
                                       int lane) {
    unsigned long long acc = 0;
    for (int base = bid - 1; base >= 0; base -= 32) {
        const int idx = base - lane;
        unsigned long long v = kFlagPrefix;                  // before CTA 0: prefix of zeros
        if (idx >= 0) {
            do { v = ld_state(state + 2 * (size_t)idx + w); } while ((v >> 62) == 0);
        }
        const unsigned pm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        const int first = pm ? __ffs(pm) - 1 : 32;           // nearest predecessor with a prefix
        unsigned long long contrib = (lane <= first) ? (v & kValueMask) : 0ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, off);
        acc += contrib;
        if (pm) break;
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------

struct ReactParams {
    int n_tri, n_optical, index_mode, n_materials;
    double intersect_eps, size_eps, ray_start_eps, new_ray_length, margin_ulps;
    const float *src_wavelength;
    TriArrays tri;
    const float4 *packed;
    int n_tiles;
    float w_one;
    const float *norm;
    const int32_t *category, *mat_in, *mat_out;
    const float *n_in, *n_out;
    const TfrtMaterial *materials;
    const double *n_table;     // (n_materials, n_lambda) or NULL
    const int32_t *lambda_index;   // per source ray, when n_table is given
    int n_lambda;
    float *cur[6];
    double *cur_d;             // 6*N: the same state in fp64 (the chain between iterations)
    uint64_t *hit;
    const int32_t *live;       // live[k]
    const int32_t *counters_k; // counters[k]
    int32_t *counters_next;    // counters[k+1]
    float *tape_f;             // tape_f + k*9*N
    int32_t *tape_tri;         // + k*N
    uint8_t *tape_status;      // + k*N
    int32_t *tape_outpos;      // + k*N
    int32_t *live_next;
    int32_t *cat_ids[3];
    unsigned long long *scan_state;   // this iteration's look-back words (2 per CTA)
    unsigned int *ticket;             // this iteration's CTA ticket
    int32_t *diag;
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

// Stage A of an iteration's O(rays) work: evaluate the search's winner in fp64, apply the
// reference's exact rule, and search again (exactly) for the rare rays whose winner fails it.
// Leaves hit[p] = validated winner (key 0) or "no hit".  A kernel of its own, so that CTAs that
// have to search again delay nobody: the compaction of k_react3d is an ordered chain over CTAs
// (look-back), and a CTA that took tens of microseconds there would stall every CTA behind it.
__global__ void __launch_bounds__(kReactThreads) k_validate3d(ReactParams P) {
    __shared__ ResearchSmem s_research;
    const int n_live = P.counters_k[0];
    if (blockIdx.x * kReactThreads >= n_live) return;   // whole CTA idle (uniform)
    const int p = blockIdx.x * kReactThreads + threadIdx.x;
    const size_t N = (size_t)P.n_rays;
    const ValidRule rule{P.intersect_eps, P.size_eps, P.ray_start_eps};
    int tri = -1;
    double S[3] = {0, 0, 0}, E[3] = {0, 0, 0};
    bool fail = false;
    if (p < n_live) {
        const int slot = P.live[p];
        const uint64_t hit = P.hit[p];
        // working state = (start, direction) in fp64; end = start + direction
        for (int c = 0; c < 3; ++c) {
            S[c] = P.cur_d[c * N + slot];
            E[c] = S[c] + P.cur_d[(3 + c) * N + slot];
        }
        if ((uint32_t)(hit >> 32) < TFRT_NOHIT_KEY) {
            tri = (int)(uint32_t)hit;
            double A[3], B[3], C[3];
            P.tri.load(tri, A, B, C);
            const HitEval e = eval_hit(S, E, A, B, C, P.intersect_eps);
            bool gross;
            fail = !revalidate(e, rule, S, E, A, B, C, P.margin_ulps, gross);
            if (fail) {
                atomicAdd(&P.diag[0], 1);
                if (gross) atomicAdd(&P.diag[1], 1);
            }
        }
    }
    SearchMesh mesh;
    mesh.tri = P.tri; mesh.packed = P.packed; mesh.n_tri = P.n_tri; mesh.n_tiles = P.n_tiles;
    mesh.w_one = P.w_one;
    const int found = block_research(fail, S, E, mesh, rule, s_research);
    if (fail) {
        tri = found;
        if (found >= 0) atomicAdd(&P.diag[2], 1);
        P.hit[p] = tri >= 0 ? (uint64_t)(uint32_t)tri : ~0ull;
    }
}

// Stage B + C: classification, fused material-index + Snell reaction, tape, and the
// order-preserving compaction of the live set (single pass, decoupled look-back).
__global__ void __launch_bounds__(kReactThreads) k_react3d(ReactParams P) {
    __shared__ int s_bid;
    __shared__ int s_warp[4][kReactWarps];
    __shared__ unsigned long long s_excl[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_live = P.counters_k[0];
    const int n_blocks = (n_live + kReactThreads - 1) / kReactThreads;
    if (tid == 0) s_bid = (int)atomicAdd(P.ticket, 1u);
    __syncthreads();
    const int bid = s_bid;
    if (n_blocks == 0) {   // nothing live: the loop has ended (tfrt/engine.py:2324-2330)
        if (bid == 0 && tid < 4) P.counters_next[tid] = tid == 0 ? 0 : P.counters_k[tid];
        return;
    }
    if (bid >= n_blocks) return;   // whole CTA idle (uniform)
    const int p = bid * kReactThreads + tid;
    const size_t N = (size_t)P.n_rays;

    int slot = -1, tri = -1;
    double S[3] = {0, 0, 0}, E[3] = {0, 0, 0};
    if (p < n_live) {
        slot = P.live[p];
        const uint64_t hit = P.hit[p];            // validated by k_validate3d
        P.hit[p] = ~0ull;                         // preset for the next iteration's atomicMin merge
        for (int c = 0; c < 3; ++c) {
            S[c] = P.cur_d[c * N + slot];
            E[c] = S[c] + P.cur_d[(3 + c) * N + slot];
        }
        if ((uint32_t)(hit >> 32) < TFRT_NOHIT_KEY) tri = (int)(uint32_t)hit;
    }

    // ---- stage B: classification, reaction, tape
    int status = -1;
    if (p < n_live) {
        float *tf = P.tape_f;
        double Pt[3] = {E[0], E[1], E[2]};
        double t = 0.0, u = 0.0, v = 0.0;
        status = TFRT_STATUS_DEAD;
        if (tri >= 0) {
            double A[3], B[3], C[3];
            P.tri.load(tri, A, B, C);
            const HitEval e = eval_hit(S, E, A, B, C, P.intersect_eps);
            t = e.t; u = e.u; v = e.v;
            Pt[0] = e.h[0]; Pt[1] = e.h[1]; Pt[2] = e.h[2];
            status = P.category[tri];   // OPTICAL=0 -> ACTIVE, STOP=1, TARGET=2
            if (status == TFRT_STATUS_ACTIVE) {
                double n_in, n_out;
                if (P.n_table) {
                    // user-supplied index functions, tabulated over the distinct wavelengths
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
                // of being re-derived from two rounded end points.  The fp64 copy carries the ray
                // from one iteration to the next, so rounding does not compound over bounces; the
                // fp32 copy feeds the intersection search.
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

    // ---- stage C: order-preserving compaction
    int rank_in_warp = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const unsigned m = __ballot_sync(0xffffffffu, status == c);
        if (status == c) rank_in_warp = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[c][warp] = __popc(m);
    }
    __syncthreads();
    if (warp == 0) {
        // this CTA's counts, published first (flag 1), then the look-back, then the prefix (flag 2)
        unsigned long long own[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            unsigned long long lo = 0, hi = 0;
#pragma unroll
            for (int q = 0; q < kReactWarps; ++q) { lo += s_warp[2 * w][q]; hi += s_warp[2 * w + 1][q]; }
            own[w] = lo | (hi << 31);
        }
        unsigned long long *mine = P.scan_state + 2 * (size_t)bid;
        if (lane == 0 && bid > 0) { st_state(mine, kFlagAgg | own[0]); st_state(mine + 1, kFlagAgg | own[1]); }
        unsigned long long excl[2];
        excl[0] = lookback(P.scan_state, 0, bid, lane);
        excl[1] = lookback(P.scan_state, 1, bid, lane);
        if (lane == 0) {
            st_state(mine, kFlagPrefix | (excl[0] + own[0]));
            st_state(mine + 1, kFlagPrefix | (excl[1] + own[1]));
            s_excl[0] = excl[0];
            s_excl[1] = excl[1];
            if (bid == n_blocks - 1) {   // totals of this iteration
                const unsigned long long t0 = excl[0] + own[0], t1 = excl[1] + own[1];
                P.counters_next[0] = (int32_t)(t0 & 0x7fffffffull);
                P.counters_next[1] = P.counters_k[1] + (int32_t)(t0 >> 31);
                P.counters_next[2] = P.counters_k[2] + (int32_t)(t1 & 0x7fffffffull);
                P.counters_next[3] = P.counters_k[3] + (int32_t)(t1 >> 31);
            }
        }
    }
    __syncthreads();
    if (status < 0) return;
    const unsigned long long ex = s_excl[status >> 1];
    int pos = (int)((status & 1) ? (ex >> 31) : (ex & 0x7fffffffull)) + rank_in_warp;
    for (int w = 0; w < warp; ++w) pos += s_warp[status][w];
    if (status == TFRT_STATUS_ACTIVE) {
        P.live_next[pos] = slot;
        P.tape_outpos[slot] = pos;
    } else {
        const int gpos = P.counters_k[status] + pos;
        P.cat_ids[status - 1][gpos] = slot;
        P.tape_outpos[slot] = gpos;
    }
}

// One launch sets up a trace: counters, diagnostics, per-iteration work counters / tickets /
// look-back words, the hit presets, the working ray state (fp32 + fp64), the first live list and
// the "not live" marks of the status tape.
struct InitParams {
    const float *src[6];
    float *cur[6];
    double *cur_d;
    int32_t *live0;
    uint64_t *hit;
    uint8_t *tape_status;
    int32_t *counters;
    int32_t *diag;
    uint32_t *sync_words;     // work counters, tickets, padding, look-back words
    size_t n_sync_words;
    int n_rays, n_iters;
};

__global__ void __launch_bounds__(256) k_trace_init(InitParams P) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (i < (size_t)(P.n_iters + 1) * 4) P.counters[i] = (i == 0) ? P.n_rays : 0;
    if (i < 4) P.diag[i] = 0;
    for (size_t w = i; w < P.n_sync_words; w += stride) P.sync_words[w] = 0u;
    const size_t N = (size_t)P.n_rays;
    if (i >= N) return;
    // working state = (start, direction); the fp64 difference of two floats is exact
    const double a0 = P.src[0][i], a1 = P.src[1][i], a2 = P.src[2][i];
    const double b0 = (double)P.src[3][i] - a0, b1 = (double)P.src[4][i] - a1,
                 b2 = (double)P.src[5][i] - a2;
    P.cur_d[i] = a0; P.cur_d[N + i] = a1; P.cur_d[2 * N + i] = a2;
    P.cur_d[3 * N + i] = b0; P.cur_d[4 * N + i] = b1; P.cur_d[5 * N + i] = b2;
    P.cur[0][i] = (float)a0; P.cur[1][i] = (float)a1; P.cur[2][i] = (float)a2;
    P.cur[3][i] = (float)b0; P.cur[4][i] = (float)b1; P.cur[5][i] = (float)b2;
    P.live0[i] = (int32_t)i;
    P.hit[i] = ~0ull;
    for (int k = 0; k < P.n_iters; ++k) P.tape_status[(size_t)k * N + i] = 0xff;
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

// b1 seam (tfrt_intersect3d): presets of the search's workspace
__global__ void k_seam_init(uint64_t *hit, int n, int32_t *diag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hit[i] = ~0ull;
    if (i == 0) hit[n] = 0ull;            // the work counter lives in the word after the hits
    if (diag && i < 4) diag[i] = 0;
}

// per-ray evaluation for the b1 seam: fp64 evaluation + re-validation of the winner
__global__ void __launch_bounds__(kReactThreads)
k_intersect_outputs(const float *sx, const float *sy, const float *sz, const float *ex,
                    const float *ey, const float *ez, SearchMesh mesh,
                    const uint64_t *hit, int n, ValidRule rule, double margin_ulps, float *x,
                    float *y, float *z, uint8_t *valid, float *ray_u, float *trig_u,
                    float *trig_v, int32_t *gather_trig, int32_t *diag) {
    __shared__ ResearchSmem s_research;
    const int i = blockIdx.x * kReactThreads + threadIdx.x;
    double S[3] = {0, 0, 0}, E[3] = {0, 0, 0};
    int t_idx = -1;
    bool fail = false;
    if (i < n) {
        const uint64_t h = hit[i];
        S[0] = sx[i]; S[1] = sy[i]; S[2] = sz[i];
        E[0] = ex[i]; E[1] = ey[i]; E[2] = ez[i];
        if ((uint32_t)(h >> 32) < TFRT_NOHIT_KEY) {
            t_idx = (int)(uint32_t)h;
            double A[3], B[3], C[3];
            mesh.tri.load(t_idx, A, B, C);
            const HitEval e = eval_hit(S, E, A, B, C, rule.intersect_eps);
            bool gross;
            fail = !revalidate(e, rule, S, E, A, B, C, margin_ulps, gross);
            if (fail && diag) {
                atomicAdd(&diag[0], 1);
                if (gross) atomicAdd(&diag[1], 1);
            }
        }
    }
    const int found = block_research(fail, S, E, mesh, rule, s_research);
    if (fail) {
        t_idx = found;
        if (found >= 0 && diag) atomicAdd(&diag[2], 1);
    }
    if (i >= n) return;
    const bool ok = t_idx >= 0;
    const int g = ok ? t_idx : 0;   // argmin of an all-invalid column is 0
    double A[3], B[3], C[3];
    mesh.tri.load(g, A, B, C);
    const HitEval e = eval_hit(S, E, A, B, C, rule.intersect_eps);
    x[i] = (float)e.h[0]; y[i] = (float)e.h[1]; z[i] = (float)e.h[2];
    valid[i] = ok ? 1 : 0;
    ray_u[i] = (float)e.t; trig_u[i] = (float)e.u; trig_v[i] = (float)e.v;
    gather_trig[i] = g;
}

// The fp32 search excludes hits nearer than this fraction of ray_start_epsilion; the fp64
// re-validation applies the exact threshold, so a hit in between costs an exact re-search
// instead of being lost to the rounding of the search's ray parameter.
constexpr float kSearchStartFraction = 0.75f;
// ... and nothing nearer than this many search margins (in ulps of the scene's coordinates, as a
// distance): below that the search cannot tell a hit from the rounding of a ray that starts on
// a surface.
constexpr float kStartFloorPerMargin = 4.0f;

}  // namespace

// ------------------------------------------------------------------------------------------------

extern "C" int tfrt_intersect3d(const float *const ray[6], int32_t n_rays,
                                const float *const tri[9], const float *packed, int32_t n_tri,
                                float intersect_eps, float size_eps, float ray_start_eps,
                                float margin_ulps, uint64_t *hit_ws, float *x, float *y, float *z,
                                uint8_t *valid, float *ray_u, float *trig_u, float *trig_v,
                                int32_t *gather_trig, int32_t *diag, const void *grid,
                                int32_t grid_max_cells, void *stream) {
    TFRT_REQUIRE(n_rays >= 0 && n_tri > 0 && packed && hit_ws, "tfrt_intersect3d: bad arguments");
    if (n_rays == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    k_seam_init<<<tfrt_div_up(n_rays + 1, 256), 256, 0, st>>>(hit_ws, n_rays, diag);
    int rc = grid
        ? tfrt_launch_grid_intersect3d(ray, nullptr, nullptr, n_rays, packed, n_tri, grid,
                                       grid_max_cells, intersect_eps,
                                       kSearchStartFraction * ray_start_eps, size_eps, hit_ws,
                                       /*dir_mode=*/0, kStartFloorPerMargin * margin_ulps, st)
        : tfrt_launch_intersect3d(ray, nullptr, nullptr, n_rays, packed, n_tri, intersect_eps,
                                  kSearchStartFraction * ray_start_eps, size_eps, hit_ws,
                                  /*dir_mode=*/0,
                                  reinterpret_cast<unsigned int *>(hit_ws + n_rays),
                                  kStartFloorPerMargin * margin_ulps, st);
    if (rc) return rc;
    SearchMesh M;
    for (int c = 0; c < 9; ++c) M.tri.c[c] = tri[c];
    M.packed = reinterpret_cast<const float4 *>(packed);
    M.n_tri = n_tri; M.n_tiles = tfrt_div_up(n_tri, TFRT_TILE_TRIS);
    M.w_one = 1.0f + 3.0f * size_eps;
    const ValidRule rule{(double)intersect_eps, (double)size_eps, (double)ray_start_eps};
    k_intersect_outputs<<<tfrt_div_up(n_rays, kReactThreads), kReactThreads, 0, st>>>(
        ray[0], ray[1], ray[2], ray[3], ray[4], ray[5], M, hit_ws, n_rays, rule,
        (double)margin_ulps, x, y, z, valid, ray_u, trig_u, trig_v, gather_trig, diag);
    tfrt_note_launch(2);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// sync workspace of a trace, in 32-bit words: K work counters, K CTA tickets, padding to 16 bytes,
// then 2 look-back words (64-bit) per react CTA per iteration
static size_t sync_header_words(int K) { return ((size_t)2 * K + 3) / 4 * 4; }

extern "C" size_t tfrt_trace3d_sync_bytes(int32_t n_rays, int32_t max_iterations) {
    if (n_rays < 0 || max_iterations <= 0) return 0;
    const size_t nb = (size_t)tfrt_div_up(n_rays > 0 ? n_rays : 1, kReactThreads);
    return 4 * (sync_header_words(max_iterations) + 4 * nb * (size_t)max_iterations);
}

extern "C" int tfrt_trace3d_forward(const TfrtTrace3DArgs *a, void *stream) {
    TFRT_REQUIRE(a, "tfrt_trace3d_forward: null args");
    TFRT_REQUIRE(a->n_rays >= 0 && a->n_tri > 0 && a->max_iterations > 0,
                 "tfrt_trace3d_forward: n_rays=%d n_tri=%d max_iterations=%d", a->n_rays,
                 a->n_tri, a->max_iterations);
    TFRT_REQUIRE(a->cur_d, "tfrt_trace3d_forward: null cur_d workspace");
    TFRT_REQUIRE(a->packed && a->category && a->hit && a->tape_f && a->tape_tri &&
                     a->tape_status && a->tape_outpos && a->live && a->sync_ws && a->counters &&
                     a->diag,
                 "tfrt_trace3d_forward: null workspace pointer");
    TFRT_REQUIRE(a->n_optical == 0 || a->n_rays == 0 ||
                     (a->n_table ? (a->lambda_index && a->mat_in && a->mat_out && a->n_lambda > 0)
                      : a->index_mode ? (a->materials && a->mat_in && a->mat_out && a->src[6])
                                      : (a->n_in && a->n_out)),
                 "tfrt_trace3d_forward: material inputs missing for index_mode=%d", a->index_mode);
    cudaStream_t st = (cudaStream_t)stream;
    const int N = a->n_rays, K = a->max_iterations;
    const size_t nb = (size_t)tfrt_div_up(N > 0 ? N : 1, kReactThreads);
    const size_t header = sync_header_words(K);
    uint32_t *sync = reinterpret_cast<uint32_t *>(a->sync_ws);
    unsigned int *work_counters = sync;
    unsigned int *tickets = sync + K;
    unsigned long long *scan = reinterpret_cast<unsigned long long *>(sync + header);

    InitParams I;
    for (int c = 0; c < 6; ++c) { I.src[c] = a->src[c]; I.cur[c] = a->cur[c]; }
    I.cur_d = a->cur_d; I.live0 = a->live; I.hit = a->hit; I.tape_status = a->tape_status;
    I.counters = a->counters; I.diag = a->diag;
    I.sync_words = sync; I.n_sync_words = header + 4 * nb * (size_t)K;
    I.n_rays = N; I.n_iters = K;
    const int init_threads = N > (K + 1) * 4 ? N : (K + 1) * 4;
    k_trace_init<<<tfrt_div_up(init_threads, 256), 256, 0, st>>>(I);
    tfrt_note_launch(1);
    if (N == 0) {
        TFRT_CHECK_CUDA(cudaGetLastError());
        return 0;
    }

    const int n_blocks = tfrt_div_up(N, kReactThreads);
    for (int k = 0; k < K; ++k) {
        const int32_t *live_k = a->live + (size_t)k * N;
        const int32_t *counters_k = a->counters + k * 4;
        const float *cur_c[6] = {a->cur[0], a->cur[1], a->cur[2], a->cur[3], a->cur[4], a->cur[5]};
        int rc = a->grid
            ? tfrt_launch_grid_intersect3d(cur_c, live_k, counters_k, N, a->packed, a->n_tri,
                                           a->grid, a->grid_max_cells, a->intersect_eps,
                                           kSearchStartFraction * a->ray_start_eps, a->size_eps,
                                           a->hit, /*dir_mode=*/1,
                                           kStartFloorPerMargin * a->margin_ulps, st)
            : tfrt_launch_intersect3d(cur_c, live_k, counters_k, N, a->packed, a->n_tri,
                                      a->intersect_eps, kSearchStartFraction * a->ray_start_eps,
                                      a->size_eps, a->hit, /*dir_mode=*/1, work_counters + k,
                                      kStartFloorPerMargin * a->margin_ulps, st);
        if (rc) return rc;
        ReactParams P;
        P.n_tri = a->n_tri; P.n_optical = a->n_optical; P.index_mode = a->index_mode;
        P.n_materials = a->n_materials;
        P.intersect_eps = a->intersect_eps; P.size_eps = a->size_eps;
        P.ray_start_eps = a->ray_start_eps; P.new_ray_length = a->new_ray_length;
        P.margin_ulps = a->margin_ulps;
        P.src_wavelength = a->src[6];
        for (int c = 0; c < 9; ++c) P.tri.c[c] = a->tri[c];
        P.packed = reinterpret_cast<const float4 *>(a->packed);
        P.n_tiles = tfrt_div_up(a->n_tri, TFRT_TILE_TRIS);
        P.w_one = 1.0f + 3.0f * a->size_eps;
        P.norm = a->norm; P.category = a->category; P.mat_in = a->mat_in; P.mat_out = a->mat_out;
        P.n_in = a->n_in; P.n_out = a->n_out; P.materials = a->materials;
        P.n_table = a->n_table; P.lambda_index = a->lambda_index; P.n_lambda = a->n_lambda;
        for (int c = 0; c < 6; ++c) P.cur[c] = a->cur[c];
        P.cur_d = a->cur_d;
        P.hit = a->hit; P.live = live_k; P.counters_k = counters_k;
        P.counters_next = a->counters + (k + 1) * 4;
        P.tape_f = a->tape_f + (size_t)k * 9 * N;
        P.tape_tri = a->tape_tri + (size_t)k * N;
        P.tape_status = a->tape_status + (size_t)k * N;
        P.tape_outpos = a->tape_outpos + (size_t)k * N;
        P.live_next = a->live + (size_t)(k + 1) * N;
        for (int c = 0; c < 3; ++c) P.cat_ids[c] = a->cat_ids[c];
        P.scan_state = scan + 2 * nb * (size_t)k;
        P.ticket = tickets + k;
        P.diag = a->diag;
        P.n_rays = N;
        k_validate3d<<<n_blocks, kReactThreads, 0, st>>>(P);
        k_react3d<<<n_blocks, kReactThreads, 0, st>>>(P);
        tfrt_note_launch(2);
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
