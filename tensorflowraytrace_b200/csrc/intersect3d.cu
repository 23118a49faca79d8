// intersect3d.cu -- K1a: brute-force nearest ray x triangle hit (sm_100a).
//
// Replaces the dense (M,N) evaluation of geometry.line_triangle_intersect
// (tfrt/geometry.py:191-320) + pruning + argmin of OpticalSystem3D._intersection
// (tfrt/engine.py:1138-1148).  Every live ray is tested against every triangle.
//
// Design
//  * a thread owns 2*R2 rays kept in registers as R2 f32x2 pairs (origin, direction);
//  * triangles stream through shared memory in tiles of TFRT_TILE_TRIS, copied by one elected
//    thread with cp.async.bulk (TMA 1-D) into a 3-stage mbarrier ring;
//  * per (ray, triangle) the projected precomputed-plane test (Wald's projection onto the
//    dominant axis k of the normal, plane scaled to a unit k-coefficient; record layout in
//    common.cuh) costs 14 fp32 ops issued as 7 FFMA2/FADD2/FMUL2 per ray (two rays per
//    instruction) + 1 MUFU.RCP + 2 integer compares:
//        det = D[k] + m1 D[k1] + m2 D[k2]      t' = dd - O[k] - m1 O[k1] - m2 O[k2]
//        t   = t'/det     (origin pre-shifted by ray_start_eps so t >= 0 <=> valid)
//        P1  = O[k1] + t D[k1]                 P2 = O[k2] + t D[k2]
//        S   = s1 P1 + s2 P2 + s0  (= 2(u+v) - C)      E = e1 P1 + e2 P2 + e0  (= u - v)
//        x   = |E| + |S - |E||     (absolute values are free operand modifiers of FADD2)
//    with u, v the barycentrics inflated by size_eps and C = 1 + 3 size_eps:
//        inside  <=>  u >= 0, v >= 0, u + v <= C  <=>  x <= C
//    (min(u,v) = (u+v-|u-v|)/2 and min(p,q) >= 0 <=> p+q >= |p-q|, applied twice), so
//    valid & nearer  <=>  bits(x) <= bits(C)  and  bits(t) < best   as unsigned integers
//    (negative t, NaN and +inf all compare high);
//    the triangles of a tile are grouped by k, so three specialised loops run back to back
//    with the roles of the ray components fixed at compile time;
//  * the running best is (bits(t), triangle index); the scan is in index order inside an axis
//    group and strict '<' keeps the first, so exact ties between triangles of one group (e.g.
//    the coplanar halves of a flat target) go to the lowest index, the behaviour of argmin in
//    the reference's CPU kernels; an exact tie between triangles of different dominant axes of
//    one tile goes to the lower axis (a bit-equal t from two different planes: a near tie by
//    the north star's rule);
//  * optional split of the triangle range over blockIdx.y, merged with a 64-bit atomicMin on
//    (key << 32 | index) -- same tie rule -- so small ray counts still fill 148 SMs.
#include <atomic>
#include <type_traits>
#include <cstdlib>

#include "common.cuh"

// Tunables.  The defaults are the measured best of the variants timed by tools/k1_variants.sh
// (profiles/r1_k1_tuning.md); tools/k1_bench.cu compiles the others with -D.
#ifndef TFRT_K1_THREADS
#define TFRT_K1_THREADS 128
#endif
#ifndef TFRT_K1_STAGES
#define TFRT_K1_STAGES 3
#endif
#ifndef TFRT_K1_UNROLL
#define TFRT_K1_UNROLL 2
#endif
#ifndef TFRT_K1_FORCE_R2    // 0: choose rays per thread from the ray count
#define TFRT_K1_FORCE_R2 0
#endif
#ifndef TFRT_K1_MINBLOCKS   // 4 CTAs of 128 threads per SM <=> at most 128 registers
#define TFRT_K1_MINBLOCKS 4
#endif

namespace {

constexpr int kThreads = TFRT_K1_THREADS;
constexpr int kStages = TFRT_K1_STAGES;
constexpr int kUnroll = TFRT_K1_UNROLL;

constexpr int kItemQueue = 8;   // published work items in flight (the producer is at most
                                // kStages + 1 items ahead of the slowest consumer warp)

struct __align__(16) IntersectSmem {
    float4 tile[kStages][TFRT_TILE_TRIS * 3];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t item_ready[kItemQueue];
    int32_t item[kItemQueue];
    // producer state, touched by thread 0 only (kept out of every thread's registers)
    uint32_t p_g;          // next tile sequence number to issue
    int32_t p_items;       // items published so far
    int32_t p_tile, p_tiles, p_tile_begin;
    int32_t p_done;
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
// packed f32x2 arithmetic (SASS FFMA2 / FMUL2 / FADD2); TFRT_K1_SCALAR=1 builds the same kernel from
// scalar FFMA / FMUL / FADD for the tuning study (profiles/r1_k1_tuning.md)
#ifndef TFRT_K1_SCALAR
#define TFRT_K1_SCALAR 0
#endif
#if TFRT_K1_SCALAR
__device__ __forceinline__ float2 k_fma2(float2 a, float2 b, float2 c) {
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
}
__device__ __forceinline__ float2 k_mul2(float2 a, float2 b) {
    return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
}
__device__ __forceinline__ float2 k_add2(float2 a, float2 b) {
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}
#else
__device__ __forceinline__ float2 k_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 k_mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 k_add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
#endif
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 abs2(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }

template <int R2, bool kQueue>
__global__ void __launch_bounds__(kThreads, (R2 >= 4 ? TFRT_K1_MINBLOCKS : 1))
k_intersect3d(const float *__restrict__ sx, const float *__restrict__ sy,
              const float *__restrict__ sz, const float *__restrict__ ex,
              const float *__restrict__ ey, const float *__restrict__ ez,
              const int32_t *__restrict__ live, const int32_t *__restrict__ n_live_ptr,
              int n_live_static, const float4 *__restrict__ packed, int n_tiles,
              int tiles_per_split, int splits, float intersect_eps, float ray_start_eps,
              float w_one, unsigned long long *__restrict__ hit, int use_atomic, int dir_mode,
              unsigned int *__restrict__ work_counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    IntersectSmem &sm = *reinterpret_cast<IntersectSmem *>(smem_raw);

    // Persistent CTAs pull work items (ray block x triangle split) from a global counter, so
    // the tail of a launch is one item, not one wave, whatever the live count turns out to be
    // (it is only known on the device).  Consecutive items are different ray blocks of the
    // same split: neighbouring CTAs stream the same triangle tiles out of L2.
    //
    // kQueue (long items, 8 rays per thread): no CTA-wide barrier separates items.  The elected
    // producer thread fetches the next item when its tile stream reaches it, publishes the id
    // through a small mbarrier-guarded queue and keeps the TMA ring kStages tiles ahead ACROSS
    // item boundaries; each warp crosses into the next item on its own (its rays, its epilogue),
    // so warps only ever wait for tiles.  The price is that a CTA claims items up to kStages
    // tiles before it starts them, which unbalances launches made of few, short items.
    // !kQueue (short items): one item at a time, fetched between two __syncthreads.
    const int n_live = n_live_ptr ? *n_live_ptr : n_live_static;
    const int rays_per_cta = kThreads * 2 * R2;
    const int ray_blocks = (n_live + rays_per_cta - 1) / rays_per_cta;
    const int total_items = ray_blocks * splits;
    const int tid = threadIdx.x;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kThreads / 32);
        }
        if constexpr (kQueue) {
#pragma unroll
            for (int q = 0; q < kItemQueue; ++q) mbar_init(&sm.item_ready[q], 1);
            sm.p_g = 0; sm.p_items = 0; sm.p_tile = 0; sm.p_tiles = 0; sm.p_tile_begin = 0;
            sm.p_done = 0;
        }
        mbar_fence_init();
    }
    if constexpr (kQueue) __syncthreads();

    // ---- producer, kQueue (thread 0 only): issue the next tile of the stream, fetching items
    // as the stream reaches them
    auto produce_one = [&]() {
        if (sm.p_done) return;
        if (sm.p_tile == sm.p_tiles) {
            const int it = (int)atomicAdd(work_counter, 1u);
            const int q = sm.p_items % kItemQueue;
            sm.item[q] = it;
            mbar_arrive(&sm.item_ready[q]);          // release: publishes sm.item[q]
            ++sm.p_items;
            if (it >= total_items) { sm.p_done = 1; return; }
            sm.p_tile_begin = (it / ray_blocks) * tiles_per_split;
            sm.p_tiles = min(n_tiles, sm.p_tile_begin + tiles_per_split) - sm.p_tile_begin;
            sm.p_tile = 0;
        }
        const uint32_t g = sm.p_g;
        const int s = g % kStages;
        if (g >= (uint32_t)kStages) mbar_wait(&sm.empty[s], ((g / kStages) - 1) & 1);
        mbar_arrive_expect_tx(&sm.full[s], TFRT_TILE_BYTES);
        bulk_g2s(sm.tile[s], packed + (size_t)(sm.p_tile_begin + sm.p_tile) * TFRT_TILE_TRIS * 3,
                 TFRT_TILE_BYTES, &sm.full[s]);
        ++sm.p_tile;
        sm.p_g = g + 1;
    };
    // ---- producer, !kQueue: tile `g` of the CTA's running sequence <- global tile `tile`
    auto issue = [&](uint32_t g, int tile) {
        const int s = g % kStages;
        if (g >= (uint32_t)kStages) mbar_wait(&sm.empty[s], ((g / kStages) - 1) & 1);
        mbar_arrive_expect_tx(&sm.full[s], TFRT_TILE_BYTES);
        bulk_g2s(sm.tile[s], packed + (size_t)tile * TFRT_TILE_TRIS * 3, TFRT_TILE_BYTES,
                 &sm.full[s]);
    };
    if constexpr (kQueue) {
        if (tid == 0)
            for (int i = 0; i < kStages; ++i) produce_one();   // fill the ring
    }

    const uint32_t c_bits = __float_as_uint(w_one);   // C = 1 + 3*size_eps (eps is folded into the planes)
    uint32_t seq = 0;                       // tiles this warp (CTA) has consumed so far
    for (int item_seq = 0;; ++item_seq) {
        int item;
        if constexpr (kQueue) {
            const int q = item_seq % kItemQueue;
            mbar_wait(&sm.item_ready[q], (uint32_t)((item_seq / kItemQueue) & 1));
            item = sm.item[q];
        } else {
            __syncthreads();                // previous item done; barriers initialised
            if (tid == 0) sm.item[0] = (int)atomicAdd(work_counter, 1u);
            __syncthreads();
            item = sm.item[0];
        }
        if (item >= total_items) break;
        const int base = (item % ray_blocks) * rays_per_cta;
        const int tile_begin = (item / ray_blocks) * tiles_per_split;
        const int my_tiles = min(n_tiles, tile_begin + tiles_per_split) - tile_begin;
        if constexpr (!kQueue) {
            if (tid == 0) {                 // producer prologue: fill the ring
                const int pre = min(kStages, my_tiles);
                for (int i = 0; i < pre; ++i) issue(seq + i, tile_begin + i);
            }
        }

        // ---- load this thread's rays: pair r = positions (2r, 2r+1) * kThreads + tid
        float2 O[3][R2], D[3][R2];
        uint32_t best[2 * R2];
        int32_t bidx[2 * R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) {
            float o[2][3], d[2][3];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = base + (2 * r + h) * kThreads + tid;
                if (p < n_live) {
                    const int slot = live ? live[p] : p;
                    const float ax = sx[slot], ay = sy[slot], az = sz[slot];
                    // dir_mode: the last three arrays hold the direction (end - start) itself,
                    // the form the trace loop keeps between iterations (no cancellation error)
                    d[h][0] = dir_mode ? ex[slot] : ex[slot] - ax;
                    d[h][1] = dir_mode ? ey[slot] : ey[slot] - ay;
                    d[h][2] = dir_mode ? ez[slot] : ez[slot] - az;
                    // shift the origin by ray_start_eps along the ray:
                    // t_shifted >= 0 <=> ray_u >= eps
                    o[h][0] = fmaf(ray_start_eps, d[h][0], ax);
                    o[h][1] = fmaf(ray_start_eps, d[h][1], ay);
                    o[h][2] = fmaf(ray_start_eps, d[h][2], az);
                } else {  // padding lane: zero direction => det == 0 => never a hit
                    o[h][0] = o[h][1] = o[h][2] = 0.f;
                    d[h][0] = d[h][1] = d[h][2] = 0.f;
                }
                best[2 * r + h] = TFRT_NOHIT_KEY;
                bidx[2 * r + h] = 0;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                O[c][r] = f2(o[0][c], o[1][c]);
                D[c][r] = f2(d[0][c], d[1][c]);
            }
        }

        for (int it = 0; it < my_tiles; ++it) {
            const uint32_t g = seq + it;
            const int s = g % kStages;
            mbar_wait(&sm.full[s], (g / kStages) & 1);
            const float4 *__restrict__ tile = sm.tile[s];

            // One specialised loop per dominant axis K of the tile's three triangle groups: the
            // ray components taking the roles (k, k1, k2) are fixed at compile time.
            auto run_group = [&](auto kc, int j0, int j1) {
                constexpr int K = decltype(kc)::value, K1 = (K + 1) % 3, K2 = (K + 2) % 3;
#pragma unroll kUnroll
                for (int j = j0; j < j1; ++j) {
                    const float4 q0 = tile[j * 3 + 0];
                    const float4 q1 = tile[j * 3 + 1];
                    const float4 q2 = tile[j * 3 + 2];
                    // broadcast operands (SASS: R.F32): one register feeds both rays of a pair
                    const float2 m1 = f2(q0.x, q0.x), m2 = f2(q0.y, q0.y), dd = f2(q0.z, q0.z);
                    const float2 s1 = f2(q1.x, q1.x), s2 = f2(q1.y, q1.y), s0 = f2(q1.z, q1.z);
                    const float2 e1 = f2(q2.x, q2.x), e2 = f2(q2.y, q2.y), e0 = f2(q2.z, q2.z);
                    float2 tt[R2], xx[R2], dets[R2];
                    bool cand = false;
#pragma unroll
                    for (int r = 0; r < R2; ++r) {
                        float2 det = k_fma2(m1, D[K1][r], D[K][r]);
                        det = k_fma2(m2, D[K2][r], det);
                        float2 tq = k_add2(dd, neg2(O[K][r]));
                        tq = k_fma2(neg2(m1), O[K1][r], tq);
                        tq = k_fma2(neg2(m2), O[K2][r], tq);
                        float2 inv;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(det.x));
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(det.y));
                        const float2 t = k_mul2(tq, inv);
                        const float2 p1 = k_fma2(t, D[K1][r], O[K1][r]);
                        const float2 p2 = k_fma2(t, D[K2][r], O[K2][r]);
                        // S = 2(u+v) - C and E = u - v straight from their own planes; then
                        //   x = |E| + |S - |E||  <=  C   <=>   u >= 0, v >= 0, u + v <= C
                        // (min(u,v) = (u+v-|u-v|)/2 and min(p,q) >= 0 <=> p+q >= |p-q|, twice);
                        // the absolute values are free operand modifiers of FADD2.
                        float2 sv = k_fma2(s1, p1, s0);
                        sv = k_fma2(s2, p2, sv);
                        float2 ev = k_fma2(e1, p1, e0);
                        ev = k_fma2(e2, p2, ev);
                        const float2 ea = abs2(ev);
                        const float2 kk = k_add2(sv, neg2(ea));
                        const float2 x = k_add2(ea, abs2(kk));
                        // valid & nearer <=> bits(x) <= bits(C) and bits(t) < best as unsigned
                        // integers (negative t, NaN and +inf all compare high)
                        cand |= (__float_as_uint(x.x) <= c_bits && __float_as_uint(t.x) < best[2 * r]) |
                                (__float_as_uint(x.y) <= c_bits && __float_as_uint(t.y) < best[2 * r + 1]);
                        tt[r] = t;
                        xx[r] = x;
                        dets[r] = det;
                    }
                    // Rare path (a few hits per ray over the whole mesh): warp-uniform branch so
                    // the common case pays one vote instead of predicated selects per test.
                    if (__any_sync(0xffffffffu, cand)) {
                        const int tri = __float_as_int(q0.w);
                        // |det * n[k]| >= intersect_eps is the reference's safe-denominator
                        // validity (tfrt/geometry.py:306)
                        const float nk = q1.w;
                        // the conditions are re-evaluated from an opaque copy of C, so that the
                        // common path does not have to keep eight per-ray predicates alive
                        uint32_t c_rare = c_bits;
                        asm volatile("" : "+r"(c_rare));
#pragma unroll
                        for (int r = 0; r < R2; ++r) {
                            if (__float_as_uint(xx[r].x) <= c_rare &&
                                __float_as_uint(tt[r].x) < best[2 * r] &&
                                fabsf(dets[r].x * nk) >= intersect_eps) {
                                best[2 * r] = __float_as_uint(tt[r].x);
                                bidx[2 * r] = tri;
                            }
                            if (__float_as_uint(xx[r].y) <= c_rare &&
                                __float_as_uint(tt[r].y) < best[2 * r + 1] &&
                                fabsf(dets[r].y * nk) >= intersect_eps) {
                                best[2 * r + 1] = __float_as_uint(tt[r].y);
                                bidx[2 * r + 1] = tri;
                            }
                        }
                    }
                }
            };
            const int header = __float_as_int(tile[2].w);
            const int c0 = header & 0xffff, c1 = c0 + (header >> 16);
            run_group(std::integral_constant<int, 0>{}, 0, c0);
            run_group(std::integral_constant<int, 1>{}, c0, c1);
            run_group(std::integral_constant<int, 2>{}, c1, TFRT_TILE_TRIS);

            // release the stage; the producer keeps the stream kStages tiles ahead
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&sm.empty[s]);
            if constexpr (kQueue) {
                if (tid == 0) produce_one();
            } else {
                if (tid == 0 && it + kStages < my_tiles)
                    issue(g + kStages, tile_begin + it + kStages);
            }
        }
        seq += (uint32_t)my_tiles;

        // ---- epilogue: (key, index) per live position
#pragma unroll
        for (int r = 0; r < R2; ++r) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = base + (2 * r + h) * kThreads + tid;
                if (p < n_live) {
                    const unsigned long long packed_hit =
                        ((unsigned long long)best[2 * r + h] << 32) | (uint32_t)bidx[2 * r + h];
                    if (use_atomic)
                        atomicMin(&hit[p], packed_hit);
                    else
                        hit[p] = packed_hit;
                }
            }
        }
    }
}

// Resident CTAs per SM of the three ray-count variants {R2 = 4, 2, 1}, asked from the runtime once
// (the persistent grid is sized from it).
const int *occupancy_table() {
    static int occ[3] = {0, 0, 0};
    if (!occ[0]) {
        const int smem = (int)sizeof(IntersectSmem);
        int v[3] = {4, 4, 6};
        cudaFuncSetAttribute(k_intersect3d<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_intersect3d<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_intersect3d<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v[0], k_intersect3d<4, false>, kThreads, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v[1], k_intersect3d<2, false>, kThreads, smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v[2], k_intersect3d<1, false>, kThreads, smem);
        for (int i = 0; i < 3; ++i) occ[i] = v[i] > 0 ? v[i] : 1;
    }
    return occ;
}

}  // namespace

// host launcher -------------------------------------------------------------------------------
// n_upper: upper bound of the live count (grid size); the kernel reads the exact count from
// n_live_ptr when given.
int tfrt_launch_intersect3d(const float *const ray[6], const int32_t *live,
                            const int32_t *n_live_ptr, int n_upper, const float *packed,
                            int n_tri, float intersect_eps, float ray_start_eps, float size_eps,
                            uint64_t *hit, int dir_mode, cudaStream_t stream) {
    if (n_upper <= 0 || n_tri <= 0) return 0;
    const int n_tiles = tfrt_div_up(n_tri, TFRT_TILE_TRIS);
    static int num_sms = 0;
    if (!num_sms) {
        int dev = 0;
        TFRT_CHECK_CUDA(cudaGetDevice(&dev));
        TFRT_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // Work decomposition.  Items = ray blocks (kThreads * 2*R2 rays) x triangle splits, handed
    // out dynamically to persistent CTAs, so the idle tail of a launch is about half an item per
    // CTA.  Long items are cheap to cut (an item's fixed cost -- ray gather, ring prologue,
    // 64-bit atomicMin merge -- is a few microseconds): aim for up to 64 items per resident CTA
    // while an item keeps >= 16 tiles, but never fewer than 8 items per CTA while it keeps >= 4
    // (measured flat between 4 and 16 items per CTA at 5-tile items, profiles/r1_k1_tuning.md).
    // Prefer 8 rays per thread (one LDS.128 feeds 8 rays); fall back to 4 / 2 rays per thread
    // and finer splits only when the ray count cannot fill the machine otherwise.
    int r2 = 1, splits = 1;
    {
        const int cand_r2[3] = {4, 2, 1};
        const int *occ = occupancy_table();           // resident CTAs per SM (registers / smem)
        static int items_lo = 0, items_hi = 0;
        if (!items_lo) {
            const char *e = getenv("TFRT_K1_ITEMS_PER_CTA");
            items_lo = e ? atoi(e) : 8;
            if (items_lo < 1) items_lo = 1;
            items_hi = e ? items_lo : 64;
        }
        bool done = false;
        for (int min_tps = 4; min_tps >= 1 && !done; min_tps >>= 1) {   // small meshes: finer splits
            for (int c = 0; c < 3 && !done; ++c) {
                const long long rb = tfrt_div_up(n_upper, kThreads * 2 * cand_r2[c]);
                const long long resident = (long long)num_sms * occ[c];
                const long long cap_lo = n_tiles / min_tps > 0 ? n_tiles / min_tps : 1;
                const long long cap_hi = n_tiles / (4 * min_tps) > 0 ? n_tiles / (4 * min_tps) : 1;
                long long sp_hi = (items_hi * resident + rb - 1) / rb;
                long long sp_lo = (items_lo * resident + rb - 1) / rb;
                if (sp_hi > cap_hi) sp_hi = cap_hi;
                if (sp_lo > cap_lo) sp_lo = cap_lo;
                long long sp = sp_hi > sp_lo ? sp_hi : sp_lo;
                if (sp < 1) sp = 1;
                r2 = cand_r2[c];
                splits = (int)sp;
                done = rb * sp >= 2 * resident;       // enough items to fill and balance
            }
        }
    }
#if TFRT_K1_FORCE_R2
    r2 = TFRT_K1_FORCE_R2;
#endif
    const int ray_blocks = tfrt_div_up(n_upper, kThreads * 2 * r2);
    const int tiles_per_split = tfrt_div_up(n_tiles, splits);
    splits = tfrt_div_up(n_tiles, tiles_per_split);
    const int use_atomic = splits > 1;
    if (use_atomic)
        TFRT_CHECK_CUDA(cudaMemsetAsync(hit, 0xff, (size_t)n_upper * sizeof(uint64_t), stream));

    // work counter of this launch: a rotating slot of a small device array, zeroed in stream
    // order (launches of one stream never overlap; 64 slots cover concurrent streams)
    static unsigned int *counters_of_device[64] = {nullptr};
    static std::atomic<unsigned int> launch_seq{0};
    int cur_dev = 0;
    TFRT_CHECK_CUDA(cudaGetDevice(&cur_dev));
    TFRT_REQUIRE(cur_dev >= 0 && cur_dev < 64, "tfrt_launch_intersect3d: device index %d", cur_dev);
    if (!counters_of_device[cur_dev])
        TFRT_CHECK_CUDA(cudaMalloc(&counters_of_device[cur_dev], 64 * sizeof(unsigned int)));
    unsigned int *counter = counters_of_device[cur_dev] + (launch_seq.fetch_add(1) % 64u);
    TFRT_CHECK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));

    const size_t smem = sizeof(IntersectSmem);
    const int occ_r2 = occupancy_table()[r2 >= 4 ? 0 : (r2 == 2 ? 1 : 2)];
    const long long total_upper = (long long)ray_blocks * splits;
    const long long resident = (long long)num_sms * occ_r2;
    dim3 grid((unsigned)(total_upper < resident ? total_upper : resident)), block(kThreads);
    // Barrier-free item crossing pays for itself on long, heavy items (8 rays per thread, >= 4
    // tiles: +1 % at 16 M rays x 99 k triangles, +4 % at 1 M x 20 k, +8 % at 1 M x 2 k); launches
    // made of short items keep the one-item-at-a-time loop (profiles/r1_k1_tuning.md section 5).
    static int queue_mode = -1;
    if (queue_mode < 0) {
        const char *e = getenv("TFRT_K1_QUEUE");     // 0 = never, 1 = auto (default), 2 = always
        queue_mode = e ? atoi(e) : 1;
    }
    const bool use_queue = queue_mode == 2 || (queue_mode == 1 && r2 == 4 && tiles_per_split >= 4);
#define TFRT_LAUNCH(R2, Q)                                                                      \
    do {                                                                                        \
        static bool attr_set = false;                                                           \
        if (!attr_set) {                                                                        \
            TFRT_CHECK_CUDA(cudaFuncSetAttribute(k_intersect3d<R2, Q>,                          \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                                 (int)smem));                                   \
            attr_set = true;                                                                    \
        }                                                                                       \
        k_intersect3d<R2, Q><<<grid, block, smem, stream>>>(                                    \
            ray[0], ray[1], ray[2], ray[3], ray[4], ray[5], live, n_live_ptr, n_upper,          \
            reinterpret_cast<const float4 *>(packed), n_tiles, tiles_per_split, splits,         \
            intersect_eps, ray_start_eps, 1.0f + 3.0f * size_eps,                               \
            reinterpret_cast<unsigned long long *>(hit), use_atomic, dir_mode, counter);        \
    } while (0)
    const bool timed = tfrt_profile_begin(stream);
#if TFRT_K1_FORCE_R2
    if (use_queue) TFRT_LAUNCH(TFRT_K1_FORCE_R2, true);
    else TFRT_LAUNCH(TFRT_K1_FORCE_R2, false);
#else
    if (r2 == 4 && use_queue) TFRT_LAUNCH(4, true);
    else if (r2 == 4) TFRT_LAUNCH(4, false);
    else if (r2 == 2 && use_queue) TFRT_LAUNCH(2, true);
    else if (r2 == 2) TFRT_LAUNCH(2, false);
    else if (use_queue) TFRT_LAUNCH(1, true);
    else TFRT_LAUNCH(1, false);
#endif
#undef TFRT_LAUNCH
    if (timed) tfrt_profile_end(stream);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}
