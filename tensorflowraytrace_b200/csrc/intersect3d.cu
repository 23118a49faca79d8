// intersect3d.cu -- K1a: brute-force nearest ray x triangle hit (sm_100a).
//
// Replaces the dense (M,N) evaluation of geometry.line_triangle_intersect
// (tfrt/geometry.py:191-320) + pruning + argmin of OpticalSystem3D._intersection
// (tfrt/engine.py:1138-1148).  Every live ray is tested against every triangle.
//
// Design
//  * a thread owns NR rays in registers (origin, direction: 6 scalars each);
//  * triangles stream through shared memory in tiles of TFRT_TILE_TRIS, copied by one elected
//    thread with cp.async.bulk (TMA 1-D) into a 3-stage mbarrier ring;
//  * the tile stores triangles in PAIRS: every packed f32x2 instruction tests one ray against two
//    triangles, with the two triangles' coefficients as the 64-bit operand and the ray's scalars
//    as 32-bit broadcast operands (SASS `R.F32`).  The register file delivers two 32-bit operands
//    per cycle per scheduler (tools/ffma2_probe.cu), so what an instruction costs is
//    max(2 FMA-pipe cycles, operands read / 2); consecutive instructions that share an operand in
//    the same slot get it from the operand reuse cache for free.  Looping the NR rays over one
//    triangle pair makes the WIDE operands (the coefficient pairs) the shared ones: ~23 operand
//    reads per test instead of ~33 for the transposed layout (two rays x one triangle), which
//    puts the loop back on the FMA pipe (profiles/r2_k1_tuning.md);
//  * per (ray, triangle) the projected precomputed-plane test (Wald's projection onto the
//    dominant axis k of the normal, plane scaled to a unit k-coefficient; record layout in
//    common.cuh) costs 14 fp32 ops = 7 packed instructions per test + 1 MUFU.RCP:
//        det = D[k] + m1 D[k1] + m2 D[k2]      t' = dd - O[k] - m1 O[k1] - m2 O[k2]
//        t   = t'/det     (origin pre-shifted by ray_start_eps so t >= 0 <=> valid)
//        P1  = O[k1] + t D[k1]                 P2 = O[k2] + t D[k2]
//        S   = s1 P1 + s2 P2 + s0  (= 2(u+v) - C)      E = e1 P1 + e2 P2 + e0  (= u - v)
//        x   = |E| + |S - |E||     (absolute values are free operand modifiers of FADD2)
//    with u, v the barycentrics inflated by size_eps and C = 1 + 3 size_eps:
//        inside  <=>  u >= 0, v >= 0, u + v <= C  <=>  x <= C
//    (min(u,v) = (u+v-|u-v|)/2 and min(p,q) >= 0 <=> p+q >= |p-q|, applied twice);
//  * x >= 0 or NaN, so "some ray of this thread is inside one of the two triangles" is
//    min over the thread's x values <= C: one FMNMX3 per ray (two x values each), one compare and
//    one warp vote per triangle pair; everything else (is the hit nearer? |den| >= eps? which
//    ray?) happens in the rare warp-uniform branch behind the vote;
//  * the triangles of a tile are grouped by k, so three specialised loops run back to back
//    with the roles of the ray components fixed at compile time;
//  * the running best is (bits(t), triangle index), updated to the lexicographic minimum: exact
//    ties go to the lowest triangle index whatever the scan order, the behaviour of argmin in the
//    reference's CPU kernels (e.g. the coplanar halves of a flat target), and the same rule the
//    64-bit atomicMin on (key << 32 | index) applies when the triangle range is split over CTAs.
#include <cstdlib>
#include <functional>
#include <mutex>
#include <type_traits>

#include "common.cuh"

// Tunables.  The defaults are the measured best of the variants timed by tools/k1_variants.sh
// (profiles/r2_k1_tuning.md); tools/k1_bench.cu compiles the others with -D.
#ifndef TFRT_K1_THREADS
#define TFRT_K1_THREADS 128
#endif
#ifndef TFRT_K1_STAGES
#define TFRT_K1_STAGES 3
#endif
#ifndef TFRT_K1_FORCE_NR    // 0: choose rays per thread from the ray count
#define TFRT_K1_FORCE_NR 0
#endif
#ifndef TFRT_K1_MINBLOCKS   // 3 CTAs of 128 threads per SM <=> at most 168 registers
#define TFRT_K1_MINBLOCKS 3
#endif

namespace {

constexpr int kThreads = TFRT_K1_THREADS;
constexpr int kStages = TFRT_K1_STAGES;

constexpr int kItemQueue = 8;   // published work items in flight (the producer is at most
                                // kStages + 1 items ahead of the slowest consumer warp)

struct __align__(16) IntersectSmem {
    float4 tile[kStages][TFRT_TILE_PAIRS * 6];
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
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }   // SASS: R.F32
__device__ __forceinline__ float2 k_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 k_mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 k_add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 abs2(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }

template <int NR, bool kQueue>
__global__ void __launch_bounds__(kThreads, ((NR >= 8 || TFRT_K1_FORCE_NR) ? TFRT_K1_MINBLOCKS : 1))
k_intersect3d(const float *__restrict__ sx, const float *__restrict__ sy,
              const float *__restrict__ sz, const float *__restrict__ ex,
              const float *__restrict__ ey, const float *__restrict__ ez,
              const int32_t *__restrict__ live, const int32_t *__restrict__ n_live_ptr,
              int n_live_static, const float4 *__restrict__ packed, int n_tiles,
              int tiles_per_split, int splits, float intersect_eps, float ray_start_eps,
              float w_one, unsigned long long *__restrict__ hit, int use_atomic, int dir_mode,
              unsigned int *__restrict__ work_counter, float start_floor_ulps) {
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
    const int rays_per_cta = kThreads * NR;
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
        bulk_g2s(sm.tile[s], packed + (size_t)(sm.p_tile_begin + sm.p_tile) * TFRT_TILE_PAIRS * 6,
                 TFRT_TILE_BYTES, &sm.full[s]);
        ++sm.p_tile;
        sm.p_g = g + 1;
    };
    // ---- producer, !kQueue: tile `g` of the CTA's running sequence <- global tile `tile`
    auto issue = [&](uint32_t g, int tile) {
        const int s = g % kStages;
        if (g >= (uint32_t)kStages) mbar_wait(&sm.empty[s], ((g / kStages) - 1) & 1);
        mbar_arrive_expect_tx(&sm.full[s], TFRT_TILE_BYTES);
        bulk_g2s(sm.tile[s], packed + (size_t)tile * TFRT_TILE_PAIRS * 6, TFRT_TILE_BYTES,
                 &sm.full[s]);
    };
    if constexpr (kQueue) {
        if (tid == 0)
            for (int i = 0; i < kStages; ++i) produce_one();   // fill the ring
    }

    const uint32_t c_bits = __float_as_uint(w_one);   // C = 1 + 3*size_eps (eps is folded into the planes)
    // Floor of the ray-start exclusion, as a distance: a ray that starts on a surface (the
    // previous hit, rounded to fp32) is off its plane by about an ulp of the scene's coordinates,
    // so the search cannot tell hits nearer than a few ulps from the self-hit.  The exact
    // exclusion (ray_start_epsilion) is applied to the winner in fp64 afterwards.
    const float start_floor =
        start_floor_ulps * 1.1920929e-7f * __ldg(reinterpret_cast<const float *>(packed) + 23);
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

        // ---- load this thread's rays: ray i = live position base + i * kThreads + tid
        float O[3][NR], D[3][NR];
        uint32_t best[NR];
        int32_t bidx[NR];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const int p = base + i * kThreads + tid;
            if (p < n_live) {
                const int slot = live ? live[p] : p;
                const float ax = sx[slot], ay = sy[slot], az = sz[slot];
                // dir_mode: the last three arrays hold the direction (end - start) itself,
                // the form the trace loop keeps between iterations (no cancellation error)
                D[0][i] = dir_mode ? ex[slot] : ex[slot] - ax;
                D[1][i] = dir_mode ? ey[slot] : ey[slot] - ay;
                D[2][i] = dir_mode ? ez[slot] : ez[slot] - az;
                // shift the origin along the ray by the search's start exclusion (in units of
                // |D|, like ray_u): t_shifted >= 0 <=> ray_u >= start
                // (explicit operation order: the grid search of grid3d.cu shifts its rays with the
                // same expression and must get the same bits)
                const float start = fmaxf(ray_start_eps, start_floor * rsqrtf(
                    fmaf(D[2][i], D[2][i], fmaf(D[1][i], D[1][i], __fmul_rn(D[0][i], D[0][i])))));
                O[0][i] = fmaf(start, D[0][i], ax);
                O[1][i] = fmaf(start, D[1][i], ay);
                O[2][i] = fmaf(start, D[2][i], az);
            } else {  // padding lane: zero direction => det == 0 => never a hit
                O[0][i] = O[1][i] = O[2][i] = 0.f;
                D[0][i] = D[1][i] = D[2][i] = 0.f;
            }
            best[i] = TFRT_NOHIT_KEY;
            bidx[i] = 0;
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
                for (int j = j0; j < j1; ++j) {
                    // one pair record = six float4 (layout: common.cuh); .x / .y halves of every
                    // float2 below belong to the pair's first / second triangle
                    const float4 c0 = tile[j * 6 + 0], c1 = tile[j * 6 + 1], c2 = tile[j * 6 + 2];
                    const float4 c3 = tile[j * 6 + 3], c4 = tile[j * 6 + 4];
                    const float2 m1 = f2(c0.x, c0.y), m2 = f2(c0.z, c0.w), dd = f2(c1.x, c1.y);
                    const float2 s1 = f2(c1.z, c1.w), s2 = f2(c2.x, c2.y), s0 = f2(c2.z, c2.w);
                    const float2 e1 = f2(c3.x, c3.y), e2 = f2(c3.z, c3.w), e0 = f2(c4.x, c4.y);
                    float2 tt[NR], xx[NR];
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        float2 det = k_fma2(m1, bc(D[K1][i]), bc(D[K][i]));
                        det = k_fma2(m2, bc(D[K2][i]), det);
                        float2 tq = k_add2(dd, bc(-O[K][i]));
                        tq = k_fma2(neg2(m1), bc(O[K1][i]), tq);
                        tq = k_fma2(neg2(m2), bc(O[K2][i]), tq);
                        float2 inv;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(det.x));
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(det.y));
                        const float2 t = k_mul2(tq, inv);
                        const float2 p1 = k_fma2(t, bc(D[K1][i]), bc(O[K1][i]));
                        const float2 p2 = k_fma2(t, bc(D[K2][i]), bc(O[K2][i]));
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
                        tt[i] = t;
                        xx[i] = x;
                    }
                    // x >= 0 or NaN (fminf drops a NaN operand): the smallest x decides whether
                    // any of the thread's rays is inside one of the two triangles.  Reduced as a
                    // tree of three-input minima (FMNMX3): depth 3 for 16 values instead of a
                    // chain of 8 dependent instructions at the end of the loop body.
                    float xmin;
                    {
                        float lvl[(2 * NR + 2) / 3 + 1];
                        int m = 0;
#pragma unroll
                        for (int i = 0; i + 1 < NR; i += 3) {     // three rays = six values -> two minima
                            if (i + 2 < NR) {
                                lvl[m++] = fminf(fminf(xx[i].x, xx[i].y), xx[i + 1].x);
                                lvl[m++] = fminf(fminf(xx[i + 1].y, xx[i + 2].x), xx[i + 2].y);
                            } else {                               // two rays left: four values
                                lvl[m++] = fminf(fminf(xx[i].x, xx[i].y), xx[i + 1].x);
                                lvl[m++] = xx[i + 1].y;
                            }
                        }
                        if (NR % 3 == 1) lvl[m++] = fminf(xx[NR - 1].x, xx[NR - 1].y);
                        // NR = 8: six values; NR = 4: three; NR = 2: two
#pragma unroll
                        for (int step = 0; step < 3; ++step) {
                            int o = 0;
#pragma unroll
                            for (int i = 0; i < m; i += 3) {
                                float v = lvl[i];
                                if (i + 1 < m) v = fminf(v, lvl[i + 1]);
                                if (i + 2 < m) v = fminf(v, lvl[i + 2]);
                                lvl[o++] = v;
                            }
                            m = o;
                        }
                        xmin = lvl[0];
                    }
                    // Rare path (a few inside hits per ray over the whole mesh): warp-uniform
                    // branch, so the common case pays the min chain, one compare and one vote per
                    // (NR rays x 2 triangles).
                    if (__any_sync(0xffffffffu, xmin <= w_one)) {
                        const float4 c5 = tile[j * 6 + 5];
                        uint32_t c_rare = c_bits;
                        asm volatile("" : "+r"(c_rare));
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int tri = __float_as_int(h ? c4.w : c4.z);
                            // |det * n[k]| >= intersect_eps is the reference's safe-denominator
                            // validity (tfrt/geometry.py:306)
                            const float nk = h ? c5.y : c5.x;
                            const float m1h = h ? m1.y : m1.x, m2h = h ? m2.y : m2.x;
#pragma unroll
                            for (int i = 0; i < NR; ++i) {
                                const float det = fmaf(m2h, D[K2][i], fmaf(m1h, D[K1][i], D[K][i]));
                                // nearer, or equally near with a lower index: the winner is the
                                // lexicographic minimum of (bits(t), index) whatever the scan
                                // order (negative t, NaN and +inf compare high as unsigned)
                                const uint32_t tb = __float_as_uint(h ? tt[i].y : tt[i].x);
                                const uint32_t xb = __float_as_uint(h ? xx[i].y : xx[i].x);
                                const bool ok = (xb <= c_rare) &
                                                ((tb < best[i]) | ((tb == best[i]) & (tri < bidx[i]))) &
                                                (tb < TFRT_NOHIT_KEY) &
                                                (fabsf(det * nk) >= intersect_eps);
                                best[i] = ok ? tb : best[i];
                                bidx[i] = ok ? tri : bidx[i];
                            }
                        }
                    }
                    // keeps the loop test below the rare-path branch (otherwise ptxas hoists it
                    // above and parks the predicate in a general register)
                    asm volatile("" : "+r"(j));
                }
            };
            // pairs per axis group of this tile (record 0 carries them)
            const int header = __float_as_int(tile[5].z);
            const int n0 = header & 0xff, n1 = n0 + ((header >> 8) & 0xff);
            const int n2 = n1 + ((header >> 16) & 0xff);
            run_group(std::integral_constant<int, 0>{}, 0, n0);
            run_group(std::integral_constant<int, 1>{}, n0, n1);
            run_group(std::integral_constant<int, 2>{}, n1, n2);

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
        for (int i = 0; i < NR; ++i) {
            const int p = base + i * kThreads + tid;
            if (p < n_live) {
                const unsigned long long packed_hit =
                    ((unsigned long long)best[i] << 32) | (uint32_t)bidx[i];
                if (use_atomic)
                    atomicMin(&hit[p], packed_hit);
                else
                    hit[p] = packed_hit;
            }
        }
    }
}

// Per-device launch configuration, initialised once per device under std::call_once: SM count,
// resident CTAs per SM of the ray-count variants {NR = 8, 4, 2} (the persistent grid is sized from
// it) and the dynamic shared-memory opt-in of every instantiation.
struct DeviceConfig {
    std::once_flag once;
    cudaError_t status = cudaSuccess;
    int num_sms = 0;
    int occ[3] = {1, 1, 1};
};
constexpr int kMaxDevices = 64;
DeviceConfig g_devices[kMaxDevices];

template <int NR, bool Q>
cudaError_t opt_in_smem() {
    return cudaFuncSetAttribute(k_intersect3d<NR, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(IntersectSmem));
}

void init_device(DeviceConfig &cfg, int dev) {
    cudaError_t e = cudaDeviceGetAttribute(&cfg.num_sms, cudaDevAttrMultiProcessorCount, dev);
#if TFRT_K1_FORCE_NR
    if (e == cudaSuccess) e = opt_in_smem<TFRT_K1_FORCE_NR, true>();
    if (e == cudaSuccess) e = opt_in_smem<TFRT_K1_FORCE_NR, false>();
#endif
    if (e == cudaSuccess) e = opt_in_smem<8, true>();
    if (e == cudaSuccess) e = opt_in_smem<8, false>();
    if (e == cudaSuccess) e = opt_in_smem<4, true>();
    if (e == cudaSuccess) e = opt_in_smem<4, false>();
    if (e == cudaSuccess) e = opt_in_smem<2, true>();
    if (e == cudaSuccess) e = opt_in_smem<2, false>();
    const int smem = (int)sizeof(IntersectSmem);
    int v[3] = {3, 4, 6};
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v[0], k_intersect3d<8, false>, kThreads, smem);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v[1], k_intersect3d<4, false>, kThreads, smem);
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v[2], k_intersect3d<2, false>, kThreads, smem);
    for (int i = 0; i < 3; ++i) cfg.occ[i] = v[i] > 0 ? v[i] : 1;
    cfg.status = e;
}

struct Tunables {
    int items_lo, items_hi, queue_mode;
    Tunables() {
        const char *e = getenv("TFRT_K1_ITEMS_PER_CTA");
        items_lo = e ? atoi(e) : 8;
        if (items_lo < 1) items_lo = 1;
        items_hi = e ? items_lo : 64;
        const char *q = getenv("TFRT_K1_QUEUE");     // 0 = never, 1 = auto (default), 2 = always
        queue_mode = q ? atoi(q) : 1;
    }
};

}  // namespace

// host launcher -------------------------------------------------------------------------------
// n_upper: upper bound of the live count (grid size); the kernel reads the exact count from
// n_live_ptr when given.  Contract (all in stream order): *work_counter == 0 and
// hit[0 .. n_upper) == ~0 on entry; the launcher itself enqueues nothing but the kernel and keeps
// no state between calls (the caller owns the counter word, so any number of launches may be in
// flight on any number of streams).
int tfrt_launch_intersect3d(const float *const ray[6], const int32_t *live,
                            const int32_t *n_live_ptr, int n_upper, const float *packed,
                            int n_tri, float intersect_eps, float ray_start_eps, float size_eps,
                            uint64_t *hit, int dir_mode, unsigned int *work_counter,
                            float start_floor_ulps, cudaStream_t stream) {
    if (n_upper <= 0 || n_tri <= 0) return 0;
    TFRT_REQUIRE(work_counter && hit, "tfrt_launch_intersect3d: null workspace");
    const int n_tiles = tfrt_div_up(n_tri, TFRT_TILE_TRIS);
    int cur_dev = 0;
    TFRT_CHECK_CUDA(cudaGetDevice(&cur_dev));
    TFRT_REQUIRE(cur_dev >= 0 && cur_dev < kMaxDevices, "tfrt_launch_intersect3d: device index %d",
                 cur_dev);
    DeviceConfig &cfg = g_devices[cur_dev];
    std::call_once(cfg.once, init_device, std::ref(cfg), cur_dev);
    TFRT_CHECK_CUDA(cfg.status);
    static const Tunables tun;                       // thread-safe (C++11 magic static), read-only
    const int num_sms = cfg.num_sms;
    // Work decomposition.  Items = ray blocks (kThreads * NR rays) x triangle splits, handed
    // out dynamically to persistent CTAs, so the idle tail of a launch is about half an item per
    // CTA.  Long items are cheap to cut (an item's fixed cost -- ray gather, ring prologue,
    // 64-bit atomicMin merge -- is a few microseconds): aim for up to 64 items per resident CTA
    // while an item keeps >= 16 tiles, but never fewer than 8 items per CTA while it keeps >= 4
    // (measured flat between 4 and 16 items per CTA at 5-tile items, profiles/r1_k1_tuning.md).
    // Prefer 8 rays per thread (one LDS.128 feeds 8 rays); fall back to 4 / 2 rays per thread
    // and finer splits only when the ray count cannot fill the machine otherwise.
    int nr = 2, splits = 1;
    {
        const int cand_nr[3] = {8, 4, 2};
        bool done = false;
        for (int min_tps = 4; min_tps >= 1 && !done; min_tps >>= 1) {   // small meshes: finer splits
            for (int c = 0; c < 3 && !done; ++c) {
                const long long rb = tfrt_div_up(n_upper, kThreads * cand_nr[c]);
                const long long resident = (long long)num_sms * cfg.occ[c];
                const long long cap_lo = n_tiles / min_tps > 0 ? n_tiles / min_tps : 1;
                const long long cap_hi = n_tiles / (4 * min_tps) > 0 ? n_tiles / (4 * min_tps) : 1;
                long long sp_hi = (tun.items_hi * resident + rb - 1) / rb;
                long long sp_lo = (tun.items_lo * resident + rb - 1) / rb;
                if (sp_hi > cap_hi) sp_hi = cap_hi;
                if (sp_lo > cap_lo) sp_lo = cap_lo;
                long long sp = sp_hi > sp_lo ? sp_hi : sp_lo;
                if (sp < 1) sp = 1;
                nr = cand_nr[c];
                splits = (int)sp;
                done = rb * sp >= 2 * resident;       // enough items to fill and balance
            }
        }
    }
#if TFRT_K1_FORCE_NR
    nr = TFRT_K1_FORCE_NR;
#endif
    const int ray_blocks = tfrt_div_up(n_upper, kThreads * nr);
    const int tiles_per_split = tfrt_div_up(n_tiles, splits);
    splits = tfrt_div_up(n_tiles, tiles_per_split);
    const int use_atomic = splits > 1;

    const size_t smem = sizeof(IntersectSmem);
    const int occ_nr = cfg.occ[nr >= 8 ? 0 : (nr == 4 ? 1 : 2)];
    const long long total_upper = (long long)ray_blocks * splits;
    const long long resident = (long long)num_sms * occ_nr;
    dim3 grid((unsigned)(total_upper < resident ? total_upper : resident)), block(kThreads);
    // Barrier-free item crossing pays for itself on long, heavy items (8 rays per thread, >= 4
    // tiles: +1 % at 16 M rays x 99 k triangles, +4 % at 1 M x 20 k, +8 % at 1 M x 2 k); launches
    // made of short items keep the one-item-at-a-time loop (profiles/r1_k1_tuning.md section 5).
    const bool use_queue =
        tun.queue_mode == 2 || (tun.queue_mode == 1 && nr == 8 && tiles_per_split >= 4);
#define TFRT_LAUNCH(NR, Q)                                                                      \
    k_intersect3d<NR, Q><<<grid, block, smem, stream>>>(                                        \
        ray[0], ray[1], ray[2], ray[3], ray[4], ray[5], live, n_live_ptr, n_upper,              \
        reinterpret_cast<const float4 *>(packed), n_tiles, tiles_per_split, splits,             \
        intersect_eps, ray_start_eps, 1.0f + 3.0f * size_eps,                                   \
        reinterpret_cast<unsigned long long *>(hit), use_atomic, dir_mode, work_counter,        \
        start_floor_ulps)
    const bool timed = tfrt_profile_begin(stream);
#if TFRT_K1_FORCE_NR
    if (use_queue) TFRT_LAUNCH(TFRT_K1_FORCE_NR, true);
    else TFRT_LAUNCH(TFRT_K1_FORCE_NR, false);
#else
    if (nr == 8 && use_queue) TFRT_LAUNCH(8, true);
    else if (nr == 8) TFRT_LAUNCH(8, false);
    else if (nr == 4 && use_queue) TFRT_LAUNCH(4, true);
    else if (nr == 4) TFRT_LAUNCH(4, false);
    else if (use_queue) TFRT_LAUNCH(2, true);
    else TFRT_LAUNCH(2, false);
#endif
#undef TFRT_LAUNCH
    if (timed) tfrt_profile_end(stream);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}
