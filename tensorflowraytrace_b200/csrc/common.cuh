// common.cuh -- shared device helpers for libtfrt_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tfrt_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtfrt_b200 is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------------------------------------
// error plumbing (host)

void tfrt_set_error(const char *fmt, ...);

#define TFRT_CHECK_CUDA(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            tfrt_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                           __LINE__);                                                          \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

#define TFRT_REQUIRE(cond, ...)                                                                \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            tfrt_set_error(__VA_ARGS__);                                                       \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

void tfrt_note_launch(int n);                     // launch counter (abi.cu)
bool tfrt_profile_begin(cudaStream_t st);          // event before the intersect kernel
void tfrt_profile_end(cudaStream_t st);            // event after it

static inline int tfrt_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// packed triangle layout
//
// Per triangle, 9 coefficients of the projected form of Wald's ray-triangle test: k = dominant
// axis of the normal n = e1 x e2, (k1, k2) = (k+1, k+2) mod 3, the plane scaled so that its
// k-coefficient is 1:
//   m1, m2, dd      det = D[k] + m1 D[k1] + m2 D[k2],  t' = dd - O[k] - m1 O[k1] - m2 O[k2]
//   s1, s2, s0      S = s1 P[k1] + s2 P[k2] + s0  =  2 (u + v) - C
//   e1, e2, e0      E = e1 P[k1] + e2 P[k2] + e0  =  u - v
// with m = n/n[k], dd = m.A, u = a3.P + ad and v = b3.P + bd the barycentric planes
// (a3 = (e2 x n)/|n|^2, b3 = (n x e1)/|n|^2, the inflation folded into ad, bd), P[k] eliminated
// through the plane equation (exact for points on the plane), C = 1 + 3 size_eps.  The kernel
// accepts a hit iff |E| + |S - |E|| <= C, which is u >= 0, v >= 0, u + v <= C: 14 FMA-pipe
// lane-ops per test.  det * n[k] is the reference's denominator (tfrt/geometry.py:300) up to
// sign.
//
// Storage: triangles travel in PAIRS (A, B) of one dominant axis, one pair record = 24 floats =
// six float4, every coefficient as the adjacent pair (A's, B's) that a packed f32x2 instruction
// takes as its 64-bit operand:
//   f4[0] = (m1A, m1B, m2A, m2B)    f4[1] = (ddA, ddB, s1A, s1B)    f4[2] = (s2A, s2B, s0A, s0B)
//   f4[3] = (e1A, e1B, e2A, e2B)    f4[4] = (e0A, e0B, indexA, indexB)
//   f4[5] = (n[k]A, n[k]B, header, 0)           (read only on the rare hit path; the last word
//           of the mesh's very first record holds the scene scale = largest |coordinate|)
// `index` is the triangle's position in the merged set.  A tile holds TFRT_TILE_TRIS consecutive
// triangles of the merged set, grouped by k (stable: index order inside a group) so that the
// kernel runs three specialised loops without a per-triangle branch; a group of odd size is
// completed with a never-hit half (S = 4 > C), hence TFRT_TILE_PAIRS = TRIS/2 + 2 pair slots;
// `header` of the tile's record 0 = pairs in group 0 | pairs in group 1 << 8 | group 2 << 16.
#define TFRT_TRI_FLOATS 12
#define TFRT_TILE_TRIS 256                      // triangles per shared-memory stage
#define TFRT_TILE_PAIRS (TFRT_TILE_TRIS / 2 + 2)
#define TFRT_TILE_BYTES (TFRT_TILE_PAIRS * 2 * TFRT_TRI_FLOATS * 4)

#define TFRT_NOHIT_KEY 0x7f800000u              // +inf bits: "no valid hit yet"

// Search margin of one triangle, in barycentric units: the fp32 search accepts a hit when the
// inflated barycentrics are >= -(size_eps + margin), and the fp64 evaluation of the winner then
// decides with the reference's exact rule (tfrt/engine.py:1138-1141).  The margin is `ulps`
// units in the last place of the largest coordinate of the triangle (the position error of an
// fp32 hit point near it), converted to barycentric units by the steepest barycentric gradient
// (1 / smallest altitude).  Scale-aware: a mesh in millimetres or far from the origin gets a
// proportionally larger margin.  A, e1 = B - A, e2 = C - A in fp64.
__host__ __device__ inline double tfrt_tri_scale(const double A[3], const double e1[3],
                                                 const double e2[3]) {
    double scale = 0.0;          // largest |coordinate| of the three vertices
    for (int c = 0; c < 3; ++c) {
        const double a = fabs(A[c]), b = fabs(A[c] + e1[c]), d = fabs(A[c] + e2[c]);
        scale = fmax(scale, fmax(a, fmax(b, d)));
    }
    return scale;
}

// steepest barycentric gradient = 1 / smallest altitude:
// |grad u| = |e2| / |N|, |grad v| = |e1| / |N|, |grad w| = |e2 - e1| / |N|
__host__ __device__ inline double tfrt_tri_bary_gradient(const double e1[3], const double e2[3]) {
    const double N[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2],
                         e1[0] * e2[1] - e1[1] * e2[0]};
    const double nn = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
    if (!(nn > 0.0)) return 0.0;
    const double l1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2];
    const double l2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
    const double d0 = e2[0] - e1[0], d1 = e2[1] - e1[1], d2 = e2[2] - e1[2];
    const double l3 = d0 * d0 + d1 * d1 + d2 * d2;
    return sqrt(fmax(l1, fmax(l2, l3)) / nn);
}

__host__ __device__ inline double tfrt_tri_margin(const double A[3], const double e1[3],
                                                  const double e2[3], double ulps) {
    return ulps * 1.1920928955078125e-7 * tfrt_tri_scale(A, e1, e2) * tfrt_tri_bary_gradient(e1, e2);
}

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk-copy (TMA 1-D) PTX wrappers

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// refractive index n(lambda) in fp64 (tfrt/materials.py:25-104)

__device__ __forceinline__ double tfrt_material_n(const TfrtMaterial &m, double x) {
    const double x2 = x * x;
    switch (m.kind) {
    case TFRT_MAT_SELLMEIER: {
        double s = 1.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) s += m.c[2 * k] * x2 / (x2 - m.c[2 * k + 1]);
        return sqrt(s);
    }
    case TFRT_MAT_LAURENT: {
        const double i2 = 1.0 / x2;
        const double i4 = i2 * i2;
        return sqrt(m.c[0] + m.c[1] * x2 + m.c[2] * x2 * x2 + m.c[3] * i2 + m.c[4] * i4 +
                    m.c[5] * i4 * i2 + m.c[6] * i4 * i4);
    }
    case TFRT_MAT_POLY:
        return m.c[0] + m.c[1] * x2 + m.c[2] / x2;
    default:
        return m.c[0];
    }
}
