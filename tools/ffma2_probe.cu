// ffma2_probe.cu -- what does a packed f32x2 instruction cost on sm_100a, by operand form?
// Tuning study for the search kernel (profiles/r2_k1_tuning.md): register-only loops of 8
// independent chains per thread, 256 threads x 8 CTAs per SM, timed with CUDA events.
// Prints cycles per warp-instruction per SM sub-partition (4 schedulers per SM, all warps busy):
// a plain scalar FFMA issues at 1.0, FFMA2 with three register-pair operands at ~2.1.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma2_probe tools/ffma2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ float2 bc(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 ab(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }

// MODE 0: a = fma2(a, b, c)            three pair operands, b / c loop invariant (reuse cache)
// MODE 1: a = fma2(bc(s[i]), a, c)     one 32-bit broadcast operand (R.F32)
// MODE 2: a = fma2(bc(s[i]), a, bc(u)) two broadcast operands
// MODE 3: a = fma2(a, b[i], c[i])      three pair operands, all distinct registers
// MODE 4: a = add2(|a|, -|b[i]|)       FADD2 with abs / neg modifiers
// MODE 5: a = fma2(bc(s[i]), b[i], a)  broadcast x pair + accumulate (the kernel's det / S / E form)
// MODE 6: a = mul2(a, b[i])            FMUL2
// MODE 7: scalar: a.x = fma(s, a.x, c.x); a.y = fma(s, a.y, c.y)   (two scalar FFMA per chain)
template <int MODE>
__global__ void __launch_bounds__(256) k_probe(int iters, const float *__restrict__ in, float *sink) {
    float2 a[8], b[8], c[8];
    in += threadIdx.x * 64;                 // per-thread values: nothing is warp-uniform
    float s[8], u = in[57];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = make_float2(in[i], in[i + 8]);
        b[i] = make_float2(in[i + 16], in[i + 24]);
        c[i] = make_float2(in[i + 32], in[i + 40]);
        s[i] = in[i + 48];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) a[i] = __ffma2_rn(a[i], b[0], c[0]);
                if (MODE == 1) a[i] = __ffma2_rn(bc(s[i]), a[i], c[0]);
                if (MODE == 2) a[i] = __ffma2_rn(bc(s[i]), a[i], bc(u));
                if (MODE == 3) a[i] = __ffma2_rn(a[i], b[i], c[i]);
                if (MODE == 4) a[i] = __fadd2_rn(ab(a[i]), make_float2(-fabsf(b[i].x), -fabsf(b[i].y)));
                if (MODE == 5) a[i] = __ffma2_rn(bc(s[i]), b[i], a[i]);
                if (MODE == 6) a[i] = __fmul2_rn(a[i], b[i]);
                if (MODE == 7) { a[i].x = fmaf(s[i], a[i].x, c[i].x); a[i].y = fmaf(s[i], a[i].y, c[i].y); }
                // MODE 8 / 9: seven FMA-bound FFMA2 (the MODE 2 form) + one / two MUFU.RCP per
                // round: does a reciprocal take issue cycles away from the FMA pipe?
                if (MODE == 8 || MODE == 9) {
                    if (i < 7) a[i] = __ffma2_rn(bc(s[i]), a[i], bc(u));
                    else {
                        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[7].x));
                        if (MODE == 9) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[7].y));
                    }
                }
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    if (r == 123.456f) sink[0] = r;
}

// MODE 10: the search kernel's 14-instruction test for 4 ray pairs against one coefficient set
// per round (coefficients rotate through 4 register sets so nothing is loop invariant), without
// the reciprocal, the min chain and the vote: the FMA-pipe work of the inner loop alone.
// MODE 11: the same with pre-duplicated coefficient pairs instead of 32-bit broadcasts.
// MODE 12: MODE 10 + the two MUFU.RCP per pair.   MODE 13: MODE 12 + FMNMX3 chain + vote.
template <int MODE>
__global__ void __launch_bounds__(128, 4) k_test(int iters, const float *__restrict__ in, float *sink) {
    float2 O[3][4], D[3][4];
    float q[2][9];
    float2 q2[2][9];
    in += threadIdx.x * 64;                 // per-thread values: nothing is warp-uniform
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cpt = 0; cpt < 3; ++cpt) {
            O[cpt][r] = make_float2(in[r * 3 + cpt], in[12 + r * 3 + cpt]);
            D[cpt][r] = make_float2(in[24 + r * 3 + cpt], in[36 + r * 3 + cpt]);
        }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int cpt = 0; cpt < 9; ++cpt) {
            q[j][cpt] = in[(j * 9 + cpt + 5) % 64] * 0.75f;
            // MODE 11: a genuine register pair (the halves differ in the last bits, so ptxas cannot
            // fold it back into a 32-bit broadcast)
            q2[j][cpt] = make_float2(q[j][cpt], q[j][cpt] * 1.0000002f);
        }
    float acc = 0.f;
    const float cone = in[63];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float2 m1, m2, dd, s1, s2, s0, e1, e2, e0;
            if (MODE == 11) {
                m1 = q2[j][0]; m2 = q2[j][1]; dd = q2[j][2]; s1 = q2[j][3]; s2 = q2[j][4]; s0 = q2[j][5];
                e1 = q2[j][6]; e2 = q2[j][7]; e0 = q2[j][8];
            } else {
                m1 = bc(q[j][0]); m2 = bc(q[j][1]); dd = bc(q[j][2]); s1 = bc(q[j][3]); s2 = bc(q[j][4]);
                s0 = bc(q[j][5]); e1 = bc(q[j][6]); e2 = bc(q[j][7]); e0 = bc(q[j][8]);
            }
            float xmin = __int_as_float(0x7f800000);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float2 det = __ffma2_rn(m1, D[1][r], D[0][r]);
                det = __ffma2_rn(m2, D[2][r], det);
                float2 tq = __fadd2_rn(dd, make_float2(-O[0][r].x, -O[0][r].y));
                tq = __ffma2_rn(make_float2(-m1.x, -m1.y), O[1][r], tq);
                tq = __ffma2_rn(make_float2(-m2.x, -m2.y), O[2][r], tq);
                float2 inv = det;
                if (MODE >= 12) {
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(det.x));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(det.y));
                }
                const float2 t = __fmul2_rn(tq, inv);
                const float2 p1 = __ffma2_rn(t, D[1][r], O[1][r]);
                const float2 p2 = __ffma2_rn(t, D[2][r], O[2][r]);
                float2 sv = __ffma2_rn(s1, p1, s0);
                sv = __ffma2_rn(s2, p2, sv);
                float2 ev = __ffma2_rn(e1, p1, e0);
                ev = __ffma2_rn(e2, p2, ev);
                const float2 ea = ab(ev);
                const float2 kk = __fadd2_rn(sv, make_float2(-ea.x, -ea.y));
                const float2 x = __fadd2_rn(ea, ab(kk));
                if (MODE >= 13) xmin = fminf(fminf(xmin, x.x), x.y);
                else { acc += x.x + x.y; }      // keep the chain alive (two scalar FADD per pair)
            }
            if (MODE >= 13) {
                if (__any_sync(0xffffffffu, xmin <= cone)) acc += xmin;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) acc += O[0][r].x + O[0][r].y;
    if (acc == 123.456f) sink[0] = acc;
}

template <typename F>
static float time_it(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    static float h_in[256 * 64];
    unsigned seed = 12345u;
    for (int i = 0; i < 256 * 64; ++i) {
        seed = seed * 1664525u + 1013904223u;
        h_in[i] = 0.5f + (seed >> 8) * (1.0f / 16777216.0f);     // [0.5, 1.5): chains stay finite
    }
    float *d_in, *sink;
    CK(cudaMalloc(&d_in, sizeof(h_in))); CK(cudaMalloc(&sink, 4));
    CK(cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice));
    const int iters = 4096;
    const double ghz = khz * 1e-6;
    printf("SMs %d, clock %.3f GHz (attribute; the achieved clock may differ)\n", sms, ghz);
    // per-instruction probes: 8 CTAs x 8 warps per SM = 16 warps per scheduler
#define PROBE(M, what)                                                                           \
    {                                                                                            \
        const float ms = time_it([&] { k_probe<M><<<sms * 8, 256>>>(iters, d_in, sink); });      \
        const double inst = (double)iters * 64.0 * (M == 7 ? 2.0 : 1.0);   /* per warp */       \
        const double cyc = ms * 1e-3 * ghz * 1e9 / (inst * 16.0);          /* 16 warps / sched */ \
        printf("probe %2d  %-58s %8.3f ms  %.3f cycles/inst/scheduler\n", M, what, ms, cyc);     \
    }
    PROBE(0, "FFMA2 pair,pair,pair (b, c invariant)");
    PROBE(1, "FFMA2 bcast32,pair,pair");
    PROBE(2, "FFMA2 bcast32,pair,bcast32");
    PROBE(3, "FFMA2 pair,pair,pair (all distinct)");
    PROBE(4, "FADD2 |pair|,-|pair|");
    PROBE(5, "FFMA2 bcast32,pair,acc");
    PROBE(6, "FMUL2 pair,pair");
    PROBE(7, "scalar FFMA x2");
    PROBE(8, "7 x FFMA2 bcast32,pair,bcast32 + 1 MUFU.RCP (per 8 slots)");
    PROBE(9, "7 x FFMA2 bcast32,pair,bcast32 + 2 MUFU.RCP (per 8 slots)");
#define TEST(M, what)                                                                            \
    {                                                                                            \
        const float ms = time_it([&] { k_test<M><<<sms * 4, 128>>>(iters, d_in, sink); });       \
        const double steps = (double)iters * 2.0;            /* (8 rays x 1 triangle) per warp */ \
        const double cyc = ms * 1e-3 * ghz * 1e9 / (steps * 4.0);          /* 4 warps / sched */  \
        printf("test %2d  %-58s %8.3f ms  %.1f cycles per (8 rays x 1 triangle)\n", M, what, ms, cyc); \
    }
    TEST(10, "14 packed ops x 4 pairs, 32-bit broadcast coefficients");
    TEST(11, "14 packed ops x 4 pairs, pre-duplicated coefficient pairs");
    TEST(12, "... broadcast + 8 MUFU.RCP");
    TEST(13, "... broadcast + 8 MUFU.RCP + FMNMX3 chain + vote");
    CK(cudaGetLastError());
    return 0;
}
