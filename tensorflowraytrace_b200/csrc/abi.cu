// abi.cu -- error plumbing, version, and the fp32 peak micro-benchmarks (roofline denominators).
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void tfrt_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *tfrt_last_error(void) { return g_err; }
extern "C" int tfrt_abi_version(void) { return TFRT_B200_ABI_VERSION; }
#ifndef TFRT_SOURCE_HASH
#define TFRT_SOURCE_HASH "unknown"
#endif
// The marker lets the loader read the hash out of the file without dlopen()ing it (a library
// that is already mapped would be returned again after a rebuild under the same name).
static const char g_source_hash[] = "TFRT_SOURCE_HASH=" TFRT_SOURCE_HASH;
extern "C" const char *tfrt_source_hash(void) { return g_source_hash + 17; }
extern "C" size_t tfrt_struct_size(int which) {
    switch (which) {
    case 0: return sizeof(TfrtMaterial);
    case 1: return sizeof(TfrtTrace3DArgs);
    case 2: return sizeof(TfrtTrace3DBwdArgs);
    case 3: return sizeof(TfrtTrace2DArgs);
    default: return 0;
    }
}

// ---------------------------------------------------------------------------------------------
// launch counter + optional event timing of the intersect kernel
#include <atomic>
#include <mutex>
#include <vector>

static std::atomic<long long> g_launches{0};
void tfrt_note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long tfrt_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_prof_events;   // pairs: begin, end
static size_t g_prof_used = 0;                   // events used (even)
static const size_t kProfMaxEvents = 8192;

bool tfrt_profile_begin(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_on || g_prof_used + 2 > kProfMaxEvents) return false;
    while (g_prof_events.size() < g_prof_used + 2) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return false;
        g_prof_events.push_back(e);
    }
    cudaEventRecord(g_prof_events[g_prof_used], st);
    return true;
}
void tfrt_profile_end(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_prof_events[g_prof_used + 1], st);
    g_prof_used += 2;
}
extern "C" int tfrt_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
    return 0;
}
extern "C" int tfrt_profile_read(double *total_ms, long long *n_launches) {
    TFRT_REQUIRE(total_ms && n_launches, "tfrt_profile_read: null argument");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double sum = 0.0;
    for (size_t i = 0; i + 1 < g_prof_used; i += 2) {
        TFRT_CHECK_CUDA(cudaEventSynchronize(g_prof_events[i + 1]));
        float ms = 0.f;
        TFRT_CHECK_CUDA(cudaEventElapsedTime(&ms, g_prof_events[i], g_prof_events[i + 1]));
        sum += ms;
    }
    *total_ms = sum;
    *n_launches = (long long)(g_prof_used / 2);
    g_prof_used = 0;
    return 0;
}

namespace {

// 8 independent accumulator chains per thread, register-only: measures the FMA pipe.
__global__ void __launch_bounds__(256) k_ffma_scalar(int iters, float seed, float *sink) {
    float a[8], b = seed, c = seed * 0.5f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) sink[0] = s;   // never true; keeps the loop alive
}

__global__ void __launch_bounds__(256) k_ffma_packed(int iters, float seed, float *sink) {
    float2 a[8], b = make_float2(seed, seed * 1.5f), c = make_float2(seed * 0.5f, seed * 0.25f);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i, seed - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], b, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) sink[0] = s;
}

// Mixed loop: 8 packed FFMA2 chains + NS scalar FFMA chains per round.  Tells whether scalar
// FFMA can issue beside FFMA2 (two fp32 data paths) or shares one pipe with it.
template <int NS>
__global__ void __launch_bounds__(256) k_ffma_mixed(int iters, float seed, float *sink) {
    float2 a[8], b = make_float2(seed, seed * 1.5f), c = make_float2(seed * 0.5f, seed * 0.25f);
    float s1[NS];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i, seed - i);
#pragma unroll
    for (int i = 0; i < NS; ++i) s1[i] = seed * (i + 2);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = __ffma2_rn(a[i], b, c);
                if (i < NS) s1[i] = fmaf(s1[i], b.x, c.y);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
#pragma unroll
    for (int i = 0; i < NS; ++i) s += s1[i];
    if (s == 123.456f) sink[0] = s;
}

// Issue-mix probe: 8 FFMA2 chains + NALU integer LOP3/IADD ops + NXU MUFU.RCP per round.  Only
// the FFMA2 flops are counted, so the result tells how much of the FMA pipe survives when the
// other pipes are fed from the same issue port (the search kernel's mix is about 8 : 3 : 1).
template <int NALU, int NXU>
__global__ void __launch_bounds__(256) k_issue_mix(int iters, float seed, float *sink) {
    float2 a[8], b = make_float2(seed, seed * 1.5f), c = make_float2(seed * 0.5f, seed * 0.25f);
    uint32_t q[NALU > 0 ? NALU : 1];
    float x[NXU > 0 ? NXU : 1];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i, seed - i);
#pragma unroll
    for (int i = 0; i < (NALU > 0 ? NALU : 1); ++i)
        q[i] = __float_as_uint(seed) + i * 77u + threadIdx.x * 2654435761u;   // per-lane values
#pragma unroll
    for (int i = 0; i < (NXU > 0 ? NXU : 1); ++i) x[i] = seed + 0.37f * i + 0.001f * threadIdx.x;
    const uint32_t m1 = __float_as_uint(seed * 3.f) ^ threadIdx.x, m2 = __float_as_uint(seed * 7.f) + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = __ffma2_rn(a[i], b, c);
                if (i < NALU) q[i] = ((q[i] | m1) & 0x80000000u) ^ (q[i] + m2);   // LOP3 + IADD
                if (i < NXU) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
#pragma unroll
    for (int i = 0; i < (NALU > 0 ? NALU : 1); ++i) s += __uint_as_float(q[i]);
#pragma unroll
    for (int i = 0; i < (NXU > 0 ? NXU : 1); ++i) s += x[i];
    if (s == 123.456f) sink[0] = s;
}

}  // namespace

extern "C" int tfrt_fp32_peak(int packed, int iters, double *flops_per_s, void *stream) {
    TFRT_REQUIRE(flops_per_s && iters > 0, "tfrt_fp32_peak: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    TFRT_CHECK_CUDA(cudaGetDevice(&dev));
    TFRT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *sink = nullptr;
    TFRT_CHECK_CUDA(cudaMalloc(&sink, sizeof(float)));
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    TFRT_CHECK_CUDA(cudaEventCreate(&e0));
    TFRT_CHECK_CUDA(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {   // first repetition is the warm-up
        TFRT_CHECK_CUDA(cudaEventRecord(e0, st));
        if (packed == 2)
            k_ffma_mixed<8><<<blocks, threads, 0, st>>>(iters, 1.0001f, sink);
        else if (packed == 3)
            k_ffma_mixed<4><<<blocks, threads, 0, st>>>(iters, 1.0001f, sink);
        else if (packed == 4)
            k_issue_mix<2, 0><<<blocks, threads, 0, st>>>(iters, 1.0001f, sink);   // 8 : 4 ALU
        else if (packed == 5)
            k_issue_mix<0, 1><<<blocks, threads, 0, st>>>(iters, 1.0001f, sink);   // 8 : 1 XU
        else if (packed == 6)
            k_issue_mix<2, 1><<<blocks, threads, 0, st>>>(iters, 1.0001f, sink);   // 8 : 4 : 1
        else if (packed == 7)
            k_issue_mix<4, 0><<<blocks, threads, 0, st>>>(iters, 1.0001f, sink);   // 8 : 8 ALU
        else if (packed)
            k_ffma_packed<<<blocks, threads, 0, st>>>(iters, 1.0001f, sink);
        else
            k_ffma_scalar<<<blocks, threads, 0, st>>>(iters, 1.0001f, sink);
        TFRT_CHECK_CUDA(cudaEventRecord(e1, st));
        TFRT_CHECK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        TFRT_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    // lane-FMAs per thread per outer iteration: 64 scalar, 128 packed, 128 + 8*NS mixed
    const double per_iter = packed == 2 ? 192.0 : packed == 3 ? 160.0 : packed ? 128.0 : 64.0;   // modes >= 4 count their FFMA2 only
    const double fmas = (double)blocks * threads * (double)iters * per_iter;
    *flops_per_s = 2.0 * fmas / (best_ms * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}
