// k1_bench.cu -- stand-alone timing harness for the ray x triangle search kernel
// (csrc/intersect3d.cu), used to compare tuning variants compiled with -D flags
// (tools/k1_variants.sh).  Synthetic scene: a cfg3-like bumpy lens sheet of T triangles hit by
// N rays travelling +x; prints best-of-5 time, tests/s, the 45-FLOP TFLOP/s figure and a
// checksum of the hit records so that variants can be checked against each other.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../tensorflowraytrace_b200/csrc/common.cuh"

int tfrt_launch_intersect3d(const float *const ray[6], const int32_t *live,
                            const int32_t *n_live_ptr, int n_upper, const float *packed,
                            int n_tri, float intersect_eps, float ray_start_eps, float size_eps,
                            uint64_t *hit, int dir_mode, unsigned int *work_counter,
                            float start_floor_ulps, cudaStream_t stream);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 2097152;
    const int side = argc > 2 ? atoi(argv[2]) : 100;   // T = 2 * side^2
    const int T = 2 * side * side;
    std::vector<float> tri(9 * (size_t)T), ray(6 * (size_t)N);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
    // sheet in the y-z plane, x = small bump; two triangles per grid cell
    auto px = [&](int i, int j) { return 0.05f * sinf(0.37f * i) * cosf(0.23f * j); };
    const float h = 2.0f / side;
    for (int i = 0; i < side; ++i)
        for (int j = 0; j < side; ++j)
            for (int k = 0; k < 2; ++k) {
                const size_t t = 2 * ((size_t)i * side + j) + k;
                int c[3][2] = {{i, j}, {i + 1, j}, {i, j + 1}};
                if (k) { c[0][0] = i + 1; c[0][1] = j + 1; c[1][0] = i; c[1][1] = j + 1; c[2][0] = i + 1; c[2][1] = j; }
                for (int v = 0; v < 3; ++v) {
                    tri[(3 * v + 0) * (size_t)T + t] = px(c[v][0], c[v][1]);
                    tri[(3 * v + 1) * (size_t)T + t] = -1.0f + h * c[v][0];
                    tri[(3 * v + 2) * (size_t)T + t] = -1.0f + h * c[v][1];
                }
            }
    for (int i = 0; i < N; ++i) {
        const float y = -0.95f + 1.9f * rnd(), z = -0.95f + 1.9f * rnd();
        ray[0 * (size_t)N + i] = -4.0f; ray[1 * (size_t)N + i] = y; ray[2 * (size_t)N + i] = z;
        ray[3 * (size_t)N + i] = -3.0f; ray[4 * (size_t)N + i] = y + 0.05f * (rnd() - 0.5f);
        ray[5 * (size_t)N + i] = z + 0.05f * (rnd() - 0.5f);
    }
    float *d_tri, *d_ray, *d_packed;
    uint64_t *d_hit;
    CK(cudaMalloc(&d_tri, tri.size() * 4)); CK(cudaMalloc(&d_ray, ray.size() * 4));
    CK(cudaMalloc(&d_packed, tfrt_packed_floats(T) * 4)); CK(cudaMalloc(&d_hit, ((size_t)N + 1) * 8));
    CK(cudaMemcpy(d_tri, tri.data(), tri.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ray, ray.data(), ray.size() * 4, cudaMemcpyHostToDevice));
    const float *trip[9], *rayp[6];
    for (int c = 0; c < 9; ++c) trip[c] = d_tri + (size_t)c * T;
    for (int c = 0; c < 6; ++c) rayp[c] = d_ray + (size_t)c * N;
    if (tfrt_pack_triangles(trip, T, 1e-5f, 0.f, d_packed, 0)) { printf("pack: %s\n", tfrt_last_error()); return 1; }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {
        // the launcher's contract: hits preset to ~0, work counter (the word after them) zero
        CK(cudaMemsetAsync(d_hit, 0xff, (size_t)N * 8));
        CK(cudaMemsetAsync(d_hit + N, 0, 8));
        CK(cudaEventRecord(e0));
        if (tfrt_launch_intersect3d(rayp, nullptr, nullptr, N, d_packed, T, 1e-10f, 1e-4f, 1e-5f, d_hit, 0,
                                    reinterpret_cast<unsigned int *>(d_hit + N), 0.f, 0)) {
            printf("launch: %s\n", tfrt_last_error()); return 1;
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    std::vector<uint64_t> hit(N);
    CK(cudaMemcpy(hit.data(), d_hit, (size_t)N * 8, cudaMemcpyDeviceToHost));
    unsigned long long sum = 0, nhit = 0;
    for (int i = 0; i < N; ++i) { sum += (hit[i] & 0xffffffffu) * (unsigned long long)(i % 7 + 1); nhit += (uint32_t)(hit[i] >> 32) < 0x7f800000u; }
    const double tests = (double)N * T;
    printf("N=%d T=%d  %.3f ms  %.4f Ttests/s  %.2f TFLOP/s(45)  hits=%llu checksum=%llu\n", N, T, best,
           tests / best * 1e-9, 45.0 * tests / best * 1e-9 / 1e3 * 1e3 / 1e3, nhit, sum);
    return 0;
}
