// pack.cu -- triangle precompute: per-triangle plane coefficients for the intersection kernel,
// and K3 (vertices -> triangle fields, and its VJP).
#include "common.cuh"

namespace {

// One CTA per tile of TFRT_TILE_TRIS triangles, one thread per triangle: computes the projected
// coefficients described in common.cuh and writes them into the tile's pair records, at the
// triangle's position in the tile's stable partition by dominant axis (k = 0 first, then 1, then
// 2; index order inside a group).  Odd groups are completed with a never-hit half, unused pair
// slots are filled with never-hit pairs; record 0 of the tile carries the pairs per group.
__global__ void __launch_bounds__(TFRT_TILE_TRIS)
k_pack_triangles(const float *__restrict__ xp, const float *__restrict__ yp,
                 const float *__restrict__ zp, const float *__restrict__ x1,
                 const float *__restrict__ y1, const float *__restrict__ z1,
                 const float *__restrict__ x2, const float *__restrict__ y2,
                 const float *__restrict__ z2, int n_tri, float size_eps, float margin_ulps,
                 float *__restrict__ packed) {
    __shared__ int s_cnt[3][TFRT_TILE_TRIS / 32];
    __shared__ float s_scale[TFRT_TILE_TRIS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * TFRT_TILE_TRIS + threadIdx.x;
    const bool real = i < n_tri;
    float scale = 0.f;     // largest |coordinate| of this thread's triangle
    // never-hit coefficients (padding, degenerate triangles): S = 4, E = 0 => x = 4 > C
    int k = 0;
    float n1 = 0.f, n2 = 0.f, dd = 0.f, nk = 0.f;
    float au1 = 0.f, au2 = 0.f, au0 = 4.f, bv1 = 0.f, bv2 = 0.f, bv0 = 0.f;
    if (real) {
        const double A[3] = {xp[i], yp[i], zp[i]};
        const double e1[3] = {x1[i] - A[0], y1[i] - A[1], z1[i] - A[2]};
        const double e2[3] = {x2[i] - A[0], y2[i] - A[1], z2[i] - A[2]};
        const double N[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2],
                             e1[0] * e2[1] - e1[1] * e2[0]};
        const double nn = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
        for (int c = 0; c < 3; ++c)
            scale = fmaxf(scale, (float)fmax(fabs(A[c]), fmax(fabs(A[c] + e1[c]), fabs(A[c] + e2[c]))));
        if (!isfinite(scale)) scale = 0.f;
        if (nn > 0.0 && isfinite(nn)) {
            const double inv = 1.0 / nn;
            // a = (e2 x N)/|N|^2 : u = a.(P-A);  b = (N x e1)/|N|^2 : v = b.(P-A)
            const double av[3] = {(e2[1] * N[2] - e2[2] * N[1]) * inv,
                                  (e2[2] * N[0] - e2[0] * N[2]) * inv,
                                  (e2[0] * N[1] - e2[1] * N[0]) * inv};
            const double bv[3] = {(N[1] * e1[2] - N[2] * e1[1]) * inv,
                                  (N[2] * e1[0] - N[0] * e1[2]) * inv,
                                  (N[0] * e1[1] - N[1] * e1[0]) * inv};
            // inflation of this triangle in the search: the reference's size_epsilion plus the
            // fp32 search margin (common.cuh); the fp64 evaluation of the winner applies the
            // exact rule afterwards
            const double eps_t = (double)size_eps +
                                 fmin(tfrt_tri_margin(A, e1, e2, (double)margin_ulps), 0.25);
            const double ad = -(av[0] * A[0] + av[1] * A[1] + av[2] * A[2]) + eps_t;
            const double bd = -(bv[0] * A[0] + bv[1] * A[1] + bv[2] * A[2]) + eps_t;
            // dominant axis of the normal; plane scaled so that its k-coefficient is 1:
            //   P[k] = dd - m1 P[k1] - m2 P[k2]  for every P on the plane
            k = fabs(N[1]) > fabs(N[0]) ? 1 : 0;
            if (fabs(N[2]) > fabs(N[k])) k = 2;
            const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
            const double m1 = N[k1] / N[k], m2 = N[k2] / N[k];
            const double dk = A[k] + m1 * A[k1] + m2 * A[k2];
            n1 = (float)m1; n2 = (float)m2; dd = (float)dk; nk = (float)N[k];
            // u = a.P + ad with P[k] eliminated (exact on the plane)
            const double u1 = av[k1] - av[k] * m1, u2 = av[k2] - av[k] * m2, u0 = ad + av[k] * dk;
            const double v1 = bv[k1] - bv[k] * m1, v2 = bv[k2] - bv[k] * m2, v0 = bd + bv[k] * dk;
            // the kernel tests the planes S = 2(u+v) - C_t and E = u - v (see intersect3d.cu)
            // against x = |E| + |S - |E|| <= C_t with C_t = 1 + 3 eps_t; x is homogeneous of
            // degree one in (S, E), so both planes are scaled by C / C_t and the kernel compares
            // every triangle with the one constant C = 1 + 3 size_eps
            const double Ct = 1.0 + 3.0 * eps_t;
            const double rho = (1.0 + 3.0 * (double)size_eps) / Ct;
            au1 = (float)(rho * 2.0 * (u1 + v1)); au2 = (float)(rho * 2.0 * (u2 + v2));
            au0 = (float)(rho * (2.0 * (u0 + v0) - Ct));
            bv1 = (float)(rho * (u1 - v1)); bv2 = (float)(rho * (u2 - v2));
            bv0 = (float)(rho * (u0 - v0));
        }
    }
    // stable partition by k inside the tile (padding threads take part in no group)
    const int kk = real ? k : 3;
    int rank_in_warp = 0;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        const unsigned m = __ballot_sync(0xffffffffu, kk == g);
        if (lane == 0) s_cnt[g][wid] = __popc(m);
        if (kk == g) rank_in_warp = __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    int total[3] = {0, 0, 0}, before = 0;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int w = 0; w < TFRT_TILE_TRIS / 32; ++w) {
            const int c = s_cnt[g][w];
            if (g == kk && w < wid) before += c;
            total[g] += c;
        }
    }
    const int pairs[3] = {(total[0] + 1) / 2, (total[1] + 1) / 2, (total[2] + 1) / 2};
    const int n_pairs = pairs[0] + pairs[1] + pairs[2];
    float *tile = packed + (size_t)blockIdx.x * TFRT_TILE_PAIRS * 24;
    // writes one half (h = 0 / 1) of pair record `pair`
    auto put = [&](int pair, int h, float m1_, float m2_, float dd_, float s1_, float s2_,
                   float s0_, float e1_, float e2_, float e0_, float idx_, float nk_) {
        float *r = tile + (size_t)pair * 24 + h;
        r[0] = m1_;  r[2] = m2_;
        r[4] = dd_;  r[6] = s1_;
        r[8] = s2_;  r[10] = s0_;
        r[12] = e1_; r[14] = e2_;
        r[16] = e0_; r[18] = idx_;
        r[20] = nk_;
    };
    if (real) {
        const int rank = before + rank_in_warp;
        const int pair = (k == 0 ? 0 : (k == 1 ? pairs[0] : pairs[0] + pairs[1])) + rank / 2;
        put(pair, rank & 1, n1, n2, dd, au1, au2, au0, bv1, bv2, bv0, __int_as_float(i), nk);
        // the last triangle of an odd group completes its pair with a never-hit half
        if ((rank & 1) == 0 && rank == total[k] - 1)
            put(pair, 1, 0.f, 0.f, 0.f, 0.f, 0.f, 4.f, 0.f, 0.f, 0.f, __int_as_float(i), 0.f);
    }
    // unused pair slots (never visited by the kernel: the header bounds the loops) stay defined
    for (int pslot = n_pairs + threadIdx.x; pslot < TFRT_TILE_PAIRS; pslot += TFRT_TILE_TRIS)
        for (int h = 0; h < 2; ++h)
            put(pslot, h, 0.f, 0.f, 0.f, 0.f, 0.f, 4.f, 0.f, 0.f, 0.f, __int_as_float(0), 0.f);
    __syncthreads();   // the header and the pad word go in after the records
    // word 23 of the very first record carries the scene scale (largest |coordinate| of the
    // mesh, zeroed by the host before this launch): the search derives its ray-start floor from it
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) scale = fmaxf(scale, __shfl_xor_sync(0xffffffffu, scale, off));
    if (lane == 0) s_scale[wid] = scale;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < TFRT_TILE_TRIS / 32; ++w) scale = fmaxf(scale, s_scale[w]);
        atomicMax(reinterpret_cast<unsigned int *>(packed + 23), __float_as_uint(scale));
    }
    if (threadIdx.x < TFRT_TILE_PAIRS && !(blockIdx.x == 0 && threadIdx.x == 0))
        tile[(size_t)threadIdx.x * 24 + 23] = 0.f;
    if (threadIdx.x < TFRT_TILE_PAIRS)
        tile[(size_t)threadIdx.x * 24 + 22] =
            __int_as_float(threadIdx.x == 0 ? (pairs[0] | (pairs[1] << 8) | (pairs[2] << 16)) : 0);
}

// K3 forward: gather face corners, unit normal normalize((p2-p1) x (p3-p2))
// (tfrt/boundaries.py:890-923).
__global__ void k_faces_forward(const float *__restrict__ vertices,
                                const int32_t *__restrict__ faces, int n_faces, float *xp,
                                float *yp, float *zp, float *x1, float *y1, float *z1, float *x2,
                                float *y2, float *z2, float *norm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_faces) return;
    const int i0 = faces[3 * i], i1 = faces[3 * i + 1], i2 = faces[3 * i + 2];
    const float p0[3] = {vertices[3 * i0], vertices[3 * i0 + 1], vertices[3 * i0 + 2]};
    const float p1[3] = {vertices[3 * i1], vertices[3 * i1 + 1], vertices[3 * i1 + 2]};
    const float p2[3] = {vertices[3 * i2], vertices[3 * i2 + 1], vertices[3 * i2 + 2]};
    xp[i] = p0[0]; yp[i] = p0[1]; zp[i] = p0[2];
    x1[i] = p1[0]; y1[i] = p1[1]; z1[i] = p1[2];
    x2[i] = p2[0]; y2[i] = p2[1]; z2[i] = p2[2];
    const double ux = (double)p1[0] - p0[0], uy = (double)p1[1] - p0[1], uz = (double)p1[2] - p0[2];
    const double vx = (double)p2[0] - p1[0], vy = (double)p2[1] - p1[1], vz = (double)p2[2] - p1[2];
    const double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
    const double inv = 1.0 / sqrt(cx * cx + cy * cy + cz * cz);
    norm[3 * i] = (float)(cx * inv);
    norm[3 * i + 1] = (float)(cy * inv);
    norm[3 * i + 2] = (float)(cz * inv);
}

// K3 backward: cotangents of (xp..z2, norm) -> cotangents of vertices (scatter-add).
//   norm = c/|c|, c = (p1-p0) x (p2-p1)
//   g_c  = (g_n - (g_n.n) n)/|c| ;  with u = p1-p0, v = p2-p1:  g_u = v x g_c, g_v = g_c x u
//   g_p0 -= g_u ; g_p1 += g_u - g_v ; g_p2 += g_v
// update_map[f][k] == 0 blocks the gradient of corner k of face f (tf.stop_gradient,
// tfrt/boundaries.py:900-913).
__global__ void k_faces_backward(const float *__restrict__ vertices,
                                 const int32_t *__restrict__ faces,
                                 const uint8_t *__restrict__ update_map, int n_faces,
                                 const float *g_xp, const float *g_yp, const float *g_zp,
                                 const float *g_x1, const float *g_y1, const float *g_z1,
                                 const float *g_x2, const float *g_y2, const float *g_z2,
                                 const float *g_norm, float *__restrict__ g_vertices) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_faces) return;
    const int idx[3] = {faces[3 * i], faces[3 * i + 1], faces[3 * i + 2]};
    double g[3][3];
    g[0][0] = g_xp ? g_xp[i] : 0.0; g[0][1] = g_yp ? g_yp[i] : 0.0; g[0][2] = g_zp ? g_zp[i] : 0.0;
    g[1][0] = g_x1 ? g_x1[i] : 0.0; g[1][1] = g_y1 ? g_y1[i] : 0.0; g[1][2] = g_z1 ? g_z1[i] : 0.0;
    g[2][0] = g_x2 ? g_x2[i] : 0.0; g[2][1] = g_y2 ? g_y2[i] : 0.0; g[2][2] = g_z2 ? g_z2[i] : 0.0;
    if (g_norm) {
        double p[3][3];
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 3; ++c) p[k][c] = vertices[3 * idx[k] + c];
        const double u[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
        const double v[3] = {p[2][0] - p[1][0], p[2][1] - p[1][1], p[2][2] - p[1][2]};
        const double c[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2],
                             u[0] * v[1] - u[1] * v[0]};
        const double len = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        const double n[3] = {c[0] / len, c[1] / len, c[2] / len};
        const double gn[3] = {g_norm[3 * i], g_norm[3 * i + 1], g_norm[3 * i + 2]};
        const double dot = gn[0] * n[0] + gn[1] * n[1] + gn[2] * n[2];
        const double gc[3] = {(gn[0] - dot * n[0]) / len, (gn[1] - dot * n[1]) / len,
                              (gn[2] - dot * n[2]) / len};
        const double gu[3] = {v[1] * gc[2] - v[2] * gc[1], v[2] * gc[0] - v[0] * gc[2],
                              v[0] * gc[1] - v[1] * gc[0]};
        const double gv[3] = {gc[1] * u[2] - gc[2] * u[1], gc[2] * u[0] - gc[0] * u[2],
                              gc[0] * u[1] - gc[1] * u[0]};
        for (int cdim = 0; cdim < 3; ++cdim) {
            g[0][cdim] -= gu[cdim];
            g[1][cdim] += gu[cdim] - gv[cdim];
            g[2][cdim] += gv[cdim];
        }
    }
    for (int k = 0; k < 3; ++k) {
        if (update_map && !update_map[3 * i + k]) continue;
        for (int c = 0; c < 3; ++c) {
            const float val = (float)g[k][c];
            if (val != 0.f) atomicAdd(&g_vertices[3 * idx[k] + c], val);
        }
    }
}

}  // namespace

extern "C" size_t tfrt_packed_floats(int32_t n_tri) {
    const size_t tiles = (size_t)tfrt_div_up(n_tri > 0 ? n_tri : 1, TFRT_TILE_TRIS);
    return tiles * TFRT_TILE_PAIRS * 2 * TFRT_TRI_FLOATS;
}

extern "C" int tfrt_pack_triangles(const float *const tri[9], int32_t n_tri, float size_eps,
                                   float margin_ulps, float *packed, void *stream) {
    TFRT_REQUIRE(n_tri >= 0 && packed && margin_ulps >= 0.f, "tfrt_pack_triangles: bad arguments");
    if (n_tri == 0) return 0;
    TFRT_CHECK_CUDA(cudaMemsetAsync(packed + 23, 0, sizeof(float), (cudaStream_t)stream));
    k_pack_triangles<<<tfrt_div_up(n_tri, TFRT_TILE_TRIS), TFRT_TILE_TRIS, 0, (cudaStream_t)stream>>>(
        tri[0], tri[1], tri[2], tri[3], tri[4], tri[5], tri[6], tri[7], tri[8], n_tri, size_eps,
        margin_ulps, packed);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfrt_faces_forward(const float *vertices, const int32_t *faces, int32_t n_faces,
                                  float *const tri_out[9], float *norm_out, void *stream) {
    TFRT_REQUIRE(vertices && faces && norm_out, "tfrt_faces_forward: null pointer");
    if (n_faces <= 0) return 0;
    k_faces_forward<<<tfrt_div_up(n_faces, 128), 128, 0, (cudaStream_t)stream>>>(
        vertices, faces, n_faces, tri_out[0], tri_out[1], tri_out[2], tri_out[3], tri_out[4],
        tri_out[5], tri_out[6], tri_out[7], tri_out[8], norm_out);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfrt_faces_backward(const float *vertices, const int32_t *faces,
                                   const uint8_t *update_map, int32_t n_faces, int32_t n_vertices,
                                   const float *const g_tri[9], const float *g_norm,
                                   float *g_vertices, void *stream) {
    TFRT_REQUIRE(vertices && faces && g_vertices, "tfrt_faces_backward: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    TFRT_CHECK_CUDA(cudaMemsetAsync(g_vertices, 0, sizeof(float) * 3 * (size_t)n_vertices, st));
    if (n_faces <= 0) return 0;
    k_faces_backward<<<tfrt_div_up(n_faces, 128), 128, 0, st>>>(
        vertices, faces, update_map, n_faces, g_tri[0], g_tri[1], g_tri[2], g_tri[3], g_tri[4],
        g_tri[5], g_tri[6], g_tri[7], g_tri[8], g_norm, g_vertices);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}
