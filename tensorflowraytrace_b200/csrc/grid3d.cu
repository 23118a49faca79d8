// grid3d.cu -- optional spatial acceleration of the nearest-hit search: a uniform grid over the
// triangles, built on the device at every system.update(), walked per ray with a 3-D DDA.
//
// This is NOT the path the tests/s metric is quoted on (that is the brute-force kernel of
// intersect3d.cu, which tests every live ray against every triangle); it is the reference's
// OpticalSystem3D._intersection (tfrt/engine.py:1102-1166) answered faster: SURVEY.md 8(f) rank 4.
//
// Same answers as the brute-force search, by construction:
//  * every candidate triangle is tested with the brute-force kernel's own arithmetic -- the same
//    projected plane coefficients (unpacked from the tile records of tfrt_pack_triangles into a
//    per-triangle table), the same operation order, fused multiply-adds where the packed kernel
//    fuses, the same MUFU.RCP -- so a (ray, triangle) pair gets bit-identical t and inside value;
//  * the winner is the lexicographic minimum of (bits(t), triangle index), like there;
//  * the walk is conservative: triangles are registered in every cell their bounding box
//    (inflated by 3 % of a cell + the search band) touches, and it only stops once the best hit
//    lies before the exit of the current cell (minus the same slack), so the brute-force winner
//    is always among the candidates.
//  The hit words then go through the same re-validation / reaction kernel (trace3d.cu).
//
// Triangles much larger than the typical one (a target plane behind a 100 k-triangle lens) would
// flood the grid; they are kept in a separate list that every ray tests directly, and the grid
// spans only the bounding box of the others.
#include "common.cuh"

namespace {

constexpr int kGridThreads = 256;
constexpr int kMaxCellsPerTriangle = 64;   // more => "large" list
constexpr float kLargeFactor = 16.f;       // extent > 16 x the typical extent => "large" list

// Device-resident header of a grid (computed on the device: no host round trip at build time).
struct GridHeader {
    float lo[3], hi[3];        // bounding box of the gridded (non-large) triangles
    float cell[3], inv_cell[3];
    int dims[3];
    int n_cells;
    int n_large;
    int n_entries;             // total (cell, triangle) pairs
    int overflow;              // entries that did not fit (0 in a valid grid)
    float log_extent_sum;      // sum of log(extent) over finite, non-degenerate triangles
    int n_extent;
    float typical;             // exp(mean log extent)
    unsigned int lo_bits[3], hi_bits[3];   // atomics scratch (ordered-int encoding)
};

__device__ __forceinline__ unsigned int f2ord(float f) {   // order-preserving float -> uint
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct Tri9 { const float *c[9]; };

__device__ __forceinline__ bool tri_box(const Tri9 &T, int i, float lo[3], float hi[3]) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float a = T.c[k][i], b = T.c[3 + k][i], c = T.c[6 + k][i];
        lo[k] = fminf(a, fminf(b, c));
        hi[k] = fmaxf(a, fmaxf(b, c));
        ok = ok && isfinite(lo[k]) && isfinite(hi[k]);
    }
    return ok;
}

__device__ __forceinline__ float tri_extent(const float lo[3], const float hi[3]) {
    return fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
}

__global__ void k_grid_reset(GridHeader *h) {
    if (threadIdx.x == 0) {
        for (int k = 0; k < 3; ++k) { h->lo_bits[k] = 0xffffffffu; h->hi_bits[k] = 0u; }
        h->n_large = 0; h->n_entries = 0; h->overflow = 0;
        h->log_extent_sum = 0.f; h->n_extent = 0; h->typical = 0.f; h->n_cells = 0;
    }
}

// pass 1: typical triangle extent = geometric mean (robust against a few huge triangles)
__global__ void k_grid_typical(Tri9 T, int n_tri, GridHeader *h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.f;
    int cnt = 0;
    if (i < n_tri) {
        float lo[3], hi[3];
        if (tri_box(T, i, lo, hi)) {
            const float e = tri_extent(lo, hi);
            if (e > 0.f) { v = logf(e); cnt = 1; }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, off);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(&h->log_extent_sum, v);
        atomicAdd(&h->n_extent, cnt);
    }
}

__global__ void k_grid_typical_finish(GridHeader *h) {
    if (threadIdx.x == 0)
        h->typical = h->n_extent > 0 ? expf(h->log_extent_sum / (float)h->n_extent) : 1.f;
}

// pass 2: bounding box of the non-large triangles
__global__ void k_grid_bounds(Tri9 T, int n_tri, GridHeader *h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tri) return;
    float lo[3], hi[3];
    if (!tri_box(T, i, lo, hi)) return;
    if (tri_extent(lo, hi) > kLargeFactor * h->typical) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        atomicMin(&h->lo_bits[k], f2ord(lo[k]));
        atomicMax(&h->hi_bits[k], f2ord(hi[k]));
    }
}

// pass 3: resolution.  Cells of about two typical extents, at most max_cells cells.
__global__ void k_grid_setup(GridHeader *h, int max_cells, float pad_world) {
    if (threadIdx.x != 0) return;
    float ext[3];
    bool empty = h->hi_bits[0] == 0u;
    for (int k = 0; k < 3; ++k) {
        float lo = empty ? 0.f : ord2f(h->lo_bits[k]), hi = empty ? 1.f : ord2f(h->hi_bits[k]);
        const float pad = 0.05f * h->typical + pad_world + 1e-6f * fmaxf(fabsf(lo), fabsf(hi));
        lo -= pad; hi += pad;
        h->lo[k] = lo; h->hi[k] = hi; ext[k] = hi - lo;
    }
    float cell = 2.f * h->typical;
    for (int iter = 0; iter < 64; ++iter) {
        double n = 1.0;
        for (int k = 0; k < 3; ++k) n *= ceil(fmax(ext[k] / cell, 1.0f));
        if (n <= (double)max_cells) break;
        cell *= 1.26f;
    }
    int total = 1;
    for (int k = 0; k < 3; ++k) {
        int d = (int)ceilf(fmaxf(ext[k] / cell, 1.f));
        d = d < 1 ? 1 : (d > 1024 ? 1024 : d);
        h->dims[k] = d;
        h->cell[k] = ext[k] / (float)d;
        h->inv_cell[k] = (float)d / ext[k];
        total *= d;
    }
    h->n_cells = total;
}

// cell range covered by a triangle's inflated box; returns false for a "large" triangle
__device__ __forceinline__ bool cell_range(const GridHeader *h, const float lo[3],
                                           const float hi[3], float band, int c0[3], int c1[3]) {
    if (tri_extent(lo, hi) > kLargeFactor * h->typical) return false;
    long long count = 1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float pad = 0.03f * h->cell[k] + band;
        int a = (int)floorf((lo[k] - pad - h->lo[k]) * h->inv_cell[k]);
        int b = (int)floorf((hi[k] + pad - h->lo[k]) * h->inv_cell[k]);
        a = a < 0 ? 0 : a;
        b = b >= h->dims[k] ? h->dims[k] - 1 : b;
        if (b < a) b = a;
        c0[k] = a; c1[k] = b;
        count *= (b - a + 1);
    }
    return count <= kMaxCellsPerTriangle;
}

// pass 4 / 6: count, then fill (same traversal; `fill` selects)
__global__ void k_grid_bin(Tri9 T, int n_tri, GridHeader *h, float size_eps, float margin_ulps,
                           const float *__restrict__ packed, int *cell_count,
                           const int *cell_start, int *entries, int max_entries, int *large,
                           int fill) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tri) return;
    float lo[3], hi[3];
    if (!tri_box(T, i, lo, hi)) return;           // never-hit record in the search as well
    int c0[3], c1[3];
    // reach of the search's acceptance band beyond the triangle: (size_eps + margin) of an
    // altitude, the margin being margin_ulps ulps of the scene's coordinates as a distance
    // (capped at a quarter altitude by the pack kernel); an extent bounds an altitude
    const float extent = tri_extent(lo, hi);
    const float band = fminf(1.5f * size_eps * extent + 2.f * margin_ulps * 1.1920929e-7f * packed[23],
                             0.35f * extent) + 1e-6f * extent;
    if (!cell_range(h, lo, hi, band, c0, c1)) {
        if (!fill) {
            const int slot = atomicAdd(&h->n_large, 1);
            large[slot] = i;
        }
        return;
    }
    for (int z = c0[2]; z <= c1[2]; ++z)
        for (int y = c0[1]; y <= c1[1]; ++y)
            for (int x = c0[0]; x <= c1[0]; ++x) {
                const int c = (z * h->dims[1] + y) * h->dims[0] + x;
                if (!fill) {
                    atomicAdd(&cell_count[c], 1);
                } else {
                    const int pos = cell_start[c] + atomicAdd(&cell_count[c], 1);
                    if (pos < max_entries) entries[pos] = i;
                    else atomicAdd(&h->overflow, 1);
                }
            }
}

// pass 5: exclusive scan of the cell counts (single CTA; cells <= a few million)
__global__ void __launch_bounds__(1024) k_grid_scan(GridHeader *h, int *cell_count,
                                                    int *cell_start) {
    __shared__ int s_part[1024];
    const int n = h->n_cells;
    const int per = (n + 1023) / 1024;
    const int b0 = threadIdx.x * per, b1 = min(n, b0 + per);
    int sum = 0;
    for (int c = b0; c < b1; ++c) sum += cell_count[c];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int v = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_part[threadIdx.x] - sum;
    for (int c = b0; c < b1; ++c) {
        const int v = cell_count[c];
        cell_start[c] = run;
        cell_count[c] = 0;          // reused as the fill cursor
        run += v;
    }
    if (threadIdx.x == 1023) {
        cell_start[n] = s_part[1023];
        h->n_entries = s_part[1023];
    }
}

// sort each cell's triangle ids ascending: the walk then tests candidates in index order (not
// needed for the result -- the winner is a lexicographic minimum -- but it makes the traversal,
// and hence timing, deterministic)
__global__ void k_grid_sort_cells(const GridHeader *h, const int *cell_start, int *entries,
                                  int max_entries) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= h->n_cells) return;
    const int a = cell_start[c], b = min(cell_start[c + 1], max_entries);
    for (int i = a + 1; i < b; ++i) {
        const int v = entries[i];
        int j = i - 1;
        while (j >= a && entries[j] > v) { entries[j + 1] = entries[j]; --j; }
        entries[j + 1] = v;
    }
}

// per-triangle coefficient table in merged-index order, copied out of the pair records
struct __align__(16) TriRecord {
    float m1, m2, dd, s1, s2, s0, e1, e2, e0, nk;
    int k;
    int pad;
};

__global__ void k_grid_unpack(const float *__restrict__ packed, int n_tiles, int n_tri,
                              TriRecord *__restrict__ table) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;      // pair record
    if (q >= n_tiles * TFRT_TILE_PAIRS) return;
    const int tile = q / TFRT_TILE_PAIRS, j = q - tile * TFRT_TILE_PAIRS;
    const float *rec0 = packed + (size_t)tile * TFRT_TILE_PAIRS * 24;
    const int header = __float_as_int(rec0[22]);
    const int n0 = header & 0xff, n1 = n0 + ((header >> 8) & 0xff), n2 = n1 + ((header >> 16) & 0xff);
    if (j >= n2) return;
    const int k = j < n0 ? 0 : (j < n1 ? 1 : 2);
    const float *r = rec0 + (size_t)j * 24;
    for (int hlf = 0; hlf < 2; ++hlf) {
        if (r[10 + hlf] == 4.f && r[0 + hlf] == 0.f && r[6 + hlf] == 0.f && r[20 + hlf] == 0.f)
            continue;                                         // never-hit half (padding)
        const int idx = __float_as_int(r[18 + hlf]);
        if (idx < 0 || idx >= n_tri) continue;
        TriRecord t;
        t.m1 = r[0 + hlf]; t.m2 = r[2 + hlf]; t.dd = r[4 + hlf];
        t.s1 = r[6 + hlf]; t.s2 = r[8 + hlf]; t.s0 = r[10 + hlf];
        t.e1 = r[12 + hlf]; t.e2 = r[14 + hlf]; t.e0 = r[16 + hlf];
        t.nk = r[20 + hlf]; t.k = k; t.pad = 0;
        table[idx] = t;
    }
}

__global__ void k_grid_table_init(TriRecord *table, int n_tri) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tri) return;
    TriRecord t;
    t.m1 = t.m2 = t.dd = t.s1 = t.s2 = 0.f; t.s0 = 4.f; t.e1 = t.e2 = t.e0 = 0.f; t.nk = 0.f;
    t.k = 0; t.pad = 0;
    table[i] = t;                                             // degenerate triangles: never hit
}

// The brute-force kernel's test of one (ray, triangle) pair, operation for operation
// (intersect3d.cu: the packed instructions are IEEE fused multiply-adds / adds / multiplies per
// lane, so the scalar forms below give the same bits).
struct RayReg { float o[3], d[3]; };

__device__ __forceinline__ void test_triangle(const TriRecord &t, const RayReg &r, int tri,
                                              uint32_t c_bits, float intersect_eps,
                                              uint32_t &best, int &bidx) {
    const int k = t.k;
    const float ok = k == 0 ? r.o[0] : (k == 1 ? r.o[1] : r.o[2]);
    const float ok1 = k == 0 ? r.o[1] : (k == 1 ? r.o[2] : r.o[0]);
    const float ok2 = k == 0 ? r.o[2] : (k == 1 ? r.o[0] : r.o[1]);
    const float dk = k == 0 ? r.d[0] : (k == 1 ? r.d[1] : r.d[2]);
    const float dk1 = k == 0 ? r.d[1] : (k == 1 ? r.d[2] : r.d[0]);
    const float dk2 = k == 0 ? r.d[2] : (k == 1 ? r.d[0] : r.d[1]);
    float det = fmaf(t.m1, dk1, dk);
    det = fmaf(t.m2, dk2, det);
    float tq = __fadd_rn(t.dd, -ok);
    tq = fmaf(-t.m1, ok1, tq);
    tq = fmaf(-t.m2, ok2, tq);
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(det));
    const float tt = __fmul_rn(tq, inv);
    const float p1 = fmaf(tt, dk1, ok1), p2 = fmaf(tt, dk2, ok2);
    float sv = fmaf(t.s1, p1, t.s0);
    sv = fmaf(t.s2, p2, sv);
    float ev = fmaf(t.e1, p1, t.e0);
    ev = fmaf(t.e2, p2, ev);
    const float ea = fabsf(ev);
    const float kk = __fadd_rn(sv, -ea);
    const float x = __fadd_rn(ea, fabsf(kk));
    const uint32_t tb = __float_as_uint(tt), xb = __float_as_uint(x);
    const bool okk = (xb <= c_bits) & ((tb < best) | ((tb == best) & (tri < bidx))) &
                     (tb < TFRT_NOHIT_KEY) & (fabsf(det * t.nk) >= intersect_eps);
    best = okk ? tb : best;
    bidx = okk ? tri : bidx;
}

__global__ void __launch_bounds__(kGridThreads)
k_grid_intersect(const float *__restrict__ sx, const float *__restrict__ sy,
                 const float *__restrict__ sz, const float *__restrict__ ex,
                 const float *__restrict__ ey, const float *__restrict__ ez,
                 const int32_t *__restrict__ live, const int32_t *__restrict__ n_live_ptr,
                 int n_live_static, const GridHeader *__restrict__ h,
                 const int *__restrict__ cell_start, const int *__restrict__ entries,
                 const int *__restrict__ large, const TriRecord *__restrict__ table,
                 const float *__restrict__ packed, float intersect_eps, float ray_start_eps,
                 float w_one, float start_floor_ulps, unsigned long long *__restrict__ hit,
                 int dir_mode) {
    const int n_live = n_live_ptr ? *n_live_ptr : n_live_static;
    const int p = blockIdx.x * kGridThreads + threadIdx.x;
    if (p >= n_live) return;
    const int slot = live ? live[p] : p;
    RayReg r;
    {
        const float ax = sx[slot], ay = sy[slot], az = sz[slot];
        r.d[0] = dir_mode ? ex[slot] : ex[slot] - ax;
        r.d[1] = dir_mode ? ey[slot] : ey[slot] - ay;
        r.d[2] = dir_mode ? ez[slot] : ez[slot] - az;
        // the brute-force kernel's start exclusion, same expressions
        const float start_floor = start_floor_ulps * 1.1920929e-7f * __ldg(packed + 23);
        const float start = fmaxf(ray_start_eps, start_floor * rsqrtf(
            fmaf(r.d[2], r.d[2], fmaf(r.d[1], r.d[1], __fmul_rn(r.d[0], r.d[0])))));
        r.o[0] = fmaf(start, r.d[0], ax);
        r.o[1] = fmaf(start, r.d[1], ay);
        r.o[2] = fmaf(start, r.d[2], az);
    }
    const uint32_t c_bits = __float_as_uint(w_one);
    uint32_t best = TFRT_NOHIT_KEY;
    int bidx = 0;

    // the large triangles: every ray tests them all
    const int n_large = h->n_large;
    for (int q = 0; q < n_large; ++q) {
        const int tri = large[q];
        test_triangle(table[tri], r, tri, c_bits, intersect_eps, best, bidx);
    }

    // clip the ray (t >= 0) against the grid's box
    float t0 = 0.f, t1 = __int_as_float(0x7f800000);
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (r.d[k] != 0.f) {
            const float inv = 1.f / r.d[k];
            float a = (h->lo[k] - r.o[k]) * inv, b = (h->hi[k] - r.o[k]) * inv;
            if (a > b) { const float s = a; a = b; b = s; }
            t0 = fmaxf(t0, a);
            t1 = fminf(t1, b);
        } else if (r.o[k] < h->lo[k] || r.o[k] > h->hi[k]) {
            inside = false;
        }
    }
    if (inside && t0 <= t1) {
        // enter slightly inside the box, then walk cell by cell
        int c[3], step[3];
        float t_next[3], t_delta[3];
        const float t_in = t0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float pos = fmaf(t_in, r.d[k], r.o[k]);
            int ci = (int)floorf((pos - h->lo[k]) * h->inv_cell[k]);
            ci = ci < 0 ? 0 : (ci >= h->dims[k] ? h->dims[k] - 1 : ci);
            c[k] = ci;
            if (r.d[k] > 0.f) {
                step[k] = 1;
                t_delta[k] = h->cell[k] / r.d[k];
                t_next[k] = (h->lo[k] + (ci + 1) * h->cell[k] - r.o[k]) / r.d[k];
            } else if (r.d[k] < 0.f) {
                step[k] = -1;
                t_delta[k] = -h->cell[k] / r.d[k];
                t_next[k] = (h->lo[k] + ci * h->cell[k] - r.o[k]) / r.d[k];
            } else {
                step[k] = 0;
                t_delta[k] = __int_as_float(0x7f800000);
                t_next[k] = __int_as_float(0x7f800000);
            }
        }
        // slack of the stop rule: 3 % of a cell along the ray, in ray-parameter units
        const float dmax = fmaxf(fabsf(r.d[0]) * h->inv_cell[0],
                                 fmaxf(fabsf(r.d[1]) * h->inv_cell[1], fabsf(r.d[2]) * h->inv_cell[2]));
        const float slack = dmax > 0.f ? 0.03f / dmax : 0.f;
        for (int guard = 0; guard < 4096; ++guard) {
            const int cell = (c[2] * h->dims[1] + c[1]) * h->dims[0] + c[0];
            const int a = cell_start[cell], b = cell_start[cell + 1];
            for (int q = a; q < b; ++q) {
                const int tri = entries[q];
                test_triangle(table[tri], r, tri, c_bits, intersect_eps, best, bidx);
            }
            const float t_exit = fminf(t_next[0], fminf(t_next[1], t_next[2]));
            // every triangle with a hit before t_exit is registered in a cell already visited
            if (best < TFRT_NOHIT_KEY && __uint_as_float(best) + slack < t_exit) break;
            if (t_exit > t1 + slack) break;
            const int axis = (t_next[0] <= t_next[1] && t_next[0] <= t_next[2]) ? 0
                             : (t_next[1] <= t_next[2] ? 1 : 2);
            // selected without indexed registers
            if (axis == 0) { c[0] += step[0]; t_next[0] += t_delta[0]; }
            else if (axis == 1) { c[1] += step[1]; t_next[1] += t_delta[1]; }
            else { c[2] += step[2]; t_next[2] += t_delta[2]; }
            if (c[0] < 0 || c[1] < 0 || c[2] < 0 || c[0] >= h->dims[0] || c[1] >= h->dims[1] ||
                c[2] >= h->dims[2])
                break;
        }
    }
    hit[p] = ((unsigned long long)best << 32) | (uint32_t)bidx;
}

// workspace layout (bytes, 256-aligned blocks): header | table | cell_count | cell_start | large | entries
struct GridLayout {
    size_t header, table, cell_count, cell_start, large, entries, total;
    int max_entries;
};

GridLayout grid_layout(int n_tri, int max_cells) {
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    GridLayout L;
    size_t off = 0;
    L.header = off; off += up(sizeof(GridHeader));
    L.table = off; off += up(sizeof(TriRecord) * (size_t)(n_tri > 0 ? n_tri : 1));
    L.cell_count = off; off += up(sizeof(int) * ((size_t)max_cells + 1));
    L.cell_start = off; off += up(sizeof(int) * ((size_t)max_cells + 1));
    L.large = off; off += up(sizeof(int) * (size_t)(n_tri > 0 ? n_tri : 1));
    const long long want = (long long)kMaxCellsPerTriangle * (n_tri > 0 ? n_tri : 1);
    L.max_entries = (int)(want > 0x7fffffffLL / 2 ? 0x7fffffffLL / 2 : want);
    L.entries = off; off += up(sizeof(int) * (size_t)L.max_entries);
    L.total = off;
    return L;
}

}  // namespace

extern "C" size_t tfrt_grid_bytes(int32_t n_tri, int32_t max_cells) {
    if (n_tri < 0 || max_cells < 1) return 0;
    return grid_layout(n_tri, max_cells).total;
}

extern "C" int tfrt_grid_build(const float *const tri[9], const float *packed, int32_t n_tri,
                               float size_eps, float margin_ulps, int32_t max_cells, void *grid_ws,
                               void *stream) {
    TFRT_REQUIRE(tri && packed && grid_ws && n_tri > 0 && max_cells >= 1,
                 "tfrt_grid_build: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const GridLayout L = grid_layout(n_tri, max_cells);
    char *base = reinterpret_cast<char *>(grid_ws);
    GridHeader *h = reinterpret_cast<GridHeader *>(base + L.header);
    TriRecord *table = reinterpret_cast<TriRecord *>(base + L.table);
    int *cell_count = reinterpret_cast<int *>(base + L.cell_count);
    int *cell_start = reinterpret_cast<int *>(base + L.cell_start);
    int *large = reinterpret_cast<int *>(base + L.large);
    int *entries = reinterpret_cast<int *>(base + L.entries);
    Tri9 T;
    for (int c = 0; c < 9; ++c) T.c[c] = tri[c];
    const int nb = tfrt_div_up(n_tri, kGridThreads);
    const int n_tiles = tfrt_div_up(n_tri, TFRT_TILE_TRIS);
    k_grid_reset<<<1, 32, 0, st>>>(h);
    TFRT_CHECK_CUDA(cudaMemsetAsync(cell_count, 0, sizeof(int) * ((size_t)max_cells + 1), st));
    k_grid_table_init<<<nb, kGridThreads, 0, st>>>(table, n_tri);
    k_grid_unpack<<<tfrt_div_up(n_tiles * TFRT_TILE_PAIRS, kGridThreads), kGridThreads, 0, st>>>(
        packed, n_tiles, n_tri, table);
    k_grid_typical<<<nb, kGridThreads, 0, st>>>(T, n_tri, h);
    k_grid_typical_finish<<<1, 32, 0, st>>>(h);
    k_grid_bounds<<<nb, kGridThreads, 0, st>>>(T, n_tri, h);
    k_grid_setup<<<1, 32, 0, st>>>(h, max_cells, 0.f);
    k_grid_bin<<<nb, kGridThreads, 0, st>>>(T, n_tri, h, size_eps, margin_ulps, packed, cell_count,
                                            cell_start, entries, L.max_entries, large, 0);
    k_grid_scan<<<1, 1024, 0, st>>>(h, cell_count, cell_start);
    k_grid_bin<<<nb, kGridThreads, 0, st>>>(T, n_tri, h, size_eps, margin_ulps, packed, cell_count,
                                            cell_start, entries, L.max_entries, large, 1);
    k_grid_sort_cells<<<tfrt_div_up(max_cells, kGridThreads), kGridThreads, 0, st>>>(
        h, cell_start, entries, L.max_entries);
    tfrt_note_launch(11);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Same contract as tfrt_launch_intersect3d: hit[p] for every live position p.
int tfrt_launch_grid_intersect3d(const float *const ray[6], const int32_t *live,
                                 const int32_t *n_live_ptr, int n_upper, const float *packed,
                                 int n_tri, const void *grid_ws, int max_cells,
                                 float intersect_eps, float ray_start_eps, float size_eps,
                                 uint64_t *hit, int dir_mode, float start_floor_ulps,
                                 cudaStream_t stream) {
    if (n_upper <= 0 || n_tri <= 0) return 0;
    TFRT_REQUIRE(grid_ws && hit && packed, "tfrt_launch_grid_intersect3d: null workspace");
    const GridLayout L = grid_layout(n_tri, max_cells);
    const char *base = reinterpret_cast<const char *>(grid_ws);
    const bool timed = tfrt_profile_begin(stream);
    k_grid_intersect<<<tfrt_div_up(n_upper, kGridThreads), kGridThreads, 0, stream>>>(
        ray[0], ray[1], ray[2], ray[3], ray[4], ray[5], live, n_live_ptr, n_upper,
        reinterpret_cast<const GridHeader *>(base + L.header),
        reinterpret_cast<const int *>(base + L.cell_start),
        reinterpret_cast<const int *>(base + L.entries),
        reinterpret_cast<const int *>(base + L.large),
        reinterpret_cast<const TriRecord *>(base + L.table), packed, intersect_eps, ray_start_eps,
        1.0f + 3.0f * size_eps, start_floor_ulps,
        reinterpret_cast<unsigned long long *>(hit), dir_mode);
    if (timed) tfrt_profile_end(stream);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Grid statistics for tests / diagnostics (host pointer out[8]): dims x,y,z, cells, entries,
// large triangles, overflow, typical extent * 1e6 (rounded).  Synchronises the stream.
extern "C" int tfrt_grid_stats(const void *grid_ws, int32_t n_tri, int32_t max_cells,
                               int64_t *out_host, void *stream) {
    TFRT_REQUIRE(grid_ws && out_host, "tfrt_grid_stats: null argument");
    const GridLayout L = grid_layout(n_tri, max_cells);
    GridHeader h;
    TFRT_CHECK_CUDA(cudaMemcpyAsync(&h, reinterpret_cast<const char *>(grid_ws) + L.header,
                                    sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    TFRT_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    out_host[0] = h.dims[0]; out_host[1] = h.dims[1]; out_host[2] = h.dims[2];
    out_host[3] = h.n_cells; out_host[4] = h.n_entries; out_host[5] = h.n_large;
    out_host[6] = h.overflow; out_host[7] = (int64_t)llround((double)h.typical * 1e6);
    return 0;
}
