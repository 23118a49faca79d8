/*
 * tfrt_b200.h -- C ABI of libtfrt_b200.so: the B200 (sm_100a) implementation of the tfrt
 * ray-trace hot path.  Plain pointers and sizes only; no torch / C++ types cross this boundary.
 *
 * The reference (ecpoppenheimer/TensorFlowRayTrace) is pure Python and has no FFI; its extension
 * seams for this path are Python call sites.  Every entry point below names the reference
 * interface it replaces (path:line under the reference tree).  INTEGRATION.md shows the ctypes
 * stub a tfrt maintainer would add at each seam.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in `_host`.
 *   - All work is enqueued on the given stream (a cudaStream_t passed as void*); no hidden
 *     synchronisation; the caller owns every buffer, including workspaces and the work counters
 *     of the persistent search kernel, so any number of calls may be in flight on any number of
 *     streams and host threads.  The library keeps no per-call state; the only memory it
 *     allocates itself is the 4-byte sink of the tfrt_fp32_peak measurement helper.
 *   - Return value 0 = success; nonzero = error, message in tfrt_last_error() (thread local).
 *   - Rays are structure-of-arrays fp32, "slot" indexed: slot i is source ray i for the whole
 *     trace (rays never split: StandardReaction emits exactly one child per active ray,
 *     tfrt/operation.py:305-307), so per-ray history is addressed by (iteration, slot).
 *   - Triangles are the merged boundary set in the reference's order optical -> stop -> target
 *     (tfrt/engine.py:1015-1018); categories OPTICAL=0, STOP=1, TARGET=2 (tfrt/engine.py:14-16).
 *   - Outputs for rays without a valid hit are unspecified, as in the reference
 *     (tfrt/engine.py:1110-1114); no entry point ever writes NaN for miss / parallel / TIR.
 */
#ifndef TFRT_B200_H
#define TFRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFRT_B200_ABI_VERSION 4

/* ray status after one projection (tfrt/engine.py:2063-2111) */
#define TFRT_STATUS_ACTIVE 0   /* hit an OPTICAL boundary: reacts and continues */
#define TFRT_STATUS_STOPPED 1  /* hit a STOP boundary */
#define TFRT_STATUS_FINISHED 2 /* hit a TARGET boundary */
#define TFRT_STATUS_DEAD 3     /* no valid intersection */

/* material kinds: refractive index n(lambda[nm]) (tfrt/materials.py:25-104) */
#define TFRT_MAT_CONSTANT 0  /* n = c[0]                                  (build_constant_material,
                                vacuum: c0=1, reflective: c0=0) */
#define TFRT_MAT_SELLMEIER 1 /* n = sqrt(1 + sum_k c[2k] x^2/(x^2 - c[2k+1])), k<3
                                (crown_glass, flint_glass, fused_silica, polycarbonate) */
#define TFRT_MAT_LAURENT 2   /* n = sqrt(c0 + c1 x^2 + c2 x^4 + c3 x^-2 + c4 x^-4 + c5 x^-6 + c6 x^-8)
                                (acrylic) */
#define TFRT_MAT_POLY 3      /* n = c0 + c1 x^2 + c2 x^-2                 (soda_lime) */

typedef struct {
    int32_t kind;
    int32_t _pad;
    double c[7];
} TfrtMaterial;

const char *tfrt_last_error(void);
int tfrt_abi_version(void);
/* hash of the sources the library was built from (set by the build; "unknown" otherwise): lets a
 * binding detect a stale library after a kernel edit that did not bump the ABI version */
const char *tfrt_source_hash(void);
/* sizeof() of the ABI structs as compiled into the library, so a binding can verify its mirror:
 * which = 0 TfrtMaterial, 1 TfrtTrace3DArgs, 2 TfrtTrace3DBwdArgs, 3 TfrtTrace2DArgs;
 * 0 for an unknown code. */
size_t tfrt_struct_size(int which);

/* ------------------------------------------------------------------------------------------
 * K3  vertices -> triangle fields.
 * Replaces TriangleBoundaryBase.update_fields_from_vertices (tfrt/boundaries.py:890-923):
 * gathers the three corners of every face and computes the unit face normal
 * normalize((p2-p1) x (p3-p2)).  faces is (F,3) int32 (no pyvista "3" prefix column).
 * tri_out[9] are the xp,yp,zp,x1,y1,z1,x2,y2,z2 field arrays (F each); norm_out is (F,3).
 */
int tfrt_faces_forward(const float *vertices, const int32_t *faces, int32_t n_faces,
                       float *const tri_out[9], float *norm_out, void *stream);

/* VJP of tfrt_faces_forward.  g_tri[9] / g_norm are cotangents of the outputs (any may be NULL);
 * update_map is the optional (F,3) uint8 vertex_update_map (0 => tf.stop_gradient on that corner,
 * tfrt/boundaries.py:900-913).  g_vertices (V,3) fp32 is OVERWRITTEN (zeroed, then accumulated). */
int tfrt_faces_backward(const float *vertices, const int32_t *faces, const uint8_t *update_map,
                        int32_t n_faces, int32_t n_vertices, const float *const g_tri[9],
                        const float *g_norm, float *g_vertices, void *stream);

/* ------------------------------------------------------------------------------------------
 * Triangle tile packing for the intersection kernel (no reference equivalent: the reference
 * re-derives edge vectors for all M x N pairs, tfrt/geometry.py:286-298).
 * tfrt_packed_floats(T) floats of workspace; size_eps is OpticalSystemBase.size_epsilion
 * (tfrt/engine.py:177-179), folded into the packed plane offsets.
 *
 * margin_ulps: the fp32 search accepts a hit when the barycentrics are inside the triangle
 * inflated by size_eps PLUS a per-triangle margin of `margin_ulps` units in the last place of the
 * triangle's largest coordinate (converted to barycentric units by the smallest altitude); the
 * winner is then re-evaluated in fp64 and must pass the reference's exact rule
 * (tfrt/engine.py:1138-1141), otherwise that ray is searched again exactly.  The margin scales
 * with the scene (millimetre units, a lens far from the origin).  0 = the fp32 band is exactly
 * size_eps (rays within fp32 rounding of an outer mesh edge may then be missed).
 */
size_t tfrt_packed_floats(int32_t n_tri);
int tfrt_pack_triangles(const float *const tri[9], int32_t n_tri, float size_eps,
                        float margin_ulps, float *packed, void *stream);

/* ------------------------------------------------------------------------------------------
 * b1  intersection seam.
 * Replaces OpticalSystem3D._intersection (tfrt/engine.py:1102-1166), i.e.
 * geometry.line_triangle_intersect (tfrt/geometry.py:191-320) + validity pruning + nearest
 * hit (argmin over triangles, lowest index on ties) + gather.
 * ray[6] = x_start,y_start,z_start,x_end,y_end,z_end (N each).  Outputs (N each): x,y,z, valid,
 * ray_u, trig_u, trig_v, gather_trig (gather_ray is arange(N) in the reference and is omitted).
 * `hit_ws` is N + 1 uint64 of workspace (the last word is the search kernel's work counter).
 * margin_ulps must equal the value given to tfrt_pack_triangles.
 * `diag` (4 int32, may be NULL) receives the re-validation counters:
 *   diag[0] = winners of the fp32 search that failed the reference's rule in fp64 and were
 *             searched again exactly (expected: a few per 100 000 rays, the rays within rounding
 *             of a triangle edge);
 *   diag[1] = of those, failures far beyond fp32 rounding -- a defect of the search, never
 *             expected (tests assert 0);
 *   diag[2] = exact re-searches that found a valid hit;  diag[3] reserved.
 */
int tfrt_intersect3d(const float *const ray[6], int32_t n_rays, const float *const tri[9],
                     const float *packed, int32_t n_tri, float intersect_eps, float size_eps,
                     float ray_start_eps, float margin_ulps, uint64_t *hit_ws, float *x, float *y,
                     float *z, uint8_t *valid, float *ray_u, float *trig_u, float *trig_v,
                     int32_t *gather_trig, int32_t *diag, const void *grid,
                     int32_t grid_max_cells, void *stream);

/* ------------------------------------------------------------------------------------------
 * Optional spatial acceleration of the search: a uniform grid over the
 * triangles, built on the device (no host round trip) by tfrt_grid_build after every
 * tfrt_pack_triangles, walked per ray with a 3-D DDA.  Every candidate is tested with the
 * brute-force kernel's own arithmetic and the winner is the same lexicographic minimum of
 * (bits(t), index), so a search through the grid returns the hit words the brute-force search
 * returns; what follows (re-validation, reaction, gradients) is shared.  Triangles much larger
 * than the typical one are kept out of the grid and tested by every ray.  Pass the workspace as
 * `grid` (with the same max_cells) to tfrt_intersect3d / TfrtTrace3DArgs; NULL = brute force, the
 * path the tests/s metric is quoted on.
 * tfrt_grid_stats (synchronises): out[8] = dims x, y, z, cells, (cell, triangle) entries, large
 * triangles, entries that did not fit (0 in a valid grid), typical triangle extent x 1e6.
 */
size_t tfrt_grid_bytes(int32_t n_tri, int32_t max_cells);
int tfrt_grid_build(const float *const tri[9], const float *packed, int32_t n_tri, float size_eps,
                    float margin_ulps, int32_t max_cells, void *grid_ws, void *stream);
int tfrt_grid_stats(const void *grid_ws, int32_t n_tri, int32_t max_cells, int64_t *out_host,
                    void *stream);

/* ------------------------------------------------------------------------------------------
 * b4  whole-loop seam: OpticalEngine.ray_trace (tfrt/engine.py:2311-2330) with
 * process_projection_3D (:1988-2191), single_pass (:2193-2302) and StandardReaction.main
 * (tfrt/operation.py:255-307) + geometry.snells_law_3D (tfrt/geometry.py:671-753) fused.
 */
typedef struct {
    int32_t n_rays;         /* N source rays */
    int32_t n_tri;          /* T merged triangles */
    int32_t n_optical;      /* the first n_optical triangles are OPTICAL */
    int32_t max_iterations; /* K */
    float intersect_eps;    /* OpticalSystemBase.intersect_epsilion  tfrt/engine.py:177 */
    float size_eps;         /* .size_epsilion (must equal the value used for `packed`) */
    float ray_start_eps;    /* .ray_start_epsilion */
    float new_ray_length;   /* OpticalEngine.new_ray_length  tfrt/engine.py:1228 */
    float margin_ulps;      /* the value given to tfrt_pack_triangles */
    int32_t index_mode;     /* 1: StandardReaction("index") mat_in/mat_out + materials table;
                               0: ("value") per-triangle n_in / n_out */
    int32_t n_materials;
    int32_t n_lambda;       /* columns of n_table */
    int32_t grid_max_cells; /* the value given to tfrt_grid_build (when `grid` is set) */

    const float *src[7];    /* x_start,y_start,z_start,x_end,y_end,z_end,wavelength; N each */
    const float *tri[9];    /* xp,yp,zp,x1,y1,z1,x2,y2,z2; T each */
    const float *norm;      /* (T,3) stored normals, or NULL: the normal of a hit triangle is then
                               normalize((p2-p1) x (p3-p2)) evaluated in fp64 from its corners
                               (what tfrt/boundaries.py:918-923 stores), and the backward pass
                               routes its cotangent straight into the corner cotangents */
    const int32_t *category; /* T */
    const int32_t *mat_in;  /* n_optical (index mode) */
    const int32_t *mat_out;
    const float *n_in;      /* n_optical (value mode) */
    const float *n_out;
    const TfrtMaterial *materials; /* n_materials, device */
    const double *n_table;  /* optional (n_materials, n_lambda): refractive indices of ARBITRARY
                               material callables (tfrt/operation.py:265-267 evaluates any
                               mat["n"](wavelength)), tabulated by the caller over the distinct
                               wavelengths of the source set; takes precedence over `materials` */
    const int32_t *lambda_index; /* N: column of n_table for every source ray */
    const float *packed;    /* tfrt_pack_triangles output */
    const void *grid;       /* optional tfrt_grid_build workspace: the search walks the grid
                               instead of testing every triangle (same hits); NULL = brute force */

    /* workspace + history ("tape"), all caller-allocated */
    float *cur[6];          /* N each: working ray state per slot: start (3) and
                               direction = end - start (3), fp32 copy read by the
                               intersection search */
    double *cur_d;          /* 6*N: the same state in fp64: the per-ray chain from one iteration
                               to the next is carried in fp64 so rounding does not compound
                               over bounces (O(N) work; the O(N*T) search is fp32) */
    uint64_t *hit;          /* N */
    float *tape_f;          /* K*9*N : per iteration Sx,Sy,Sz,Px,Py,Pz,ray_u,trig_u,trig_v */
    int32_t *tape_tri;      /* K*N : merged triangle index of the hit (-1: none) */
    uint8_t *tape_status;   /* K*N : TFRT_STATUS_*; 255 = slot not live at that iteration */
    int32_t *tape_outpos;   /* K*N : position of the slot in its category list at that iteration */
    int32_t *live;          /* (K+1)*N : live[k] = slots entering iteration k, ascending */
    int32_t *cat_ids[3];    /* N each: slots that STOPPED / FINISHED / DIED, iteration-major */
    void *sync_ws;          /* tfrt_trace3d_sync_bytes(N, K) bytes, 16-byte aligned: per-iteration
                               work counters of the search kernel, CTA tickets and look-back
                               words of the single-pass compaction (zeroed by the call) */
    int32_t *counters;      /* (K+1)*4: [k][0]=live entering k, [k][1..3]=cumulative
                               stopped/finished/dead before k */
    int32_t *diag;          /* 4: re-validation counters, summed over iterations (meaning as in
                               tfrt_intersect3d) */
} TfrtTrace3DArgs;

size_t tfrt_trace3d_sync_bytes(int32_t n_rays, int32_t max_iterations);

/* Enqueues 1 + 3 K kernels (set-up, then search + re-validation + reaction per iteration); no
 * memsets, no host round trip: the live counts stay on the device. */
int tfrt_trace3d_forward(const TfrtTrace3DArgs *args, void *stream);

/* Materialise one history set in the reference's order (concat over iterations of the
 * order-preserving masks, tfrt/engine.py:2074-2111, :1379-1399).
 * which: 0 active, 1 stopped, 2 finished, 3 dead.  `counters_host` is the (K+1)*4 counters
 * array copied to the host.  out[6] = x_start..z_end (total each), out_slot = source slot of
 * every output ray (for inherited fields, tfrt/engine.py:2242-2257), out_iter = iteration.
 * dead_ray_length < 0 => unchanged (tfrt/engine.py:2036-2046).  Returns the total through
 * *n_out (host). */
int tfrt_history_count(const TfrtTrace3DArgs *args, const int32_t *counters_host, int which,
                       int64_t *n_out);
int tfrt_materialize3d(const TfrtTrace3DArgs *args, const int32_t *counters_host, int which,
                       float dead_ray_length, float *const out[6], int32_t *out_slot,
                       int32_t *out_iter, void *stream);

/* ------------------------------------------------------------------------------------------
 * b5  gradient seam: tape.gradient(error, parameters) (tfrt/optimizer.py:216-220) restricted to
 * the trace: cotangents of the materialised history -> cotangents of triangle fields / norm /
 * source rays.  g_hist[which][6] are cotangents of the arrays produced by tfrt_materialize3d
 * (NULL where absent).  Per triangle six fp64 sums (the corner cotangents are per-triangle linear
 * maps of them) are accumulated in g_acc -- a workspace of T 128-byte rows (T*16 doubles, 128-byte
 * aligned, initialised by the call) -- then mapped to fp32 g_tri_out[9] (T each) and g_norm_out (T,3).
 * g_src[6] (N each, may be NULL) receives d/d(source start,end).
 */
typedef struct {
    const float *g_hist[4][6];
    double *g_acc;            /* T*16 doubles, 128-byte aligned */
    float *g_ray_ws;          /* unused (kept for layout compatibility; may be NULL): the running
                                 cotangent of a ray now lives in registers */
    float *g_tri_out[9];
    float *g_norm_out;
    float *g_src[6];
    float dead_ray_length;    /* value used when materialising dead rays; < 0 => unchanged */
    int32_t n_grad_tri;       /* only the first n_grad_tri merged triangles receive cotangents
                                 (<= 0: all).  The merged order is optical -> stop -> target and
                                 stop / target boundaries are rarely differentiable
                                 (tfrt/engine.py:971-1018); skipping them removes the atomics of
                                 every finished ray into the few target triangles. */
} TfrtTrace3DBwdArgs;

/* One launch walks every ray's iterations in reverse (plus a memset and the fp64 -> fp32
 * rounding of the accumulators); the iteration counts are read on the device, `counters_host` is
 * ignored (may be NULL). */
int tfrt_trace3d_backward(const TfrtTrace3DArgs *fwd, const int32_t *counters_host,
                          const TfrtTrace3DBwdArgs *bwd, void *stream);

/* ------------------------------------------------------------------------------------------
 * K4  2D path: segments + arcs.
 *
 * tfrt_intersect2d replaces OpticalSystem2D.intersect (tfrt/engine.py:523-616), i.e.
 * _segment_intersection (:687-749, geometry.line_intersect tfrt/geometry.py:27-167),
 * _arc_intersection (:767-866, geometry.line_circle_intersect tfrt/geometry.py:338-547,
 * angle_in_interval :766-802), the segment norm atan2(dy,dx)+pi/2 (:580-586), _get_arc_norm
 * (:666-670) and _seg_or_arc (:626-657).  All arithmetic is fp64 on the fp32 inputs: the 2D
 * path is launch-latency bound, so the search itself runs at the reference's
 * precision and hit indices need no near-tie exemption.
 *
 * ray[4] = x_start,y_start,x_end,y_end (N).  seg[4] = x_start,y_start,x_end,y_end (n_seg, merged
 * optical -> stop -> target); arc[5] = x_center,y_center,angle_start,angle_end,radius (n_arc).
 * Outputs (N each, any may be NULL): seg_out[5] / arc_out[5] = x, y, ray_u, segment_u|arc_u,
 * norm;  seg_valid / arc_valid (after _seg_or_arc when both kinds exist);  gather_segment /
 * gather_arc (argmin index, 0 when nothing is valid, like the reference's argmin).
 */
int tfrt_intersect2d(const float *const ray[4], int32_t n_rays, const float *const seg[4],
                     int32_t n_seg, const float *const arc[5], int32_t n_arc,
                     float intersect_eps, float size_eps, float ray_start_eps,
                     float *const seg_out[5], uint8_t *seg_valid, int32_t *gather_segment,
                     float *const arc_out[5], uint8_t *arc_valid, int32_t *gather_arc,
                     void *stream);

/* Whole-loop 2D trace: OpticalEngine.ray_trace (tfrt/engine.py:2311-2330) with
 * process_projection_2D (:1544-1986), single_pass (:2193-2302), StandardReaction.main
 * (tfrt/operation.py:255-307) and geometry.snells_law_2D (tfrt/geometry.py:565-653) fused: one
 * thread follows one source ray through all iterations (rays never split or interact).
 * The result is a per-(iteration, slot) tape; the caller orders it into the reference's history
 * sets (per iteration: segment hits, then arc hits, each order-preserving).
 */
#define TFRT_HIT2D_NONE (-1)
typedef struct {
    int32_t n_rays;
    int32_t max_iterations;
    int32_t n_seg, n_seg_optical, n_seg_stop;   /* merged order optical -> stop -> target */
    int32_t n_arc, n_arc_optical, n_arc_stop;
    float intersect_eps, size_eps, ray_start_eps, new_ray_length;
    int32_t index_mode;     /* as TfrtTrace3DArgs */
    int32_t n_materials;
    const float *src[5];    /* x_start,y_start,x_end,y_end,wavelength; N each */
    const float *seg[4];
    const float *arc[5];
    const int32_t *seg_mat_in, *seg_mat_out;   /* n_seg_optical (index mode) */
    const float *seg_n_in, *seg_n_out;         /* n_seg_optical (value mode) */
    const int32_t *arc_mat_in, *arc_mat_out;   /* n_arc_optical */
    const float *arc_n_in, *arc_n_out;
    const TfrtMaterial *materials;
    /* tape, caller-allocated */
    float *tape_f;          /* K*4*N: Sx,Sy,Px,Py (ray start, projected end) per iteration */
    double *tape_d;         /* K*2*N: ray_u, boundary parameter (segment_u | arc_u) */
    int32_t *tape_hit;      /* K*N: merged segment index s >= 0, arc index a as -2-a, or
                               TFRT_HIT2D_NONE */
    uint8_t *tape_status;   /* K*N: TFRT_STATUS_*; 255 = slot not live at that iteration */
    int32_t *live_counts;   /* K: live rays entering each iteration */
} TfrtTrace2DArgs;

int tfrt_trace2d_forward(const TfrtTrace2DArgs *args, void *stream);

/* VJP of tfrt_trace2d_forward.  g_tape is the cotangent of tape_f (K*4*N, zeros where unused).
 * g_seg[4] (n_seg each) / g_arc[5] (n_arc each; angle_start / angle_end get zeros, they only
 * enter comparisons) / g_src[4] (N each, may be NULL) are overwritten.  g_acc is
 * (4*n_seg + 3*n_arc) doubles of workspace. */
int tfrt_trace2d_backward(const TfrtTrace2DArgs *fwd, const float *g_tape, double *g_acc,
                          float *const g_seg[4], float *const g_arc[5], float *const g_src[4],
                          void *stream);

/* ------------------------------------------------------------------------------------------
 * Rays born on the device: the sampling the reference's sources repeat at every system.update()
 * (tfrt/optimizer.py:217; tfrt/distributions.py:1361-1392 squares, :1726-1850 spheres).
 * Counter-based: sample i of a distribution is Philox4x32-10 keyed by (seed, stream) at counter
 * (global index, draw, epoch), so a rank that generates indices [first, first + n) of a sharded
 * source reproduces that slice of the single-rank stream bit for bit.  `epoch` numbers the updates
 * of one distribution object, `stream` the objects.  All outputs are fp32, computed in fp64.
 *
 * tfrt_sample_uniform: out[i] = lo + (hi - lo) U_i, variable `which` of the sample.
 * tfrt_sample_square : RandomUniformSquare -- points (n,2) in [-x_size,x_size] x [-y_size,y_size],
 *                      ranks (n,2) = points / max(x_size, y_size).
 * tfrt_sample_sphere : unit vectors about +x times `radius`, points (n,3), ranks (n,2) =
 *                      (phi, theta mod 2 pi).  mode 0 RandomUniformSphere, 1 RandomLambertianSphere,
 *                      2 StaticUniformSphere (golden spiral; `total` = the distribution's
 *                      sample_count), 3 StaticLambertianSphere.
 */
int tfrt_sample_uniform(uint32_t seed, uint32_t stream, uint32_t epoch, uint32_t which,
                        int64_t first, int32_t n, double lo, double hi, float *out, void *stream_);
int tfrt_sample_square(uint32_t seed, uint32_t stream, uint32_t epoch, int64_t first, int32_t n,
                       double x_size, double y_size, float *points, float *ranks, void *stream_);
int tfrt_sample_sphere(uint32_t seed, uint32_t stream, uint32_t epoch, int32_t mode, int64_t first,
                       int32_t n, int64_t total, double angular_size, double radius, float *points,
                       float *ranks, void *stream_);

/* The ray fields of a directed 3D source with paired domains (tfrt/sources.py:590-635 PointSource,
 * :820-866 AngularSource) in one launch: dirs (n,3) unit vectors about +x, base (n,base_dim) points
 * in the y-z plane (base_dim 2) or in space (3), or NULL (all rays start at `center`).  Both are
 * rotated by the unit quaternion quat = (w,x,y,z), the base is offset by center,
 * end = start + ray_length * direction; forward = 0 swaps the two ends.  rows: (7,n) fp32 block
 * x_start, y_start, z_start, x_end, y_end, z_end, wavelength (wavelength (n,) copied, 0 if NULL) --
 * the layout tfrt_trace3d_forward reads as src[0..6]. */
int tfrt_rays_from_directions(int32_t n, const float *dirs, const float *base, int32_t base_dim,
                              const double quat[4], const double center[3], double ray_length,
                              int32_t forward, const float *wavelength, float *rows, void *stream_);

/* ------------------------------------------------------------------------------------------
 * Measurement helpers (roofline denominators): run an FFMA-only loop on every SM and return the
 * achieved fp32 FLOP/s through *flops_per_s (host).  packed=0: scalar FFMA, 1: FFMA2 (f32x2),
 * 2 / 3: FFMA2 interleaved with scalar FFMA at 1:1 / 2:1 (do the two forms share one pipe?). */
int tfrt_fp32_peak(int packed, int iters, double *flops_per_s, void *stream);

/* Number of kernel launches libtfrt_b200 has issued in this process since the last reset
 * (bench.py reports it as gpu_launches); reset != 0 zeroes the counter after reading. */
long long tfrt_launch_count(int reset);

/* Optional device timing of the dominant kernel (the ray x triangle intersection): when enabled,
 * every launch of it is bracketed by CUDA events on its own stream.  tfrt_profile_read waits for
 * the recorded events, returns their summed duration and the launch count since the last read,
 * and clears the records.  At most 4096 launches are recorded between reads; later ones run
 * untimed. */
int tfrt_profile_enable(int on);
int tfrt_profile_read(double *total_ms, long long *n_launches);

#ifdef __cplusplus
}
#endif
#endif /* TFRT_B200_H */
