// sources.cu -- rays born on the device: counter-based sampling kernels for the distributions the
// reference's sources draw from every system.update() (tfrt/optimizer.py:217 regenerates the rays
// inside the tape; tfrt/distributions.py:1361-1392 squares, :1726-1850 spheres).
//
// Random numbers are Philox4x32-10 keyed by (seed, stream) with the counter
// (index_lo, index_hi, draw, epoch): sample `i` of a distribution is a pure function of its GLOBAL
// index, so rank r of an N-rank job that generates indices [lo, hi) reproduces slice r of the
// single-rank stream bit for bit, and nothing depends on the launch shape.  `draw` separates the
// variables of one sample (x / y, cos phi / theta), `epoch` the successive updates of one object.
// oracle/sampling.py restates the generator and the transforms in NumPy.
#include "common.cuh"

namespace {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

struct Philox4 { uint32_t v[4]; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
        const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += kPhiloxW0; k1 += kPhiloxW1;
    }
    return Philox4{{c0, c1, c2, c3}};
}

// uniform in (0, 1) from 32 random bits, in fp64: (x + 0.5) 2^-32
__device__ __forceinline__ double u01(uint32_t x) { return ((double)x + 0.5) * 2.3283064365386963e-10; }

struct SampleKey {
    uint32_t seed, stream, epoch;
    long long first;      // global index of this launch's first sample
};

__device__ __forceinline__ double draw(const SampleKey &k, long long i, uint32_t which) {
    const unsigned long long g = (unsigned long long)(k.first + i);
    return u01(philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), which, k.epoch, k.seed, k.stream).v[0]);
}

__global__ void k_sample_uniform(SampleKey key, uint32_t which, int n, double lo, double hi,
                                 float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)(lo + (hi - lo) * draw(key, i, which));
}

// RandomUniformSquare (tfrt/distributions.py:1379-1392): points in [-x_size, x_size] x
// [-y_size, y_size]; rank = point / max(x_size, y_size).  points / ranks are (n, 2) row-major.
__global__ void k_sample_square(SampleKey key, int n, double x_size, double y_size,
                                float *__restrict__ points, float *__restrict__ ranks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = (2.0 * draw(key, i, 0) - 1.0) * x_size;
    const double y = (2.0 * draw(key, i, 1) - 1.0) * y_size;
    const double inv = 1.0 / fmax(x_size, y_size);
    points[2 * i] = (float)x; points[2 * i + 1] = (float)y;
    ranks[2 * i] = (float)(x * inv); ranks[2 * i + 1] = (float)(y * inv);
}

// Sphere distributions about +x (tfrt/distributions.py:1601-1850):
//   vector = radius (cos phi, sin phi cos theta, sin phi sin theta), ranks = (phi, theta mod 2 pi)
// mode 0 RandomUniformSphere     cos phi = U(cos a, 1),            theta = golden * U
// mode 1 RandomLambertianSphere  cos phi = sqrt(U(cos^2 a, 1)),    theta = golden * U
// mode 2 StaticUniformSphere     cos phi = linspace(1, cos a, total)[i],        theta = golden (i + 1/2)
// mode 3 StaticLambertianSphere  cos phi = sqrt(linspace(1, cos^2 a, total)[i]), theta = golden (i + 1/2)
// golden = pi (1 + sqrt 5); theta is reduced mod 2 pi in fp64 before the trigonometry.
__global__ void k_sample_sphere(SampleKey key, int mode, int n, long long total,
                                double angular_size, double radius, float *__restrict__ points,
                                float *__restrict__ ranks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double kTwoPi = 6.283185307179586476925, kGolden = 3.141592653589793238463 * (1.0 + 2.2360679774997896964);
    const double ca = cos(angular_size);
    double c, theta;
    if (mode < 2) {
        const double u = draw(key, i, 0), v = draw(key, i, 1);
        c = mode == 0 ? ca + (1.0 - ca) * u : sqrt(ca * ca + (1.0 - ca * ca) * u);
        theta = kGolden * v;
    } else {
        const double g = (double)(key.first + i);
        const double frac = total > 1 ? g / (double)(total - 1) : 0.0;
        c = mode == 2 ? 1.0 + (ca - 1.0) * frac : sqrt(1.0 + (ca * ca - 1.0) * frac);
        theta = kGolden * (g + 0.5);
    }
    c = fmin(fmax(c, -1.0), 1.0);
    theta = fmod(theta, kTwoPi);
    const double phi = acos(c), s = sqrt(fmax(0.0, 1.0 - c * c));
    double st, ct;
    sincos(theta, &st, &ct);
    points[3 * i] = (float)(radius * c);
    points[3 * i + 1] = (float)(radius * s * ct);
    points[3 * i + 2] = (float)(radius * s * st);
    ranks[2 * i] = (float)phi;
    ranks[2 * i + 1] = (float)theta;
}

// Directed sources (tfrt/sources.py:590-635 PointSource, :820-866 AngularSource, 3D, paired
// domains): rotate the unit directions (and the base points, given in the y-z plane or in 3D) by
// the unit quaternion q = (w, r) with v' = v + 2 r x (r x v + w v), offset the base by `center`,
// end = start + ray_length * direction; `forward` = 0 swaps the two ends.  fp64 inside, one
// rounding to fp32 on the way out.  Writes the seven rows x_start .. z_end, wavelength of the
// (7, n) block the trace kernels read.
struct RayFrame { double q[4], center[3], ray_length; };

__device__ __forceinline__ void rotate(const double q[4], const double v[3], double out[3]) {
    const double w = q[0], rx = q[1], ry = q[2], rz = q[3];
    const double tx = ry * v[2] - rz * v[1] + w * v[0];
    const double ty = rz * v[0] - rx * v[2] + w * v[1];
    const double tz = rx * v[1] - ry * v[0] + w * v[2];
    out[0] = v[0] + 2.0 * (ry * tz - rz * ty);
    out[1] = v[1] + 2.0 * (rz * tx - rx * tz);
    out[2] = v[2] + 2.0 * (rx * ty - ry * tx);
}

__global__ void k_rays_from_directions(RayFrame f, int n, const float *__restrict__ dirs,
                                       const float *__restrict__ base, int base_dim, int forward,
                                       const float *__restrict__ wavelength,
                                       float *__restrict__ rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d[3] = {dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]};
    double dir[3], start[3] = {f.center[0], f.center[1], f.center[2]};
    rotate(f.q, d, dir);
    if (base) {
        double b[3], rb[3];
        if (base_dim == 2) { b[0] = 0.0; b[1] = base[2 * i]; b[2] = base[2 * i + 1]; }
        else { b[0] = base[3 * i]; b[1] = base[3 * i + 1]; b[2] = base[3 * i + 2]; }
        rotate(f.q, b, rb);
        for (int c = 0; c < 3; ++c) start[c] += rb[c];
    }
    const size_t N = (size_t)n;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // the reference forms both ends in float32: start, then start + length * direction
        const float s = (float)start[c];
        const float e = (float)((double)s + f.ray_length * dir[c]);
        rows[(forward ? c : 3 + c) * N + i] = s;
        rows[(forward ? 3 + c : c) * N + i] = e;
    }
    rows[6 * N + i] = wavelength ? wavelength[i] : 0.f;
}

}  // namespace

static SampleKey make_key(uint32_t seed, uint32_t stream, uint32_t epoch, long long first) {
    SampleKey k;
    k.seed = seed; k.stream = stream; k.epoch = epoch; k.first = first;
    return k;
}

extern "C" int tfrt_sample_uniform(uint32_t seed, uint32_t stream, uint32_t epoch, uint32_t which,
                                   int64_t first, int32_t n, double lo, double hi, float *out,
                                   void *stream_) {
    TFRT_REQUIRE(n >= 0 && first >= 0 && (out || n == 0), "tfrt_sample_uniform: bad arguments");
    if (n == 0) return 0;
    k_sample_uniform<<<tfrt_div_up(n, 256), 256, 0, (cudaStream_t)stream_>>>(
        make_key(seed, stream, epoch, first), which, n, lo, hi, out);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfrt_sample_square(uint32_t seed, uint32_t stream, uint32_t epoch, int64_t first,
                                  int32_t n, double x_size, double y_size, float *points,
                                  float *ranks, void *stream_) {
    TFRT_REQUIRE(n >= 0 && first >= 0 && x_size > 0 && y_size > 0 && ((points && ranks) || n == 0),
                 "tfrt_sample_square: bad arguments");
    if (n == 0) return 0;
    k_sample_square<<<tfrt_div_up(n, 256), 256, 0, (cudaStream_t)stream_>>>(
        make_key(seed, stream, epoch, first), n, x_size, y_size, points, ranks);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfrt_sample_sphere(uint32_t seed, uint32_t stream, uint32_t epoch, int32_t mode,
                                  int64_t first, int32_t n, int64_t total, double angular_size,
                                  double radius, float *points, float *ranks, void *stream_) {
    TFRT_REQUIRE(n >= 0 && first >= 0 && mode >= 0 && mode <= 3 && total >= first + n &&
                     ((points && ranks) || n == 0),
                 "tfrt_sample_sphere: bad arguments");
    if (n == 0) return 0;
    k_sample_sphere<<<tfrt_div_up(n, 256), 256, 0, (cudaStream_t)stream_>>>(
        make_key(seed, stream, epoch, first), mode, n, total, angular_size, radius, points, ranks);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tfrt_rays_from_directions(int32_t n, const float *dirs, const float *base,
                                         int32_t base_dim, const double quat[4],
                                         const double center[3], double ray_length,
                                         int32_t forward, const float *wavelength, float *rows,
                                         void *stream_) {
    TFRT_REQUIRE(n >= 0 && quat && center && ((dirs && rows) || n == 0) &&
                     (!base || base_dim == 2 || base_dim == 3),
                 "tfrt_rays_from_directions: bad arguments");
    if (n == 0) return 0;
    RayFrame f;
    for (int c = 0; c < 4; ++c) f.q[c] = quat[c];
    for (int c = 0; c < 3; ++c) f.center[c] = center[c];
    f.ray_length = ray_length;
    k_rays_from_directions<<<tfrt_div_up(n, 256), 256, 0, (cudaStream_t)stream_>>>(
        f, n, dirs, base, base_dim, forward, wavelength, rows);
    tfrt_note_launch(1);
    TFRT_CHECK_CUDA(cudaGetLastError());
    return 0;
}
