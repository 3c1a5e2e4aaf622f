// rasterize.cu -- range-image rasterization of raw sweeps (K1 scatter + K1b resolve).
//
// Replaces (paths relative to /root/reference):
//   math/range_view.py:14-44            build_range_view
//   math/numpy/conversions.py:46-73     cart_to_sph
//   math/numpy/conversions.py:9-43      build_range_view_coordinates (library column formula)
//   converters/av2/utils.py:108-153     build_range_view_coordinates (converter column formula)
//   math/numpy/conversions.py:106-128   z_buffer (numba serial loop)
//
// Design: the reference's z-buffer is a serial, order-dependent loop with a float32 depth
// buffer and float64 distances.  Its fixed point has a closed form (DESIGN.md "K1"):
//   m      = min_i f32(d_i)                      (f32 = round-to-nearest)
//   class0 = { i : f32(d_i) == m and d_i <  m }  (rounded up: always re-trigger the `<` test)
//   class1 = { i : f32(d_i) == m and d_i >= m }
//   winner = max index of class0 if class0 is non-empty, else min index of class1.
// That is the minimum of one packed 64-bit key per point, so a single atomicMin per point
// gives a deterministic, order-independent, bit-exact winner:
//   key = f32bits(m) << 32 | class << 31 | (class ? i : ~i & 0x7fffffff)
// K1  (scatter): 1 thread / point, float4 load, fp64 index math, atomicMin.u64 (REDG).
// K1b (resolve): 1 thread / pixel, reads the key, gathers the winning point, recomputes its
//                spherical coordinates and writes the 7 channel planes coalesced.
#include <cmath>
#include <cstdlib>
#include <math_constants.h>

#include "common.cuh"
#include "fastmath.cuh"

namespace rv3d {

constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;

struct RasterArgs {
  int32_t B, max_points, H, W, az_bins, num_lasers, col_mode, fast_col;
  double ox, oy, oz, min_distance, bin_scale;  // bin_scale = az_bins / tau (computed on the host in double)
  // float32 side of the azimuth-bin fast path (make_args)
  float ox_hi, ox_lo, oy_hi, oy_lo;   // lidar offset as hi + lo float32 pairs
  float min_xy, scale_hi, scale_lo;   // smallest |cx| + |cy| the float32 offsets resolve; bin_scale as hi + lo
  float frac0, sgn, col0, half_minus_band, col_max;
};

// azimuth -> column, both formulas, in fp64 with round-half-even (np.round == rint)
__device__ __forceinline__ double column_of(double az, const RasterArgs &a) {
  double t = az + CUDART_PI;   // azimuth += math.pi
  t = t * a.bin_scale;         // azimuth *= n_azimuth_bins / math.tau
  const double nb = static_cast<double>(a.az_bins);
  double c = (a.col_mode == RV3D_COL_LIBRARY) ? rint((nb - t) - 1.0) : (nb - rint(t));
  return fmin(fmax(c, 0.0), nb - 1.0);  // np.clip(col, 0, W-1)
}

// The exact column of a point whose float32 estimate was too close to a bin boundary (~0.2 % of the points): fp64
// table atan2 (fastmath.cuh, <= 2 ulp: 1e-12 bins) unless THAT lands within 1e-7 bins of a boundary; then libm.
static __device__ __noinline__ int column_exact(double cx, double cy, RasterArgs a) {
  const double nb = static_cast<double>(a.az_bins);
  const double t = (fast_atan2(cy, cx) + CUDART_PI) * a.bin_scale;
  const double pre = (a.col_mode == RV3D_COL_LIBRARY) ? ((nb - t) - 1.0) : t;   // value handed to rint()
  if (fabs((pre - floor(pre)) - 0.5) > 1.0e-7) {
    const double r = rint(pre);
    const double c = (a.col_mode == RV3D_COL_LIBRARY) ? r : (nb - r);
    return static_cast<int>(fmin(fmax(c, 0.0), nb - 1.0));
  }
  return static_cast<int>(column_of(atan2(cy, cx), a));
}

// Column of a point in float32, with a proof obligation instead of an fp64 atan2.  With v = az * bin_scale (bins from
// the -x axis ... ) both reference formulas are `integer - k + rint(z)`-shaped once k = rint(v) is split off:
//   library    rint((nb - (az + pi) s) - 1) = (nb/2 - 1) - v   = B_i - k + rint(B_f - res)
//   converter  nb - rint((az + pi) s)       = nb - (nb/2 + v)  = (nb - H_i) - k - rint(H_f + res)
// (res = v - k in [-0.5, 0.5]; B_i + B_f = nb/2 - 1, H_i + H_f = nb/2, fractional parts 0 or 0.5; pi * bin_scale = nb/2
// to 1e-13 bins).  az comes from atan2f_lite on float32 copies of the offsets (|error| < 5e-7 rad in total, measured:
// tests/test_gpu_fastmath.py), res from one FMA against the hi + lo split of bin_scale (no cancellation error), so the
// float32 z is within `band` = 2e-6 rad * bin_scale + 2e-6 bins of the fp64 value: when z is farther than that from a
// rounding boundary both paths round to the same integer.  Returns false when it cannot promise that.
__device__ __forceinline__ bool column_fast(const RasterArgs &a, float x32, float y32, int &col) {
  const float mag = fabsf(x32) + fabsf(y32);
  const float az = atan2f_lite(y32, x32);
  const float kMagic = 12582912.0f;                      // 1.5 * 2^23: (v + kMagic) - kMagic == rintf(v) for |v| < 2^22
  const float kf = (az * a.scale_hi + kMagic) - kMagic;
  float res = fmaf(az, a.scale_hi, -kf);
  res = fmaf(az, a.scale_lo, res);
  const float z = fmaf(-a.sgn, res, a.frac0);            // in [-0.5, 1.0]
  const float rz = z > 0.5f ? 1.0f : 0.0f;
  float c = (a.col0 - kf) + a.sgn * rz;
  c = fminf(fmaxf(c, 0.0f), a.col_max);                  // np.clip(col, 0, n_azimuth_bins - 1)
  col = __float_as_int(c + kMagic) - 0x4B400000;
  return fabsf(z - rz) < a.half_minus_band && mag >= a.min_xy && mag < 1.0e30f;   // NaN fails every test
}

// Radius of a point, bit-compatible with the reference where it matters.  The reference computes
// r = hypot(hypot(x, y), z) with numpy / libm (numpy/conversions.py:61-62); the z-buffer then only looks at
// float32(r), at the comparison r < float32(r) (the class bit of the key) and at r < min_distance.  The fast form
// sqrt(x^2 + y^2 + z^2) (fastmath.cuh fast_sqrt) is within a few fp64 ulps of libm's value, so those three outcomes
// can only differ when r lies within a few ulps of a float32 value, of a midpoint between two float32 values, or of
// min_distance.  Exactly then (about 1e-6 of the points) the radius is recomputed with the restatement of libm's hypot
// (fastmath.cuh libm_hypot): pixel assignment and the range channel are bit-exact by construction.
// The float32 neighbourhood test and the rounding itself are integer operations on the fp64 bit pattern: the low 29
// mantissa bits of r are its position between two float32 values (0 = exact, 2^28 = midpoint).
constexpr uint32_t kUlpBand = 64;   // fp64 ulps; fast_sqrt of the fused sum of squares vs hypot(hypot()): <= ~6

__device__ __forceinline__ unsigned long long pack_key_bits(uint32_t f32_bits, uint32_t cls, uint32_t i) {
  const uint32_t low = cls ? (0x80000000u | i) : (~i & 0x7fffffffu);
  return (static_cast<unsigned long long>(f32_bits) << 32) | low;
}

__device__ __forceinline__ unsigned long long pack_key(double d, uint32_t i) {
  const float m = __double2float_rn(d);
  return pack_key_bits(__float_as_uint(m), (d < static_cast<double>(m)) ? 0u : 1u, i);
}
__device__ __forceinline__ uint32_t key_index(unsigned long long key) {
  const uint32_t low = static_cast<uint32_t>(key);
  return (low & 0x80000000u) ? (low & 0x7fffffffu) : (~low & 0x7fffffffu);
}

// the reference's own arithmetic for the few points the fast forms cannot decide; returns kEmptyKey for "no write"
static __device__ __noinline__ unsigned long long key_exact(double cx, double cy, double cz, double min_distance, uint32_t i) {
  const double r = libm_hypot(libm_hypot(cx, cy), cz);
  // z_buffer: `d < min_distance -> continue`, then `d < buffer` with buffer starting at +inf; NaN and +inf never write
  if (!(r >= min_distance) || !(r < CUDART_INF)) return kEmptyKey;
  return pack_key(r, i);
}

// Four points per thread.  (1) Every load is issued before the first use.  (2) The common path of a point is
// BRANCH-FREE (eval_point: ~100 instructions ending in a state flag), so the compiler interleaves the four independent
// dependency chains -- with one point at a time a warp waited ~10 cycles between instructions on fp64 latencies
// (ncu r02c: stall_wait 2.8 + long_scoreboard 3.1 per issue at 43 % occupancy).  (3) The rare exits -- radius within 64
// fp64 ulps of a float32 decision point, azimuth within 2e-6 rad of a bin boundary (~0.2 % of the points), a row outside
// the image -- are flagged and handled after the straight-line part by out-of-line routines that restate the
// reference's own arithmetic.
constexpr int kScatterThreads = 128;
constexpr int kScatterPerThread = 4;
enum : uint32_t { kPtDrop = 0u, kPtReady = 1u, kPtSlowKey = 2u, kPtSlowCol = 4u, kPtSlowRow = 8u };

struct PointEval {
  unsigned long long key;
  uint32_t pix, state;
  int row;
};

__device__ __forceinline__ PointEval eval_point(const RasterArgs &a, int i, int l, float4 p, const int32_t *s_map) {
  PointEval e;
  const double cx = static_cast<double>(p.x) - a.ox;  // range_view.py:29
  const double cy = static_cast<double>(p.y) - a.oy;
  const double cz = static_cast<double>(p.z) - a.oz;
  // ---- radius -> (float32 bits, class bit)
  const double s = fma(cx, cx, fma(cy, cy, cz * cz));
  const double r = fast_sqrt(s);                                                       // garbage unless fast_sqrt_ok(s)
  const double diff = r - a.min_distance;
  const uint32_t lo = static_cast<uint32_t>(__double2loint(r));
  const uint32_t lo29 = lo & 0x1FFFFFFFu;
  const bool near_f32 = ((lo29 + kUlpBand) & 0x0FFFFFFFu) <= 2u * kUlpBand;            // exact value or midpoint
  const bool near_min = fabs(diff) <= r * (static_cast<double>(kUlpBand) * 0x1p-52);
  const bool key_ok = fast_sqrt_ok(s) && !near_f32 && !near_min;
  const uint32_t up = lo29 > 0x10000000u ? 1u : 0u;                                    // round to nearest: up => r < float32(r)
  const uint32_t bits = (((static_cast<uint32_t>(__double2hiint(r)) - 0x38000000u) << 3) | (lo >> 29)) + up;
  e.key = pack_key_bits(bits, up ^ 1u, static_cast<uint32_t>(i));
  // ---- column
  int col;
  const float x32 = (p.x - a.ox_hi) - a.ox_lo, y32 = (p.y - a.oy_hi) - a.oy_lo;
  const bool col_ok = column_fast(a, x32, y32, col) && a.fast_col;
  // ---- row, raveled index row * W + col (col may exceed W when azimuth_bins > W: the reference ravels the same way)
  e.row = a.H - s_map[l & 255] - 1;  // conversions.py:37
  const bool row_ok = static_cast<uint32_t>(e.row) < static_cast<uint32_t>(a.H);
  e.pix = static_cast<uint32_t>(e.row) * a.W + col;                                    // H * W < 2^31 (checked on the host)
  // z_buffer: `d < min_distance -> continue`
  const bool drop = (l >= a.num_lasers) || (key_ok && diff < 0.0) || (key_ok && col_ok && row_ok && e.pix >= static_cast<uint32_t>(a.H * a.W));
  e.state = drop ? kPtDrop : ((key_ok && col_ok && row_ok) ? kPtReady : ((key_ok ? 0u : kPtSlowKey) | (col_ok ? 0u : kPtSlowCol) | (row_ok ? 0u : kPtSlowRow)));
  return e;
}

// the flagged points: the reference's arithmetic where the fast form declined
static __device__ __noinline__ void finish_point_slow(RasterArgs a, int i, float4 p, PointEval e, unsigned long long *keys) {
  const double cx = static_cast<double>(p.x) - a.ox, cy = static_cast<double>(p.y) - a.oy, cz = static_cast<double>(p.z) - a.oz;
  unsigned long long key = e.key;
  if (e.state & kPtSlowKey) {
    key = key_exact(cx, cy, cz, a.min_distance, static_cast<uint32_t>(i));
    if (key == kEmptyKey) return;
  }
  int col;
  const float x32 = (p.x - a.ox_hi) - a.ox_lo, y32 = (p.y - a.oy_hi) - a.oy_lo;
  if ((e.state & kPtSlowCol) || !a.fast_col || !column_fast(a, x32, y32, col)) col = column_exact(cx, cy, a);
  const long long p64 = static_cast<long long>(e.row) * a.W + col;
  if (p64 < 0 || p64 >= static_cast<long long>(a.H) * a.W) return;
  atomicMin(keys + p64, key);
}

__global__ void __launch_bounds__(kScatterThreads, 7)
raster_scatter_kernel(RasterArgs a, const float4 *__restrict__ points, const uint8_t *__restrict__ laser,
                      const int32_t *__restrict__ n_points, const int32_t *__restrict__ laser_mapping,
                      unsigned long long *__restrict__ keys) {
  __shared__ int32_t s_map[256];
  for (int t = threadIdx.x; t < 256; t += kScatterThreads) s_map[t] = t < a.num_lasers ? laser_mapping[t] : 0;
  const int b = blockIdx.y;
  const int n = n_points[b];
  const int i0 = blockIdx.x * (kScatterThreads * kScatterPerThread) + threadIdx.x;
  const float4 *pts = points + static_cast<size_t>(b) * a.max_points;
  const uint8_t *las = laser + static_cast<size_t>(b) * a.max_points;
  int l[kScatterPerThread];
  float4 p[kScatterPerThread];
  const uint64_t keep = l2_policy_evict_last();   // resolve gathers the winners from these lines: keep them in L2
#pragma unroll
  for (int u = 0; u < kScatterPerThread; ++u) {
    const int i = i0 + u * kScatterThreads;
    const int gi = i < n ? i : 0;
    l[u] = las[gi];
    p[u] = ldg_f4_hint(pts + gi, keep);
  }
  __syncthreads();
  unsigned long long *k = keys + static_cast<size_t>(b) * a.H * a.W;
  PointEval e[kScatterPerThread];
#pragma unroll
  for (int u = 0; u < kScatterPerThread; ++u) {
    e[u] = eval_point(a, i0 + u * kScatterThreads, l[u], p[u], s_map);
    if (i0 + u * kScatterThreads >= n) e[u].state = kPtDrop;
  }
#pragma unroll
  for (int u = 0; u < kScatterPerThread; ++u) {
    if (e[u].state == kPtReady) atomicMin(k + e[u].pix, e[u].key);
    else if (e[u].state != kPtDrop) finish_point_slow(a, i0 + u * kScatterThreads, p[u], e[u], k);
  }
}

// K1b: key -> winning point (a random 16-byte gather; synthetic sweeps are in random point order, the worst case) ->
// azimuth / inclination in fp64 -> 7 coalesced plane stores.  The range channel is NOT recomputed: the key's high word
// is float32(r) of the winner, bit for bit (scatter_point).  The kernel is bound by the latency of its two dependent
// loads (ncu: 53 % of the stall samples on the gather and its first use at one pixel per thread), so a thread owns
// four pixels, reads the four keys, then issues the four gathers unconditionally (an empty pixel gathers point 0 and
// ignores it) before any arithmetic.
constexpr int kResolvePx = 4;
constexpr int kResolveThreads = 128;
struct PixelOut { float az, inc, rr, x, y, z, it; int32_t w; };

__device__ __forceinline__ PixelOut resolve_pixel(const RasterArgs &a, unsigned long long key, float4 p) {
  PixelOut o{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, -1};
  if (key != kEmptyKey) {
    o.w = static_cast<int32_t>(key_index(key));
    const double cx = static_cast<double>(p.x) - a.ox;
    const double cy = static_cast<double>(p.y) - a.oy;
    const double cz = static_cast<double>(p.z) - a.oz;
    const double s2 = fma(cx, cx, cy * cy);
    const double hxy = fast_sqrt_ok(s2) ? fast_sqrt(s2) : sqrt(s2);   // planar radius: inclination channel only (1-ulp bar)
    // features are snapshotted BEFORE the in-place azimuth rescale (range_view.py:33, H3)
    o.az = static_cast<float>(fast_atan2(cy, cx));    // fastmath.cuh: <= 2 ulp of fp64 before the cast
    o.inc = static_cast<float>(fast_atan2(cz, hxy));
    o.rr = __uint_as_float(static_cast<uint32_t>(key >> 32));   // == float32(hypot(hypot(x, y), z)), see scatter_point
    o.x = p.x; o.y = p.y; o.z = p.z; o.it = p.w;
  }
  return o;
}

__global__ void __launch_bounds__(kResolveThreads, 8)
raster_resolve_kernel(RasterArgs a, const float4 *__restrict__ points, const unsigned long long *__restrict__ keys,
                      float *__restrict__ image, int32_t *__restrict__ winner) {
  const int b = blockIdx.y;
  const int HW = a.H * a.W;
  const int pix0 = blockIdx.x * (kResolveThreads * kResolvePx) + threadIdx.x;
  const unsigned long long *kb = keys + static_cast<size_t>(b) * HW;
  const float4 *pts = points + static_cast<size_t>(b) * a.max_points;
  unsigned long long key[kResolvePx];
  float4 p[kResolvePx];
#pragma unroll
  for (int j = 0; j < kResolvePx; ++j) {
    const int pix = pix0 + j * kResolveThreads;
    key[j] = pix < HW ? __ldg(kb + pix) : kEmptyKey;
  }
#pragma unroll
  for (int j = 0; j < kResolvePx; ++j)
    p[j] = ldg_stream_f4(pts + (key[j] != kEmptyKey ? key_index(key[j]) : 0u));
  const uint64_t stream = l2_policy_evict_first();   // the 76 MB image must not push the points out of L2
  float *out = image + static_cast<size_t>(b) * 7 * HW;
#pragma unroll
  for (int j = 0; j < kResolvePx; ++j) {
    const int pix = pix0 + j * kResolveThreads;
    if (pix >= HW) break;
    const PixelOut o = resolve_pixel(a, key[j], p[j]);
    float *q = out + pix;
    st_f_hint(q, o.az, stream);
    st_f_hint(q + HW, o.inc, stream);
    st_f_hint(q + 2 * HW, o.rr, stream);
    st_f_hint(q + 3 * HW, o.x, stream);
    st_f_hint(q + 4 * HW, o.y, stream);
    st_f_hint(q + 5 * HW, o.z, stream);
    st_f_hint(q + 6 * HW, o.it, stream);
    if (winner) winner[static_cast<size_t>(b) * HW + pix] = o.w;
  }
}

// ---------------------------------------------------------------------------------------------
// generic z_buffer(indices, distances, features, H, W, min_distance)
// ---------------------------------------------------------------------------------------------
template <typename D>
__global__ void __launch_bounds__(256)
zbuffer_scatter_kernel(const int64_t *__restrict__ rows, const int64_t *__restrict__ cols,
                       const D *__restrict__ dist, int64_t n, int H, int W, double min_distance,
                       unsigned long long *__restrict__ keys) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = static_cast<double>(dist[i]);
  if (!(d >= min_distance) || !(d < CUDART_INF)) return;
  const int64_t pix = rows[i] * W + cols[i];
  if (pix < 0 || pix >= static_cast<int64_t>(H) * W) return;
  atomicMin(keys + pix, pack_key(d, static_cast<uint32_t>(i)));
}

template <typename F>
__global__ void __launch_bounds__(256)
zbuffer_resolve_kernel(const F *__restrict__ feat, int C, int64_t n, int HW,
                       const unsigned long long *__restrict__ keys, float *__restrict__ image,
                       int32_t *__restrict__ winner) {
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HW) return;
  const unsigned long long key = keys[pix];
  const int32_t w = (key == kEmptyKey) ? -1 : static_cast<int32_t>(key_index(key));
  for (int c = 0; c < C; ++c)
    image[static_cast<size_t>(c) * HW + pix] = (w < 0) ? 0.f : static_cast<float>(feat[static_cast<size_t>(c) * n + w]);
  if (winner) winner[pix] = w;
}

// ---------------------------------------------------------------------------------------------
// cart_to_sph / build_range_view_coordinates as free-standing operators
// ---------------------------------------------------------------------------------------------
__global__ void cart_to_sph_kernel(const double *__restrict__ cart, double *__restrict__ sph, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = cart[3 * i], y = cart[3 * i + 1], z = cart[3 * i + 2];
  const double hxy = libm_hypot(x, y);   // bit-compatible with numpy.hypot (fastmath.cuh)
  sph[3 * i] = atan2(y, x);
  sph[3 * i + 1] = atan2(z, hxy);
  sph[3 * i + 2] = libm_hypot(hxy, z);
}

__global__ void rv_coordinates_kernel(double *__restrict__ sph, const int64_t *__restrict__ laser,
                                      const int64_t *__restrict__ mapping, int n_mapping, int64_t n,
                                      int n_inc, RasterArgs a, double *__restrict__ hybrid) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t = sph[3 * i] + CUDART_PI;
  t = t * a.bin_scale;
  sph[3 * i] = t;  // the reference mutates its input (conversions.py:33-34)
  const double nb = static_cast<double>(a.az_bins);
  double c = (a.col_mode == RV3D_COL_LIBRARY) ? rint((nb - t) - 1.0) : (nb - rint(t));
  c = fmin(fmax(c, 0.0), nb - 1.0);
  double row;
  if (a.col_mode == RV3D_COL_CONVERTER_UNIFORM) {
    // converters/av2/utils.py:138-145: rows uniform in inclination over a +-10 degree field of view
    const double fov = (10.0 / 180.0) * CUDART_PI;   // |(-10.0 / 180.0) * pi| == (10 / 180.0) * pi
    double r = 1.0 - (sph[3 * i + 1] + fov) / (fov + fov);
    r = r * static_cast<double>(n_inc);
    row = fmin(fmax(rint(r), 0.0), static_cast<double>(n_inc - 1));
  } else {
    int64_t l = laser[i];
    if (l < 0) l += n_mapping;  // numpy negative indexing
    row = (l >= 0 && l < n_mapping) ? static_cast<double>(n_inc - mapping[l] - 1) : CUDART_NAN;
  }
  hybrid[3 * i] = row;
  hybrid[3 * i + 1] = c;
  hybrid[3 * i + 2] = sph[3 * i + 2];
}

// ---------------------------------------------------------------------------------------------
// loader-side post-processing: feature / cart / mask assembly + subsample_range_view, one pass
// ---------------------------------------------------------------------------------------------
struct InputsArgs {
  int B, H, W, Wo, stride, pad, mode, F, tanh_ch;
  int ch[8];
};

__global__ void __launch_bounds__(256)
range_view_inputs_kernel(InputsArgs a, const float *__restrict__ image, float *__restrict__ features,
                         float *__restrict__ cart, uint8_t *__restrict__ mask) {
  const int b = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;   // output pixel h * Wo + wo
  if (o >= a.H * a.Wo) return;
  const int h = o / a.Wo, wo = o - h * a.Wo;
  int w = wo * a.stride - a.pad;                          // column in the unpadded image
  bool inside = true;
  if (w < 0 || w >= a.W) {
    if (a.mode == RV3D_PAD_CIRCULAR) { w %= a.W; if (w < 0) w += a.W; }
    else inside = false;                                  // constant (zero) padding
  }
  const int HW = a.H * a.W, HWo = a.H * a.Wo;
  const float *src = image + static_cast<size_t>(b) * 7 * HW + h * a.W + w;
  const bool valid = inside && src[2 * HW] > 0.0f;        // mask = range > 0 (loader.py:645-650)
  for (int f = 0; f < a.F; ++f) {
    float v = 0.f;
    if (inside) {
      v = src[static_cast<size_t>(a.ch[f]) * HW];
      if (a.ch[f] == a.tanh_ch) v = tanhf(v);             // Waymo: intensity.tanh() (loader.py:625-626)
      v = valid ? v : v * 0.0f;                           // range_view *= mask (loader.py:808)
    }
    features[(static_cast<size_t>(b) * a.F + f) * HWo + o] = v;
  }
  for (int k = 0; k < 3; ++k)
    cart[(static_cast<size_t>(b) * 3 + k) * HWo + o] = inside ? src[static_cast<size_t>(3 + k) * HW] : 0.f;
  mask[static_cast<size_t>(b) * HWo + o] = valid ? 1 : 0;
}

__global__ void __launch_bounds__(256)
subsample_range_view_kernel(const float *__restrict__ rv, const uint8_t *__restrict__ mask,
                            const float *__restrict__ cart, int C, int H, int W, int Wo, int stride, int pad,
                            int mode, float *__restrict__ o_rv, uint8_t *__restrict__ o_mask,
                            float *__restrict__ o_cart) {
  const int b = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= H * Wo) return;
  const int h = o / Wo, wo = o - h * Wo;
  int w = wo * stride - pad;
  bool inside = true;
  if (w < 0 || w >= W) {
    if (mode == RV3D_PAD_CIRCULAR) { w %= W; if (w < 0) w += W; }
    else inside = false;
  }
  const int HW = H * W, HWo = H * Wo, p = h * W + w;
  const uint8_t m = inside ? mask[static_cast<size_t>(b) * HW + p] : 0;
  for (int c = 0; c < C; ++c) {
    float v = 0.f;
    if (inside) { v = rv[(static_cast<size_t>(b) * C + c) * HW + p]; v = m ? v : v * 0.0f; }   // range_view *= mask
    o_rv[(static_cast<size_t>(b) * C + c) * HWo + o] = v;
  }
  for (int k = 0; k < 3; ++k)
    o_cart[(static_cast<size_t>(b) * 3 + k) * HWo + o] = inside ? cart[(static_cast<size_t>(b) * 3 + k) * HW + p] : 0.f;
  o_mask[static_cast<size_t>(b) * HWo + o] = m ? 1 : 0;
}


// ---------------------------------------------------------------------------------------------
// rasterize straight into the network's inputs: K1b with the loader's assembly folded in.  One
// thread per OUTPUT pixel (after padding / striding): key -> winning point -> only the channels the feature list
// asks for (the two fp64 arctangents are skipped when neither azimuth nor inclination is a feature, the usual
// case), tanh / mask / zero padding applied on the way out.  The 7-plane image is never written or re-read.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kResolveThreads, 8)
raster_resolve_inputs_kernel(RasterArgs a, InputsArgs ia, int need_angles, const float4 *__restrict__ points,
                             const unsigned long long *__restrict__ keys, float *__restrict__ features,
                             float *__restrict__ cart, uint8_t *__restrict__ mask) {
  const int b = blockIdx.y;
  const int HW = a.H * a.W, HWo = ia.H * ia.Wo;
  const int o0 = blockIdx.x * (kResolveThreads * kResolvePx) + threadIdx.x;
  const unsigned long long *kb = keys + static_cast<size_t>(b) * HW;
  const float4 *pts = points + static_cast<size_t>(b) * a.max_points;
  unsigned long long key[kResolvePx];
  float4 p[kResolvePx];
  bool inside[kResolvePx];
#pragma unroll
  for (int j = 0; j < kResolvePx; ++j) {
    const int o = o0 + j * kResolveThreads;
    key[j] = kEmptyKey;
    inside[j] = false;
    if (o < HWo) {
      const int h = o / ia.Wo, wo = o - h * ia.Wo;
      int w = wo * ia.stride - ia.pad;                      // column in the unpadded image
      inside[j] = true;
      if (w < 0 || w >= ia.W) {
        if (ia.mode == RV3D_PAD_CIRCULAR) { w %= ia.W; if (w < 0) w += ia.W; }
        else inside[j] = false;                             // constant (zero) padding
      }
      if (inside[j]) key[j] = __ldg(kb + h * ia.W + w);
    }
  }
#pragma unroll
  for (int j = 0; j < kResolvePx; ++j)
    p[j] = ldg_stream_f4(pts + (key[j] != kEmptyKey ? key_index(key[j]) : 0u));
#pragma unroll
  for (int j = 0; j < kResolvePx; ++j) {
    const int o = o0 + j * kResolveThreads;
    if (o >= HWo) break;
    float ch[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (key[j] != kEmptyKey) {
      if (need_angles) {
        const PixelOut px = resolve_pixel(a, key[j], p[j]);
        ch[0] = px.az; ch[1] = px.inc;
      }
      ch[2] = __uint_as_float(static_cast<uint32_t>(key[j] >> 32));
      ch[3] = p[j].x; ch[4] = p[j].y; ch[5] = p[j].z; ch[6] = p[j].w;
    }
    const bool valid = inside[j] && ch[2] > 0.0f;           // mask = range > 0 (loader.py:645-650)
    for (int f = 0; f < ia.F; ++f) {
      float v = 0.f;
      if (inside[j]) {
        const int c = ia.ch[f];
        v = c == 0 ? ch[0] : c == 1 ? ch[1] : c == 2 ? ch[2] : c == 3 ? ch[3] : c == 4 ? ch[4] : c == 5 ? ch[5] : ch[6];
        if (c == ia.tanh_ch) v = tanhf(v);                  // Waymo: intensity.tanh() (loader.py:625-626)
        v = valid ? v : v * 0.0f;                           // range_view *= mask (loader.py:808)
      }
      features[(static_cast<size_t>(b) * ia.F + f) * HWo + o] = v;
    }
    cart[(static_cast<size_t>(b) * 3 + 0) * HWo + o] = inside[j] ? ch[3] : 0.f;
    cart[(static_cast<size_t>(b) * 3 + 1) * HWo + o] = inside[j] ? ch[4] : 0.f;
    cart[(static_cast<size_t>(b) * 3 + 2) * HWo + o] = inside[j] ? ch[5] : 0.f;
    mask[static_cast<size_t>(b) * HWo + o] = valid ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// test hooks: the fast math routines and the rasterizer's float32 column decision, exposed so the GPU suite can
// measure their error / check the proof obligation directly (tests/test_gpu_fastmath.py)
// ---------------------------------------------------------------------------------------------
__global__ void debug_fastmath_kernel(int op, const double *__restrict__ a, const double *__restrict__ b,
                                      double *__restrict__ out, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = a[i], y = b ? b[i] : 0.0;
  double r;
  switch (op) {
    case 0: r = fast_atan2(x, y); break;
    case 1: r = fast_exp(x); break;
    case 2: r = fast_sqrt_ok(x) ? fast_sqrt(x) : sqrt(x); break;
    default: r = static_cast<double>(atan2f_lite(static_cast<float>(x), static_cast<float>(y))); break;
  }
  out[i] = r;
}

__global__ void debug_column_kernel(RasterArgs a, const float4 *__restrict__ points, int64_t n,
                                    int32_t *__restrict__ col_fast, int32_t *__restrict__ col_exact) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = points[i];
  const double cx = static_cast<double>(p.x) - a.ox, cy = static_cast<double>(p.y) - a.oy;
  const float x32 = (p.x - a.ox_hi) - a.ox_lo, y32 = (p.y - a.oy_hi) - a.oy_lo;
  int col = -1;
  const bool ok = a.fast_col && column_fast(a, x32, y32, col);
  col_fast[i] = ok ? col : -1 - column_exact(cx, cy, a);      // undecided: -1 - (the two-level fallback's answer)
  col_exact[i] = static_cast<int>(column_of(atan2(cy, cx), a));   // the reference's arithmetic, libm atan2
}

static void split_f32(double v, float &hi, float &lo) {
  hi = static_cast<float>(v);
  lo = static_cast<float>(v - static_cast<double>(hi));
}

static RasterArgs make_args(const rv3d_raster_params *p) {
  RasterArgs a{};
  a.B = p->batch; a.max_points = p->max_points; a.H = p->height; a.W = p->width;
  a.az_bins = p->azimuth_bins; a.num_lasers = p->num_lasers; a.col_mode = p->col_mode;
  a.ox = p->lidar_offset[0]; a.oy = p->lidar_offset[1]; a.oz = p->lidar_offset[2];
  a.min_distance = p->min_distance;
  a.bin_scale = static_cast<double>(p->azimuth_bins) / 6.283185307179586;  // n_azimuth_bins / math.tau
  // float32 column path (column_fast): constants of `col0 - k + sgn * rint(frac0 - sgn * res)`
  const double nb = static_cast<double>(p->azimuth_bins);
  const double base = (p->col_mode == RV3D_COL_LIBRARY) ? nb / 2.0 - 1.0 : nb / 2.0;   // B = nb/2 - 1 | H = nb/2
  const double base_i = std::floor(base);
  split_f32(a.ox, a.ox_hi, a.ox_lo);
  split_f32(a.oy, a.oy_hi, a.oy_lo);
  split_f32(a.bin_scale, a.scale_hi, a.scale_lo);
  a.frac0 = static_cast<float>(base - base_i);                                          // 0 or 0.5
  a.sgn = (p->col_mode == RV3D_COL_LIBRARY) ? 1.0f : -1.0f;
  a.col0 = static_cast<float>((p->col_mode == RV3D_COL_LIBRARY) ? base_i : nb - base_i);
  a.col_max = static_cast<float>(nb - 1.0);
  const double band = 2.0e-6 * a.bin_scale + 2.0e-6;                                    // bins
  a.half_minus_band = static_cast<float>(0.5 - band);
  // below this |cx| + |cy| the residual of the float32 hi + lo offsets (2^-48 relative) is no longer negligible
  a.min_xy = static_cast<float>(std::fmax(1.0e-30, (std::fabs(a.ox) + std::fabs(a.oy)) * 0x1p-18));
  const bool finite_off = std::isfinite(a.ox) && std::isfinite(a.oy) && std::fabs(a.ox) < 1e30 && std::fabs(a.oy) < 1e30;
  a.fast_col = (p->azimuth_bins <= (1 << 20) && band < 0.25 && finite_off) ? 1 : 0;
  return a;
}

}  // namespace rv3d

using namespace rv3d;

extern "C" size_t rv3d_rasterize_scratch_bytes(const rv3d_raster_params *p) {
  if (!p) return 0;
  return static_cast<size_t>(p->batch) * p->height * p->width * sizeof(unsigned long long);
}

// memset of the z-keys + K1 for the whole batch; shared by rv3d_rasterize and rv3d_rasterize_inputs
static int launch_scatter(const rv3d_raster_params *p, const RasterArgs &a, const float *points, const uint8_t *laser,
                          const int32_t *n_points, const int32_t *laser_mapping, unsigned long long *keys, cudaStream_t s) {
  const int64_t HW = static_cast<int64_t>(p->height) * p->width;
  RV3D_CHECK_CUDA(cudaMemsetAsync(keys, 0xFF, static_cast<size_t>(p->batch) * HW * sizeof(unsigned long long), s));
  dim3 g1(ceil_div(p->max_points, kScatterThreads * kScatterPerThread), p->batch);
  raster_scatter_kernel<<<g1, kScatterThreads, 0, s>>>(a, reinterpret_cast<const float4 *>(points), laser, n_points,
                                                       laser_mapping, keys);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

static int check_raster_args(const rv3d_raster_params *p, const void *points, const void *scratch, size_t scratch_bytes) {
  RV3D_CHECK_ARG(p->batch > 0 && p->max_points > 0 && p->height > 0 && p->width > 0 && p->azimuth_bins > 0);
  RV3D_CHECK_ARG(p->num_lasers > 0 && p->num_lasers <= 256 && p->reserved == 0);
  RV3D_CHECK_ARG(p->col_mode == RV3D_COL_LIBRARY || p->col_mode == RV3D_COL_CONVERTER);
  RV3D_CHECK_ARG(static_cast<int64_t>(p->height) * p->width < (int64_t(1) << 31));
  if (!aligned(points, 16) || !aligned(scratch, 8)) return RV3D_ERR_ALIGN;
  if (scratch_bytes < rv3d_rasterize_scratch_bytes(p)) return RV3D_ERR_SCRATCH;
  return RV3D_OK;
}

// Measured and dropped (profiles/r02_raster_chunks.md): processing the batch in L2-sized chunks of sweeps (scatter +
// resolve per chunk, so that the keys and points a resolve pass gathers are still L2-resident) is SLOWER at every chunk
// size -- the half-size kernels lose more to their tails and launch gaps than the gathers win.
extern "C" int rv3d_rasterize(const rv3d_raster_params *p, const float *points, const uint8_t *laser,
                              const int32_t *n_points, const int32_t *laser_mapping, float *image,
                              int32_t *winner, void *scratch, size_t scratch_bytes, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(p && points && laser && n_points && laser_mapping && image && scratch);
  if (const int st = check_raster_args(p, points, scratch, scratch_bytes)) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const RasterArgs a = make_args(p);
  auto *keys = static_cast<unsigned long long *>(scratch);
  if (const int st = launch_scatter(p, a, points, laser, n_points, laser_mapping, keys, s)) return st;
  const int64_t HW = static_cast<int64_t>(p->height) * p->width;
  dim3 g2(ceil_div(HW, kResolveThreads * kResolvePx), p->batch);
  raster_resolve_kernel<<<g2, kResolveThreads, 0, s>>>(a, reinterpret_cast<const float4 *>(points), keys, image, winner);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

static int make_inputs_args(const rv3d_inputs_params *p, InputsArgs &a) {
  RV3D_CHECK_ARG(p->batch > 0 && p->height > 0 && p->width > 0 && p->x_stride > 0 && p->pad >= 0);
  RV3D_CHECK_ARG(p->pad_mode == RV3D_PAD_CIRCULAR || p->pad_mode == RV3D_PAD_CONSTANT);
  RV3D_CHECK_ARG(p->n_features >= 0 && p->n_features <= 8 && p->tanh_channel >= -1 && p->tanh_channel < 7);
  RV3D_CHECK_ARG(p->pad_mode != RV3D_PAD_CIRCULAR || p->pad <= p->width);   // torch's circular pad limit
  a = InputsArgs{};
  a.B = p->batch; a.H = p->height; a.W = p->width; a.stride = p->x_stride; a.pad = p->pad; a.mode = p->pad_mode;
  a.Wo = (p->width + 2 * p->pad + p->x_stride - 1) / p->x_stride;
  a.F = p->n_features; a.tanh_ch = p->tanh_channel;
  for (int f = 0; f < a.F; ++f) {
    RV3D_CHECK_ARG(p->feature_channel[f] >= 0 && p->feature_channel[f] < 7);
    a.ch[f] = p->feature_channel[f];
  }
  return RV3D_OK;
}

extern "C" int rv3d_rasterize_inputs(const rv3d_raster_params *p, const rv3d_inputs_params *ip, const float *points,
                                     const uint8_t *laser, const int32_t *n_points, const int32_t *laser_mapping,
                                     float *features, float *cart, uint8_t *mask, void *scratch, size_t scratch_bytes,
                                     rv3d_stream_t stream) {
  RV3D_CHECK_ARG(p && ip && points && laser && n_points && laser_mapping && features && cart && mask && scratch);
  if (const int st = check_raster_args(p, points, scratch, scratch_bytes)) return st;
  RV3D_CHECK_ARG(ip->batch == p->batch && ip->height == p->height && ip->width == p->width);
  InputsArgs ia;
  if (const int st = make_inputs_args(ip, ia)) return st;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const RasterArgs a = make_args(p);
  auto *keys = static_cast<unsigned long long *>(scratch);
  if (const int st = launch_scatter(p, a, points, laser, n_points, laser_mapping, keys, s)) return st;
  int need_angles = 0;
  for (int f = 0; f < ia.F; ++f) need_angles |= (ia.ch[f] == 0 || ia.ch[f] == 1) ? 1 : 0;
  dim3 g2(ceil_div(static_cast<int64_t>(ia.H) * ia.Wo, kResolveThreads * kResolvePx), p->batch);
  raster_resolve_inputs_kernel<<<g2, kResolveThreads, 0, s>>>(a, ia, need_angles, reinterpret_cast<const float4 *>(points), keys,
                                                              features, cart, mask);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" size_t rv3d_zbuffer_scratch_bytes(int32_t height, int32_t width) {
  return static_cast<size_t>(height) * width * sizeof(unsigned long long);
}

extern "C" int rv3d_zbuffer(const int64_t *rows, const int64_t *cols, const void *dist, int32_t dist_is_f64,
                            const void *feat, int32_t feat_is_f64, int32_t channels, int64_t n,
                            int32_t height, int32_t width, double min_distance, float *image,
                            int32_t *winner, void *scratch, size_t scratch_bytes, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(image && scratch && height > 0 && width > 0 && channels >= 0 && n >= 0);
  RV3D_CHECK_ARG(n == 0 || (rows && cols && dist && feat));
  RV3D_CHECK_ARG(n < (int64_t(1) << 31) && static_cast<int64_t>(height) * width < (int64_t(1) << 31));
  const size_t need = rv3d_zbuffer_scratch_bytes(height, width);
  if (scratch_bytes < need) return RV3D_ERR_SCRATCH;
  if (!aligned(scratch, 8)) return RV3D_ERR_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto *keys = static_cast<unsigned long long *>(scratch);
  const int HW = height * width;
  RV3D_CHECK_CUDA(cudaMemsetAsync(keys, 0xFF, need, s));
  if (n > 0) {
    const int g = ceil_div(n, 256);
    if (dist_is_f64)
      zbuffer_scatter_kernel<double><<<g, 256, 0, s>>>(rows, cols, static_cast<const double *>(dist), n, height,
                                                       width, min_distance, keys);
    else
      zbuffer_scatter_kernel<float><<<g, 256, 0, s>>>(rows, cols, static_cast<const float *>(dist), n, height,
                                                      width, min_distance, keys);
    RV3D_CHECK_LAUNCH();
  }
  const int g2 = ceil_div(HW, 256);
  if (feat_is_f64)
    zbuffer_resolve_kernel<double><<<g2, 256, 0, s>>>(static_cast<const double *>(feat), channels, n, HW, keys,
                                                      image, winner);
  else
    zbuffer_resolve_kernel<float><<<g2, 256, 0, s>>>(static_cast<const float *>(feat), channels, n, HW, keys,
                                                     image, winner);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_cart_to_sph(const double *cart, double *sph, int64_t n, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && (n == 0 || (cart && sph)));
  if (n == 0) return RV3D_OK;
  cart_to_sph_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(cart, sph, n);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_range_view_coordinates(double *sph, const int64_t *laser, const int64_t *laser_mapping,
                                           int32_t n_mapping, int64_t n, int32_t n_inclination_bins,
                                           int32_t n_azimuth_bins, int32_t col_mode, double *hybrid,
                                           rv3d_stream_t stream) {
  const bool uniform = col_mode == RV3D_COL_CONVERTER_UNIFORM;   // rows from the inclination: no laser tables needed
  RV3D_CHECK_ARG(n >= 0 && n_azimuth_bins > 0 && n_inclination_bins > 0 && (n == 0 || (sph && hybrid)));
  RV3D_CHECK_ARG(uniform || (n_mapping > 0 && (n == 0 || (laser && laser_mapping))));
  RV3D_CHECK_ARG(col_mode == RV3D_COL_LIBRARY || col_mode == RV3D_COL_CONVERTER || uniform);
  if (n == 0) return RV3D_OK;
  RasterArgs a{};
  a.az_bins = n_azimuth_bins;
  a.col_mode = col_mode;
  a.bin_scale = static_cast<double>(n_azimuth_bins) / 6.283185307179586;
  rv_coordinates_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      sph, laser, laser_mapping, n_mapping, n, n_inclination_bins, a, hybrid);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_range_view_inputs(const rv3d_inputs_params *p, const float *image, float *features,
                                      float *cart, uint8_t *mask, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(p && image && features && cart && mask);
  InputsArgs a;
  if (const int st = make_inputs_args(p, a)) return st;
  dim3 grid(ceil_div(static_cast<int64_t>(a.H) * a.Wo, 256), a.B);
  range_view_inputs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, image, features, cart, mask);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_subsample_range_view(const float *range_view, const uint8_t *mask, const float *cart,
                                         int32_t batch, int32_t channels, int32_t height, int32_t width,
                                         int32_t x_stride, int32_t pad, int32_t pad_mode, float *out_range_view,
                                         uint8_t *out_mask, float *out_cart, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(range_view && mask && cart && out_range_view && out_mask && out_cart);
  RV3D_CHECK_ARG(batch > 0 && channels >= 0 && height > 0 && width > 0 && x_stride > 0 && pad >= 0);
  RV3D_CHECK_ARG(pad_mode == RV3D_PAD_CIRCULAR || pad_mode == RV3D_PAD_CONSTANT);
  RV3D_CHECK_ARG(pad_mode != RV3D_PAD_CIRCULAR || pad <= width);
  const int Wo = (width + 2 * pad + x_stride - 1) / x_stride;
  dim3 grid(ceil_div(static_cast<int64_t>(height) * Wo, 256), batch);
  subsample_range_view_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      range_view, mask, cart, channels, height, width, Wo, x_stride, pad, pad_mode, out_range_view, out_mask, out_cart);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_debug_fastmath(int32_t op, const double *a, const double *b, double *out, int64_t n,
                                   rv3d_stream_t stream) {
  RV3D_CHECK_ARG(op >= 0 && op <= 3 && n >= 0 && (n == 0 || (a && out)) && (b || op == 1 || op == 2 || n == 0));
  if (n == 0) return RV3D_OK;
  debug_fastmath_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(op, a, b, out, n);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_debug_column(const rv3d_raster_params *p, const float *points, int64_t n, int32_t *col_fast,
                                 int32_t *col_exact, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(p && n >= 0 && p->azimuth_bins > 0 && (n == 0 || (points && col_fast && col_exact)));
  RV3D_CHECK_ARG(p->col_mode == RV3D_COL_LIBRARY || p->col_mode == RV3D_COL_CONVERTER);
  if (n == 0) return RV3D_OK;
  if (!aligned(points, 16)) return RV3D_ERR_ALIGN;
  debug_column_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      make_args(p), reinterpret_cast<const float4 *>(points), n, col_fast, col_exact);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}
