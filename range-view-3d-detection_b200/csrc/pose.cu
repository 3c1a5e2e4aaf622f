// pose.cu -- sweep preparation in front of the rasterizer.
//
// Replaces (paths relative to /root/reference):
//   converters/av2/utils.py:229-295  unmotion_compensate   (per-point pose: scipy Slerp for the rotation, a
//                                     linear blend for the translation; inverse SE(3); compose with the
//                                     sweep's reference pose; apply)
//   converters/av2/utils.py:43-57    sensor_SE3_egovehicle = SE3(R, t).inverse(); .transform_point_cloud(cart)
//                                     (av2.geometry.se3, un-vendored: restated from its published definition)
//   converters/av2/utils.py:211-226  correct_laser_numbers (two table look-ups per point)
//
// One thread per point, everything in fp64 like the numpy code, nothing but the point's own row is read
// besides a pose table of a few thousand rows that lives in L1 / L2.  The kernels are HBM streams:
// 32 B in + 25 B out (unmotion), 24 + 24 B (transform), 8 + 8 B (laser numbers) per point.
#include <math_constants.h>

#include "common.cuh"
#include "fastmath.cuh"

namespace rv3d {

struct Quat { double x, y, z, w; };   // scalar-last, like scipy

__device__ __forceinline__ Quat q_normalized(Quat q) {
  const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return Quat{q.x / n, q.y / n, q.z / n, q.w / n};
}
// Hamilton product p * q (rotation q first, then p)
__device__ __forceinline__ Quat q_mul(Quat p, Quat q) {
  return Quat{p.w * q.x + q.w * p.x + (p.y * q.z - p.z * q.y),
              p.w * q.y + q.w * p.y + (p.z * q.x - p.x * q.z),
              p.w * q.z + q.w * p.z + (p.x * q.y - p.y * q.x),
              p.w * q.w - (p.x * q.x + p.y * q.y + p.z * q.z)};
}
// scipy Rotation.as_rotvec: canonical w >= 0, angle = 2 atan2(|v|, w), series below 1e-3
__device__ __forceinline__ void q_to_rotvec(Quat q, double (&rv)[3]) {
  if (q.w < 0.0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
  const double nv = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  const double angle = 2.0 * atan2(nv, q.w);
  double scale;
  if (angle <= 1e-3) {
    const double a2 = angle * angle;
    scale = 2.0 + a2 / 12.0 + 7.0 * a2 * a2 / 2880.0;
  } else {
    scale = angle / sin(angle / 2.0);
  }
  rv[0] = scale * q.x; rv[1] = scale * q.y; rv[2] = scale * q.z;
}
// scipy Rotation.from_rotvec
__device__ __forceinline__ Quat q_from_rotvec(const double (&rv)[3]) {
  const double n = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
  double scale;
  if (n <= 1e-3) {
    const double n2 = n * n;
    scale = 0.5 - n2 / 48.0 + n2 * n2 / 3840.0;
  } else {
    scale = sin(n / 2.0) / n;
  }
  return Quat{scale * rv[0], scale * rv[1], scale * rv[2], cos(n / 2.0)};
}
// unit quaternion -> row-major 3x3 (scipy Rotation.as_matrix)
__device__ __forceinline__ void q_to_matrix(Quat q, double (&m)[9]) {
  const double x2 = q.x * q.x, y2 = q.y * q.y, z2 = q.z * q.z, w2 = q.w * q.w;
  const double xy = q.x * q.y, zw = q.z * q.w, xz = q.x * q.z, yw = q.y * q.w, yz = q.y * q.z, xw = q.x * q.w;
  m[0] = x2 - y2 - z2 + w2; m[1] = 2.0 * (xy - zw);    m[2] = 2.0 * (xz + yw);
  m[3] = 2.0 * (xy + zw);   m[4] = -x2 + y2 - z2 + w2; m[5] = 2.0 * (yz - xw);
  m[6] = 2.0 * (xz - yw);   m[7] = 2.0 * (yz + xw);    m[8] = -x2 - y2 + z2 + w2;
}

__device__ __forceinline__ Quat load_quat(const double *p) { return q_normalized(Quat{p[0], p[1], p[2], p[3]}); }

struct UnmotionArgs {
  long long n;
  long long timestamp_ns;
  int n_poses;
  double target_rot[9];   // city_SE3_roll: the pose at the sweep's own timestamp (utils.py:258-273)
  double target_t[3];
};

// first index with a[i] >= v
template <typename T, typename F>
__device__ __forceinline__ int lower_bound(int n, T v, F at) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (at(mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
unmotion_kernel(UnmotionArgs a, const double *__restrict__ xyz, const long long *__restrict__ offset_ns,
                const long long *__restrict__ pose_ts, const double *__restrict__ pose_quat,
                const double *__restrict__ pose_t, double *__restrict__ out, uint8_t *__restrict__ valid) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const long long t = a.timestamp_ns + offset_ns[i];                          // :236
  const int M = a.n_poses;
  const bool ok = t > pose_ts[0] && t < pose_ts[M - 1];                       // :237-240 (strict on both sides)
  valid[i] = ok ? 1 : 0;
  if (!ok) {   // the reference drops the row; the host mirror compacts with this mask
    out[3 * i] = CUDART_NAN; out[3 * i + 1] = CUDART_NAN; out[3 * i + 2] = CUDART_NAN;
    return;
  }
  // translation: integer search, side = "left" (:244), rows idx - 1 and idx (:247-248)
  const int idx = lower_bound(M, t, [&](int k) { return pose_ts[k]; });
  const long long ts_lo = pose_ts[idx - 1], ts_hi = pose_ts[idx];
  const double alpha = static_cast<double>(t - ts_lo) / static_cast<double>(ts_hi - ts_lo);   // :275
  double tp[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)   // :276 -- the weights are the reference's (alpha on the LOWER pose)
    tp[k] = pose_t[3 * (idx - 1) + k] * alpha + (1.0 - alpha) * pose_t[3 * idx + k];

  // rotation: scipy Slerp works on float64 timestamps (ns since the epoch lose their low ~8 bits) and
  // searches those; mirror that instead of reusing idx
  const double tf = static_cast<double>(t);
  int ind = lower_bound(M, tf, [&](int k) { return static_cast<double>(pose_ts[k]); }) - 1;
  if (tf == static_cast<double>(pose_ts[0])) ind = 0;
  ind = ind < 0 ? 0 : (ind > M - 2 ? M - 2 : ind);
  const double t0 = static_cast<double>(pose_ts[ind]), t1 = static_cast<double>(pose_ts[ind + 1]);
  const double beta = (tf - t0) / (t1 - t0);
  const Quat q0 = load_quat(pose_quat + 4 * ind), q1 = load_quat(pose_quat + 4 * (ind + 1));
  double rv[3];
  q_to_rotvec(q_normalized(q_mul(Quat{-q0.x, -q0.y, -q0.z, q0.w}, q1)), rv);
  rv[0] *= beta; rv[1] *= beta; rv[2] *= beta;
  const Quat qp = q_normalized(q_mul(q0, q_from_rotvec(rv)));
  double R[9];
  q_to_matrix(qp, R);

  // laser_SE3_roll = inv(city_SE3_laser) @ city_SE3_roll applied to the point (:279-294):
  //   R_p^T (R_target x + t_target - t_p)
  const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  double v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    v[k] = (a.target_rot[3 * k] * x + a.target_rot[3 * k + 1] * y + a.target_rot[3 * k + 2] * z) + (a.target_t[k] - tp[k]);
#pragma unroll
  for (int k = 0; k < 3; ++k) out[3 * i + k] = R[k] * v[0] + R[3 + k] * v[1] + R[6 + k] * v[2];
}


// ------------------------------------------------------------------------------------------
// Table form (the production shape: one log = one pose table, many sweeps).  Everything about a Slerp step that
// depends on the pose PAIR only -- the normalised lower quaternion and the rotation vector of q0^-1 q1 (scipy
// Slerp.__init__ computes exactly this once per interval) -- is computed once per interval by pose_intervals_kernel
// with the same device functions as above, so a point's work shrinks to: locate the interval (interpolation guess +
// a short walk instead of two 12-step binary searches), blend, one sincos, one quaternion product, apply.
// ------------------------------------------------------------------------------------------

// ---- the per-point arithmetic of the table form: same formulas, cheaper primitives (all within ~2 fp64 ulps of the IEEE /
// libm forms above, far below the 1e-9 m the parity tests allow; everything outside the guarded ranges takes the forms
// above).  An IEEE fp64 division is ~20 instructions with a slow-path branch, libm's sin + cos ~80: a point has 7 of the
// former and one pair of the latter. ----
__device__ __forceinline__ double fast_rcp(double x) {      // 2^-500 < |x| < 2^500 (callers guard)
  double r = rcp_seed(x);
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ bool rcp_ok(double x) {
  return (static_cast<uint32_t>(__double2hiint(x)) >> 20 & 0x7ffu) - 523u < 1000u;
}
// sin / cos of a small angle (|x| <= 0.5: a Slerp step between neighbouring poses is milliradians): Taylor to x^15 / x^16
__device__ __forceinline__ void sincos_small(double x, double &s, double &c) {
  const double x2 = x * x;
  double ps = fma(x2, -1.0 / 1307674368000.0, 1.0 / 6227020800.0);
  ps = fma(x2, ps, -1.0 / 39916800.0);
  ps = fma(x2, ps, 1.0 / 362880.0);
  ps = fma(x2, ps, -1.0 / 5040.0);
  ps = fma(x2, ps, 1.0 / 120.0);
  ps = fma(x2, ps, -1.0 / 6.0);
  s = fma(x * x2, ps, x);
  double pc = fma(x2, 1.0 / 20922789888000.0, -1.0 / 87178291200.0);
  pc = fma(x2, pc, 1.0 / 479001600.0);
  pc = fma(x2, pc, -1.0 / 3628800.0);
  pc = fma(x2, pc, 1.0 / 40320.0);
  pc = fma(x2, pc, -1.0 / 720.0);
  pc = fma(x2, pc, 1.0 / 24.0);
  pc = fma(x2, pc, -0.5);
  c = fma(x2, pc, 1.0);
}
__device__ __forceinline__ Quat q_normalized_fast(Quat q) {
  const double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  if (!(n2 > 0x1p-200 && n2 < 0x1p200)) return q_normalized(q);
  const double y = rsqrt_seed(n2);                         // 1 / sqrt(n2): seed + coupled Newton steps (fast_sqrt's core)
  double g = n2 * y, h = 0.5 * y;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g); h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  g = fma(g, e, g); h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  h = fma(h, e, h);
  const double r = 2.0 * h;
  return Quat{q.x * r, q.y * r, q.z * r, q.w * r};
}
__device__ __forceinline__ Quat q_from_rotvec_fast(const double (&rv)[3]) {
  const double n2 = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2];
  if (!(n2 > 1e-6 && n2 <= 1.0)) return q_from_rotvec(rv);    // n <= 1e-3 takes scipy's series, n > 1 the library sincos
  const double n = fast_sqrt(n2);
  double sn, cs;
  sincos_small(0.5 * n, sn, cs);
  const double scale = sn * fast_rcp(n);
  return Quat{scale * rv[0], scale * rv[1], scale * rv[2], cs};
}

__global__ void __launch_bounds__(128)
pose_intervals_kernel(int n_poses, const double *__restrict__ pose_quat, double *__restrict__ iv) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_poses - 1) return;
  const Quat q0 = load_quat(pose_quat + 4 * k), q1 = load_quat(pose_quat + 4 * (k + 1));
  double rv[3];
  q_to_rotvec(q_normalized(q_mul(Quat{-q0.x, -q0.y, -q0.z, q0.w}, q1)), rv);
  double *o = iv + 8 * static_cast<size_t>(k);
  o[0] = q0.x; o[1] = q0.y; o[2] = q0.z; o[3] = q0.w; o[4] = rv[0]; o[5] = rv[1]; o[6] = rv[2]; o[7] = 0.0;
}

struct UnmotionTableArgs {
  long long n;
  long long timestamp_ns;
  long long first_ns, last_ns;   // pose_ts[0], pose_ts[M - 1]
  double guess_scale;            // (M - 1) / (last - first)
  int n_poses;
  double target_rot[9];
  double target_t[3];
};

__global__ void __launch_bounds__(256)
unmotion_table_kernel(UnmotionTableArgs a, const double *__restrict__ xyz, const long long *__restrict__ offset_ns,
                      const long long *__restrict__ pose_ts, const double *__restrict__ pose_t,
                      const double *__restrict__ iv, double *__restrict__ out, uint8_t *__restrict__ valid,
                      int *__restrict__ n_dropped) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool live = i < a.n;
  const long long t = live ? a.timestamp_ns + offset_ns[i] : 0;               // :236
  const bool ok = live && t > a.first_ns && t < a.last_ns;                     // :237-240 (strict on both sides)
  if (n_dropped) {                                                             // rows the reference's filter removes
    const unsigned bad = __ballot_sync(0xffffffffu, live && !ok);
    if (bad && (threadIdx.x & 31) == 0) atomicAdd(n_dropped, __popc(bad));
  }
  if (!live) return;
  valid[i] = ok ? 1 : 0;
  if (!ok) {
    out[3 * i] = CUDART_NAN; out[3 * i + 1] = CUDART_NAN; out[3 * i + 2] = CUDART_NAN;
    return;
  }
  const int M = a.n_poses;
  // idx = first k with pose_ts[k] >= t (searchsorted side = "left", :244); ts[0] < t < ts[M-1] => 1 <= idx <= M-1.
  // Pose tables are (nearly) uniform in time: start at the interpolated position and walk; a table that defeats the
  // guess falls back to the binary search.
  int idx = static_cast<int>(static_cast<double>(t - a.first_ns) * a.guess_scale) + 1;
  idx = idx < 1 ? 1 : (idx > M - 1 ? M - 1 : idx);
  int steps = 0;
  while (steps < 6 && idx > 1 && pose_ts[idx - 1] >= t) { --idx; ++steps; }
  while (steps < 6 && pose_ts[idx] < t) { ++idx; ++steps; }
  if (steps >= 6) idx = lower_bound(M, t, [&](int k) { return pose_ts[k]; });
  const long long ts_lo = pose_ts[idx - 1], ts_hi = pose_ts[idx];
  const double span = static_cast<double>(ts_hi - ts_lo);
  const double alpha = rcp_ok(span) ? static_cast<double>(t - ts_lo) * fast_rcp(span)                // :275
                                    : static_cast<double>(t - ts_lo) / span;
  double tp[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)   // :276 -- the weights are the reference's (alpha on the LOWER pose)
    tp[k] = pose_t[3 * (idx - 1) + k] * alpha + (1.0 - alpha) * pose_t[3 * idx + k];

  // scipy Slerp searches float64 timestamps.  The conversion is monotone, so its lower bound is idx unless earlier
  // timestamps round to the same double as t (t within ~256 ns of a pose): walk down over those.
  const double tf = static_cast<double>(t);
  int lb = idx;
  while (lb > 0 && static_cast<double>(pose_ts[lb - 1]) >= tf) --lb;
  int ind = lb - 1;
  if (tf == static_cast<double>(a.first_ns)) ind = 0;
  ind = ind < 0 ? 0 : (ind > M - 2 ? M - 2 : ind);
  const double t0 = static_cast<double>(pose_ts[ind]), t1 = static_cast<double>(pose_ts[ind + 1]);
  const double beta = rcp_ok(t1 - t0) ? (tf - t0) * fast_rcp(t1 - t0) : (tf - t0) / (t1 - t0);
  const double *v8 = iv + 8 * static_cast<size_t>(ind);
  const Quat q0{v8[0], v8[1], v8[2], v8[3]};
  double rv[3] = {v8[4] * beta, v8[5] * beta, v8[6] * beta};
  const Quat qp = q_normalized_fast(q_mul(q0, q_from_rotvec_fast(rv)));
  double R[9];
  q_to_matrix(qp, R);
  const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
  double v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    v[k] = (a.target_rot[3 * k] * x + a.target_rot[3 * k + 1] * y + a.target_rot[3 * k + 2] * z) + (a.target_t[k] - tp[k]);
#pragma unroll
  for (int k = 0; k < 3; ++k) out[3 * i + k] = R[k] * v[0] + R[3 + k] * v[1] + R[6 + k] * v[2];
}

struct RigidArgs { double r[9]; double t[3]; };

// out = p @ R^T + t  (av2 SE3.transform_point_cloud)
__global__ void __launch_bounds__(256)
transform_kernel(RigidArgs a, const double *__restrict__ xyz, long long n, double *__restrict__ out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
#pragma unroll
  for (int k = 0; k < 3; ++k) out[3 * i + k] = (a.r[3 * k] * x + a.r[3 * k + 1] * y + a.r[3 * k + 2] * z) + a.t[k];
}

// two rows per thread (16-byte loads / stores when the arrays are 16-byte aligned); the two small tables sit in shared memory
__global__ void __launch_bounds__(256)
laser_numbers_kernel(const long long *__restrict__ laser, long long n, const long long *__restrict__ laser_mapping,
                     const long long *__restrict__ row_mapping, int n_rows, long long *__restrict__ out,
                     int *__restrict__ bad, int vec_ok) {
  __shared__ long long s_rows[256];
  __shared__ long long s_map[32];
  const int nr = n_rows < 256 ? n_rows : 256;
  for (int k = threadIdx.x; k < nr; k += blockDim.x) s_rows[k] = row_mapping[k];
  if (laser_mapping && threadIdx.x < 32) s_map[threadIdx.x] = laser_mapping[threadIdx.x];
  __syncthreads();
  auto one = [&](long long l) -> long long {
    if (laser_mapping) {                        // :214-220 (log in LOG_IDS): upper and lower block of 32 beams
      if (l >= 32 && l < 64) l = s_map[l - 32] + 32;
      else if (l >= 0 && l < 32) l = s_map[l];
    }
    if (l < 0 || l >= n_rows) {                 // numpy would raise IndexError
      if (bad) atomicExch(bad, 1);
      return -1;
    }
    return l < 256 ? s_rows[l] : row_mapping[l];   // :222-226
  };
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i >= n) return;
  if (vec_ok && i + 1 < n) {
    const longlong2 v = *reinterpret_cast<const longlong2 *>(laser + i);
    *reinterpret_cast<longlong2 *>(out + i) = make_longlong2(one(v.x), one(v.y));
  } else {
    out[i] = one(laser[i]);
    if (i + 1 < n) out[i + 1] = one(laser[i + 1]);
  }
}

}  // namespace rv3d

using namespace rv3d;

static void quat_to_matrix_host(const double *q_xyzw, double *m) {
  double x = q_xyzw[0], y = q_xyzw[1], z = q_xyzw[2], w = q_xyzw[3];
  const double n = sqrt(x * x + y * y + z * z + w * w);
  x /= n; y /= n; z /= n; w /= n;
  const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
  const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
  m[0] = x2 - y2 - z2 + w2; m[1] = 2.0 * (xy - zw);    m[2] = 2.0 * (xz + yw);
  m[3] = 2.0 * (xy + zw);   m[4] = -x2 + y2 - z2 + w2; m[5] = 2.0 * (yz - xw);
  m[6] = 2.0 * (xz - yw);   m[7] = 2.0 * (yz + xw);    m[8] = -x2 - y2 + z2 + w2;
}

extern "C" int rv3d_unmotion_compensate(const double *xyz, const int64_t *offset_ns, int64_t n, int64_t timestamp_ns,
                                        const int64_t *pose_timestamps_ns, const double *pose_quat_xyzw,
                                        const double *pose_translation, int32_t n_poses,
                                        const double *target_quat_xyzw, const double *target_translation,
                                        double *out_xyz, uint8_t *out_valid, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && n_poses >= 2 && pose_timestamps_ns && pose_quat_xyzw && pose_translation);
  RV3D_CHECK_ARG(target_quat_xyzw && target_translation);
  if (n == 0) return RV3D_OK;
  RV3D_CHECK_ARG(xyz && offset_ns && out_xyz && out_valid);
  UnmotionArgs a;
  a.n = n; a.timestamp_ns = timestamp_ns; a.n_poses = n_poses;
  quat_to_matrix_host(target_quat_xyzw, a.target_rot);
  for (int k = 0; k < 3; ++k) a.target_t[k] = target_translation[k];
  unmotion_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, xyz, reinterpret_cast<const long long *>(offset_ns), reinterpret_cast<const long long *>(pose_timestamps_ns),
      pose_quat_xyzw, pose_translation, out_xyz, out_valid);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}


extern "C" int rv3d_pose_intervals(const double *pose_quat_xyzw, int32_t n_poses, double *intervals, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n_poses >= 2 && pose_quat_xyzw && intervals);
  pose_intervals_kernel<<<ceil_div(n_poses - 1, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(n_poses, pose_quat_xyzw,
                                                                                                 intervals);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_unmotion_compensate_table(const double *xyz, const int64_t *offset_ns, int64_t n, int64_t timestamp_ns,
                                              const int64_t *pose_timestamps_ns, const double *pose_translation,
                                              const double *intervals, int32_t n_poses, int64_t first_ns, int64_t last_ns,
                                              const double *target_quat_xyzw, const double *target_translation,
                                              double *out_xyz, uint8_t *out_valid, int32_t *n_dropped,
                                              rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && n_poses >= 2 && pose_timestamps_ns && pose_translation && intervals && last_ns > first_ns);
  RV3D_CHECK_ARG(target_quat_xyzw && target_translation);
  if (n == 0) return RV3D_OK;
  RV3D_CHECK_ARG(xyz && offset_ns && out_xyz && out_valid);
  UnmotionTableArgs a;
  a.n = n; a.timestamp_ns = timestamp_ns; a.n_poses = n_poses; a.first_ns = first_ns; a.last_ns = last_ns;
  a.guess_scale = static_cast<double>(n_poses - 1) / static_cast<double>(last_ns - first_ns);
  quat_to_matrix_host(target_quat_xyzw, a.target_rot);
  for (int k = 0; k < 3; ++k) a.target_t[k] = target_translation[k];
  unmotion_table_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, xyz, reinterpret_cast<const long long *>(offset_ns), reinterpret_cast<const long long *>(pose_timestamps_ns),
      pose_translation, intervals, out_xyz, out_valid, n_dropped);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_transform_points(const double *xyz, int64_t n, const double *rotation, const double *translation,
                                     int32_t inverse, double *out, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && rotation && translation);
  if (n == 0) return RV3D_OK;
  RV3D_CHECK_ARG(xyz && out);
  RigidArgs a;
  if (inverse) {   // SE3.inverse(): rotation^T, rotation^T . (-translation)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) a.r[3 * r + c] = rotation[3 * c + r];
    for (int r = 0; r < 3; ++r)
      a.t[r] = a.r[3 * r] * -translation[0] + a.r[3 * r + 1] * -translation[1] + a.r[3 * r + 2] * -translation[2];
  } else {
    for (int k = 0; k < 9; ++k) a.r[k] = rotation[k];
    for (int k = 0; k < 3; ++k) a.t[k] = translation[k];
  }
  transform_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, xyz, n, out);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_correct_laser_numbers(const int64_t *laser_numbers, int64_t n, const int64_t *laser_mapping,
                                          const int64_t *row_mapping, int32_t n_rows, int64_t *out,
                                          int32_t *out_of_range, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && row_mapping && n_rows > 0);
  if (n == 0) return RV3D_OK;
  RV3D_CHECK_ARG(laser_numbers && out);
  const int vec_ok = aligned(laser_numbers, 16) && aligned(out, 16);
  laser_numbers_kernel<<<ceil_div(n, 512), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long *>(laser_numbers), n, reinterpret_cast<const long long *>(laser_mapping),
      reinterpret_cast<const long long *>(row_mapping), n_rows, reinterpret_cast<long long *>(out), out_of_range, vec_ok);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}
