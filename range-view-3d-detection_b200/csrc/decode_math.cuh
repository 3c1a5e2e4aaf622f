// decode_math.cuh -- typed loads and the per-pixel box decoder shared by decode.cu (dense / compacting decode kernels)
// and assign.cu (training-time targets decode only the foreground pixels they need).
//
// Replaces (paths relative to /root/reference): math/ops/coding.py:110-144 decode_range_view (+ :79-107).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.cuh"
#include "fastmath.cuh"

namespace rv3d {

// ------------------------------------------------------------------------------------------
// typed loads: everything is widened to float on load (exact for f16 / bf16)
// ------------------------------------------------------------------------------------------
template <typename T> struct Ld;
template <> struct Ld<float> {
  static __device__ __forceinline__ float one(const float *p) { return __ldg(p); }
  static __device__ __forceinline__ void four(const float *p, float (&v)[4]) {
    const float4 t = ldg_stream_f4(reinterpret_cast<const float4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ float cast(double x) { return static_cast<float>(x); }
  static __device__ __forceinline__ float round_f32(float x) { return x; }
};
template <> struct Ld<__half> {
  static __device__ __forceinline__ float one(const __half *p) { return __half2float(*p); }
  static __device__ __forceinline__ void four(const __half *p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    const __half2 a = *reinterpret_cast<const __half2 *>(&t.x), b = *reinterpret_cast<const __half2 *>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ __half cast(double x) { return __double2half(x); }
  static __device__ __forceinline__ float round_f32(float x) { return __half2float(__float2half_rn(x)); }
};
template <> struct Ld<__nv_bfloat16> {
  static __device__ __forceinline__ float one(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void four(const __nv_bfloat16 *p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
  static __device__ __forceinline__ __nv_bfloat16 cast(double x) { return __double2bfloat16(x); }
  static __device__ __forceinline__ float round_f32(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};

// sigmoid the way the reference's own CUDA path evaluates it (range_decoder.py:49: torch's CUDA sigmoid
// computes 1 / (1 + exp(-x)) in float32 opmath for float32 / float16 / bfloat16 tensors and rounds to the
// tensor dtype), returned widened to float.  Different libms differ by an ulp here.
template <typename T>
__device__ __forceinline__ float sigmoid_t(float x) {
  return static_cast<float>(Ld<T>::cast(static_cast<double>(1.0f / (1.0f + expf(-x)))));
}

// coding.py:110-144 in fp64.  reg[8], cart[3] -> out[7] (double)
__device__ __forceinline__ void decode_box(const float (&reg)[8], const float (&cart)[3], bool az_inv,
                                           double (&out)[7]) {
  double ox = reg[0], oy = reg[1];
  const double oz = reg[2];
  const double sy = reg[6], cyw = reg[7];
  const double cx = cart[0], cy = cart[1], cz = cart[2];
  double yaw;
  if (az_inv) {                                                                    // :79-107
    // cos / sin of phi = atan2(cy, cx) are cx / h and cy / h (one reciprocal square root instead of a sincos);
    // degenerate or non-finite rays take the library path
    const double h2 = cx * cx + cy * cy;
    double s, c;
    if (h2 > 1e-60 && h2 < 1e60) {
      double rh = rsqrt_seed(h2);                  // 1 / sqrt(h2): MUFU seed + two Newton steps (<= 1 ulp)
      double e = fma(-h2 * rh, rh, 1.0);
      rh = fma(0.5 * rh, e, rh);
      e = fma(-h2 * rh, rh, 1.0);
      rh = fma(0.5 * rh, e, rh);
      c = cx * rh; s = cy * rh;
    } else {
      sincos(atan2(cy, cx), &s, &c);
    }
    const double x = c * ox - s * oy;
    const double y = s * ox + c * oy;
    ox = x; oy = y;
    // yaw = atan2(sy, cyw) + atan2(cy, cx) (:136, :104) with ONE arctangent: the sum of the two angles is the angle of
    // the product of the two complex numbers, up to a multiple of 2 pi that the signs of the two angles determine
    // (both in [0, pi] -> sum in [0, 2 pi]; both negative -> [-2 pi, 0); mixed -> (-pi, pi)).  Next to the branch cut
    // of the merged arctangent (|Y| tiny against |X|) rounding could pick the wrong sheet: those take the two-call form.
    const double Y = fma(sy, cx, cyw * cy), X = fma(cyw, cx, -(sy * cy));
    if (fabs(Y) > 1.0e-9 * fabs(X)) {
      const double m = fast_atan2(Y, X);
      const bool n1 = __double2hiint(sy) < 0, n2 = __double2hiint(cy) < 0;   // sign bits: atan2(-0, .) is negative too
      double wrap = 0.0;
      if (!n1 && !n2 && m < 0.0) wrap = 6.283185307179586;
      if (n1 && n2 && m > 0.0) wrap = -6.283185307179586;
      yaw = m + wrap;
    } else {
      yaw = fast_atan2(sy, cyw) + fast_atan2(cy, cx);
    }
  } else {
    yaw = fast_atan2(sy, cyw);                                                     // :136 (fastmath.cuh: <= 2 ulp)
  }
  out[0] = cx + ox; out[1] = cy + oy; out[2] = cz + oz;                            // :142
  fast_exp3(static_cast<double>(reg[3]), static_cast<double>(reg[4]), static_cast<double>(reg[5]), out[3], out[4], out[5]);  // :132
  out[6] = yaw;
}

}  // namespace rv3d
