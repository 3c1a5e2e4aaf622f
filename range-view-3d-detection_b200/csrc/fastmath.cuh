// fastmath.cuh -- fp64 atan2 at ~1/3 of libm's instruction count, accurate to ~3e-16 absolute (2e-16 relative for
// small angles): the path's azimuth / inclination / yaw values are all rounded to float32 (or half) afterwards,
// where this differs from libm's result in fewer than 1e-6 of the values, by one float32 ulp (the same class of
// difference as device libm vs host libm, which the parity bars already allow: DESIGN.md "Parity bars").
//
//   t = min(|x|,|y|) / max(|x|,|y|) in [0, 1];  k = round(256 t) from a float32 estimate;  t_k = k / 256
//   atan(t) = atan(t_k) + atan(u),  u = (mn - t_k mx) / (mx + t_k mn),  |u| <= 1/512 + 1e-6
//   atan(u) = u - u^3/3 + u^5/5 - u^7/7   (next term < 1e-26)
// atan(t_k) comes from a 257-entry table; the reciprocal is a float32 seed + three Newton steps.
#pragma once
#include "atan_table.cuh"

namespace rv3d {

__device__ __forceinline__ double fast_atan2(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const double mx = fmax(ax, ay), mn = fmin(ax, ay);
  if (!(mx > 1e-30 && mx < 1e30)) return atan2(y, x);   // zeros, infinities, NaN, extreme magnitudes: libm
  const float tf = __fdividef(static_cast<float>(mn), static_cast<float>(mx));
  const int k = __float2int_rn(tf * 256.0f);             // 0 .. 256
  const double tk = static_cast<double>(k) * 0.00390625;
  const double num = fma(-tk, mx, mn);
  const double den = fma(tk, mn, mx);
  double r = static_cast<double>(__frcp_rn(static_cast<float>(den)));
  r = r * fma(-den, r, 2.0);
  r = r * fma(-den, r, 2.0);
  r = r * fma(-den, r, 2.0);
  const double u = num * r;
  const double u2 = u * u;
  double q = fma(u2, -1.0 / 7.0, 0.2);
  q = fma(u2, q, -1.0 / 3.0);
  double a = kAtanTable[k] + fma(u * u2, q, u);
  if (ay > ax) a = 1.5707963267948966 - a;
  if (signbit(x)) a = 3.141592653589793 - a;
  return copysign(a, y);
}

}  // namespace rv3d
