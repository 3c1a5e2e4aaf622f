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

// hypot(x, y) exactly as the reference's numpy evaluates it: numpy.hypot is libm's hypot, and glibc (>= 2.35; 2.39 in this
// image) implements it, when no FMA is used, as  h = sqrt(ax^2 + ay^2)  followed by one correction step
// (sysdeps/ieee754/dbl-64/e_hypot.c `kernel`, after C. Borges, "An Improved Algorithm for hypot(a,b)").  The result is
// NOT the correctly rounded value (it differs from it in ~0.6 % of the cases), so being bit-exact with the reference
// means restating it: IEEE add / mul / div / sqrt only, no contraction (the library is built with --fmad=false).
// oracle/libm_hypot.py holds the same restatement in numpy; tests/test_oracle_golden.py pins it against numpy.hypot.
// Inputs here are differences of float32-origin values: the scaling branches for |x| > 2^511 or |y| < 2^-459 never apply
// (y == 0 takes the `ax >= ay / EPS` exit like upstream).
static __device__ __noinline__ double libm_hypot(double x, double y) {
  x = fabs(x); y = fabs(y);
  const double ax = x < y ? y : x, ay = x < y ? x : y;
  if (ax >= ay / 0x1p-54) return ax + ay;
  double h = sqrt(ax * ax + ay * ay);
  double t1, t2;
  if (h <= 2.0 * ay) {
    const double delta = h - ay;
    t1 = ax * (2.0 * delta - ax);
    t2 = (delta - 2.0 * (ax - ay)) * delta;
  } else {
    const double delta = h - ax;
    t1 = 2.0 * delta * (ax - 2.0 * ay);
    t2 = (4.0 * delta - ay) * ay + delta * delta;
  }
  h -= (t1 + t2) / (2.0 * h);
  return h;
}

}  // namespace rv3d
