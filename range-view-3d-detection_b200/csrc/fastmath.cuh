// fastmath.cuh -- fp64 atan2 / exp / sqrt and a float32 atan2 at a fraction of libm's instruction count.
//
// The path's azimuth / inclination / yaw / size values are all rounded to float32 (or half) afterwards; the fp64
// routines here are accurate to ~1 fp64 ulp (atan2: ~3e-16 absolute, exp: ~2e-16 relative), so after that rounding they
// differ from libm's result in fewer than 1e-6 of the values, by one float32 ulp -- the same class of difference as
// device libm vs host libm, which the parity bars already allow (DESIGN.md "Parity bars").  Measured error
// distributions: tests/test_gpu_fastmath.py through rv3d_debug_fastmath.
//
// What is avoided, and why (ncu, profiles/r02_*): on sm_100 every f32<->f64 / int<->float conversion and every MUFU
// runs on the XU pipe at 16 lanes per SM (8 cycles per warp instruction per scheduler); the previous forms spent 8 XU
// instructions per atan2 and the rasterizer kernels sat at 40-44 % XU utilisation next to 63-69 % issue utilisation.
//   * reciprocal / rsqrt seeds come from MUFU.RCP64H / RSQ64H on the fp64 value itself (no round trip through float32);
//   * "round to a multiple of 2^-8" and "round to integer" are one DADD with a 1.5 * 2^k constant, the integer is read
//     from the low word (no F2I / I2F);
//   * quadrant fix-ups are one DADD with a selected constant and a sign flip.
//
// fast_atan2:
//   t = min(|x|,|y|) / max(|x|,|y|) in [0, 1];  k = round(256 t) from the 20-bit seed quotient;  t_k = k / 256
//   atan(t) = atan(t_k) + atan(u),  u = (mn - t_k mx) / (mx + t_k mn),  |u| <= 1/512 + 2^-19
//   atan(u) = u - u^3/3 + u^5/5   (next term < 2^-56 relative)
// atan(t_k) comes from a 257-entry table; 1 / den is the MUFU seed + two Newton steps.
#pragma once
#include "atan_table.cuh"

namespace rv3d {

// MUFU.RCP64H / MUFU.RSQ64H: ~20 good bits, computed from the high word of the operand
__device__ __forceinline__ double rcp_seed(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}
__device__ __forceinline__ double rsqrt_seed(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

__device__ __forceinline__ double fast_atan2(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const bool swap = ay > ax;
  const double mx = swap ? ay : ax, mn = swap ? ax : ay;
  // 2^-100 <= mx < 2^100 (biased exponent in [923, 1123); NaN / inf: 2047, zero / subnormal: 0); a NaN in the smaller
  // operand fails mn <= mx.  Everything else: libm.
  const uint32_t ex = static_cast<uint32_t>(__double2hiint(mx)) >> 20;
  if (ex - 923u >= 200u || !(mn <= mx)) return atan2(y, x);
  const double tq = mn * rcp_seed(mx);            // t to ~2^-19
  const double d = tq + 0x1.8p44;                 // ulp(2^44) = 2^-8: d = 1.5 * 2^44 + k / 256, k in the low word
  const int k = __double2loint(d);                // 0 .. 256
  const double tk = d - 0x1.8p44;                 // k / 256, exact
  const double num = fma(-tk, mx, mn);
  const double den = fma(tk, mn, mx);
  double r = rcp_seed(den);
  double e = fma(-den, r, 1.0);
  r = fma(r, e, r);
  e = fma(-den, r, 1.0);
  r = fma(r, e, r);
  const double u = num * r;
  const double u2 = u * u;
  const double q = fma(u2, 0.2, -1.0 / 3.0);
  double a = kAtanTable[k] + fma(u * u2, q, u);
  // quadrant: (swap, x < 0) -> a | pi/2 - a | pi - a | pi/2 + a  =  c0 + (+-a); pi and pi/2 share their low word
  const bool negx = __double2hiint(x) < 0;
  const int c0_hi = swap ? 0x3FF921FB : (negx ? 0x400921FB : 0);
  const int c0_lo = (swap || negx) ? 0x54442D18 : 0;
  if (swap != negx) a = -a;
  a = __hiloint2double(c0_hi, c0_lo) + a;
  return copysign(a, y);
}

// exp(x) = 2^(k/64) * e^r,  k = round(64 x / ln 2),  r = x - k ln2/64 (two-constant reduction), |r| <= ln2/128:
//   e^r = 1 + r + r^2 (1/2 + r/6 + r^2/24 + r^3/120)   (next term 3.4e-17 relative)
// 2^((k mod 64)/64) from a 64-entry table, 2^(k div 64) added to the exponent field (|x| < 700: always normal).
__device__ __forceinline__ double fast_exp_core(double x) {   // |x| < 700
  const double d = fma(x, 92.33248261689366, 0x1.8p52);   // 64 / ln 2; integer k in the low word (two's complement)
  const int k = __double2loint(d);
  const double kd = d - 0x1.8p52;
  double r = fma(kd, -0x1.62e42fee00000p-7, x);   // ln2/64 high part: 32 significant bits, kd * hi is exact
  r = fma(kd, -0x1.a39ef35793c76p-39, r);
  double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(r, p, 1.0 / 6.0);
  p = fma(r, p, 0.5);
  const double r2 = r * r;
  const double em1 = fma(r2, p, r);               // e^r - 1
  const double t = kExp2Table[k & 63];
  const double v = fma(t, em1, t);
  return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
}
__device__ __forceinline__ double fast_exp(double x) {
  if (!(fabs(x) < 700.0)) return exp(x);          // overflow / underflow / NaN: libm
  return fast_exp_core(x);
}
// three at once (the decoder's l, w, h): one range test, three independent chains for the scheduler to interleave
__device__ __forceinline__ void fast_exp3(double x0, double x1, double x2, double &e0, double &e1, double &e2) {
  if (!(fmax(fmax(fabs(x0), fabs(x1)), fabs(x2)) < 700.0) || x0 != x0 || x1 != x1 || x2 != x2) {
    e0 = exp(x0); e1 = exp(x1); e2 = exp(x2);
    return;
  }
  e0 = fast_exp_core(x0); e1 = fast_exp_core(x1); e2 = fast_exp_core(x2);
}

// sqrt(s) to <= 2 fp64 ulps for 2^-200 <= s < 2^200 (callers guard): RSQ64H seed + two coupled Newton steps + one
// residual correction; no special-case branches (libm's sqrt spends ~1/3 of its instructions there).
__device__ __forceinline__ double fast_sqrt(double s) {
  const double y = rsqrt_seed(s);
  double g = s * y, h = 0.5 * y;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  const double dd = fma(-g, g, s);
  return fma(dd, h, g);
}
__device__ __forceinline__ bool fast_sqrt_ok(double s) {   // 2^-200 <= s < 2^200, finite, positive
  return (static_cast<uint32_t>(__double2hiint(s)) >> 20) - 823u < 400u;
}

// float32 atan2 for decisions that are re-checked in fp64 when close (the rasterizer's azimuth bin): |error| < 4e-7 rad
// against the exact atan2 of its float32 arguments (odd minimax polynomial of degree 15 on [0, 1], 4.9e-8; approximate
// division, 1.2e-7; float32 pi / pi/2 and the roundings of the fix-ups, 2e-7).  Callers guard the operand range
// (`ax + ay` finite and in [1e-30, 1e30]); no special cases here.
__device__ __forceinline__ float atan2f_lite(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = __fdividef(mn, mx);
  const float s = t * t;
  float p = fmaf(s, -0.0043553938157856464f, 0.02304009348154068f);
  p = fmaf(s, p, -0.05777352675795555f);
  p = fmaf(s, p, 0.0979423001408577f);
  p = fmaf(s, p, -0.13976579904556274f);
  p = fmaf(s, p, 0.19962704181671143f);
  p = fmaf(s, p, -0.3333165943622589f);
  float a = fmaf(p * s, t, t);
  if (ay > ax) a = 1.57079637f - a;
  if (x < 0.f) a = 3.14159274f - a;
  return copysignf(a, y);
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
