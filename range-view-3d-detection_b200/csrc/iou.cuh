// iou.cuh -- rotated BEV IoU device functions.
//
// Two routines, because the reference binds two different third-party kernels:
//   rot_iou   : detectron2 / mmcv single_box_iou_rotated<float> on (xc, yc, w, h, angle)
//               (call sites math/ops/nms.py:41-45 and math/ops/iou.py:15).
//   bev_iou   : mmdet3d / OpenPCDet iou_bev on (x1, y1, x2, y2, ry), the routine TorchEx's
//               weighted NMS derives from (call site math/ops/nms.py:161-170).
// Both third-party sources are absent from the reference tree ("parity unpinned"); these
// functions are written against oracle/csrc/oracle.c and must agree with it BIT FOR BIT:
// same operation order, no FMA contraction (the library is built with --fmad=false), IEEE
// division, trig evaluated in fp64 and rounded once to float.
#pragma once
#include "common.cuh"

namespace rv3d {

struct P2 { float x, y; };

__device__ __forceinline__ float cross2(P2 a, P2 b) { return a.x * b.y - b.x * a.y; }
__device__ __forceinline__ float dot2(P2 a, P2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ P2 sub2(P2 a, P2 b) { return P2{a.x - b.x, a.y - b.y}; }

// ------------------------------------------------------------------------------------------
// record formats (one per candidate, in sorted order)
// ------------------------------------------------------------------------------------------
struct __align__(16) HardRec {  // 32 B
  float x, y, w, h;             // centre, extent along the box axis / across it
  float c2, s2;                 // (float)cos(theta) * 0.5f, (float)sin(theta) * 0.5f (mmcv flavour: -sin, see make_hard_rec)
  float r;                      // padded circumscribed radius (pruning only)
  float wv;                     // w as the vertex formula sees it: w (detectron2) or -w (mmcv)
};

struct __align__(16) WRec {     // 64 B
  float x1, y1, x2, y2;         // axis-aligned extents before rotation
  float ca, sa;                 // (float)cos(ry), (float)sin(ry)
  float r;                      // padded circumscribed radius (pruning only)
  float pad;
  float qx[4], qy[4];           // corners rotated by +ry about the centre
};

__device__ __forceinline__ float padded_radius(float w, float h) {
  // A pair whose padded circles are disjoint yields zero intersection points in either IoU
  // routine, so skipping it is exact.  The pads cover the routines' own tolerances:
  //   - edge/edge hits accept t in (-1e-5, 1+1e-5): edges grow by 1e-5 * length  -> 0.2 % of r
  //   - vertex-inside tests accept projections down to -1e-5 (units m^2): a box of extent e
  //     "contains" points up to 1e-5 / e outside it                             -> 2e-5 / min(w,h)
  // Degenerate (tiny / NaN / inf) boxes get an infinite or NaN radius: inf is never pruned,
  // NaN is always pruned, which matches an IoU of NaN never exceeding a threshold.
  const float r = 0.5f * sqrtf(w * w + h * h);
  return r * 1.002f + 1e-3f + 2e-5f / fminf(fabsf(w), fabsf(h));
}

// mmcv = false: detectron2's get_rotated_vertices (the w-axis points along (cos, -sin); nms_rotated, degrees in).
// mmcv = true: mmcv.ops.box_iou_rotated with its default clockwise=True (radians): the same routine with the OTHER rotation
// direction and its own vertex order,
//     v0 = (xc - s2 h - c2 w, yc + c2 h - s2 w),  v1 = (xc + s2 h - c2 w, yc - c2 h - s2 w),
// which is detectron2's formula evaluated with s2 -> -s2 and w -> -w: sign changes are exact in floating point, so storing
// the record that way reproduces mmcv's arithmetic bit for bit with ONE vertex routine (area and pruning keep the true w).
// Which direction mmcv uses is pinned by its published unit-test vector (tests/test_oracle_iou.py::test_mmcv_published_vectors).
__device__ __forceinline__ HardRec make_hard_rec(float xc, float yc, float w, float h, float angle,
                                                 double angle_scale, bool mmcv = false) {
  HardRec r;
  const double theta = static_cast<double>(angle) * angle_scale;
  r.x = xc; r.y = yc; r.w = w; r.h = h;
  r.c2 = static_cast<float>(cos(theta)) * 0.5f;
  const float s2 = static_cast<float>(sin(theta)) * 0.5f;
  r.s2 = mmcv ? -s2 : s2;
  r.r = padded_radius(w, h);
  r.wv = mmcv ? -w : w;
  return r;
}

__device__ __forceinline__ WRec make_w_rec(float x1, float y1, float x2, float y2, float ry) {
  WRec r;
  r.x1 = x1; r.y1 = y1; r.x2 = x2; r.y2 = y2;
  r.ca = static_cast<float>(cos(static_cast<double>(ry)));
  r.sa = static_cast<float>(sin(static_cast<double>(ry)));
  r.r = padded_radius(x2 - x1, y2 - y1);
  r.pad = 0.f;
  const float cx = (x1 + x2) / 2, cy = (y1 + y2) / 2;
  const float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    r.qx[k] = (px[k] - cx) * r.ca - (py[k] - cy) * r.sa + cx;
    r.qy[k] = (px[k] - cx) * r.sa + (py[k] - cy) * r.ca + cy;
  }
  return r;
}

__device__ __forceinline__ float rec_cx(const HardRec &r) { return r.x; }
__device__ __forceinline__ float rec_cy(const HardRec &r) { return r.y; }
__device__ __forceinline__ float rec_cx(const WRec &r) { return (r.x1 + r.x2) * 0.5f; }
__device__ __forceinline__ float rec_cy(const WRec &r) { return (r.y1 + r.y2) * 0.5f; }

// ------------------------------------------------------------------------------------------
// detectron2-style rotated IoU.  a = higher-ranked box (box1), b = lower-ranked (box2).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void hard_vertices(float xc, float yc, const HardRec &b, P2 (&v)[4]) {
  v[0].x = xc + b.s2 * b.h + b.c2 * b.wv;
  v[0].y = yc + b.c2 * b.h - b.s2 * b.wv;
  v[1].x = xc - b.s2 * b.h + b.c2 * b.wv;
  v[1].y = yc - b.c2 * b.h - b.s2 * b.wv;
  v[2].x = 2 * xc - v[0].x;
  v[2].y = 2 * yc - v[0].y;
  v[3].x = 2 * xc - v[1].x;
  v[3].y = 2 * yc - v[1].y;
}

// The upstream routine compares float32 values with double constants (1e-5, 1e-14, 1e-6, 1e-8) and halves / subtracts
// through double.  On sm_100 every float <-> double conversion is an XU-pipe instruction (16 lanes per SM): the literal
// form spent 85 F2F + 134 DSETP per call, on the suppression kernel's critical path.  Each of those operations has an
// exactly equivalent float32 form:
//   (double)t > D  <=>  t > FL(D),  FL(D) = the largest float32 <= D;      (double)t < D  <=>  t < FC(D),  FC(D) = the
//   smallest float32 >= D   (no float32 lies strictly between D and its float32 neighbour);  a bound computed at run time
//   in double, (double)u + 1e-5, becomes its round-up conversion;  halving is exact in either type;  the difference of two
//   float32 values rounded through double equals their float32 difference (the double difference is exact unless the
//   operands are > 2^29 apart, where both forms return the larger one).
// The CPU oracle keeps the literal form; tests/test_gpu_nms.py and tests/test_gpu_iou_decisions.py compare bit for bit.
constexpr float kNegEpsLE = -0x1.4f8b5ap-17f;      // FL(-1e-5)
constexpr float kOnePlusEpsGE = 0x1.0000a8p+0f;    // FC(1 + 1e-5)
constexpr float kDetLE = 0x1.6849b8p-47f;          // FL(1e-14)
constexpr float kAreaGE = 0x1.6849bap-47f;         // FC(1e-14)
constexpr float kNegCpGE = -0x1.0c6f7ap-20f;       // FC(-1e-6)
constexpr float kCpGE = 0x1.0c6f7cp-20f;           // FC(1e-6)
constexpr float kDistLE = 0x1.5798eep-27f;         // FL(1e-8)

static __device__ __noinline__ float rot_iou(const HardRec &a, const HardRec &b) {
  const float area1 = a.w * a.h, area2 = b.w * b.h;
  if (area1 < kAreaGE || area2 < kAreaGE) return 0.f;
  // shift both centres by their midpoint (computed in double upstream)
  const float sx = (a.x + b.x) * 0.5f;
  const float sy = (a.y + b.y) * 0.5f;
  const float ax = a.x - sx;
  const float ay = a.y - sy;
  const float bx = b.x - sx;
  const float by = b.y - sy;
  P2 p1[4], p2[4];
  hard_vertices(ax, ay, a, p1);
  hard_vertices(bx, by, b, p2);

  P2 v1[4], v2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v1[i] = sub2(p1[(i + 1) & 3], p1[i]);
    v2[i] = sub2(p2[(i + 1) & 3], p2[i]);
  }
  const double EPS = 1e-5;
  P2 ip[24];
  int n = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float det = cross2(v2[j], v1[i]);
      if (fabsf(det) <= kDetLE) continue;
      const P2 v12 = sub2(p2[j], p1[i]);
      const float t1 = cross2(v2[j], v12) / det;
      const float t2 = cross2(v1[i], v12) / det;
      if (t1 > kNegEpsLE && t1 < kOnePlusEpsGE && t2 > kNegEpsLE && t2 < kOnePlusEpsGE) {
        ip[n].x = p1[i].x + v1[i].x * t1;
        ip[n].y = p1[i].y + v1[i].y * t1;
        ++n;
      }
    }
  }
  {
    const P2 AB = v2[0], DA = v2[3];
    const float ubAB = __double2float_ru(static_cast<double>(dot2(AB, AB)) + EPS);   // FC(ABdotAB + EPS)
    const float ubAD = __double2float_ru(static_cast<double>(dot2(DA, DA)) + EPS);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const P2 AP = sub2(p1[i], p2[0]);
      const float APdotAB = dot2(AP, AB);
      const float APdotAD = -dot2(AP, DA);
      if (APdotAB > kNegEpsLE && APdotAD > kNegEpsLE && APdotAB < ubAB && APdotAD < ubAD) ip[n++] = p1[i];
    }
  }
  {
    const P2 AB = v1[0], DA = v1[3];
    const float ubAB = __double2float_ru(static_cast<double>(dot2(AB, AB)) + EPS);
    const float ubAD = __double2float_ru(static_cast<double>(dot2(DA, DA)) + EPS);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const P2 AP = sub2(p2[i], p1[0]);
      const float APdotAB = dot2(AP, AB);
      const float APdotAD = -dot2(AP, DA);
      if (APdotAB > kNegEpsLE && APdotAD > kNegEpsLE && APdotAB < ubAB && APdotAD < ubAD) ip[n++] = p2[i];
    }
  }
  float inter = 0.f;
  if (n > 2) {
    // Graham scan (shift_to_zero variant), CUDA flavour of the angular sort
    // (ip / q / dist are indexed dynamically and live in local memory: the loops below keep the element they work on
    // in registers, so the exchange sort loads one element per step instead of re-reading q[i] and both distances --
    // the same operations in the same order as upstream)
    int t = 0;
    P2 start = ip[0];
    for (int i = 1; i < n; ++i) {
      const P2 c = ip[i];
      if (c.y < start.y || (c.y == start.y && c.x < start.x)) { t = i; start = c; }
    }
    P2 q[24];
    float dist[24];
    for (int i = 0; i < n; ++i) {
      const P2 d = sub2(ip[i], start);
      q[i] = d;
      dist[i] = dot2(d, d);
    }
    {
      const P2 tmp = q[0]; q[0] = q[t]; q[t] = tmp;
      const float dt = dist[0]; dist[0] = dist[t]; dist[t] = dt;
    }
    for (int i = 1; i < n - 1; ++i) {
      P2 qi = q[i];
      float di = dist[i];
      for (int j = i + 1; j < n; ++j) {
        const P2 qj = q[j];
        const float cp = cross2(qi, qj);
        bool exchange = cp < kNegCpGE;
        if (!exchange && fabsf(cp) < kCpGE) exchange = di > dist[j];
        if (exchange) {
          const float dj = dist[j];
          q[j] = qi; dist[j] = di;
          qi = qj; di = dj;
        }
      }
      q[i] = qi;
      dist[i] = di;
    }
    int k = 1;
    for (; k < n; ++k)
      if (dist[k] > kDistLE) break;
    int m = 1;
    if (k < n) {
      q[1] = q[k];
      m = 2;
      for (int i = k + 1; i < n; ++i) {
        while (m > 1) {
          const P2 q1 = sub2(q[i], q[m - 2]), q2 = sub2(q[m - 1], q[m - 2]);
          if (q1.x * q2.y >= q2.x * q1.y) m--; else break;
        }
        q[m++] = q[i];
      }
    }
    if (m > 2) {
      float area = 0.f;
      for (int i = 1; i < m - 1; ++i) area += fabsf(cross2(sub2(q[i], q[0]), sub2(q[i + 1], q[0])));
      inter = area * 0.5f;
    }
  }
  return inter / (area1 + area2 - inter);
}

// ------------------------------------------------------------------------------------------
// mmdet3d / OpenPCDet-style iou_bev on (x1,y1,x2,y2,ry); rotation counter-clockwise by +ry.
// ------------------------------------------------------------------------------------------
#define RV3D_BEV_EPS 1e-8f

__device__ __forceinline__ float cross3(P2 p1, P2 p2, P2 p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

__device__ __forceinline__ bool seg_isect(P2 p1, P2 p0, P2 q1, P2 q0, P2 &ans) {
  const bool rc = fminf(p0.x, p1.x) <= fmaxf(q0.x, q1.x) && fminf(q0.x, q1.x) <= fmaxf(p0.x, p1.x) &&
                  fminf(p0.y, p1.y) <= fmaxf(q0.y, q1.y) && fminf(q0.y, q1.y) <= fmaxf(p0.y, p1.y);
  if (!rc) return false;
  const float s1 = cross3(q0, p1, p0);
  const float s2 = cross3(p1, q1, p0);
  const float s3 = cross3(p0, q1, q0);
  const float s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
  const float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > RV3D_BEV_EPS) {
    ans.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    const float D = a0 * b1 - a1 * b0;
    ans.x = (b0 * c1 - b1 * c0) / D;
    ans.y = (a1 * c0 - a0 * c1) / D;
  }
  return true;
}

__device__ __forceinline__ bool in_box2d(const WRec &b, P2 p) {
  const float MARGIN = 1e-5f;
  const float cx = (b.x1 + b.x2) / 2, cy = (b.y1 + b.y2) / 2;
  const float rx = (p.x - cx) * b.ca + (p.y - cy) * b.sa + cx;
  const float ry = -(p.x - cx) * b.sa + (p.y - cy) * b.ca + cy;
  return rx > b.x1 - MARGIN && rx < b.x2 + MARGIN && ry > b.y1 - MARGIN && ry < b.y2 + MARGIN;
}

static __device__ __noinline__ float bev_iou(const WRec &a, const WRec &b) {
  P2 A[5], B[5];
#pragma unroll
  for (int k = 0; k < 4; ++k) { A[k] = P2{a.qx[k], a.qy[k]}; B[k] = P2{b.qx[k], b.qy[k]}; }
  A[4] = A[0];
  B[4] = B[0];
  P2 cp[16];
  P2 pc{0.f, 0.f};
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      P2 ans;
      if (seg_isect(A[i + 1], A[i], B[j + 1], B[j], ans)) {
        cp[cnt] = ans;
        pc.x += ans.x; pc.y += ans.y;
        ++cnt;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (in_box2d(a, B[k])) { pc.x += B[k].x; pc.y += B[k].y; cp[cnt++] = B[k]; }
    if (in_box2d(b, A[k])) { pc.x += A[k].x; pc.y += A[k].y; cp[cnt++] = A[k]; }
  }
  float overlap = 0.f;
  if (cnt > 0) {
    pc.x /= cnt; pc.y /= cnt;
    float ang[16];
    for (int i = 0; i < cnt; ++i)
      ang[i] = static_cast<float>(atan2(static_cast<double>(cp[i].y - pc.y), static_cast<double>(cp[i].x - pc.x)));
    for (int j = 0; j < cnt - 1; ++j)
      for (int i = 0; i < cnt - j - 1; ++i)
        if (ang[i] > ang[i + 1]) {
          const P2 t = cp[i]; cp[i] = cp[i + 1]; cp[i + 1] = t;
          const float ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
        }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) area += cross2(sub2(cp[k], cp[0]), sub2(cp[k + 1], cp[0]));
    overlap = fabsf(area) / 2.0f;
  }
  const float sa = (a.x2 - a.x1) * (a.y2 - a.y1);
  const float sb = (b.x2 - b.x1) * (b.y2 - b.y1);
  return overlap / fmaxf(sa + sb - overlap, RV3D_BEV_EPS);
}

__device__ __forceinline__ float pair_iou(const HardRec &a, const HardRec &b) { return rot_iou(a, b); }
__device__ __forceinline__ float pair_iou(const WRec &a, const WRec &b) { return bev_iou(a, b); }

}  // namespace rv3d
