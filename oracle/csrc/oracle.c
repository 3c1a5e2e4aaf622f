/*
 * oracle.c -- CPU restatement of the native arithmetic the reference's hot path
 * delegates to un-vendored third-party extensions.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under range-view-3d-detection_b200/ may
 * link, import or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC oracle.c -lm
 * (-ffp-contract=off: every float op rounds once, no FMA contraction, so the
 * CUDA side, compiled with --fmad=false, can match bit for bit.)
 *
 * What is restated, and from where (paths relative to /root/reference):
 *  - orc_zbuffer            src/torchbox3d/math/numpy/conversions.py:106-128
 *                           (numba serial nearest-return scatter; copy at
 *                           converters/av2/utils.py:186-208)
 *  - orc_rot_iou            detectron2 box_iou_rotated_utils.h
 *                           single_box_iou_rotated<float>, CUDA flavour of the
 *                           hull sort (call site: math/ops/nms.py:41-45).
 *                           detectron2 is an un-pinned dependency
 *                           (conda/environment.yml:6) that is NOT in the
 *                           reference tree: restated from its published
 *                           algorithm => "parity unpinned" (see DESIGN.md).
 *  - orc_nms_rotated        detectron2 nms_rotated_cuda.cu host loop: sort by
 *                           score desc, greedy, suppress on iou > thr.
 *  - orc_iou_bev            mmdet3d / OpenPCDet iou3d kernel on
 *                           (x1,y1,x2,y2,ry), the routine TorchEx's
 *                           weighted_nms_ext derives from (call site
 *                           math/ops/nms.py:161-170).  TorchEx is un-pinned and
 *                           absent => "parity unpinned".
 *  - orc_wnms               TorchEx wnms_gpu contract as pinned by the wrapper
 *                           math/ops/nms.py:126-177 + RangeDet py_weighted_nms
 *                           semantics.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* z-buffer: numpy/conversions.py:106-128                              */
/* ------------------------------------------------------------------ */
/* rows/cols: (N,) int64.  dist: (N,) f64 or f32.  feat: (C,N) f64 or f32.
 * image: (C,H*W) f32, winner: (H*W,) int32 (extra output: the index of the
 * point whose features ended up in the pixel, -1 if none).                 */
void orc_zbuffer(const int64_t *rows, const int64_t *cols, const void *dist,
                 int dist_is_f64, const void *feat, int feat_is_f64, int C,
                 int64_t N, int H, int W, double min_distance, float *image,
                 int32_t *winner) {
  int64_t P = (int64_t)H * W;
  float *buffer = (float *)malloc(sizeof(float) * (size_t)P);
  for (int64_t p = 0; p < P; ++p) buffer[p] = INFINITY;            /* :118 */
  memset(image, 0, sizeof(float) * (size_t)C * (size_t)P);          /* :117 */
  if (winner)
    for (int64_t p = 0; p < P; ++p) winner[p] = -1;
  for (int64_t i = 0; i < N; ++i) {
    int64_t p = rows[i] * (int64_t)W + cols[i];                     /* :121 */
    double d = dist_is_f64 ? ((const double *)dist)[i]
                           : (double)((const float *)dist)[i];
    if (d < min_distance) continue;                                 /* :123 */
    if (p < 0 || p >= P) continue; /* numba: UB; we drop the point */
    if (d < (double)buffer[p]) {                                    /* :125 */
      for (int c = 0; c < C; ++c) {                                 /* :126 */
        image[(int64_t)c * P + p] =
            feat_is_f64 ? (float)((const double *)feat)[(int64_t)c * N + i]
                        : ((const float *)feat)[(int64_t)c * N + i];
      }
      buffer[p] = dist_is_f64 ? (float)d : ((const float *)dist)[i]; /* :127 */
      if (winner) winner[p] = (int32_t)i;
    }
  }
  free(buffer);
}

/* ------------------------------------------------------------------ */
/* detectron2 single_box_iou_rotated<float> (CUDA flavour)             */
/* ------------------------------------------------------------------ */
typedef struct { float x, y; } pt;

static inline float cross2(pt a, pt b) { return a.x * b.y - b.x * a.y; }
static inline float dot2(pt a, pt b) { return a.x * b.x + a.y * b.y; }
static inline pt sub(pt a, pt b) { pt r = {a.x - b.x, a.y - b.y}; return r; }

/* vertices of a box (xc,yc,w,h,theta[rad as double]).
 * detectron2 (box_iou_rotated_utils.h get_rotated_vertices): the w-axis points along (cos, -sin).
 * mmcv (ops/csrc/common/box_iou_rotated_utils.hpp, "y: top --> down; x: left --> right", its default
 * clockwise=True): the same routine with the OTHER rotation direction, w-axis along (cos, +sin), and its own
 * vertex order.  mmcv's published unit-test vector (tests/test_oracle_iou.py) tells the two apart. */
static void rot_vertices(float xc, float yc, float w, float h, double theta,
                         int mmcv, pt v[4]) {
  float c2 = (float)cos(theta) * 0.5f;
  float s2 = (float)sin(theta) * 0.5f;
  if (mmcv) {
    v[0].x = xc - s2 * h - c2 * w;
    v[0].y = yc + c2 * h - s2 * w;
    v[1].x = xc + s2 * h - c2 * w;
    v[1].y = yc - c2 * h - s2 * w;
  } else {
    v[0].x = xc + s2 * h + c2 * w;
    v[0].y = yc + c2 * h - s2 * w;
    v[1].x = xc - s2 * h + c2 * w;
    v[1].y = yc - c2 * h - s2 * w;
  }
  v[2].x = 2 * xc - v[0].x;
  v[2].y = 2 * yc - v[0].y;
  v[3].x = 2 * xc - v[1].x;
  v[3].y = 2 * yc - v[1].y;
}

static int isect_points(const pt p1[4], const pt p2[4], pt out[24]) {
  pt v1[4], v2[4];
  for (int i = 0; i < 4; ++i) {
    v1[i] = sub(p1[(i + 1) % 4], p1[i]);
    v2[i] = sub(p2[(i + 1) % 4], p2[i]);
  }
  const double EPS = 1e-5;
  int n = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float det = cross2(v2[j], v1[i]);
      if (fabs((double)det) <= 1e-14) continue;
      pt v12 = sub(p2[j], p1[i]);
      float t1 = cross2(v2[j], v12) / det;
      float t2 = cross2(v1[i], v12) / det;
      if (t1 > -EPS && t1 < 1.0f + EPS && t2 > -EPS && t2 < 1.0f + EPS) {
        pt r = {p1[i].x + v1[i].x * t1, p1[i].y + v1[i].y * t1};
        out[n++] = r;
      }
    }
  { /* vertices of rect1 inside rect2 */
    pt AB = v2[0], DA = v2[3];
    float ABdotAB = dot2(AB, AB), ADdotAD = dot2(DA, DA);
    for (int i = 0; i < 4; ++i) {
      pt AP = sub(p1[i], p2[0]);
      float APdotAB = dot2(AP, AB);
      float APdotAD = -dot2(AP, DA);
      if (APdotAB > -EPS && APdotAD > -EPS && APdotAB < ABdotAB + EPS &&
          APdotAD < ADdotAD + EPS)
        out[n++] = p1[i];
    }
  }
  { /* vertices of rect2 inside rect1 */
    pt AB = v1[0], DA = v1[3];
    float ABdotAB = dot2(AB, AB), ADdotAD = dot2(DA, DA);
    for (int i = 0; i < 4; ++i) {
      pt AP = sub(p2[i], p1[0]);
      float APdotAB = dot2(AP, AB);
      float APdotAD = -dot2(AP, DA);
      if (APdotAB > -EPS && APdotAD > -EPS && APdotAB < ABdotAB + EPS &&
          APdotAD < ADdotAD + EPS)
        out[n++] = p2[i];
    }
  }
  return n;
}

/* Graham scan, shift_to_zero=true variant; returns hull size, hull in q. */
static int hull_graham(const pt p[24], int n_in, pt q[24]) {
  int t = 0;
  for (int i = 1; i < n_in; ++i)
    if (p[i].y < p[t].y || (p[i].y == p[t].y && p[i].x < p[t].x)) t = i;
  pt start = p[t];
  for (int i = 0; i < n_in; ++i) q[i] = sub(p[i], start);
  pt tmp = q[0]; q[0] = q[t]; q[t] = tmp;
  float dist[24];
  for (int i = 0; i < n_in; ++i) dist[i] = dot2(q[i], q[i]);
  /* CUDA flavour: O(n^2) exchange sort by angle, ties by distance */
  for (int i = 1; i < n_in - 1; ++i)
    for (int j = i + 1; j < n_in; ++j) {
      float cp = cross2(q[i], q[j]);
      if ((cp < -1e-6) || (fabs((double)cp) < 1e-6 && dist[i] > dist[j])) {
        pt qt = q[i]; q[i] = q[j]; q[j] = qt;
        float dt = dist[i]; dist[i] = dist[j]; dist[j] = dt;
      }
    }
  int k;
  for (k = 1; k < n_in; ++k)
    if (dist[k] > 1e-8) break;
  if (k == n_in) { q[0] = p[t]; return 1; }
  q[1] = q[k];
  int m = 2;
  for (int i = k + 1; i < n_in; ++i) {
    while (m > 1) {
      pt q1 = sub(q[i], q[m - 2]), q2 = sub(q[m - 1], q[m - 2]);
      if (q1.x * q2.y >= q2.x * q1.y) m--; else break;
    }
    q[m++] = q[i];
  }
  return m;
}

static float poly_area(const pt q[24], int m) {
  if (m <= 2) return 0.f;
  float area = 0.f;
  for (int i = 1; i < m - 1; ++i)
    area += fabsf(cross2(sub(q[i], q[0]), sub(q[i + 1], q[0])));
  return (float)(area / 2.0);
}

/* box = (xc, yc, w, h, angle); angle_scale converts the stored angle to
 * radians in double: detectron2 uses 0.01745329251 (degrees in), mmcv uses 1. */
static float rot_iou_flavour(const float *b1, const float *b2, double angle_scale, int mmcv) {
  float sx = (float)((double)(b1[0] + b2[0]) / 2.0);
  float sy = (float)((double)(b1[1] + b2[1]) / 2.0);
  float x1 = (float)((double)b1[0] - (double)sx), y1 = (float)((double)b1[1] - (double)sy);
  float x2 = (float)((double)b2[0] - (double)sx), y2 = (float)((double)b2[1] - (double)sy);
  float area1 = b1[2] * b1[3], area2 = b2[2] * b2[3];
  if (area1 < 1e-14 || area2 < 1e-14) return 0.f;
  pt p1[4], p2[4], ip[24], hp[24];
  rot_vertices(x1, y1, b1[2], b1[3], (double)b1[4] * angle_scale, mmcv, p1);
  rot_vertices(x2, y2, b2[2], b2[3], (double)b2[4] * angle_scale, mmcv, p2);
  int n = isect_points(p1, p2, ip);
  float inter = 0.f;
  if (n > 2) {
    int m = hull_graham(ip, n, hp);
    inter = poly_area(hp, m);
  }
  return inter / (area1 + area2 - inter);
}

float orc_rot_iou(const float *b1, const float *b2, double angle_scale) {   /* detectron2 */
  return rot_iou_flavour(b1, b2, angle_scale, 0);
}

void orc_rot_iou_aligned(const float *a, const float *b, int64_t n,
                         double angle_scale, float *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = orc_rot_iou(a + 5 * i, b + 5 * i, angle_scale);
}

/* mmcv.ops.box_iou_rotated(aligned=True, clockwise=True): angles in radians */
void orc_rot_iou_aligned_mmcv(const float *a, const float *b, int64_t n, float *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = rot_iou_flavour(a + 5 * i, b + 5 * i, 1.0, 1);
}

/* order[] must hold the indices sorted by score descending (ties: index
 * ascending).  Greedy scan as in nms_rotated_cuda.cu; suppression on
 * iou > thr (CUDA comparison; the CPU kernel of detectron2 uses >=).
 * Returns the number kept; keep[] receives ORIGINAL indices in score order. */
int64_t orc_nms_rotated(const float *boxes, const int64_t *order, int64_t n,
                        double thr, double angle_scale, int64_t *keep,
                        int64_t *n_iou_evals, int64_t max_keep) {
  uint8_t *removed = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
  int64_t nk = 0, evals = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (removed[i]) continue;
    /* max_keep >= 0: stop once that many boxes are kept.  The caller only uses the
     * first num_post_nms kept boxes (nms.py:53-56), so the output is unchanged. */
    if (max_keep >= 0 && nk >= max_keep) break;
    keep[nk++] = order[i];
    const float *bi = boxes + 5 * order[i];
    for (int64_t j = i + 1; j < n; ++j) {
      if (removed[j]) continue;
      ++evals;
      if ((double)orc_rot_iou(bi, boxes + 5 * order[j], angle_scale) > thr)
        removed[j] = 1;
    }
  }
  free(removed);
  if (n_iou_evals) *n_iou_evals = evals;
  return nk;
}

/* ------------------------------------------------------------------ */
/* iou_bev on (x1,y1,x2,y2,ry): mmdet3d / OpenPCDet lineage            */
/* ------------------------------------------------------------------ */
#define BEV_EPS 1e-8f

static inline float cross3(pt p1, pt p2, pt p0) {
  return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

static inline int rect_cross(pt p1, pt p2, pt q1, pt q2) {
  return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) &&
         fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
         fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) &&
         fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}

static int seg_isect(pt p1, pt p0, pt q1, pt q0, pt *ans) {
  if (!rect_cross(p0, p1, q0, q1)) return 0;
  float s1 = cross3(q0, p1, p0);
  float s2 = cross3(p1, q1, p0);
  float s3 = cross3(p0, q1, q0);
  float s4 = cross3(q1, p1, q0);
  if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
  float s5 = cross3(q1, p1, p0);
  if (fabsf(s5 - s1) > BEV_EPS) {
    ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
    ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
  } else {
    float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
    float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
    float D = a0 * b1 - a1 * b0;
    ans->x = (b0 * c1 - b1 * c0) / D;
    ans->y = (a1 * c0 - a0 * c1) / D;
  }
  return 1;
}

/* Rotation is COUNTER-CLOCKWISE by +ry about the box centre (our choice; the
 * TorchEx source is unavailable -- see DESIGN.md "parity unpinned").        */
static inline pt rot_about(pt c, float ca, float sa, pt p) {
  pt r;
  r.x = (p.x - c.x) * ca - (p.y - c.y) * sa + c.x;
  r.y = (p.x - c.x) * sa + (p.y - c.y) * ca + c.y;
  return r;
}

static int in_box2d(const float *box, float ca, float sa, pt p) {
  const float MARGIN = 1e-5f;
  float cx = (box[0] + box[2]) / 2, cy = (box[1] + box[3]) / 2;
  /* rotate the point by -ry into the box frame */
  float rx = (p.x - cx) * ca + (p.y - cy) * sa + cx;
  float ry = -(p.x - cx) * sa + (p.y - cy) * ca + cy;
  return rx > box[0] - MARGIN && rx < box[2] + MARGIN && ry > box[1] - MARGIN &&
         ry < box[3] + MARGIN;
}

static float bev_overlap(const float *a, const float *b) {
  pt ca_ = {(a[0] + a[2]) / 2, (a[1] + a[3]) / 2};
  pt cb_ = {(b[0] + b[2]) / 2, (b[1] + b[3]) / 2};
  pt A[5] = {{a[0], a[1]}, {a[2], a[1]}, {a[2], a[3]}, {a[0], a[3]}};
  pt B[5] = {{b[0], b[1]}, {b[2], b[1]}, {b[2], b[3]}, {b[0], b[3]}};
  float aca = (float)cos((double)a[4]), asa = (float)sin((double)a[4]);
  float bca = (float)cos((double)b[4]), bsa = (float)sin((double)b[4]);
  for (int k = 0; k < 4; ++k) {
    A[k] = rot_about(ca_, aca, asa, A[k]);
    B[k] = rot_about(cb_, bca, bsa, B[k]);
  }
  A[4] = A[0];
  B[4] = B[0];
  pt cp[16];
  pt pc = {0.f, 0.f};
  int cnt = 0;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      if (seg_isect(A[i + 1], A[i], B[j + 1], B[j], &cp[cnt])) {
        pc.x += cp[cnt].x; pc.y += cp[cnt].y;
        cnt++;
      }
    }
  for (int k = 0; k < 4; ++k) {
    if (in_box2d(a, aca, asa, B[k])) { pc.x += B[k].x; pc.y += B[k].y; cp[cnt++] = B[k]; }
    if (in_box2d(b, bca, bsa, A[k])) { pc.x += A[k].x; pc.y += A[k].y; cp[cnt++] = A[k]; }
  }
  if (cnt == 0) return 0.f;
  pc.x /= cnt; pc.y /= cnt;
  /* bubble sort by angle about the centroid (atan2 evaluated in double, cast
   * to float so host libm and device libm agree after rounding) */
  float ang[16];
  for (int i = 0; i < cnt; ++i)
    ang[i] = (float)atan2((double)(cp[i].y - pc.y), (double)(cp[i].x - pc.x));
  for (int j = 0; j < cnt - 1; ++j)
    for (int i = 0; i < cnt - j - 1; ++i)
      if (ang[i] > ang[i + 1]) {
        pt t = cp[i]; cp[i] = cp[i + 1]; cp[i + 1] = t;
        float ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
      }
  float area = 0.f;
  for (int k = 0; k < cnt - 1; ++k)
    area += cross2(sub(cp[k], cp[0]), sub(cp[k + 1], cp[0]));
  return fabsf(area) / 2.0f;
}

float orc_iou_bev(const float *a, const float *b) {
  float sa = (a[2] - a[0]) * (a[3] - a[1]);
  float sb = (b[2] - b[0]) * (b[3] - b[1]);
  float so = bev_overlap(a, b);
  return so / fmaxf(sa + sb - so, BEV_EPS);
}

/* Weighted NMS.  Inputs are ALREADY sorted by score descending, exactly as
 * math/ops/nms.py:148-154 hands them to wnms_gpu:
 *   boxes (n,5) f32 (x1,y1,x2,y2,ry); data (n,D) f32 with the score in the
 *   LAST column; output (n,D) f32 zero-initialised; keep (n,) int64;
 *   count (n,) int64 zero-initialised.  Returns num_out.
 * For each kept k (greedy, suppression on iou > nms_thr), the merge set is
 * {k} + {j > k alive when k is kept, iou(k,j) > merge_thr}; output row =
 * score-weighted mean over the set for columns 0..D-2, score column = s_k. */
int64_t orc_wnms(const float *boxes, const float *data, int64_t n, int D,
                 float nms_thr, float merge_thr, float *output, int64_t *keep,
                 int64_t *count, int64_t max_out) {
  uint8_t *removed = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
  double *acc = (double *)malloc(sizeof(double) * (size_t)D);
  int64_t nk = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (removed[i]) continue;
    if (max_out >= 0 && nk >= max_out) break;
    double wsum = (double)data[i * D + D - 1];
    for (int c = 0; c < D - 1; ++c) acc[c] = wsum * (double)data[i * D + c];
    int64_t cnt = 1;
    for (int64_t j = i + 1; j < n; ++j) {
      if (removed[j]) continue;
      float iou = orc_iou_bev(boxes + 5 * i, boxes + 5 * j);
      if (iou > merge_thr) {
        double s = (double)data[j * D + D - 1];
        wsum += s;
        for (int c = 0; c < D - 1; ++c) acc[c] += s * (double)data[j * D + c];
        cnt++;
      }
      if (iou > nms_thr) removed[j] = 1;
    }
    for (int c = 0; c < D - 1; ++c) output[nk * D + c] = (float)(acc[c] / wsum);
    output[nk * D + D - 1] = data[i * D + D - 1];
    keep[nk] = i;
    count[nk] = cnt;
    nk++;
  }
  free(removed);
  free(acc);
  return nk;
}
