// nms.cu -- segment sort (K3), exact greedy rotated NMS / weighted NMS (K4 / K5), output pack.
//
// Replaces (paths relative to /root/reference):
//   math/ops/nms.py:181-266  batched_multiclass_nms  (per-sweep Python loop, host sync each)
//   math/ops/nms.py:11-61    hard_multiclass_nms     -> detectron2 nms_rotated (:41-45)
//   math/ops/nms.py:64-123   weighted_multiclass_nms -> weighted_nms (:126-177) -> TorchEx wnms_gpu
//   nn/decoders/range_decoder.py:110-123  threshold-only branch, yaw_to_quat, final cat
//
// The third-party kernels build an N x N/64 suppression mask (313 MB at N = 50 k), copy it to
// the host and scan it serially.  Here the greedy scan stays on the device and only tests a
// candidate against boxes that were actually KEPT ("frontier" algorithm, DESIGN.md "K4"):
//
//   one CTA per (sweep, class) segment, candidates sorted by (score desc, index asc);
//   alive bitmap in shared memory; repeat
//     1. frontier  = the first F alive candidates in rank order
//     2. all pairs inside the frontier -> suppression bit-matrix (circle pre-test, then exact IoU
//        through a shared-memory work queue so the expensive routine runs with full warps)
//     3. one warp resolves the frontier greedily from the bit-matrix (<= F dependent steps)
//     4. every alive candidate behind the frontier is tested against the NEWLY kept boxes only
//        (circle pre-test -> work queue -> IoU), and cleared from the bitmap if suppressed
//   until num_post_nms boxes are kept or nothing is alive.
//
// The result is identical to the sequential greedy scan: a candidate enters a frontier only if
// no earlier kept box suppressed it, the frontier is resolved in rank order, and pruned pairs
// have IoU exactly 0 (iou.cuh padded_radius).  IoU evaluations drop from O(N * kept) on
// every candidate to (alive candidates near a kept box).
#include <cooperative_groups.h>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include <math_constants.h>

#include "common.cuh"
#include "iou.cuh"

namespace cg = cooperative_groups;

namespace rv3d {

constexpr int kNmsThreads = 512;
// frontier size: 512 boxes per round in hard mode (one 32 KB bit-matrix), 256 in weighted mode (two
// matrices, 64-byte records): fewer, fuller rounds cost fewer barriers and leader-only phases
template <bool kWeighted> struct Frontier { static constexpr int kF = kWeighted ? 256 : 512; };
constexpr int kMaxD = 16;               // max data columns of the weighted merge

// ------------------------------------------------------------------------------------------
// small kernels around the sort
// ------------------------------------------------------------------------------------------
__global__ void iota_kernel(uint32_t *v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// sorted position -> suppression record (+ merge row for the weighted mode); also seg_begin / seg_end from the
// sorted keys (both zero-initialised: empty segments stay [0, 0))
template <bool kWeighted>
__global__ void __launch_bounds__(256)
prepare_records_kernel(const uint32_t *__restrict__ order, const float *__restrict__ boxes, int n,
                       void *__restrict__ recs, float *__restrict__ data,
                       const unsigned long long *__restrict__ keys, int seg_shift, int *__restrict__ seg_begin,
                       int *__restrict__ seg_end) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (keys) {   // segment boundaries of the sorted keys, in the same pass (was a kernel of its own)
    const uint32_t sg = static_cast<uint32_t>(keys[i] >> seg_shift);
    if (i == 0 || static_cast<uint32_t>(keys[i - 1] >> seg_shift) != sg) seg_begin[sg] = i;
    if (i == n - 1 || static_cast<uint32_t>(keys[i + 1] >> seg_shift) != sg) seg_end[sg] = i + 1;
  }
  const float4 *src = reinterpret_cast<const float4 *>(boxes + static_cast<size_t>(order[i]) * 8);
  const float4 b0 = src[0], b1 = src[1];  // x y z l | w h yaw score
  if (!kWeighted) {
    // nms.py:33,40: boxes [x, y, l, w, -rad2deg(yaw)] with rad2deg evaluated in float32
    const float angle = -(b1.z * 57.29577951308232f);
    static_cast<HardRec *>(recs)[i] = make_hard_rec(b0.x, b0.y, b0.w, b1.x, angle, 0.01745329251);
  } else {
    // nms.py:87-100: [x - l/2, y - w/2, x + l/2, y + w/2, yaw]; merge row [x,y,z,l,w,h,sin,cos,score]
    static_cast<WRec *>(recs)[i] =
        make_w_rec(b0.x - b0.w / 2, b0.y - b1.x / 2, b0.x + b0.w / 2, b0.y + b1.x / 2, b1.z);
    float *d = data + static_cast<size_t>(i) * 9;
    d[0] = b0.x; d[1] = b0.y; d[2] = b0.z; d[3] = b0.w; d[4] = b1.x; d[5] = b1.y;
    d[6] = static_cast<float>(sin(static_cast<double>(b1.z)));
    d[7] = static_cast<float>(cos(static_cast<double>(b1.z)));
    d[8] = b1.w;
  }
}

// ------------------------------------------------------------------------------------------
// the suppression kernel
// ------------------------------------------------------------------------------------------
struct NmsArgs {
  const void *recs;        // HardRec / WRec, sorted order
  const int *seg_begin, *seg_end, *kept_base;
  int num_pre, num_post;
  float thr, mthr;
  int prune;               // 0: thresholds < 0 make even disjoint boxes interact -> test every pair
  int *kept_pos;           // [kept_base[s] + i] = position inside the segment
  int *kept_count;         // [s]
  // static candidate grid (scratch, per segment at 4*seg_begin / seg_begin)
  void *grid_entries;      // GridEntry[n_total]
  uint32_t *oversize;      // [n_total] candidates that are not in the grid
  int *firstsup;           // [n_total] weighted: first suppressor rank of a candidate in the current round
  // weighted merge
  const float *data;       // (n, D) rows in sorted order, score last
  int D;
  double *acc;             // [(kept_base[s] + i) * D + c]: c < D-1 weighted sums, c = D-1 weight sum
  int *merge_count;        // [kept_base[s] + i]
  unsigned long long *stats;
};

// ---- static dense grid over a segment's candidates --------------------------------------------
// Built once per segment by the leader CTA.  kG x kG cells cover mean +- 3.5 sigma of the candidate
// centres (cell >= half the mean padded radius); a candidate is registered ONCE, in the cell of its
// centre, coordinates clamped to the grid (clamping is monotone, so a clamped window still covers every
// clamped candidate: far outliers just land in a border cell and cost a wasted circle test).  Entries are
// counting-sorted by row-major cell index, so the cells cx0..cx1 of one grid row are ONE contiguous run.
// Candidates with a padded radius above r_cap = 2 x mean (or non-finite) go to an "oversize" list that
// every query scans.  A kept box (x, y, r) must look at centres within r + r_cap: rows cy0..cy1, one
// run each; no hashing, no duplicates, no per-entry cell checks.
constexpr int kG = 64;
constexpr int kCells = kG * kG;
constexpr float kPosCap = 1.0e6f;
constexpr int kQ2Cap = 8192;             // exact-IoU work queue (pairs that passed circle + bound)
constexpr int kWBuf = 64;                // per-warp hit buffer
static_assert(kCells % kNmsThreads == 0, "whole cells per scan lane");

struct __align__(16) GridEntry { float x, y, r; uint32_t idx; };

struct GridGeom {
  float x0, y0, inv_cell, r_cap;
  __device__ __forceinline__ int cell_x(float x) const {
    return static_cast<int>(fminf(fmaxf((x - x0) * inv_cell, 0.f), static_cast<float>(kG - 1)));
  }
  __device__ __forceinline__ int cell_y(float y) const {
    return static_cast<int>(fminf(fmaxf((y - y0) * inv_cell, 0.f), static_cast<float>(kG - 1)));
  }
  __device__ __forceinline__ bool gridded(float x, float y, float r) const {
    return (r <= r_cap) && (fabsf(x) <= kPosCap) && (fabsf(y) <= kPosCap);   // false for NaN
  }
};

// ---- cheap, safe upper bound on the IoU (separating axes + projected overlap) -----------------
// The intersection lies inside box A and inside B's bounding box in A's frame, so its area is at
// most (overlap of the projections on A's two axes); same in B's frame.  Used only to SKIP the
// exact routine when even the bound (inflated by 2 % + 1e-5) cannot exceed the threshold.
struct Obb { float x, y, w, h, c, s; };  // centre, full extents, axis u = (c, s), v = (-s, c)
__device__ __forceinline__ Obb obb_of(const HardRec &r) { return Obb{r.x, r.y, r.w, r.h, 2.f * r.c2, -2.f * r.s2}; }
__device__ __forceinline__ Obb obb_of(const WRec &r) {
  return Obb{(r.x1 + r.x2) * 0.5f, (r.y1 + r.y2) * 0.5f, r.x2 - r.x1, r.y2 - r.y1, r.ca, r.sa};
}
__device__ __forceinline__ float overlap_in_frame(const Obb &a, const Obb &b) {
  const float dx = b.x - a.x, dy = b.y - a.y;
  const float du = dx * a.c + dy * a.s, dv = -dx * a.s + dy * a.c;
  const float cd = fabsf(a.c * b.c + a.s * b.s), sd = fabsf(a.c * b.s - a.s * b.c);
  const float bw = fabsf(b.w) * 0.5f, bh = fabsf(b.h) * 0.5f, aw = fabsf(a.w) * 0.5f, ah = fabsf(a.h) * 0.5f;
  const float eu = cd * bw + sd * bh, ev = sd * bw + cd * bh;
  const float ou = fminf(aw, du + eu) - fmaxf(-aw, du - eu);
  const float ov = fminf(ah, dv + ev) - fmaxf(-ah, dv - ev);
  return fmaxf(ou, 0.f) * fmaxf(ov, 0.f);
}
template <typename Rec>
__device__ __forceinline__ bool iou_may_exceed(const Rec &ra, const Rec &rb, float thr) {
  const Obb a = obb_of(ra), b = obb_of(rb);
  const float inter = fminf(overlap_in_frame(a, b), overlap_in_frame(b, a));
  const float uni = fabsf(a.w * a.h) + fabsf(b.w * b.h) - inter;
  if (!(uni > 0.f) || !(inter == inter)) return true;  // degenerate / NaN: let the exact routine decide
  return (inter / uni) * 1.02f + 1e-5f >= thr;
}

// ---- fast approximate IoU (float32 Sutherland-Hodgman clip in box A's frame) --------------------
// NMS never needs the IoU value, only the comparisons `iou > thr` (and `> merge_thr`).  The clip below
// is accurate to ~1e-5, the exact routines deviate from true geometry by <= ~1e-4 for sane boxes, so a
// pair whose approximate IoU is outside a +-(2 % + 1e-3) band around a threshold is decided without the
// bit-exact routine; pairs inside the band, and degenerate boxes, still run it.  ~98 % of the pairs that
// survive the circle + bound filters are decided here at ~1/10 of the instructions.
__device__ __forceinline__ float approx_iou(const Obb &A, const Obb &B) {
  const float aw = fabsf(A.w) * 0.5f, ah = fabsf(A.h) * 0.5f, bw = fabsf(B.w) * 0.5f, bh = fabsf(B.h) * 0.5f;
  const float dx = B.x - A.x, dy = B.y - A.y;
  const float cx = dx * A.c + dy * A.s, cy = -dx * A.s + dy * A.c;          // B's centre in A's frame
  const float cd = A.c * B.c + A.s * B.s, sd = A.c * B.s - A.s * B.c;       // B's axis in A's frame
  const float ux = cd * bw, uy = sd * bw, vx = -sd * bh, vy = cd * bh;
  float px[10], py[10], qx[10], qy[10];
  px[0] = cx + ux + vx; py[0] = cy + uy + vy;
  px[1] = cx - ux + vx; py[1] = cy - uy + vy;
  px[2] = cx - ux - vx; py[2] = cy - uy - vy;
  px[3] = cx + ux - vx; py[3] = cy + uy - vy;
  int n = 4;
#pragma unroll
  for (int plane = 0; plane < 4; ++plane) {
    // planes: x <= aw, -x <= aw, y <= ah, -y <= ah  (coordinate c, sign sg, limit lim)
    const float sg = (plane & 1) ? -1.f : 1.f;
    const float lim = (plane < 2) ? aw : ah;
    int m = 0;
    float prx = px[n - 1], pry = py[n - 1];
    float prd = sg * ((plane < 2) ? prx : pry) - lim;     // signed distance, inside when <= 0
    for (int i = 0; i < n; ++i) {
      const float cxp = px[i], cyp = py[i];
      const float cd2 = sg * ((plane < 2) ? cxp : cyp) - lim;
      if ((cd2 <= 0.f) != (prd <= 0.f)) {
        const float t = __fdividef(prd, prd - cd2);
        qx[m] = prx + t * (cxp - prx); qy[m] = pry + t * (cyp - pry); ++m;
      }
      if (cd2 <= 0.f) { qx[m] = cxp; qy[m] = cyp; ++m; }
      prx = cxp; pry = cyp; prd = cd2;
    }
    if (m < 3) return 0.f;
    n = m;
    for (int i = 0; i < n; ++i) { px[i] = qx[i]; py[i] = qy[i]; }
  }
  float area2 = 0.f;
  for (int i = 0; i < n; ++i) {
    const int k = (i + 1 == n) ? 0 : i + 1;
    area2 += px[i] * py[k] - px[k] * py[i];
  }
  const float inter = 0.5f * fabsf(area2);
  const float uni = 4.f * (aw * ah + bw * bh) - inter;
  return uni > 0.f ? inter / uni : CUDART_NAN_F;
}

// -1: certainly below / equal, +1: certainly above, 0: too close to call (or unusual boxes) -> exact routine
__device__ __forceinline__ bool obb_sane(const Obb &o) {
  const float w = fabsf(o.w), h = fabsf(o.h);
  return fminf(w, h) >= 0.05f && fmaxf(w, h) <= 1.0e4f && fabsf(o.x) <= kPosCap && fabsf(o.y) <= kPosCap;
}
__device__ __forceinline__ int decide_vs(float approx, float thr) {
  if (!(approx == approx)) return 0;
  const float m = 0.02f * fabsf(thr) + 1e-3f;
  return approx > thr + m ? 1 : (approx < thr - m ? -1 : 0);
}

template <typename Rec, bool kWeighted, int kF>
__host__ __device__ inline size_t nms_smem_bytes(int nwords) {
  constexpr int kFW = kF / 32;
  size_t b = 0;
  b += align_up_c(sizeof(uint32_t) * nwords, 16);                  // alive
  b += sizeof(Rec) * kF * 2;                                       // frec, krec
  b += sizeof(float) * kF * 6;                                     // fx, fy, fr, kx, ky, kr
  b += sizeof(int) * kF;                                           // front_pos
  b += sizeof(uint32_t) * kF * kFW * (kWeighted ? 2 : 1);          // sup (+ mrg)
  b += sizeof(int) * (kCells + 1);                                 // cell_start (the build's cursors alias queue2 + wbuf)
  b += sizeof(uint32_t) * kQ2Cap;                                  // queue2
  b += sizeof(uint32_t) * (kNmsThreads / 32) * kWBuf;              // per-warp hit buffers
  b += align_up_c(sizeof(uint16_t) * kF + sizeof(int16_t) * kF, 16);  // keptf, keptrank
  if (kWeighted) b += sizeof(uint32_t) * kQ2Cap + sizeof(int) * kF;  // qflag, killer
  return b;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int &total) {
  // kNmsThreads threads; s_warp has kNmsThreads/32 + 1 ints
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();  // protect s_warp from the previous use
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = lane < kNmsThreads / 32 ? s_warp[lane] : 0;
    int inc2 = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc2, o);
      if (lane >= o) inc2 += t;
    }
    if (lane < kNmsThreads / 32) s_warp[lane] = inc2 - w;
    if (lane == 31) s_warp[kNmsThreads / 32] = inc2;
  }
  __syncthreads();
  total = s_warp[kNmsThreads / 32];
  return s_warp[wid] + incl - v;
}

template <typename Rec, bool kWeighted, int kF>
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_segment_kernel(NmsArgs a) {
  constexpr int kFW = kF / 32;   // words per bit-matrix row
  static_assert(kF <= 1024 && kFW <= 32 && kNmsThreads % kFW == 0, "frontier size");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_warp[kNmsThreads / 32 + 1];
  __shared__ int s_qn, s_qvalid, s_nk, s_nos, s_overflow;
  __shared__ int s_next;       // leader: next kept box to hand out in the kill scan (dynamic load balance)
  __shared__ int s_round[8];   // leader -> cluster: {nf, nk, cursor_word, -, n oversize}
  __shared__ GridGeom s_geom;  // leader -> cluster
  __shared__ float s_red[kNmsThreads / 32 * 6];
  __shared__ uint32_t s_haspred[kFW], s_removed[kFW];

  // A cluster of P CTAs works on one segment.  CTA 0 (the leader) owns the alive bitmap and runs the
  // serial-ish phases (frontier, pairs, greedy); the kill scan and the exact IoU -- 80 % of the work --
  // are split over all P CTAs, which read the leader's state and clear bits in its bitmap through
  // distributed shared memory.  P = 1 degenerates to a plain CTA per segment.
  cg::cluster_group cluster = cg::this_cluster();
  const int P = static_cast<int>(cluster.num_blocks());
  const int crank = static_cast<int>(cluster.block_rank());
  const bool leader = crank == 0;
  const int seg = blockIdx.x / P;
  const int beg = a.seg_begin[seg];
  const int n = min(a.seg_end[seg] - beg, a.num_pre);  // top num_pre_nms by score (nms.py:29-32)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (n <= 0) {   // uniform over the cluster
    if (tid == 0 && leader) a.kept_count[seg] = 0;
    return;
  }
  const int nwords = (n + 31) >> 5;
  const Rec *recs = static_cast<const Rec *>(a.recs) + beg;
  GridEntry *entries = static_cast<GridEntry *>(a.grid_entries) + beg;
  uint32_t *os_list = a.oversize + beg;
  int *firstsup = kWeighted ? a.firstsup + beg : nullptr;
  const int kbase = a.kept_base[seg];
  const float thr_any = kWeighted ? fminf(a.thr, a.mthr) : a.thr;  // smallest IoU that matters
  const bool prune = a.prune != 0;

  // ---- carve shared memory
  unsigned char *p = smem_raw;
  uint32_t *alive = reinterpret_cast<uint32_t *>(p); p += align_up_c(sizeof(uint32_t) * nwords, 16);
  Rec *frec = reinterpret_cast<Rec *>(p); p += sizeof(Rec) * kF;
  float *fx = reinterpret_cast<float *>(p); p += sizeof(float) * kF;
  float *fy = reinterpret_cast<float *>(p); p += sizeof(float) * kF;
  float *fr = reinterpret_cast<float *>(p); p += sizeof(float) * kF;
  Rec *krec = reinterpret_cast<Rec *>(p); p += sizeof(Rec) * kF;   // this round's kept boxes, rank order
  float *kx = reinterpret_cast<float *>(p); p += sizeof(float) * kF;
  float *ky = reinterpret_cast<float *>(p); p += sizeof(float) * kF;
  float *kr = reinterpret_cast<float *>(p); p += sizeof(float) * kF;
  int *front_pos = reinterpret_cast<int *>(p); p += sizeof(int) * kF;
  uint32_t *sup = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * kF * kFW;
  uint32_t *mrg = sup;
  if (kWeighted) { mrg = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * kF * kFW; }
  int *cell_start = reinterpret_cast<int *>(p); p += sizeof(int) * (kCells + 1);
  uint32_t *queue2 = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * kQ2Cap;
  int *cell_cursor = reinterpret_cast<int *>(queue2);   // kCells ints, build phase only (kQ2Cap >= kCells)
  uint32_t *wbuf = reinterpret_cast<uint32_t *>(p) + wid * kWBuf; p += sizeof(uint32_t) * (kNmsThreads / 32) * kWBuf;
  uint16_t *keptf = reinterpret_cast<uint16_t *>(p);
  int16_t *keptrank = reinterpret_cast<int16_t *>(keptf + kF);
  p += align_up_c(sizeof(uint16_t) * kF + sizeof(int16_t) * kF, 16);
  uint32_t *qflag = kWeighted ? reinterpret_cast<uint32_t *>(p) : nullptr;   // per queued pair: bit0 iou > thr, bit1 iou > merge_thr
  int *killer = kWeighted ? reinterpret_cast<int *>(qflag + kQ2Cap) : nullptr;  // frontier box -> rank of its first suppressor

  unsigned long long st_iou = 0, st_circle = 0, st_hit = 0, st_approx = 0;   // exact IoUs, circle tests, pairs above thr, approximate IoUs
  // per-phase cycle counters (thread 0, only when stats are requested):
  // [0] grid build + frontier gather, [1] frontier load, [2] frontier pairs, [3] greedy, [4] kill scan, [5] exact IoU
  long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // [6] publish (+ weighted frontier merges), [7] cluster sync A + helper copies
  long long t_mark = clock64();
  auto lap = [&](int k) {
    if (a.stats && tid == 0) { const long long t = clock64(); ph[k] += t - t_mark; t_mark = t; }
  };

  // The three passes of the grid build read (centre, radius) of every candidate; each is a chain of dependent
  // global loads unless several are in flight per thread: 4 records are loaded before the first is used.
  auto for_each_centre = [&](auto &&fn) {
    constexpr int kU = 4;
    for (int i0 = tid; i0 < n; i0 += kU * kNmsThreads) {
      float x[kU], y[kU], r[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = i0 + u * kNmsThreads;
        if (i < n) { x[u] = rec_cx(recs[i]); y[u] = rec_cy(recs[i]); r[u] = recs[i].r; }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int i = i0 + u * kNmsThreads;
        if (i < n) fn(i, x[u], y[u], r[u]);
      }
    }
  };

  // ======================= 0. alive bitmap + static candidate grid (leader) =======================
  for (int w = tid; w < nwords; w += kNmsThreads)
    alive[w] = (w == nwords - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
  if (tid == 0) { s_nos = 0; s_overflow = 0; s_qn = 0; }
  if (leader) {
    for (int c = tid; c < kCells; c += kNmsThreads) cell_cursor[c] = 0;
    {
      // first and second moments of the sane candidates -> grid origin / cell size / radius cap
      float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // sum x, y, x^2, y^2, r, count
      for_each_centre([&](int, float x, float y, float r) {
        if (r > 0.f && r < 1.0e4f && fabsf(x) <= kPosCap && fabsf(y) <= kPosCap) {
          acc[0] += x; acc[1] += y; acc[2] += x * x; acc[3] += y * y; acc[4] += r; acc[5] += 1.f;
        }
      });
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        for (int o = 16; o; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
        if (lane == 0) s_red[k * (kNmsThreads / 32) + wid] = acc[k];
      }
      __syncthreads();
      if (tid == 0) {
        float t[6];
        for (int k = 0; k < 6; ++k) { t[k] = 0.f; for (int w = 0; w < kNmsThreads / 32; ++w) t[k] += s_red[k * (kNmsThreads / 32) + w]; }
        const float cnt = fmaxf(t[5], 1.f);
        const float mx = t[0] / cnt, my = t[1] / cnt, mr = t[5] > 0.f ? t[4] / cnt : 1.f;
        const float sx = sqrtf(fmaxf(t[2] / cnt - mx * mx, 0.f)), sy = sqrtf(fmaxf(t[3] / cnt - my * my, 0.f));
        const float cell = fminf(fmaxf(fmaxf(7.f * fmaxf(sx, sy) / kG, 0.5f * mr), 1e-3f), 1.0e5f);
        s_geom.inv_cell = 1.0f / cell;
        s_geom.x0 = mx - 0.5f * kG * cell;
        s_geom.y0 = my - 0.5f * kG * cell;
        s_geom.r_cap = 2.0f * mr;
      }
      __syncthreads();
    }
    const GridGeom g = s_geom;
    if (prune) {
      for_each_centre([&](int i, float x, float y, float r) {   // count
        if (g.gridded(x, y, r)) atomicAdd(&cell_cursor[g.cell_y(y) * kG + g.cell_x(x)], 1);
        else os_list[atomicAdd(&s_nos, 1)] = static_cast<uint32_t>(i);
      });
    } else {
      for (int i = tid; i < n; i += kNmsThreads) os_list[i] = static_cast<uint32_t>(i);
      if (tid == 0) s_nos = n;
    }
    __syncthreads();
    {
      constexpr int kPer = kCells / kNmsThreads;
      int cnt[kPer], sum = 0;
#pragma unroll
      for (int k = 0; k < kPer; ++k) { cnt[k] = cell_cursor[tid * kPer + k]; sum += cnt[k]; }
      int total;
      int off = block_exclusive_scan(sum, s_warp, total);
#pragma unroll
      for (int k = 0; k < kPer; ++k) { cell_start[tid * kPer + k] = off; cell_cursor[tid * kPer + k] = off; off += cnt[k]; }
      if (tid == 0) cell_start[kCells] = total;
    }
    __syncthreads();
    if (prune) {
      for_each_centre([&](int i, float x, float y, float r) {   // fill
        if (!g.gridded(x, y, r)) return;
        entries[atomicAdd(&cell_cursor[g.cell_y(y) * kG + g.cell_x(x)], 1)] = GridEntry{x, y, r, static_cast<uint32_t>(i)};
      });
    }
    if (kWeighted)
      for (int i = tid; i < n; i += kNmsThreads) firstsup[i] = 0x7fffffff;
    if (tid == 0) s_round[4] = s_nos;
    __threadfence();
  }
  cluster.sync();   // grid entries (global) + cell offsets (leader's shared memory) are ready
  uint32_t *lead_alive = cluster.map_shared_rank(alive, 0);
  const int *lead_round = cluster.map_shared_rank(s_round, 0);
  if (!leader) {
    const int *lead_cs = cluster.map_shared_rank(cell_start, 0);
    for (int c = tid; c <= kCells; c += kNmsThreads) cell_start[c] = lead_cs[c];
  }
  const int nos = lead_round[4];
  const GridGeom geom = *cluster.map_shared_rank(&s_geom, 0);
  __syncthreads();

  int kept_total = 0;
  int cursor_word = 0;
  int rounds = 0;

  // record one evaluated pair in the frontier bit-matrices
  auto mark_pair = [&](int i, int j, float iou) {
    if (iou > a.thr) { ++st_hit; atomicOr(&sup[i * kFW + (j >> 5)], 1u << (j & 31)); }
    if (kWeighted && iou > a.mthr) atomicOr(&mrg[i * kFW + (j >> 5)], 1u << (j & 31));
  };
  auto accumulate = [&](int slot, int row_in_seg) {  // merge candidate `row_in_seg` into kept slot
    const float *row = a.data + static_cast<size_t>(beg + row_in_seg) * a.D;
    float v[kMaxD];
#pragma unroll
    for (int c = 0; c < kMaxD; ++c) v[c] = c < a.D ? row[c] : 0.f;   // independent loads first: one latency, not D
    const double sj = row[a.D - 1];
    double *acc = a.acc + static_cast<size_t>(kbase + slot) * a.D;
#pragma unroll
    for (int c = 0; c < kMaxD; ++c)
      if (c < a.D - 1) atomicAdd(acc + c, sj * static_cast<double>(v[c]));
    atomicAdd(acc + a.D - 1, sj);
    atomicAdd(a.merge_count + kbase + slot, 1);
  };
  // a candidate is suppressed: clear it in the leader's bitmap (the truth) and in the local snapshot
  auto kill = [&](int j) {
    const uint32_t m = ~(1u << (j & 31));
    atomicAnd(&lead_alive[j >> 5], m);
    if (!leader) atomicAnd(&alive[j >> 5], m);
  };
  // warp-converged append of `item` (valid where `pred`) to the exact-IoU queue; returns false for the
  // lanes whose item did not fit (caller handles them in place)
  auto q2_push = [&](bool pred, uint32_t item) -> bool {
    const uint32_t m = __ballot_sync(0xffffffffu, pred);
    if (!m) return true;
    int base = 0;
    if (lane == 0) base = atomicAdd(&s_qn, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!pred) return true;
    const int slot = base + __popc(m & ((1u << lane) - 1u));
    if (slot >= kQ2Cap) return false;
    queue2[slot] = item;
    return true;
  };

  // bound -> approximate -> exact for ONE pair, in place (overflow paths; divergent, so only a fallback)
  auto classify_pair = [&](const Rec &ra, const Rec &rb, bool &above, bool &above_m) {
    int d1 = 0, d2 = 0;
    if (prune) {
      if (!iou_may_exceed(ra, rb, thr_any)) { above = false; above_m = false; return; }
      const Obb oa = obb_of(ra), ob = obb_of(rb);
      if (obb_sane(oa) && obb_sane(ob)) {
        const float ap = approx_iou(oa, ob);
        ++st_approx;
        d1 = decide_vs(ap, a.thr);
        d2 = kWeighted ? decide_vs(ap, a.mthr) : 1;
      }
    }
    if (d1 != 0 && d2 != 0) { above = d1 > 0; above_m = kWeighted && d2 > 0; return; }
    const float iou = pair_iou(ra, rb);
    ++st_iou;
    above = iou > a.thr;
    above_m = kWeighted && iou > a.mthr;
  };

  // Two-stage evaluation of a queue of pairs, warp-converged: every lane classifies its pair with the
  // approximate IoU; the few undecided pairs are compacted per warp (wbuf) so that the exact routine
  // always runs on (nearly) full warps.  get(q, ra, rb) loads the pair, emit(q, above_thr, above_mthr)
  // consumes the two comparisons.
  auto eval_queue = [&](int qn, auto &&get, auto &&emit) {
    int nbuf = 0;   // warp-uniform
    auto exact32 = [&](int count) {
      __syncwarp();
      if (lane < count) {
        const int q = static_cast<int>(wbuf[lane]);
        Rec ra, rb;
        get(q, ra, rb);
        const float iou = pair_iou(ra, rb);
        ++st_iou;
        emit(q, iou > a.thr, kWeighted && iou > a.mthr);
      }
      __syncwarp();
    };
    const int qn_pad = (qn + 31) & ~31;
    for (int q0 = wid * 32; q0 < qn_pad; q0 += kNmsThreads) {   // each warp owns 32 consecutive entries per pass
      __syncwarp();
      const int q = q0 + lane;
      bool undecided = false;
      if (q < qn) {
        Rec ra, rb;
        get(q, ra, rb);
        const Obb oa = obb_of(ra), ob = obb_of(rb);
        int d1 = 0, d2 = 0;
        if (prune && !iou_may_exceed(ra, rb, thr_any)) {
          d1 = -1; d2 = -1;                       // even the upper bound stays below both thresholds
        } else if (prune && obb_sane(oa) && obb_sane(ob)) {
          const float ap = approx_iou(oa, ob);
          ++st_approx;
          d1 = decide_vs(ap, a.thr);
          d2 = kWeighted ? decide_vs(ap, a.mthr) : 1;
        }
        if (d1 != 0 && d2 != 0) emit(q, d1 > 0, kWeighted && d2 > 0);
        else undecided = true;
      }
      const uint32_t m = __ballot_sync(0xffffffffu, undecided);
      if (m) {
        if (undecided) wbuf[nbuf + __popc(m & ((1u << lane) - 1u))] = static_cast<uint32_t>(q);
        nbuf += __popc(m);
        if (nbuf >= 32) {
          exact32(32);
          if (lane < nbuf - 32) wbuf[lane] = wbuf[32 + lane];
          nbuf -= 32;
          __syncwarp();
        }
      }
    }
    if (nbuf > 0) exact32(nbuf);
  };

  while (true) {
    int nf = 0, nk = 0;
    if (leader) {
      // ================= 1. frontier: first kF alive candidates =================
      for (int base = cursor_word; base < nwords && nf < kF; base += kNmsThreads) {
        const int wi = base + tid;
        const uint32_t word = wi < nwords ? alive[wi] : 0u;
        int total;
        const int off = block_exclusive_scan(__popc(word), s_warp, total);
        if (word && nf + off < kF) {
          uint32_t m = word;
          int r = nf + off;
          while (m && r < kF) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            front_pos[r++] = (wi << 5) + bit;
          }
        }
        nf = min(kF, nf + total);
      }
      __syncthreads();
      lap(0);
      if (nf > 0) {
        ++rounds;
        // the frontier is decided this round: clear its bits, reset the per-round state
        if (tid < nf) {
          const int pos = front_pos[tid];
          atomicAnd(&alive[pos >> 5], ~(1u << (pos & 31)));
          keptrank[tid] = -1;
        }
        for (int i = tid; i < kF * kFW; i += kNmsThreads) {
          sup[i] = 0u;
          if (kWeighted) mrg[i] = 0u;
        }
        if (tid < kFW) { s_haspred[tid] = 0u; s_removed[tid] = 0u; }
        cursor_word = front_pos[0] >> 5;
      }
      if (tid == 0) { s_round[0] = nf; s_round[2] = cursor_word; }
    }
    cluster.sync();   // (A0) the frontier (positions) is published, the leader's bit-matrices are zeroed
    nf = lead_round[0];
    cursor_word = lead_round[2];
    if (nf == 0) break;

    // ================= 2. interacting pairs inside the frontier (whole cluster) =================
    // Every CTA loads the frontier's records; the rows of the upper-triangular pair matrix are dealt
    // round-robin to the warps of the cluster (row i costs nf - i tests, so interleaving balances):
    // warp per row, lanes over the columns j > i, circle test -> work queue; then bound / approximate /
    // exact IoU on dense lanes, results OR-ed into the leader's bit-matrices through DSMEM.
    {
      const int *lead_front = cluster.map_shared_rank(front_pos, 0);
      if (tid < nf) {
        const int pos = lead_front[tid];
        const Rec r = recs[pos];
        frec[tid] = r;
        fx[tid] = rec_cx(r); fy[tid] = rec_cy(r); fr[tid] = r.r;
      }
      if (tid == 0) { s_qn = 0; s_qvalid = kQ2Cap; }
      __syncthreads();
      if (leader) lap(1);
      uint32_t *lead_sup = cluster.map_shared_rank(sup, 0);
      uint32_t *lead_mrg = cluster.map_shared_rank(mrg, 0);
      auto mark_remote = [&](int i, int j, bool above, bool above_m) {
        if (above) { ++st_hit; atomicOr(&lead_sup[i * kFW + (j >> 5)], 1u << (j & 31)); }
        if (kWeighted && above_m) atomicOr(&lead_mrg[i * kFW + (j >> 5)], 1u << (j & 31));
      };
      // One queue reservation per ROW (a single-address shared atomic per 32-column batch serialises the
      // CTA on dense frontiers): lane b keeps the hit mask of the row's b-th batch, the warp reserves the
      // row's total once and every lane writes out its own batch.  A row that does not fit waits for the
      // next drain: reservations are handed out in order, so everything before the first refused one
      // (s_qvalid) is written and everything after it is refused too.
      static_assert(kF <= 1024, "a row's batches must fit one mask per lane");
      constexpr int kWarps = kNmsThreads / 32;
      int i = crank * kWarps + wid;
      for (;;) {
        while (i < nf - 1) {
          const float xi = fx[i], yi = fy[i], ri = fr[i];
          uint32_t mymask = 0;
          for (int jb = i + 1, b = 0; jb < nf; jb += 32, ++b) {
            const int j = jb + lane;
            bool hit = false;
            if (j < nf) {
              ++st_circle;
              hit = true;
              if (prune) {
                const float dx = xi - fx[j], dy = yi - fy[j], rr = ri + fr[j];
                hit = dx * dx + dy * dy <= rr * rr;   // the IoU bound runs later, on dense lanes (eval_queue)
              }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (lane == b) mymask = m;
          }
          const int mine = __popc(mymask);
          int incl = mine;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
          }
          const int total = __shfl_sync(0xffffffffu, incl, 31);
          if (total) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_qn, total);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + total > kQ2Cap) {
              if (lane == 0) atomicMin(&s_qvalid, base);
              break;   // this row is redone after the drain
            }
            int slot = base + incl - mine;
            const int j0 = i + 1 + 32 * lane;
            for (uint32_t m = mymask; m; m &= m - 1)
              queue2[slot++] = static_cast<uint32_t>((i << 10) | (j0 + __ffs(m) - 1));
          }
          i += P * kWarps;
        }
        const int more = __syncthreads_or(i < nf - 1);
        eval_queue(min(s_qn, s_qvalid),
                   [&](int q, Rec &ra, Rec &rb) { ra = frec[queue2[q] >> 10]; rb = frec[queue2[q] & 1023]; },
                   [&](int q, bool above, bool above_m) {
                     mark_remote(static_cast<int>(queue2[q] >> 10), static_cast<int>(queue2[q] & 1023), above, above_m);
                   });
        if (!more) break;
        __syncthreads();
        if (tid == 0) { s_qn = 0; s_qvalid = kQ2Cap; }
        __syncthreads();
      }
    }
    cluster.sync();   // (A1) every CTA's pair results are in the leader's bit-matrices
    if (leader) {
      lap(2);
      {
        // ================= 3. greedy resolution of the frontier =================
        // Boxes that no earlier frontier box can suppress (empty column in `sup`) are kept outright, in
        // parallel; only the rest needs the dependent scan, done by one warp.
        {
          const int w = tid & (kFW - 1);
          uint32_t colbits = 0u;
          for (int i = tid / kFW; i < nf; i += kNmsThreads / kFW) colbits |= sup[i * kFW + w];
          if (colbits) atomicOr(&s_haspred[w], colbits);
        }
        __syncthreads();
        {
          const int w = tid & (kFW - 1);
          uint32_t rm = 0u;
          for (int i = tid / kFW; i < nf; i += kNmsThreads / kFW)
            if (!((s_haspred[i >> 5] >> (i & 31)) & 1u)) rm |= sup[i * kFW + w];   // rows of the free boxes
          if (rm) atomicOr(&s_removed[w], rm);
        }
        __syncthreads();
        if (tid < 32) {
          // lane w < kFW owns word w.  valid = bits < nf; free = valid & ~haspred (kept for sure)
          uint32_t valid = 0u, removed = 0u, pending = 0u, kept = 0u;
          if (lane < kFW) {
            const int lo = lane << 5;
            valid = (nf >= lo + 32) ? 0xffffffffu : (nf <= lo ? 0u : ((1u << (nf - lo)) - 1u));
            removed = s_removed[lane];
            pending = valid & s_haspred[lane];       // must be visited in rank order
            kept = valid & ~s_haspred[lane];
          }
          while (true) {
            const uint32_t cand = pending & ~removed;
            const uint32_t have = __ballot_sync(0xffffffffu, cand != 0u);
            if (!have) break;
            const int wsel = __ffs(have) - 1;
            const uint32_t cw = __shfl_sync(0xffffffffu, cand, wsel);
            const int bit = __ffs(cw) - 1;
            const int i = (wsel << 5) + bit;
            if (lane == wsel) kept |= 1u << bit;
            if (lane < kFW) {
              // everything up to and including i is decided now
              const int lo = lane << 5;
              if (i >= lo + 31) pending = 0u;
              else if (i >= lo) pending &= ~((2u << (i - lo)) - 1u);
              removed |= sup[i * kFW + lane];
            }
          }
          // kept boxes in rank order, truncated to the room left under num_post_nms
          const int room = a.num_post - kept_total;
          const int cnt = lane < kFW ? __popc(kept) : 0;
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < kFW; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
          }
          int rank = incl - cnt;
          if (lane < kFW) {
            uint32_t m = kept;
            while (m && rank < room) {
              const int bit = __ffs(m) - 1;
              m &= m - 1;
              const int i = (lane << 5) + bit;
              keptf[rank] = static_cast<uint16_t>(i);
              keptrank[i] = static_cast<int16_t>(rank);
              ++rank;
            }
          }
          const int total = __shfl_sync(0xffffffffu, incl, kFW - 1);
          if (lane == 0) s_nk = min(total, room);
        }
        __syncthreads();
        nk = s_nk;
        lap(3);

        // ================= 4. publish the newly kept boxes =================
        if (tid < nk) {
          const int fi = keptf[tid];
          a.kept_pos[kbase + kept_total + tid] = front_pos[fi];
          krec[tid] = frec[fi];
          kx[tid] = fx[fi]; ky[tid] = fy[fi]; kr[tid] = fr[fi];
        }
        if (kWeighted) {
          // merge sets inside the frontier: candidate j joins every kept i < j with iou > merge_thr
          // that comes no later than its first suppressor; kept boxes join themselves.  The (j, slot)
          // pairs are queued first and accumulated afterwards, one pair per lane: done in place, each
          // lane's global loads + fp64 atomics would run alone (lanes reach their merges at different t).
          // Work is proportional to the set bits of the kept rows: (1) first suppressor of every frontier
          // box = min rank over the kept rows that contain it; (2) each kept row queues itself and the
          // boxes of its merge row that it reaches no later than their first suppressor.
          if (tid == 0) s_qn = 0;
          if (tid < nf) killer[tid] = 0x7fffffff;
          __syncthreads();
          if (tid < nk) {
            const int i = keptf[tid];
            for (int w = 0; w < kFW; ++w) {
              uint32_t bits = sup[i * kFW + w];
              while (bits) { const int j = (w << 5) + __ffs(bits) - 1; bits &= bits - 1; atomicMin(&killer[j], tid); }
            }
          }
          __syncthreads();
          if (tid < nk) {
            const int i = keptf[tid];
            auto push = [&](int j) {
              const int slot = atomicAdd(&s_qn, 1);
              if (slot < kQ2Cap) queue2[slot] = (static_cast<uint32_t>(j) << 16) | static_cast<uint32_t>(tid);
              else accumulate(kept_total + tid, front_pos[j]);   // (cannot happen: <= kF * kF / 2 only if every pair merges)
            };
            push(i);                                             // a kept box joins its own merge set
            for (int w = 0; w < kFW; ++w) {
              uint32_t bits = mrg[i * kFW + w];
              while (bits) { const int j = (w << 5) + __ffs(bits) - 1; bits &= bits - 1; if (tid <= killer[j]) push(j); }
            }
          }
          __syncthreads();
          const int nq = min(s_qn, kQ2Cap);
          for (int q = tid; q < nq; q += kNmsThreads)
            accumulate(kept_total + static_cast<int>(queue2[q] & 0xffffu), front_pos[queue2[q] >> 16]);
        }
      }
      __syncthreads();
      lap(6);
      if (tid == 0) { s_round[0] = nf; s_round[1] = nk; s_qn = 0; s_next = 0; }
      if (tid == 1) s_round[2] = cursor_word;
    }
    cluster.sync();   // (A) the leader's round state, kept boxes and bitmap are final
    nf = lead_round[0];
    nk = lead_round[1];
    cursor_word = lead_round[2];
    if (nf == 0) break;
    const int slot0 = kept_total;
    kept_total += nk;
    // nms.py:53-56: only the first num_post_nms kept survive, so the scan can stop there.  The
    // weighted mode still owes the last kept boxes their merge sets from the candidates behind
    // the frontier, so it runs one more kill phase before leaving.
    const bool last_round = kept_total >= a.num_post;
    if (last_round && !kWeighted) break;
    if (!leader) {
      // helpers: this round's kept boxes and a snapshot of the alive bitmap, through DSMEM
      const Rec *lrec = cluster.map_shared_rank(krec, 0);
      const float *lx = cluster.map_shared_rank(kx, 0), *ly = cluster.map_shared_rank(ky, 0), *lr = cluster.map_shared_rank(kr, 0);
      if (tid < nk) { krec[tid] = lrec[tid]; kx[tid] = lx[tid]; ky[tid] = ly[tid]; kr[tid] = lr[tid]; }
      for (int w = tid; w < nwords; w += kNmsThreads) alive[w] = lead_alive[w];   // all words: decided bits too
      if (tid == 0) s_qn = 0;
    }
    __syncthreads();
    if (leader) lap(7);

    // ================= 5. kill scan: one warp per newly kept box over the static grid ==============
    // Lanes stride over the contiguous entries of each cell the kept box's circle touches (+ the
    // oversize list): alive test, circle test, hits compacted per warp so the IoU bound runs on full
    // warps; pairs that pass go to the exact-IoU queue.
    // Kept boxes are handed out one at a time from a counter in the leader's shared memory: a box in a dense
    // cell costs 10x one at the periphery, a static split leaves most warps idle.
    int *lead_next = cluster.map_shared_rank(&s_next, 0);
    while (true) {
      int t = 0;
      if (lane == 0) t = atomicAdd(lead_next, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= nk) break;
      const Rec &rk = krec[t];
      const float qx = kx[t], qy = ky[t], qr = kr[t];
      int nbuf = 0;  // warp-uniform count of buffered hits
      auto flush = [&](int count) {   // run the bound on `count` (<= 32) buffered hits, lanes < count
        __syncwarp();
        bool pass = false;
        uint32_t j = 0;
        if (lane < count) {
          j = wbuf[lane];
          pass = !prune || iou_may_exceed(rk, recs[j], thr_any);
        }
        if (!q2_push(pass, (j << 10) | static_cast<uint32_t>(t))) {
          if (kWeighted) {
            s_overflow = 1;   // redo this round's kill phase with the exact serial fallback
          } else {            // hard mode only needs ANY suppressor: evaluate in place
            bool above, above_m;
            classify_pair(rk, recs[j], above, above_m);
            if (above) { ++st_hit; kill(j); }
          }
        }
        __syncwarp();
      };
      auto offer = [&](bool hit, uint32_t j) {  // warp-converged: buffer the hits, flush full warps
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (!m) return;
        if (hit) wbuf[nbuf + __popc(m & ((1u << lane) - 1u))] = j;
        nbuf += __popc(m);
        if (nbuf >= 32) {
          flush(32);
          if (lane < nbuf - 32) wbuf[lane] = wbuf[32 + lane];   // disjoint halves: no hazard
          nbuf -= 32;
          __syncwarp();
        }
      };
      auto scan_range = [&](const GridEntry *ents, int e0, int e1) {
        // kUnroll independent 16-byte loads in flight per lane: the scan is L2-latency bound otherwise
        constexpr int kUnroll = 4;
        for (int eb = e0; eb < e1; eb += 32 * kUnroll) {
          GridEntry ge[kUnroll];
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const int e = eb + u * 32 + lane;
            ge[u] = (e < e1) ? ents[e] : GridEntry{0.f, 0.f, 0.f, 0u};
          }
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            if (eb + u * 32 >= e1) break;   // warp-uniform
            const uint32_t j = ge[u].idx;
            bool hit = false;
            if (eb + u * 32 + lane < e1 && ((alive[j >> 5] >> (j & 31)) & 1u)) {
              ++st_circle;
              const float dx = ge[u].x - qx, dy = ge[u].y - qy, rr = ge[u].r + qr;
              hit = dx * dx + dy * dy <= rr * rr;
            }
            offer(hit, j);
          }
        }
      };
      if (prune) {
        // every gridded candidate whose circle can touch has its centre within qr + r_cap
        const float reach = qr + geom.r_cap;
        const int cx0 = geom.cell_x(qx - reach), cx1 = geom.cell_x(qx + reach);
        const int cy0 = geom.cell_y(qy - reach), cy1 = geom.cell_y(qy + reach);
        for (int cy = cy0; cy <= cy1; ++cy) scan_range(entries, cell_start[cy * kG + cx0], cell_start[cy * kG + cx1 + 1]);
        for (int ob = 0; ob < nos; ob += 32) {               // oversize candidates
          const int o = ob + lane;
          bool hit = false;
          uint32_t j = 0;
          if (o < nos) {
            j = os_list[o];
            if ((alive[j >> 5] >> (j & 31)) & 1u) {
              ++st_circle;
              const Rec &rj = recs[j];
              const float dx = rec_cx(rj) - qx, dy = rec_cy(rj) - qy, rr = rj.r + qr;
              hit = dx * dx + dy * dy <= rr * rr;
            }
          }
          offer(hit, j);
        }
      } else {
        // pruning disabled (negative thresholds): every alive candidate interacts
        for (int jb = cursor_word << 5; jb < n; jb += 32) {
          const int j = jb + lane;
          const bool hit = j < n && ((alive[j >> 5] >> (j & 31)) & 1u);
          if (hit) ++st_circle;
          offer(hit, static_cast<uint32_t>(j));
        }
      }
      if (nbuf > 0) flush(nbuf);
    }
    __syncthreads();
    if (leader) lap(4);

    // ================= 6. exact IoU of the queued (candidate, kept) pairs =================
    {
      const int qn = min(s_qn, kQ2Cap);
      if (!kWeighted) {
        eval_queue(qn, [&](int q, Rec &ra, Rec &rb) { ra = krec[queue2[q] & 1023]; rb = recs[queue2[q] >> 10]; },
                   [&](int q, bool above, bool) {
                     if (above) { ++st_hit; kill(static_cast<int>(queue2[q] >> 10)); }
                   });
      } else {
        // the queue of ANY CTA overflowing sends the whole round to the exact serial fallback
        cluster.sync();
        bool overflow = false;
        for (int r = 0; r < P; ++r) overflow |= *cluster.map_shared_rank(&s_overflow, r) != 0;
        if (!overflow) {
          // pass 1: the two comparisons of every queued pair; remember each candidate's FIRST suppressor
          eval_queue(qn, [&](int q, Rec &ra, Rec &rb) { ra = krec[queue2[q] & 1023]; rb = recs[queue2[q] >> 10]; },
                     [&](int q, bool above, bool above_m) {
                       qflag[q] = (above ? 1u : 0u) | (above_m ? 2u : 0u);
                       if (above) atomicMin(&firstsup[queue2[q] >> 10], static_cast<int>(queue2[q] & 1023));
                     });
          __threadfence();
          cluster.sync();   // (B) every CTA's first-suppressor votes are in
          // pass 2: merges up to and including the first suppressor; the suppressor clears the bit
          for (int q = tid; q < qn; q += kNmsThreads) {
            const uint32_t e = queue2[q];
            const int j = e >> 10, t = e & 1023;
            const int fs = __ldcg(&firstsup[j]);
            if (t > fs) continue;
            if (qflag[q] & 2u) accumulate(slot0 + t, j);
            if (t == fs) kill(j);
          }
        } else {
          cluster.sync();   // keep the barrier count uniform
          if (leader) {
            // exact serial fallback: one thread per alive candidate, kept boxes visited in rank order
            for (int j = (cursor_word << 5) + tid; j < n; j += kNmsThreads) {
              if (!((alive[j >> 5] >> (j & 31)) & 1u)) continue;
              const Rec rj = recs[j];
              const float jx = rec_cx(rj), jy = rec_cy(rj);
              for (int t = 0; t < nk; ++t) {
                if (prune) {
                  const float dx = kx[t] - jx, dy = ky[t] - jy, rr = kr[t] + rj.r;
                  if (!(dx * dx + dy * dy <= rr * rr)) continue;
                }
                bool above, above_m;
                classify_pair(krec[t], rj, above, above_m);
                if (above_m) accumulate(slot0 + t, j);
                if (above) { kill(j); break; }
              }
            }
          }
        }
      }
    }
    cluster.sync();   // (C) every kill has reached the leader's bitmap
    if (tid == 0) s_overflow = 0;
    if (leader) lap(5);
    if (last_round) break;
  }
  cluster.sync();     // nobody leaves while its shared memory may still be read

  if (tid == 0 && leader) a.kept_count[seg] = kept_total;
  if (a.stats) {
    // warp-reduce then one atomic per warp
    for (int o = 16; o; o >>= 1) {
      st_iou += __shfl_xor_sync(0xffffffffu, st_iou, o);
      st_circle += __shfl_xor_sync(0xffffffffu, st_circle, o);
      st_hit += __shfl_xor_sync(0xffffffffu, st_hit, o);
      st_approx += __shfl_xor_sync(0xffffffffu, st_approx, o);
    }
    if (lane == 0) {
      atomicAdd(a.stats + 0, st_iou);
      atomicAdd(a.stats + 3, st_circle);
      atomicAdd(a.stats + 18, st_hit);
      atomicAdd(a.stats + 19, st_approx);
    }
    if (tid == 0 && leader) {
      atomicAdd(a.stats + 1, static_cast<unsigned long long>(kept_total));
      atomicAdd(a.stats + 2, static_cast<unsigned long long>(rounds));
      unsigned long long tot = 0;
      for (int k = 0; k < 6; ++k) { atomicAdd(a.stats + 4 + k, static_cast<unsigned long long>(ph[k])); tot += ph[k]; }
      const unsigned long long prev = atomicMax(a.stats + 10, tot);  // slowest segment (cycles)
      if (tot > prev)                                                // (diagnostic, last writer wins) its phases
        for (int k = 0; k < 6; ++k) a.stats[12 + k] = static_cast<unsigned long long>(ph[k]);
      atomicAdd(a.stats + 20, static_cast<unsigned long long>(ph[6]));
      atomicAdd(a.stats + 21, static_cast<unsigned long long>(ph[7]));
      atomicMax(a.stats + 11, static_cast<unsigned long long>(n));   // largest segment (candidates)
    }
  }
}

// ------------------------------------------------------------------------------------------
// output pack
// ------------------------------------------------------------------------------------------
__global__ void kept_scan_kernel(const int *__restrict__ kept_count, int S, int *__restrict__ out_off,
                                 int *__restrict__ out_count, int out_capacity) {
  // single thread block; S is small (sweeps x classes)
  __shared__ int s_part[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (S + nt - 1) / nt;
  const int lo = min(S, tid * per), hi = min(S, lo + per);
  int sum = 0;
  for (int s = lo; s < hi; ++s) sum += kept_count[s];
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < nt; ++t) { const int v = s_part[t]; s_part[t] = run; run += v; }
    *out_count = min(run, out_capacity);
  }
  __syncthreads();
  int run = s_part[tid];
  for (int s = lo; s < hi; ++s) { out_off[s] = run; run += kept_count[s]; }
}

struct PackArgs {
  const int *seg_begin, *kept_base, *kept_count, *out_off, *kept_pos;
  const uint32_t *order;
  const float *boxes;
  const double *acc;
  int total_classes, out_capacity, weighted, yaw_layout;
  float *out_params, *out_scores, *out_cats, *out_batch;
  // fused gather: every rank's buffer is (world, peer_capacity + 1, 16) f32; this rank writes slot `peer_rank` of each
  const int *out_count;
  int *out_count_w;
  int n_segments, fused_scan;
  int n_peers, peer_rank, peer_capacity, sweep_offset;
  float *peer_rows[RV3D_MAX_PEERS];
};

__global__ void __launch_bounds__(128) pack_kernel(PackArgs a) {
  const int seg = blockIdx.x;
  const int cnt = a.kept_count[seg];
  const int kb = a.kept_base[seg], beg = a.seg_begin[seg];
  // Output offset of the segment = kept boxes of the segments before it.  With few segments every block sums that
  // prefix itself (<= 4 loads per thread) instead of waiting for a separate one-block scan kernel; block (0, 0) also
  // publishes the total.  Many segments (n_segments > kPackFusedScan): kept_scan_kernel ran before, out_off is ready.
  __shared__ int s_part[8];
  int off, total = 0;
  const bool first = blockIdx.x == 0 && blockIdx.y == 0;
  if (a.fused_scan) {
    int part = 0, tot = 0;
    const int upto = first ? a.n_segments : seg;
    for (int s = threadIdx.x; s < upto; s += blockDim.x) {
      const int v = a.kept_count[s];
      tot += v;
      if (s < seg) part += v;
    }
    for (int o = 16; o; o >>= 1) {
      part += __shfl_xor_sync(0xffffffffu, part, o);
      tot += __shfl_xor_sync(0xffffffffu, tot, o);
    }
    if ((threadIdx.x & 31) == 0) { s_part[threadIdx.x >> 5] = part; s_part[4 + (threadIdx.x >> 5)] = tot; }
    __syncthreads();
    off = s_part[0] + s_part[1] + s_part[2] + s_part[3];
    total = s_part[4] + s_part[5] + s_part[6] + s_part[7];
    total = total < a.out_capacity ? total : a.out_capacity;
    if (first && threadIdx.x == 0) *a.out_count_w = total;
  } else {
    off = a.out_off[seg];
    if (first) total = *a.out_count;
  }
  // grid = (segments, chunks of the kept list): every kept box has its own thread (the kernel is three dependent
  // loads and a sincos per row, i.e. pure latency)
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < cnt; t += gridDim.y * blockDim.x) {
    const int row = off + t;
    if (row >= a.out_capacity) return;
    const float4 *src = reinterpret_cast<const float4 *>(a.boxes + static_cast<size_t>(a.order[beg + a.kept_pos[kb + t]]) * 8);
    const float4 b0 = src[0], b1 = src[1];
    float p[7] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z};
    if (a.weighted) {
      // nms.py:109-111: merged [x,y,z,l,w,h] and yaw = atan2(merged sin, merged cos) in float32
      const double *acc = a.acc + static_cast<size_t>(kb + t) * 9;
      const double ws = acc[8];
      float m[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) m[c] = static_cast<float>(acc[c] / ws);
#pragma unroll
      for (int c = 0; c < 6; ++c) p[c] = m[c];
      p[6] = static_cast<float>(atan2(static_cast<double>(m[6]), static_cast<double>(m[7])));
    }
    if (a.yaw_layout) {
      float *o = a.out_params + static_cast<size_t>(row) * 7;
#pragma unroll
      for (int c = 0; c < 7; ++c) o[c] = p[c];
    } else {
      double s, c;
      sincos(static_cast<double>(p[6] * 0.5f), &s, &c);  // SO3.py:122-134
      float *o = a.out_params + static_cast<size_t>(row) * 10;
      o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3]; o[4] = p[4]; o[5] = p[5];
      o[6] = static_cast<float>(c); o[7] = 0.f; o[8] = 0.f; o[9] = static_cast<float>(s);
    }
    a.out_scores[row] = b1.w;
    a.out_cats[row] = static_cast<float>(seg % a.total_classes);   // nms.py:51: full_like(scores, j)
    a.out_batch[row] = static_cast<float>(seg / a.total_classes);  // nms.py:242
    if (a.n_peers > 0 && row < a.peer_capacity) {
      // the path's one exchange step, fused: the detection goes straight into every rank's gather buffer with
      // 16-byte stores through the NVLink-mapped peer pointers (no staging copy, no collective call)
      double s, c;
      sincos(static_cast<double>(p[6] * 0.5f), &s, &c);
      const float4 r0 = make_float4(static_cast<float>(seg / a.total_classes + a.sweep_offset),
                                    static_cast<float>(seg % a.total_classes), b1.w, 0.f);
      const float4 r1 = make_float4(p[0], p[1], p[2], p[3]);
      const float4 r2 = make_float4(p[4], p[5], static_cast<float>(c), 0.f);
      const float4 r3 = make_float4(0.f, static_cast<float>(s), 0.f, 0.f);
      const size_t at = (static_cast<size_t>(a.peer_rank) * (a.peer_capacity + 1) + 1 + row) * 16;
      for (int q = 0; q < a.n_peers; ++q) {
        float4 *dst = reinterpret_cast<float4 *>(a.peer_rows[q] + at);
        dst[0] = r0; dst[1] = r1; dst[2] = r2; dst[3] = r3;
      }
    }
  }
  if (a.n_peers > 0 && first && threadIdx.x == 0) {   // header row: [rows written, rows kept]
    const float4 h = make_float4(static_cast<float>(total < a.peer_capacity ? total : a.peer_capacity),
                                 static_cast<float>(total), 0.f, 0.f);
    for (int q = 0; q < a.n_peers; ++q)
      *reinterpret_cast<float4 *>(a.peer_rows[q] + static_cast<size_t>(a.peer_rank) * (a.peer_capacity + 1) * 16) = h;
  }
}

// ------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------
struct Carver {
  unsigned char *p;
  size_t used = 0;
  explicit Carver(void *base) : p(static_cast<unsigned char *>(base)) {}
  template <typename T> T *take(size_t count) {
    used = align_up(used, 256);
    T *r = p ? reinterpret_cast<T *>(p + used) : nullptr;
    used += sizeof(T) * count;
    return r;
  }
};

struct NmsLayout {
  unsigned long long *keys_alt;
  uint32_t *order, *order_alt;
  int *seg_begin, *seg_end, *kept_base, *kept_count, *out_off, *kept_pos, *merge_count;
  void *recs;
  float *data;
  double *acc;
  void *grid_entries;
  uint32_t *oversize;
  int *firstsup;
  void *cub_tmp;
  size_t cub_bytes, total;
};

static size_t cub_sort_bytes(int n, int end_bit) {
  size_t bytes = 0;
  cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, n, 0, end_bit, nullptr);
  return bytes;
}

static NmsLayout nms_layout(void *scratch, int n, int S, bool weighted, int D, int end_bit) {
  Carver c(scratch);
  NmsLayout L{};
  const int nn = n > 0 ? n : 1;
  L.keys_alt = c.take<unsigned long long>(nn);
  L.order = c.take<uint32_t>(nn);
  L.order_alt = c.take<uint32_t>(nn);
  L.seg_begin = c.take<int>(2 * static_cast<size_t>(S));  // seg_begin + seg_end contiguous (one memset)
  L.seg_end = L.seg_begin ? L.seg_begin + S : nullptr;
  L.kept_base = L.seg_begin;   // see rv3d_nms: a segment's kept rows live in its own [seg_begin, seg_end) slice
  L.kept_count = c.take<int>(S);
  L.out_off = c.take<int>(S);
  L.kept_pos = c.take<int>(nn);
  L.recs = c.take<unsigned char>(static_cast<size_t>(nn) * (weighted ? sizeof(WRec) : sizeof(HardRec)));
  L.data = nullptr; L.acc = nullptr; L.merge_count = nullptr;
  if (weighted) {
    L.data = c.take<float>(static_cast<size_t>(nn) * D);
    L.acc = c.take<double>(static_cast<size_t>(nn) * D);   // acc + merge_count contiguous (one memset)
    L.merge_count = c.take<int>(nn);
  }
  L.grid_entries = c.take<GridEntry>(nn);
  L.oversize = c.take<uint32_t>(nn);
  L.firstsup = weighted ? c.take<int>(nn) : nullptr;
  L.cub_bytes = cub_sort_bytes(nn, end_bit);
  L.cub_tmp = c.take<unsigned char>(L.cub_bytes);
  L.total = align_up(c.used, 256);
  return L;
}

template <typename Rec, bool kWeighted>
static int launch_nms_segments(const NmsArgs &a, int S, int max_seg_n, cudaStream_t s) {
  const int nwords = (max_seg_n + 31) / 32;
  constexpr int kF = Frontier<kWeighted>::kF;
  const size_t smem = nms_smem_bytes<Rec, kWeighted, kF>(nwords);
  // the alive bitmap lives in shared memory and grid entries index candidates with 20 bits
  if (smem > 200 * 1024 || max_seg_n >= (1 << 20)) return RV3D_ERR_ARG;
  RV3D_CHECK_CUDA(cudaFuncSetAttribute(nms_segment_kernel<Rec, kWeighted, kF>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  // One cluster per segment, 1 CTA / SM, P = ceil(148 / S) CTAs per cluster (at most 4).  Rounding UP
  // oversubscribes the machine a little (S = 48 -> 192 CTAs; GPC boundaries only let 45 clusters of 3 or
  // 36 of 4 be co-resident anyway): segments finish at different times, so late clusters start in the
  // gaps, and the larger cluster shortens every segment.  Measured at S = 48: P = 2 / 3 / 4 / 6 / 8 ->
  // 1.03 / 0.99 / 0.96 / 1.00 / 1.12 ms for the whole sort + NMS + pack stage; at S = 24: P = 3 / 4 / 7 / 8
  // -> 0.80 / 0.71 / 0.78 / 0.77 ms (the leader-only phases and the cluster barriers stop paying beyond 4).
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kNmsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int P = (kNumSMs + S - 1) / (S > 0 ? S : 1);
  P = P < 1 ? 1 : (P > 4 ? 4 : P);
  if (const char *e = getenv("RV3D_NMS_CLUSTER")) {   // experiments only
    const int v = atoi(e);
    if (v >= 1 && v <= 8) P = v;
  }
  attr[0].val.clusterDim.x = static_cast<unsigned>(P);
  cfg.gridDim = dim3(static_cast<unsigned>(S * P));
  RV3D_CHECK_CUDA(cudaLaunchKernelEx(&cfg, nms_segment_kernel<Rec, kWeighted, kF>, a));
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

}  // namespace rv3d

using namespace rv3d;

extern "C" size_t rv3d_nms_scratch_bytes(const rv3d_nms_params *p) {
  if (!p || p->batch <= 0 || p->total_classes <= 0 || p->n_candidates < 0) return 0;
  const int S = p->batch * p->total_classes;
  const int end_bit = bits_for(S) + (p->score_bits ? p->score_bits : 32) + bits_for(p->total_candidates);
  return nms_layout(nullptr, p->n_candidates, S, p->mode == RV3D_NMS_WEIGHTED, 9, end_bit > 64 ? 64 : end_bit).total;
}

extern "C" int rv3d_nms(const rv3d_nms_params *p, uint64_t *keys_in, const float *boxes, float *out_params,
                        float *out_scores, float *out_categories, float *out_batch, int32_t *out_count,
                        int64_t *stats, void *scratch, size_t scratch_bytes, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(p && out_count && scratch);
  RV3D_CHECK_ARG(p->batch > 0 && p->total_classes > 0 && p->total_candidates > 0 && p->n_candidates >= 0);
  RV3D_CHECK_ARG(p->num_pre_nms > 0 && p->num_post_nms > 0 && p->out_capacity >= 0);
  RV3D_CHECK_ARG(p->mode == RV3D_NMS_HARD || p->mode == RV3D_NMS_WEIGHTED);
  RV3D_CHECK_ARG(p->out_layout == RV3D_OUT_QUAT || p->out_layout == RV3D_OUT_YAW);
  RV3D_CHECK_ARG(p->peer_world >= 0 && p->peer_world <= RV3D_MAX_PEERS);
  if (p->peer_world > 0) {
    RV3D_CHECK_ARG(p->out_layout == RV3D_OUT_QUAT && p->peer_rank >= 0 && p->peer_rank < p->peer_world && p->peer_capacity > 0);
    for (int q = 0; q < p->peer_world; ++q) RV3D_CHECK_ARG(p->peer_rows[q] && aligned(p->peer_rows[q], 16));
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n = p->n_candidates;
  if (n == 0) {
    RV3D_CHECK_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t), s));
    return RV3D_OK;
  }
  RV3D_CHECK_ARG(keys_in && boxes && out_params && out_scores && out_categories && out_batch);
  if (!aligned(boxes, 16) || !aligned(keys_in, 8) || !aligned(scratch, 256)) return RV3D_ERR_ALIGN;
  const int S = p->batch * p->total_classes;
  const int idx_bits = bits_for(p->total_candidates);
  const int score_bits = p->score_bits ? p->score_bits : 32;
  RV3D_CHECK_ARG(score_bits == 31 || score_bits == 32);
  const int end_bit = bits_for(S) + score_bits + idx_bits;
  if (end_bit > 64) return RV3D_ERR_KEYBITS;
  const bool weighted = p->mode == RV3D_NMS_WEIGHTED;
  const NmsLayout L = nms_layout(scratch, n, S, weighted, 9, end_bit);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  auto *keys = reinterpret_cast<unsigned long long *>(keys_in);

  // K3: sort (segment asc, score desc, candidate asc)
  iota_kernel<<<ceil_div(n, 256), 256, 0, s>>>(L.order, n);
  RV3D_CHECK_LAUNCH();
  cub::DoubleBuffer<unsigned long long> kb(keys, L.keys_alt);
  cub::DoubleBuffer<uint32_t> vb(L.order, L.order_alt);
  size_t cub_bytes = L.cub_bytes;
  RV3D_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, cub_bytes, kb, vb, n, 0, end_bit, s));
  const unsigned long long *skeys = kb.Current();
  const uint32_t *order = vb.Current();

  RV3D_CHECK_CUDA(cudaMemsetAsync(L.seg_begin, 0, sizeof(int) * 2 * S, s));

  // a segment keeps at most as many boxes as it has candidates, so its slice [seg_begin, seg_end) of the n-sized
  // kept_pos / acc / merge_count arrays is its private, sufficient region: kept_base == seg_begin (no scan kernel)
  NmsArgs a{};
  a.recs = L.recs; a.seg_begin = L.seg_begin; a.seg_end = L.seg_end; a.kept_base = L.seg_begin;
  a.num_pre = p->num_pre_nms; a.num_post = p->num_post_nms; a.thr = p->iou_threshold; a.mthr = p->merge_threshold;
  a.prune = (p->iou_threshold >= 0.f && (p->mode != RV3D_NMS_WEIGHTED || p->merge_threshold >= 0.f)) ? 1 : 0;
  a.kept_pos = L.kept_pos; a.kept_count = L.kept_count; a.data = L.data; a.D = 9; a.acc = L.acc;
  a.merge_count = L.merge_count; a.stats = reinterpret_cast<unsigned long long *>(stats);
  a.grid_entries = L.grid_entries; a.oversize = L.oversize; a.firstsup = L.firstsup;
  const int max_seg_n = n < p->num_pre_nms ? n : p->num_pre_nms;
  int rc;
  if (weighted) {
    RV3D_CHECK_CUDA(cudaMemsetAsync(L.acc, 0, sizeof(double) * static_cast<size_t>(n) * 9, s));
    RV3D_CHECK_CUDA(cudaMemsetAsync(L.merge_count, 0, sizeof(int) * static_cast<size_t>(n), s));
    prepare_records_kernel<true><<<ceil_div(n, 256), 256, 0, s>>>(order, boxes, n, L.recs, L.data, skeys, score_bits + idx_bits,
                                                                  L.seg_begin, L.seg_end);
    RV3D_CHECK_LAUNCH();
    rc = launch_nms_segments<WRec, true>(a, S, max_seg_n, s);
  } else {
    prepare_records_kernel<false><<<ceil_div(n, 256), 256, 0, s>>>(order, boxes, n, L.recs, nullptr, skeys, score_bits + idx_bits,
                                                                   L.seg_begin, L.seg_end);
    RV3D_CHECK_LAUNCH();
    rc = launch_nms_segments<HardRec, false>(a, S, max_seg_n, s);
  }
  if (rc != RV3D_OK) return rc;

  constexpr int kPackFusedScan = 512;   // up to this many segments the pack blocks sum their own output offset
  const bool fused_scan = S <= kPackFusedScan;
  if (!fused_scan) {
    kept_scan_kernel<<<1, 1024, 0, s>>>(L.kept_count, S, L.out_off, out_count, p->out_capacity);
    RV3D_CHECK_LAUNCH();
  }
  PackArgs pa{};
  pa.out_count_w = out_count; pa.n_segments = S; pa.fused_scan = fused_scan ? 1 : 0;
  pa.seg_begin = L.seg_begin; pa.kept_base = L.seg_begin; pa.kept_count = L.kept_count; pa.out_off = L.out_off;
  pa.kept_pos = L.kept_pos; pa.order = order; pa.boxes = boxes; pa.acc = L.acc;
  pa.total_classes = p->total_classes; pa.out_capacity = p->out_capacity; pa.weighted = weighted ? 1 : 0; pa.yaw_layout = p->out_layout == RV3D_OUT_YAW;
  pa.out_params = out_params; pa.out_scores = out_scores; pa.out_cats = out_categories; pa.out_batch = out_batch;
  pa.out_count = out_count;
  pa.n_peers = p->peer_world; pa.peer_rank = p->peer_rank; pa.peer_capacity = p->peer_capacity; pa.sweep_offset = p->sweep_offset;
  for (int q = 0; q < p->peer_world; ++q) pa.peer_rows[q] = p->peer_rows[q];
  {
    const int chunks = ceil_div(p->num_post_nms < (1 << 20) ? p->num_post_nms : (1 << 20), 128);
    pack_kernel<<<dim3(S, chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks)), 128, 0, s>>>(pa);
  }
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

// ==========================================================================================
// free-standing operator forms
// ==========================================================================================
namespace rv3d {

__global__ void score_keys_kernel(const float *__restrict__ scores, int n, int idx_bits,
                                  unsigned long long *__restrict__ keys, uint32_t *__restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t desc = ~orderable_f32(__float_as_uint(scores[i]));
  keys[i] = (static_cast<unsigned long long>(desc) << idx_bits) | static_cast<uint32_t>(i);
  order[i] = i;
}

__global__ void single_segment_kernel(int n, int *seg_begin, int *seg_end, int *kept_base) {
  seg_begin[0] = 0; seg_end[0] = n; kept_base[0] = 0;
}

__global__ void hard_recs_from_boxes5_kernel(const float *__restrict__ boxes5, const uint32_t *__restrict__ order,
                                             int n, double angle_scale, HardRec *__restrict__ recs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *b = boxes5 + static_cast<size_t>(order ? order[i] : i) * 5;
  recs[i] = make_hard_rec(b[0], b[1], b[2], b[3], b[4], angle_scale);
}

__global__ void w_recs_from_boxes5_kernel(const float *__restrict__ boxes5, int n, WRec *__restrict__ recs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *b = boxes5 + static_cast<size_t>(i) * 5;
  recs[i] = make_w_rec(b[0], b[1], b[2], b[3], b[4]);
}

__global__ void keep_indices_kernel(const int *__restrict__ kept_pos, const int *__restrict__ kept_count,
                                    const uint32_t *__restrict__ order, int64_t *__restrict__ keep,
                                    int32_t *__restrict__ n_keep) {
  const int cnt = kept_count[0];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_keep = cnt;
  if (i < cnt) keep[i] = order ? order[kept_pos[i]] : kept_pos[i];
}

__global__ void wnms_finalize_kernel(const int *__restrict__ kept_pos, const int *__restrict__ kept_count,
                                     const double *__restrict__ acc, const int *__restrict__ merge_count,
                                     const float *__restrict__ data, int D, float *__restrict__ output,
                                     int64_t *__restrict__ keep, int64_t *__restrict__ count,
                                     int32_t *__restrict__ n_out) {
  const int cnt = kept_count[0];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_out = cnt;
  if (i >= cnt) return;
  const double ws = acc[static_cast<size_t>(i) * D + D - 1];
  for (int c = 0; c < D - 1; ++c)
    output[static_cast<size_t>(i) * D + c] = static_cast<float>(acc[static_cast<size_t>(i) * D + c] / ws);
  output[static_cast<size_t>(i) * D + D - 1] = data[static_cast<size_t>(kept_pos[i]) * D + D - 1];
  keep[i] = kept_pos[i];
  count[i] = merge_count[i];
}

__device__ __forceinline__ float nan_to_num_f(float v) {
  if (v != v) return 0.f;
  if (v == CUDART_INF_F) return 3.4028234663852886e38f;
  if (v == -CUDART_INF_F) return -3.4028234663852886e38f;
  return v;
}

// iou.py:11-47
__global__ void __launch_bounds__(128)
iou3d_aligned_kernel(const float *__restrict__ A, const float *__restrict__ B, int64_t n,
                     float *__restrict__ iou3d, float *__restrict__ iou_bev, int32_t *__restrict__ status) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *a = A + i * 7, *b = B + i * 7;
  const HardRec ra = make_hard_rec(a[0], a[1], a[3], a[4], a[6], 1.0);  // mmcv: angle in radians
  const HardRec rb = make_hard_rec(b[0], b[1], b[3], b[4], b[6], 1.0);
  float bev = rot_iou(ra, rb);
  bev = (bev != bev) ? bev : fminf(fmaxf(bev, 0.0f), 1.0f);  // clamp keeps NaN
  bev = nan_to_num_f(bev);
  const float area_a = a[3] * a[4], area_b = b[3] * b[4];
  const float ov_bev = bev * (area_a + area_b) / (1.0f + bev);
  const float a_top = a[2] + a[5] / 2.0f, a_btm = a[2] - a[5] / 2.0f;
  const float b_top = b[2] + b[5] / 2.0f, b_btm = b[2] - b[5] / 2.0f;
  // torch.max / torch.min propagate NaN
  const float hi_btm = (a_btm != a_btm || b_btm != b_btm) ? CUDART_NAN_F : fmaxf(a_btm, b_btm);
  const float lo_top = (a_top != a_top || b_top != b_top) ? CUDART_NAN_F : fminf(a_top, b_top);
  float ov_h = lo_top - hi_btm;
  ov_h = (ov_h != ov_h) ? ov_h : fmaxf(ov_h, 0.0f);
  const float ov3 = ov_bev * ov_h;
  const float va = a[3] * a[4] * a[5], vb = b[3] * b[4] * b[5];
  float den = va + vb - ov3;
  den = (den != den) ? den : fmaxf(den, 1e-8f);
  float v = nan_to_num_f(ov3 / den);
  if (!(fabsf(v) <= 3.4028234663852886e38f)) atomicExch(status, 1);
  iou3d[i] = v;
  iou_bev[i] = bev;
}

// mmcv box_iou_rotated(bboxes1 (N,5), bboxes2 (M,5), aligned): (xc,yc,w,h,angle rad) -> (N,M) or (N,) IoU.
// One thread per output element, consecutive threads along M so box b is read coalesced and box a broadcast.
__global__ void __launch_bounds__(128)
box_iou_rotated_kernel(const float *__restrict__ A, int64_t n, const float *__restrict__ B, int64_t m, int aligned,
                       float *__restrict__ out) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = aligned ? n : n * m;
  if (t >= total) return;
  const int64_t i = aligned ? t : t / m, j = aligned ? t : t - i * m;
  const float *a = A + i * 5, *b = B + j * 5;
  out[t] = rot_iou(make_hard_rec(a[0], a[1], a[2], a[3], a[4], 1.0), make_hard_rec(b[0], b[1], b[2], b[3], b[4], 1.0));
}

// threshold-only branch: rows ordered by (sweep, candidate index)
__global__ void remap_keys_kernel(const unsigned long long *__restrict__ keys, int n, int idx_bits, int score_bits,
                                  int total_classes, unsigned long long *__restrict__ out, uint32_t *__restrict__ order,
                                  uint32_t *__restrict__ seg_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  const uint32_t seg = static_cast<uint32_t>(k >> (score_bits + idx_bits));
  const unsigned long long cand = k & ((1ull << idx_bits) - 1ull);
  out[i] = (static_cast<unsigned long long>(seg / total_classes) << idx_bits) | cand;
  order[i] = i;
  seg_out[i] = seg;
}

__global__ void pack_candidates_kernel(const uint32_t *__restrict__ order, const uint32_t *__restrict__ segs,
                                       const float *__restrict__ boxes, int n, int total_classes,
                                       float *__restrict__ out_params, float *__restrict__ out_scores,
                                       int64_t *__restrict__ out_cats, int64_t *__restrict__ out_batch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t src = order[i];
  const float4 *bp = reinterpret_cast<const float4 *>(boxes + static_cast<size_t>(src) * 8);
  const float4 b0 = bp[0], b1 = bp[1];
  double s, c;
  sincos(static_cast<double>(b1.z * 0.5f), &s, &c);
  float *o = out_params + static_cast<size_t>(i) * 10;
  o[0] = b0.x; o[1] = b0.y; o[2] = b0.z; o[3] = b0.w; o[4] = b1.x; o[5] = b1.y;
  o[6] = static_cast<float>(c); o[7] = 0.f; o[8] = 0.f; o[9] = static_cast<float>(s);
  out_scores[i] = b1.w;
  out_cats[i] = segs[src] % total_classes;
  out_batch[i] = segs[src] / total_classes;
}

struct OpLayout {
  unsigned long long *keys, *keys_alt;
  uint32_t *order, *order_alt, *segs;
  int *seg_begin, *seg_end, *kept_base, *kept_count, *kept_pos, *merge_count;
  void *recs;
  double *acc;
  void *grid_entries;
  uint32_t *oversize;
  int *firstsup;
  void *cub_tmp;
  size_t cub_bytes, total;
};

static OpLayout op_layout(void *scratch, int n, bool weighted, int D, bool sort) {
  Carver c(scratch);
  OpLayout L{};
  const int nn = n > 0 ? n : 1;
  L.keys = c.take<unsigned long long>(nn);
  L.keys_alt = c.take<unsigned long long>(nn);
  L.order = c.take<uint32_t>(nn);
  L.order_alt = c.take<uint32_t>(nn);
  L.segs = c.take<uint32_t>(nn);
  L.seg_begin = c.take<int>(4);
  L.seg_end = L.seg_begin ? L.seg_begin + 1 : nullptr;
  L.kept_base = L.seg_begin ? L.seg_begin + 2 : nullptr;
  L.kept_count = L.seg_begin ? L.seg_begin + 3 : nullptr;
  L.kept_pos = c.take<int>(nn);
  L.recs = c.take<unsigned char>(static_cast<size_t>(nn) * (weighted ? sizeof(WRec) : sizeof(HardRec)));
  L.acc = nullptr; L.merge_count = nullptr;
  if (weighted) {
    L.acc = c.take<double>(static_cast<size_t>(nn) * D);
    L.merge_count = c.take<int>(nn);
  }
  L.grid_entries = c.take<GridEntry>(nn);
  L.oversize = c.take<uint32_t>(nn);
  L.firstsup = weighted ? c.take<int>(nn) : nullptr;
  L.cub_bytes = sort ? cub_sort_bytes(nn, 64) : 0;
  L.cub_tmp = c.take<unsigned char>(L.cub_bytes ? L.cub_bytes : 1);
  L.total = align_up(c.used, 256);
  return L;
}

}  // namespace rv3d

extern "C" size_t rv3d_nms_rotated_scratch_bytes(int32_t n) { return op_layout(nullptr, n, false, 0, true).total; }

extern "C" int rv3d_nms_rotated(const float *boxes, const float *scores, int32_t n, float iou_threshold,
                                int64_t *keep, int32_t *n_keep, void *scratch, size_t scratch_bytes,
                                rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && n_keep && scratch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    RV3D_CHECK_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int32_t), s));
    return RV3D_OK;
  }
  RV3D_CHECK_ARG(boxes && scores && keep);
  if (!aligned(scratch, 256)) return RV3D_ERR_ALIGN;
  const OpLayout L = op_layout(scratch, n, false, 0, true);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  const int idx_bits = bits_for(n);
  score_keys_kernel<<<ceil_div(n, 256), 256, 0, s>>>(scores, n, idx_bits, L.keys, L.order);
  RV3D_CHECK_LAUNCH();
  cub::DoubleBuffer<unsigned long long> kb(L.keys, L.keys_alt);
  cub::DoubleBuffer<uint32_t> vb(L.order, L.order_alt);
  size_t cub_bytes = L.cub_bytes;
  RV3D_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, cub_bytes, kb, vb, n, 0, 32 + idx_bits, s));
  const uint32_t *order = vb.Current();
  hard_recs_from_boxes5_kernel<<<ceil_div(n, 256), 256, 0, s>>>(boxes, order, n, 0.01745329251,
                                                                static_cast<HardRec *>(L.recs));
  RV3D_CHECK_LAUNCH();
  single_segment_kernel<<<1, 1, 0, s>>>(n, L.seg_begin, L.seg_end, L.kept_base);
  RV3D_CHECK_LAUNCH();
  NmsArgs a{};
  a.recs = L.recs; a.seg_begin = L.seg_begin; a.seg_end = L.seg_end; a.kept_base = L.kept_base;
  a.num_pre = n; a.num_post = n; a.thr = iou_threshold; a.mthr = 0.f; a.prune = iou_threshold >= 0.f ? 1 : 0;
  a.kept_pos = L.kept_pos; a.kept_count = L.kept_count;
  a.grid_entries = L.grid_entries; a.oversize = L.oversize; a.firstsup = L.firstsup;
  const int rc = launch_nms_segments<HardRec, false>(a, 1, n, s);
  if (rc != RV3D_OK) return rc;
  keep_indices_kernel<<<ceil_div(n, 256), 256, 0, s>>>(L.kept_pos, L.kept_count, order, keep, n_keep);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" size_t rv3d_wnms_scratch_bytes(int32_t n, int32_t d) { return op_layout(nullptr, n, true, d, false).total; }

extern "C" int rv3d_wnms(const float *boxes, const float *data, int32_t n, int32_t d, float nms_threshold,
                         float merge_threshold, float *output, int64_t *keep, int64_t *count, int32_t *n_out,
                         void *scratch, size_t scratch_bytes, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && d >= 1 && d <= kMaxD && n_out && scratch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    RV3D_CHECK_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), s));
    return RV3D_OK;
  }
  RV3D_CHECK_ARG(boxes && data && output && keep && count);
  if (!aligned(scratch, 256)) return RV3D_ERR_ALIGN;
  const OpLayout L = op_layout(scratch, n, true, d, false);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  w_recs_from_boxes5_kernel<<<ceil_div(n, 256), 256, 0, s>>>(boxes, n, static_cast<WRec *>(L.recs));
  RV3D_CHECK_LAUNCH();
  single_segment_kernel<<<1, 1, 0, s>>>(n, L.seg_begin, L.seg_end, L.kept_base);
  RV3D_CHECK_LAUNCH();
  RV3D_CHECK_CUDA(cudaMemsetAsync(L.acc, 0, sizeof(double) * static_cast<size_t>(n) * d, s));
  RV3D_CHECK_CUDA(cudaMemsetAsync(L.merge_count, 0, sizeof(int) * static_cast<size_t>(n), s));
  RV3D_CHECK_CUDA(cudaMemsetAsync(output, 0, sizeof(float) * static_cast<size_t>(n) * d, s));  // nms.py:155,173
  RV3D_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int64_t) * static_cast<size_t>(n), s));
  NmsArgs a{};
  a.recs = L.recs; a.seg_begin = L.seg_begin; a.seg_end = L.seg_end; a.kept_base = L.kept_base;
  a.num_pre = n; a.num_post = n; a.thr = nms_threshold; a.mthr = merge_threshold;
  a.prune = (nms_threshold >= 0.f && merge_threshold >= 0.f) ? 1 : 0;
  a.kept_pos = L.kept_pos; a.kept_count = L.kept_count; a.data = data; a.D = d; a.acc = L.acc;
  a.merge_count = L.merge_count;
  a.grid_entries = L.grid_entries; a.oversize = L.oversize; a.firstsup = L.firstsup;
  const int rc = launch_nms_segments<WRec, true>(a, 1, n, s);
  if (rc != RV3D_OK) return rc;
  wnms_finalize_kernel<<<ceil_div(n, 256), 256, 0, s>>>(L.kept_pos, L.kept_count, L.acc, L.merge_count, data, d,
                                                        output, keep, count, n_out);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_iou3d_aligned(const float *cuboids_a, const float *cuboids_b, int64_t n, float *iou3d,
                                  float *iou_bev, int32_t *status, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && status);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RV3D_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
  if (n == 0) return RV3D_OK;
  RV3D_CHECK_ARG(cuboids_a && cuboids_b && iou3d && iou_bev);
  iou3d_aligned_kernel<<<ceil_div(n, 128), 128, 0, s>>>(cuboids_a, cuboids_b, n, iou3d, iou_bev, status);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_box_iou_rotated(const float *boxes_a, int64_t n, const float *boxes_b, int64_t m, int32_t aligned,
                                    float *out, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && m >= 0 && (!aligned || n == m));
  const int64_t total = aligned ? n : n * m;
  if (total == 0) return RV3D_OK;
  RV3D_CHECK_ARG(boxes_a && boxes_b && out && total < (int64_t(1) << 37));
  box_iou_rotated_kernel<<<ceil_div(total, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(boxes_a, n, boxes_b, m,
                                                                                            aligned, out);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" size_t rv3d_pack_candidates_scratch_bytes(int32_t n) { return op_layout(nullptr, n, false, 0, true).total; }

extern "C" int rv3d_pack_candidates(uint64_t *keys, const float *boxes, int32_t n, int32_t batch,
                                    int32_t total_classes, int32_t total_candidates, int32_t score_bits, float *out_params,
                                    float *out_scores, int64_t *out_categories, int64_t *out_batch, void *scratch,
                                    size_t scratch_bytes, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && batch > 0 && total_classes > 0 && total_candidates > 0 && scratch);
  RV3D_CHECK_ARG(score_bits == 31 || score_bits == 32);
  if (n == 0) return RV3D_OK;
  RV3D_CHECK_ARG(keys && boxes && out_params && out_scores && out_categories && out_batch);
  if (!aligned(scratch, 256) || !aligned(boxes, 16)) return RV3D_ERR_ALIGN;
  const OpLayout L = op_layout(scratch, n, false, 0, true);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int idx_bits = bits_for(total_candidates);
  remap_keys_kernel<<<ceil_div(n, 256), 256, 0, s>>>(reinterpret_cast<unsigned long long *>(keys), n, idx_bits, score_bits,
                                                     total_classes, L.keys, L.order, L.segs);
  RV3D_CHECK_LAUNCH();
  cub::DoubleBuffer<unsigned long long> kb(L.keys, L.keys_alt);
  cub::DoubleBuffer<uint32_t> vb(L.order, L.order_alt);
  size_t cub_bytes = L.cub_bytes;
  RV3D_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, cub_bytes, kb, vb, n, 0, bits_for(batch) + idx_bits, s));
  pack_candidates_kernel<<<ceil_div(n, 256), 256, 0, s>>>(vb.Current(), L.segs, boxes, n, total_classes, out_params,
                                                          out_scores, out_categories, out_batch);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}
