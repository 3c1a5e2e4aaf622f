// nms.cu -- score bucketing (K3), exact greedy rotated NMS / weighted NMS (K4 / K5), output pack.
//
// Replaces (paths relative to /root/reference):
//   math/ops/nms.py:181-266  batched_multiclass_nms  (per-sweep Python loop, host sync each)
//   math/ops/nms.py:11-61    hard_multiclass_nms     -> detectron2 nms_rotated (:41-45)
//   math/ops/nms.py:64-123   weighted_multiclass_nms -> weighted_nms (:126-177) -> TorchEx wnms_gpu
//   nn/decoders/range_decoder.py:110-123  threshold-only branch, yaw_to_quat, final cat
//
// The third-party kernels build an N x N/64 suppression mask (313 MB at N = 50 k), copy it to the host and scan
// it serially.  Here nothing leaves the device and NO count is needed on the host: the candidate count, the
// per-segment counts and the detection count live in device memory, every kernel below has a fixed launch
// geometry, so rasterize -> decode -> NMS -> pack replays as one CUDA graph (DESIGN.md "K3", "K4").
//
//   K3  hist / bin_scan / scatter_records: a counting sort of the compacted candidates by
//       (segment = sweep x class, coarse score bin).  No global multi-pass radix sort: the exact order inside a
//       bin is established lazily by the NMS kernel, only for the candidates it actually consumes.
//   K4  nms_pull_kernel, one CTA per segment, "pull" form of the greedy scan:
//         repeat: take the next window (<= 2048 candidates = whole score bins), sort it exactly in shared
//                 memory by (score desc, candidate asc);
//           a. PULL (lazy): the candidates the next frontier needs -- the leftovers of the previous chunk against the
//              NEW kept boxes, the next chunk of the window against ALL kept boxes -- look up the kept boxes in a
//              spatial grid (circle pre-test -> work queue -> IoU upper bound -> bit-exact IoU on dense lanes) and die
//              if suppressed;
//           b. the first nf survivors form a frontier (nf <= kF, and no more than the room under num_post_nms can
//              use): all touching pairs inside it -> suppression bit-matrix -> greedy resolution in rank order ->
//              newly kept boxes join the grid; back to a.
//         until num_post_nms boxes are kept or the candidates run out.
//       A candidate is only ever tested against kept boxes, candidates behind the last pulled chunk are never
//       touched in hard mode, and nothing is tested twice.  The result equals the sequential greedy scan: a
//       candidate reaches a frontier only after it has met every kept box that ranks above it, frontiers are
//       resolved in rank order, pruned pairs have IoU exactly 0 (iou.cuh padded_radius).
//   K5  weighted mode: the same scan, every (candidate, kept) comparison also feeds the merge sets; after the
//       last kept box the candidates behind it are pulled once more for their merge contributions only.
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <math_constants.h>

#include "common.cuh"
#include "iou.cuh"

#include <cstdio>
#include <type_traits>
namespace rv3d {

#ifndef RV3D_NMS_THREADS
#define RV3D_NMS_THREADS 512
#endif
constexpr int kNmsThreads = RV3D_NMS_THREADS;
constexpr int kNmsWarps = kNmsThreads / 32;
// frontier size: 512 boxes per round in hard mode (one 32 KB bit-matrix), 256 in weighted mode (two matrices,
// 64-byte records)
template <bool kWeighted> struct Frontier { static constexpr int kF = kWeighted ? 256 : 512; };
constexpr int kMaxD = 16;            // max data columns of the weighted merge
constexpr int kWin = 2048;           // window: candidates sorted and pulled together
constexpr int kMaxBins = 2048;       // coarse score bins per segment (bin offsets live in shared memory)
constexpr int kQ2Cap = 8192;         // IoU work queue (pairs that passed the circle test)
constexpr int kHitCap = 12;          // per-thread buffer (shared memory) of the hits of one grid walk; longer lists walk twice
constexpr int kKeptSmem = 2048;      // kept boxes tracked in shared memory when num_post_nms <= this, else in global memory
constexpr int kBucketsSmem = 4096;   // hash buckets of the kept-box grid (shared-memory form)
constexpr float kPosCap = 1.0e6f;
constexpr int kMaxCellsPerQuery = 25;
constexpr int kFrontBuckets = 1024;   // hash buckets of the per-round frontier grid
constexpr int kRankSortMax = 48;      // bins up to this size are ordered by rank counting, larger ones by the bitonic network

// ------------------------------------------------------------------------------------------
// K3: counting sort by (segment, coarse score bin)
// ------------------------------------------------------------------------------------------
// Keys are [segment | desc | candidate] with desc = ~orderable(score) truncated to score_bits, so ascending desc ==
// descending score.  Bin of a key = min((desc - lo) >> shift, nb - 1), 0 below lo: ascending bin == descending
// score.  (lo, shift) come from the score range when the caller knows it (decode: [min_confidence, 1]) or from a
// device-side min / max reduction; a bad range only unbalances the bins, the order stays exact.
struct KeyGeom { int idx_bits, score_bits, nb; };
struct BinMap { uint32_t lo; int shift; };
struct BinMapArg {
  BinMap m;
  const BinMap *dev;   // non-null: read the map from device memory (computed by binmap_from_minmax_kernel)
  __device__ __forceinline__ BinMap get() const { return dev ? *dev : m; }
};
__device__ __forceinline__ int bin_of(uint32_t desc, const BinMap &m, int nb) {
  if (desc <= m.lo) return 0;
  const uint32_t b = (desc - m.lo) >> m.shift;
  return b < static_cast<uint32_t>(nb) ? static_cast<int>(b) : nb - 1;
}
__device__ __forceinline__ int live_count(const int32_t *n_ptr, int capacity) {
  const int n = *n_ptr;
  return n < 0 ? 0 : (n < capacity ? n : capacity);
}

__global__ void __launch_bounds__(256)
minmax_kernel(const unsigned long long *__restrict__ keys, const int32_t *__restrict__ n_ptr, int capacity, KeyGeom g,
              uint32_t *__restrict__ mm) {   // mm[0] = min desc (init 0xffffffff), mm[1] = max desc (init 0)
  const int n = live_count(n_ptr, capacity);
  const uint32_t smask = g.score_bits >= 32 ? 0xffffffffu : ((1u << g.score_bits) - 1u);
  uint32_t lo = 0xffffffffu, hi = 0u;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t d = static_cast<uint32_t>(keys[i] >> g.idx_bits) & smask;
    lo = min(lo, d); hi = max(hi, d);
  }
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(mm, lo); atomicMax(mm + 1, hi); }
}

__host__ __device__ inline BinMap make_binmap(uint32_t lo, uint32_t hi, int nb) {
  BinMap m{lo, 0};
  const uint32_t span = hi >= lo ? hi - lo : 0u;
  while ((span >> m.shift) >= static_cast<uint32_t>(nb)) ++m.shift;
  return m;
}
__global__ void binmap_from_minmax_kernel(const uint32_t *__restrict__ mm, int nb, BinMap *__restrict__ out) {
  *out = make_binmap(mm[0], mm[1] >= mm[0] ? mm[1] : mm[0], nb);
}

__global__ void __launch_bounds__(256)
hist_kernel(const unsigned long long *__restrict__ keys, const int32_t *__restrict__ n_ptr, int capacity, KeyGeom g,
            BinMapArg bma, int n_segments, int *__restrict__ hist) {
  const int n = live_count(n_ptr, capacity);
  const BinMap bm = bma.get();
  const uint32_t smask = g.score_bits >= 32 ? 0xffffffffu : ((1u << g.score_bits) - 1u);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    const uint32_t seg = static_cast<uint32_t>(k >> (g.score_bits + g.idx_bits));
    if (seg >= static_cast<uint32_t>(n_segments)) continue;   // corrupt key: ignored everywhere
    atomicAdd(hist + static_cast<size_t>(seg) * g.nb + bin_of(static_cast<uint32_t>(k >> g.idx_bits) & smask, bm, g.nb), 1);
  }
}

// One CTA per segment: exclusive scan of the segment's histogram -> bin_start[seg][0..nb] (relative to the segment),
// seg_count[seg]; the histogram is zeroed again (scatter uses it as the per-bin cursor).  The last CTA to finish
// turns the per-segment counts into seg_begin (exclusive prefix over the segments).
__global__ void __launch_bounds__(256)
bin_scan_kernel(int *__restrict__ hist, int nb, int S, int *__restrict__ bin_start, int *__restrict__ seg_count,
                int *__restrict__ seg_begin, int *__restrict__ ticket) {
  __shared__ int s_w[9];
  __shared__ int s_last;
  const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int per = nb >= 256 ? nb / 256 : 1;
  int *h = hist + static_cast<size_t>(seg) * nb;
  int *bs = bin_start + static_cast<size_t>(seg) * (nb + 1);
  int cnt[kMaxBins / 256];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < kMaxBins / 256; ++k) {
    const int b = tid * per + k;
    cnt[k] = (k < per && b < nb) ? h[b] : 0;
    sum += cnt[k];
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_w[wid] = incl;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int w = 0; w < 8; ++w) { const int v = s_w[w]; s_w[w] = run; run += v; }
    s_w[8] = run;
  }
  __syncthreads();
  int off = s_w[wid] + incl - sum;
#pragma unroll
  for (int k = 0; k < kMaxBins / 256; ++k) {
    const int b = tid * per + k;
    if (k < per && b < nb) { bs[b] = off; h[b] = 0; off += cnt[k]; }
  }
  if (tid == 0) { bs[nb] = s_w[8]; seg_count[seg] = s_w[8]; }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(ticket, 1) == S - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // exclusive prefix over the S segment counts (S is small: sweeps x classes)
  const int chunk = (S + 255) / 256;
  const int lo = min(S, tid * chunk), hi = min(S, lo + chunk);
  int part = 0;
  for (int s = lo; s < hi; ++s) part += __ldcg(seg_count + s);
  int inc2 = part;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc2, o);
    if (lane >= o) inc2 += t;
  }
  __syncthreads();
  if (lane == 31) s_w[wid] = inc2;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int w = 0; w < 8; ++w) { const int v = s_w[w]; s_w[w] = run; run += v; }
  }
  __syncthreads();
  int run = s_w[wid] + inc2 - part;
  for (int s = lo; s < hi; ++s) { seg_begin[s] = run; run += __ldcg(seg_count + s); }
}

// candidate row -> suppression record (+ merge row for the weighted mode)
struct RecFromBoxes8 {   // compaction rows [x,y,z,l,w,h,yaw,score]
  const float *boxes;
  __device__ __forceinline__ void hard(int i, HardRec &r) const {
    const float4 *src = reinterpret_cast<const float4 *>(boxes + static_cast<size_t>(i) * 8);
    const float4 b0 = src[0], b1 = src[1];
    // nms.py:33,40: boxes [x, y, l, w, -rad2deg(yaw)] with rad2deg evaluated in float32
    const float angle = -(b1.z * 57.29577951308232f);
    r = make_hard_rec(b0.x, b0.y, b0.w, b1.x, angle, 0.01745329251);
  }
  __device__ __forceinline__ void weighted(int i, WRec &r, float *d) const {
    const float4 *src = reinterpret_cast<const float4 *>(boxes + static_cast<size_t>(i) * 8);
    const float4 b0 = src[0], b1 = src[1];
    // nms.py:87-100: [x - l/2, y - w/2, x + l/2, y + w/2, yaw]; merge row [x,y,z,l,w,h,sin,cos,score]
    r = make_w_rec(b0.x - b0.w / 2, b0.y - b1.x / 2, b0.x + b0.w / 2, b0.y + b1.x / 2, b1.z);
    d[0] = b0.x; d[1] = b0.y; d[2] = b0.z; d[3] = b0.w; d[4] = b1.x; d[5] = b1.y;
    d[6] = static_cast<float>(sin(static_cast<double>(b1.z)));
    d[7] = static_cast<float>(cos(static_cast<double>(b1.z)));
    d[8] = b1.w;
  }
};
struct RecFromBoxes5 {   // detectron2 rows (xc, yc, w, h, angle in degrees)
  const float *boxes;
  __device__ __forceinline__ void hard(int i, HardRec &r) const {
    const float *b = boxes + static_cast<size_t>(i) * 5;
    r = make_hard_rec(b[0], b[1], b[2], b[3], b[4], 0.01745329251);
  }
  __device__ __forceinline__ void weighted(int, WRec &, float *) const {}
};

template <bool kWeighted, typename Builder>
__global__ void __launch_bounds__(256)
scatter_records_kernel(const unsigned long long *__restrict__ keys, Builder build, const int32_t *__restrict__ n_ptr,
                       int capacity, KeyGeom g, BinMapArg bma, int n_segments, int *__restrict__ cursor,
                       const int *__restrict__ bin_start, const int *__restrict__ seg_begin, void *__restrict__ recs,
                       unsigned long long *__restrict__ skey, uint32_t *__restrict__ src, float *__restrict__ data,
                       int lazy_records = 0) {
  const int n = live_count(n_ptr, capacity);
  const BinMap bm = bma.get();
  const uint32_t smask = g.score_bits >= 32 ? 0xffffffffu : ((1u << g.score_bits) - 1u);
  const unsigned long long lowmask = (1ull << (g.score_bits + g.idx_bits)) - 1ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    const uint32_t seg = static_cast<uint32_t>(k >> (g.score_bits + g.idx_bits));
    if (seg >= static_cast<uint32_t>(n_segments)) continue;
    const int bin = bin_of(static_cast<uint32_t>(k >> g.idx_bits) & smask, bm, g.nb);
    const int slot = atomicAdd(cursor + static_cast<size_t>(seg) * g.nb + bin, 1);
    const int pos = seg_begin[seg] + bin_start[static_cast<size_t>(seg) * (g.nb + 1) + bin] + slot;
    skey[pos] = k & lowmask;
    src[pos] = static_cast<uint32_t>(i);
    if (!kWeighted) {
      // lazy_records: the suppression kernel builds the records of the windows it actually consumes (a hard scan stops
      // at num_post_nms kept boxes: 12 % of the candidates at the bench workload) from `src` -- this pass then neither
      // reads the 32-byte box rows nor writes the 32-byte records
      if (!lazy_records) {
        HardRec r;
        build.hard(i, r);
        static_cast<HardRec *>(recs)[pos] = r;
      }
    } else {
      WRec r;
      float d[9];
      build.weighted(i, r, d);
      static_cast<WRec *>(recs)[pos] = r;
      float *o = data + static_cast<size_t>(pos) * 9;
#pragma unroll
      for (int c = 0; c < 9; ++c) o[c] = d[c];
    }
  }
}

// ------------------------------------------------------------------------------------------
// the suppression kernel
// ------------------------------------------------------------------------------------------
// Weighted mode, hand-over from the per-segment scan to the grid-wide tail pass (wnms_tail_kernel): once a segment has
// kept num_post_nms boxes, every candidate behind the scan position still owes the kept boxes its merge contribution.
// The scan publishes the record range [pos_lo, pos_hi) that is left (whole score bins inside the num_pre_nms cut), the
// number of kept boxes and the geometry of its kept-box grid; the grid itself is in global memory by then.
// Buckets of the spatial grids: a TORUS of the cell grid, bucket = (iy mod 2^by) << bx | (ix mod 2^bx), both dimensions
// >= 32 cells.  A query window spans at most kMaxCellsPerQuery (25) cells per axis, so (1) its cells map to distinct
// buckets -- an entry is met at most once per query -- and (2) an entry that aliases into a visited bucket from another
// cell lies >= 32 - 12 - 1 = 19 cells away from the query while the query's reach is < 13 cells: it fails the circle
// test.  The circle test alone therefore decides, exactly; the walks need no "does this entry belong to the cell I am
// visiting" check (two float -> int conversions per hit with a multiplicative hash, where cells of one window could
// share a bucket).
__device__ __forceinline__ uint32_t torus_bucket(int ix, int iy, int bx, uint32_t bmask) {
  return ((static_cast<uint32_t>(iy) << bx) | (static_cast<uint32_t>(ix) & ((1u << bx) - 1u))) & bmask;
}
__device__ __forceinline__ int torus_bits_x(int n_buckets) { return (31 - __clz(n_buckets)) >> 1; }
static_assert(kMaxCellsPerQuery < 32 && kFrontBuckets == 1024 && kBucketsSmem == 4096, "torus dimensions (32 x 32, 64 x 64, >= 64 x 64)");

struct TailInfo {
  int pos_lo, pos_hi, kept, nos;
  float inv_cell, r_cap;
};

struct NmsArgs {
  TailInfo *tail;                 // [seg], zeroed per call; null: the scan kernel does the tail itself
  const float *lazy_boxes;        // hard mode, non-null: records are built per window from these (n, 8) rows via src[pos]
  const uint32_t *src;            // [pos] source row of a record
  const void *recs;               // HardRec / WRec, grouped by (segment, score bin)
  unsigned long long *skey;       // [pos] low key bits (score desc | candidate), the fine sort key
  uint32_t *gorder;               // scratch for bins larger than a window (sorted through global memory)
  const int *seg_begin, *seg_count;
  const int *bin_start;           // [seg][nb + 1], null when presorted
  int nb, presorted;              // presorted: position == rank (rv3d_wnms: the caller sorted)
  int num_pre, num_post;
  float thr, mthr;
  int prune;                      // 0: thresholds < 0 make even disjoint boxes interact -> test every pair
  int kept_stride;                // > 0: kept rows of segment s start at s * kept_stride; 0: at seg_begin[s]
  int *kept_pos;                  // [kept_base + i] = position inside the segment
  int *kept_count;                // [s]
  // kept-box grid in global memory (num_post_nms > kKeptSmem), per segment at kept_base / seg * n_buckets
  float4 *kxyr_g;
  int *knext_g, *heads_g;
  int n_buckets_g;
  int *kos;                       // [kept_base + i] kept boxes that are not in the grid (oversize / far away)
  // weighted merge
  const float *data;              // (n, D) rows in record order, score last
  int D;
  double *acc;                    // [(kept_base + i) * D + c]: c < D-1 weighted sums, c = D-1 weight sum
  int *merge_count;               // [kept_base + i]
  unsigned long long *stats;
  float rcap_mult, cell_mult;     // grid geometry in units of the mean padded radius
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
// The bound is on the TRUE IoU.  The reference routines work with absolute tolerances (1e-5 on dot products in m^2, 1e-6
// on cross products), which for boxes of a few decimetres become commensurate with the geometry: there the routine's
// value can exceed the true IoU by 10 % and more (tests/test_gpu_iou_decisions.py found 0.327 for a true 0.289 on
// 0.15 m x 0.07 m boxes).  The bound is therefore only trusted when both boxes have extents >= kBoundMinExtent, where
// those tolerances move the result by < 1e-4 relative; smaller boxes always run the exact routine.
constexpr float kBoundMinExtent = 0.5f;
template <typename Rec>
__device__ __forceinline__ bool iou_may_exceed(const Rec &ra, const Rec &rb, float thr) {
  const Obb a = obb_of(ra), b = obb_of(rb);
  if (!(fminf(fminf(fabsf(a.w), fabsf(a.h)), fminf(fabsf(b.w), fabsf(b.h))) >= kBoundMinExtent)) return true;
  const float inter = fminf(overlap_in_frame(a, b), overlap_in_frame(b, a));
  const float uni = fabsf(a.w * a.h) + fabsf(b.w * b.h) - inter;
  if (!(uni > 0.f) || !(inter == inter)) return true;  // degenerate / NaN: let the exact routine decide
  return (inter / uni) * 1.02f + 1e-5f >= thr;
}

// There is deliberately NO approximate IoU stage.  Round 1 decided ~98 % of the pairs with a float32 polygon clip and a
// +-(2 % + 1e-3) band around the threshold; tests/test_gpu_iou_decisions.py showed why that cannot be bit-exact: the
// detectron2 routine the reference binds is itself irregular -- for some near-parallel pairs its tolerance-based
// angular sort drops hull points and it returns e.g. 0.133 where the true IoU is 0.307 (and 0.307 with the boxes
// swapped).  Reproducing the reference's keep-set means reproducing that, so every pair that survives the circle
// test and the upper bound runs the bit-exact routine.

template <typename Rec, bool kWeighted, int kF>
__host__ __device__ inline size_t nms_smem_bytes(bool kept_in_smem) {
  constexpr int kFW = kF / 32;
  size_t b = 0;
  b += align_up_c(sizeof(int) * (kMaxBins + 1), 16);                // s_bins
  b += sizeof(uint32_t) * kWin;                                    // wpos
  b += sizeof(uint16_t) * kWin * 2;                                // survivor lists (double buffer)
  b += kWin;                                                       // walive
  b += sizeof(Rec) * kF;                                           // frec
  b += sizeof(float4) * kF;                                        // fq
  b += sizeof(int) * kFrontBuckets + sizeof(uint16_t) * kF;        // fheads, fos
  b += sizeof(uint16_t) * kF;                                      // front_w
  b += sizeof(uint32_t) * kF * kFW * (kWeighted ? 2 : 1);          // sup (+ mrg)
  b += sizeof(uint32_t) * kQ2Cap;                                  // queue2 (the window sort's keys + indices alias it)
  b += (kept_in_smem ? sizeof(uint16_t) : sizeof(uint32_t)) * kNmsThreads * kHitCap;   // hitbuf
  b += align_up_c(sizeof(uint16_t) * kF + sizeof(int16_t) * kF, 16);  // keptf, keptrank
  if (kWeighted) b += sizeof(int) * kWin + kQ2Cap + sizeof(int) * kF; // wfs, qflag, killer
  if (kept_in_smem) b += sizeof(float4) * kKeptSmem + sizeof(int) * kBucketsSmem + sizeof(int) * kKeptSmem;
  return align_up_c(b, 16);
}
static_assert(sizeof(unsigned long long) * kWin + sizeof(uint16_t) * kWin <= sizeof(uint32_t) * kQ2Cap, "sort buffers alias the queue");

__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int &total) {
  // kNmsThreads threads; s_warp has kNmsWarps + 1 ints
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
    int w = lane < kNmsWarps ? s_warp[lane] : 0;
    int inc2 = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc2, o);
      if (lane >= o) inc2 += t;
    }
    if (lane < kNmsWarps) s_warp[lane] = inc2 - w;
    if (lane == 31) s_warp[kNmsWarps] = inc2;
  }
  __syncthreads();
  total = s_warp[kNmsWarps];
  return s_warp[wid] + incl - v;
}

// Bitonic sort (normalised form: every compare-exchange is ascending, the first step of a stage mirrors), which
// sorts ANY length without padding: a partner index >= n stands for +inf and never moves.  keys / vals may live
// in shared or global memory; all kNmsThreads threads of the CTA call it.
template <typename V>
__device__ __forceinline__ void cta_bitonic_sort(unsigned long long *keys, V *vals, int n) {
  if (n < 2) return;
  int np2 = 2;
  while (np2 < n) np2 <<= 1;
  const int pairs = np2 >> 1;
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const bool mirror = (j == (k >> 1));
      for (int p = threadIdx.x; p < pairs; p += kNmsThreads) {
        int l, r;
        if (mirror) {
          const int blk = p / j, o = p - blk * j;
          l = blk * k + o;
          r = blk * k + (k - 1 - o);
        } else {
          l = 2 * p - (p & (j - 1));
          r = l + j;
        }
        if (r < n) {
          const unsigned long long a = keys[l], b = keys[r];
          if (a > b) {
            keys[l] = b; keys[r] = a;
            const V t = vals[l]; vals[l] = vals[r]; vals[r] = t;
          }
        }
      }
      __syncthreads();
    }
  }
}

template <typename Rec, bool kWeighted, int kF, bool kSm>
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_pull_kernel(NmsArgs a) {
  constexpr int kFW = kF / 32;   // words per bit-matrix row
  static_assert(kF <= 1024 && kFW <= 32 && kNmsThreads % kFW == 0 && kF <= kNmsThreads, "frontier size");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_warp[kNmsWarps + 1];
  __shared__ int s_qn, s_nk, s_nos, s_nfos, s_xn;
  __shared__ float s_red[kNmsWarps * 2];
  __shared__ float s_cell[2];   // inv_cell, r_cap
  __shared__ uint32_t s_haspred[kFW], s_removed[kFW];

  const int seg = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n_seg = a.seg_count[seg];
  const int n_use = min(n_seg, a.num_pre);   // top num_pre_nms by score (nms.py:29-32)
  if (n_use <= 0) {
    if (tid == 0) a.kept_count[seg] = 0;
    return;
  }
  const int beg = a.seg_begin[seg];
  const Rec *recs = static_cast<const Rec *>(a.recs) + beg;
  unsigned long long *skey = a.skey ? a.skey + beg : nullptr;
  uint32_t *gorder = a.gorder ? a.gorder + beg : nullptr;
  const int kbase = a.kept_stride > 0 ? seg * a.kept_stride : beg;
  int *kept_pos = a.kept_pos + kbase;
  const float thr_any = kWeighted ? fminf(a.thr, a.mthr) : a.thr;  // smallest IoU that matters
  const bool prune = a.prune != 0;

  // ---- carve shared memory
  unsigned char *p = smem_raw;
  int *s_bins = reinterpret_cast<int *>(p); p += align_up_c(sizeof(int) * (kMaxBins + 1), 16);
  uint32_t *wpos = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * kWin;
  uint16_t *surv_a = reinterpret_cast<uint16_t *>(p); p += sizeof(uint16_t) * kWin;
  uint16_t *surv_b = reinterpret_cast<uint16_t *>(p); p += sizeof(uint16_t) * kWin;
  uint8_t *walive = reinterpret_cast<uint8_t *>(p); p += kWin;
  Rec *frec = reinterpret_cast<Rec *>(p); p += sizeof(Rec) * kF;
  float4 *fq = reinterpret_cast<float4 *>(p); p += sizeof(float4) * kF;            // frontier (x, y, padded radius, next in chain)
  int *fheads = reinterpret_cast<int *>(p); p += sizeof(int) * kFrontBuckets;
  uint16_t *fos = reinterpret_cast<uint16_t *>(p); p += sizeof(uint16_t) * kF;     // frontier boxes that are not in the frontier grid
  uint16_t *front_w = reinterpret_cast<uint16_t *>(p); p += sizeof(uint16_t) * kF;
  uint32_t *sup = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * kF * kFW;
  uint32_t *mrg = sup;
  if (kWeighted) { mrg = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * kF * kFW; }
  uint32_t *queue2 = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * kQ2Cap;
  unsigned long long *wkey = reinterpret_cast<unsigned long long *>(queue2);        // window sort only
  uint16_t *widx = reinterpret_cast<uint16_t *>(wkey + kWin);
  // hits of the walk a thread is doing, [c * kNmsThreads + tid] (conflict-free); kept indices fit 16 bits in the kSm form
  using Hit = typename std::conditional<kSm, uint16_t, uint32_t>::type;
  Hit *hitbuf = reinterpret_cast<Hit *>(p) + tid;
  uint16_t *xq = reinterpret_cast<uint16_t *>(p);   // eval_queue's list of pairs for the exact routine: the walks' hits are in the queue by then
  p += sizeof(Hit) * kNmsThreads * kHitCap;
  uint16_t *keptf = reinterpret_cast<uint16_t *>(p);
  int16_t *keptrank = reinterpret_cast<int16_t *>(keptf + kF);
  p += align_up_c(sizeof(uint16_t) * kF + sizeof(int16_t) * kF, 16);
  int *wfs = nullptr; uint8_t *qflag = nullptr; int *killer = nullptr;
  if (kWeighted) {
    wfs = reinterpret_cast<int *>(p); p += sizeof(int) * kWin;       // window candidate -> first suppressor (kept index) in the current pull
    qflag = reinterpret_cast<uint8_t *>(p); p += kQ2Cap;             // per queued pair: bit0 iou > thr, bit1 iou > merge_thr
    killer = reinterpret_cast<int *>(p); p += sizeof(int) * kF;      // frontier box -> rank of its first suppressor
  }
  // Kept boxes: a spatial grid (torus_bucket) with chaining.  Shared-memory form (kSm, num_post_nms <= kKeptSmem): ONE 16-byte entry
  // per kept box, (x, y, padded radius, next + 1 << 20 | position in the segment), so a chain step is a single LDS.128.
  // Global form: (x, y, r, position) + a separate next array, for calls that may keep more than kKeptSmem boxes.
  float4 *kxyr; int *knext = nullptr, *heads, *kos; uint32_t bmask;
  if (kSm) {
    kxyr = reinterpret_cast<float4 *>(p); p += sizeof(float4) * kKeptSmem;
    heads = reinterpret_cast<int *>(p); p += sizeof(int) * kBucketsSmem;
    kos = reinterpret_cast<int *>(p); p += sizeof(int) * kKeptSmem;
    bmask = kBucketsSmem - 1;
    for (int i = tid; i < kBucketsSmem; i += kNmsThreads) heads[i] = -1;
  } else {
    kxyr = a.kxyr_g + kbase; knext = a.knext_g + kbase; kos = a.kos + kbase;
    heads = a.heads_g + static_cast<size_t>(seg) * a.n_buckets_g;   // pre-set to -1 by the host (memset 0xFF)
    bmask = static_cast<uint32_t>(a.n_buckets_g - 1);
  }
  if (!a.presorted)
    for (int i = tid; i <= a.nb; i += kNmsThreads) s_bins[i] = a.bin_start[static_cast<size_t>(seg) * (a.nb + 1) + i];
  if (tid == 0) { s_nos = 0; s_qn = 0; }
  __syncthreads();

  uint32_t st_iou = 0, st_circle = 0, st_hit = 0, st_bound = 0;   // per thread: exact IoUs, circle tests, pairs above thr, upper-bound tests
  // per-phase cycle counters (thread 0, only when stats are requested):
  // [0] window assembly + sort, [1] pull: grid walks, [2] pull: IoU, [3] frontier pairs, [4] greedy, [5] publish (+ weighted merges)
  long long ph[6] = {0, 0, 0, 0, 0, 0};
  long long t_mark = clock64();
  long long sub[3] = {0, 0, 0}, t_sub = 0;   // [0] frontier walk, [1] frontier evaluation, [2] pull walk + scan (subsets of ph[3], ph[3], ph[1])
  auto sub_begin = [&]() { if (a.stats && tid == 0) t_sub = clock64(); };
  auto sub_end = [&](int k) { if (a.stats && tid == 0) { const long long t = clock64(); sub[k] += t - t_sub; t_sub = t; } };
  auto lap = [&](int k) {
    if (a.stats && tid == 0) { const long long t = clock64(); ph[k] += t - t_mark; t_mark = t; }
  };

  int kept = 0;          // boxes kept so far (uniform)
  float surv_rate = 1.f; // share of the last pulled chunk that survived (uniform; sizes the next chunk)
  int rounds = 0;
  float inv_cell = 0.f, r_cap = 0.f;
  bool geom_set = false;

  auto cell_of = [&](float v) -> int { return static_cast<int>(fminf(fmaxf(floorf(v * inv_cell), -32768.f), 32767.f)); };
  const int kbx = kSm ? 6 : torus_bits_x(a.n_buckets_g);
  auto kept_bucket = [&](int ix, int iy) -> uint32_t { return torus_bucket(ix, iy, kbx, bmask); };
  auto front_bucket = [&](int ix, int iy) -> uint32_t { return torus_bucket(ix, iy, 5, kFrontBuckets - 1); };
  auto in_grid = [&](float x, float y, float r) -> bool {
    return (r <= r_cap) && (fabsf(x) <= kPosCap) && (fabsf(y) <= kPosCap);   // false for NaN
  };
  auto kept_entry = [&](int k, float &x, float &y, float &r, int &next) {
    const float4 q = kxyr[k];
    x = q.x; y = q.y; r = q.z;
    if (kSm) next = static_cast<int>(__float_as_uint(q.w) >> 20) - 1;
    else next = knext[k];
  };
  auto kept_position = [&](int k) -> int {
    const uint32_t w = __float_as_uint(kxyr[k].w);
    return static_cast<int>(kSm ? (w & 0xfffffu) : w);
  };
  // does the circle (x, y, r) touch (qx, qy, qr)?  (the pads of iou.cuh padded_radius make "no" exact)
  auto touches = [&](float x, float y, float r, float qx, float qy, float qr) -> bool {
    const float dx = qx - x, dy = qy - y, rr = qr + r;
    return dx * dx + dy * dy <= rr * rr;
  };

  // Query window of a box in cell coordinates; false -> the caller scans everything (pruning off, a radius that is
  // infinite or covers too many cells).  NaN boxes never interact: `skip`.
  auto query_cells = [&](float x, float y, float r, int &ix0, int &ix1, int &iy0, int &iy1, bool &skip) -> bool {
    skip = false;
    if (!prune) return false;
    if (!(x == x) || !(y == y) || !(r == r)) { skip = true; return false; }   // its circle test is false against anything
    const float reach = r + r_cap;
    if (!(reach <= kPosCap) || !(fabsf(x) <= kPosCap) || !(fabsf(y) <= kPosCap)) return false;
    ix0 = cell_of(x - reach); ix1 = cell_of(x + reach); iy0 = cell_of(y - reach); iy1 = cell_of(y + reach);
    return (ix1 - ix0 + 1) * (iy1 - iy0 + 1) <= kMaxCellsPerQuery;
  };

  // Every kept box k >= since whose padded circle touches (x, y, r): fn(k).  Chains are newest-first (insertion
  // prepends, kept indices only grow), so a walk stops at the first index below `since`.  The circle test alone decides
  // (torus_bucket: the cells of one window never share a bucket, aliases from other cells are too far away to touch).
  auto for_each_near = [&](float x, float y, float r, int since, auto &&fn) {
    int ix0 = 0, ix1 = 0, iy0 = 0, iy1 = 0;
    bool skip;
    if (!query_cells(x, y, r, ix0, ix1, iy0, iy1, skip)) {
      if (skip) return;
      for (int k = since; k < kept; ++k) {
        const float4 q = kxyr[k];
        ++st_circle;
        if (!prune || touches(x, y, r, q.x, q.y, q.z)) fn(k);
      }
      return;
    }
    for (int iy = iy0; iy <= iy1; ++iy)
      for (int ix = ix0; ix <= ix1; ++ix) {
        int k = heads[kept_bucket(ix, iy)];
        while (k >= since) {
          float qx, qy, qr; int next;
          kept_entry(k, qx, qy, qr, next);
          ++st_circle;
          if (touches(x, y, r, qx, qy, qr)) fn(k);
          k = next;
        }
      }
    const int nos = s_nos;
    for (int o = 0; o < nos; ++o) {
      const int k = kos[o];
      if (k >= since) {
        const float4 q = kxyr[k];
        ++st_circle;
        if (touches(x, y, r, q.x, q.y, q.z)) fn(k);
      }
    }
  };

  auto accumulate = [&](int slot, int pos) {  // merge the candidate at `pos` into kept slot
    const float *row = a.data + static_cast<size_t>(beg + pos) * a.D;
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

  // bound -> exact for ONE pair, in place (overflow paths; divergent, so only a fallback)
  auto classify_pair = [&](const Rec &ra, const Rec &rb, bool &above, bool &above_m) {
    if (prune && !iou_may_exceed(ra, rb, thr_any)) { above = false; above_m = false; return; }
    const float iou = pair_iou(ra, rb);
    ++st_iou;
    above = iou > a.thr;
    above_m = kWeighted && iou > a.mthr;
  };

  // Evaluation of a queue of pairs.  Phase 1: every lane tests its pair against the cheap upper bound; the pairs that may
  // exceed a threshold go to a CTA-wide list (xq, warp-aggregated reservation).  Phase 2: the bit-exact routine runs over
  // that list on dense lanes, 32 pairs per warp.  The routine is a ~2000-instruction dependent chain, so what an
  // evaluation costs is the number of times the busiest warp has to run it: ceil(pending / (32 * warps)) with the shared
  // list -- once for up to 512 pairs -- where per-warp lists ran it twice as soon as one warp's share exceeded 32.
  // get(q, ra, rb) loads the pair, emit(q, above_thr, above_mthr) consumes the two comparisons (only called for pairs that
  // reach the exact routine: a pair stopped by the bound is below both thresholds).  Callers sync before and after.
  constexpr int kXq = kNmsThreads * kHitCap;   // uint16 entries that fit the hit buffer in either form
  static_assert(kXq % 32 == 0 && kQ2Cap <= 65536, "queue slices are whole warps; queue indices fit 16 bits");
  auto eval_queue = [&](int qn, auto &&get, auto &&emit) {
    for (int q_lo = 0; q_lo < qn; q_lo += kXq) {   // (one slice unless more than kXq pairs are queued)
      const int q_hi = min(qn, q_lo + kXq);
      if (tid == 0) s_xn = 0;
      __syncthreads();
      const int q_pad = q_lo + ((q_hi - q_lo + 31) & ~31);
      for (int q = q_lo + tid; q < q_pad; q += kNmsThreads) {   // warp-uniform trip count
        bool pending = q < q_hi;
        if (pending && prune) {
          Rec ra, rb;
          get(q, ra, rb);
          ++st_bound;
          pending = iou_may_exceed(ra, rb, thr_any);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, pending);
        if (m) {
          int at = 0;
          if (lane == 0) at = atomicAdd(&s_xn, __popc(m));
          at = __shfl_sync(0xffffffffu, at, 0);
          if (pending) xq[at + __popc(m & ((1u << lane) - 1u))] = static_cast<uint16_t>(q);
        }
      }
      __syncthreads();
      const int xn = s_xn;
      for (int x = tid; x < xn; x += kNmsThreads) {
        const int q = static_cast<int>(xq[x]);
        Rec ra, rb;
        get(q, ra, rb);
        const float iou = pair_iou(ra, rb);
        ++st_iou;
        emit(q, iou > a.thr, kWeighted && iou > a.mthr);
      }
      if (q_hi < qn) __syncthreads();   // the next slice reuses xq and the counter
    }
  };

  // ---- PULL: the candidates list[0, ns) (window indices, rank order) against the kept boxes [since, kept).
  // One candidate per thread: walk the grid once (the first kHitCap hits go to the thread's slots of hitbuf), block scan of the hit
  // counts, write the (candidate, kept) pairs to the queue at the scanned offsets, evaluate the queue on dense lanes.
  // A batch whose pairs do not fit the queue is cut at the last thread that fits (the offsets are a prefix sum, so the
  // threads that fit form a prefix of the batch).
  // merges_only: the scan is over (weighted, num_post_nms reached) -- candidates only contribute to merge sets.
  // The list has two zones: entries [0, n_a) have met the kept boxes below since_a already, the entries behind them
  // none of the kept boxes (lazy pulls, see the round loop).
  auto pull = [&](const uint16_t *list, int ns, int n_a, int since_a, bool merges_only) {
    if (kept <= 0 || ns <= 0 || (n_a >= ns && since_a >= kept)) return;
    int bi = 0;
    int base = 0;   // pairs waiting in the queue (uniform).  Entries are (window index << 20 | kept index).
    // The batches of a pull only ENQUEUE; the queue is evaluated once at the end (or when it is full): the bit-exact
    // routine is a ~2000-instruction dependent chain, so every evaluation costs its latency once however few pairs it
    // has -- one evaluation per pull instead of one per batch of kNmsThreads candidates, on fuller warps.
    auto flush = [&]() {
      const int qn = base;
      auto get = [&](int q, Rec &ra, Rec &rb) {
        const uint32_t e = queue2[q];
        ra = recs[kept_position(static_cast<int>(e & 0xfffffu))];   // the kept box ranks higher: box1 of the routine
        rb = recs[wpos[e >> 20]];
      };
      if (!kWeighted) {
        eval_queue(qn, get, [&](int q, bool above, bool) {
          if (above) { ++st_hit; walive[queue2[q] >> 20] = 0; }
        });
        __syncthreads();
      } else {
        for (int q = tid; q < qn; q += kNmsThreads) qflag[q] = 0;
        __syncthreads();
        // pass 1: the two comparisons of every queued pair; each candidate's FIRST suppressor
        eval_queue(qn, get, [&](int q, bool above, bool above_m) {
          qflag[q] = static_cast<uint8_t>((above ? 1u : 0u) | (above_m ? 2u : 0u));
          if (above) atomicMin(&wfs[queue2[q] >> 20], static_cast<int>(queue2[q] & 0xfffffu));
        });
        __syncthreads();
        // pass 2: merges up to and including the first suppressor; a suppressed candidate leaves the window
        for (int q = tid; q < qn; q += kNmsThreads) {
          const uint32_t e = queue2[q];
          const int jj = static_cast<int>(e >> 20), k = static_cast<int>(e & 0xfffffu);
          const int fs = wfs[jj];
          if (k > fs) continue;
          if (qflag[q] & 2u) accumulate(k, static_cast<int>(wpos[jj]));
          if (k == fs && !merges_only) walive[jj] = 0;
        }
        __syncthreads();
      }
      lap(2);
      base = 0;
    };
    while (bi < ns) {
      sub_begin();
      const int t = bi + tid;
      const bool active = t < ns;
      const int since = t < n_a ? since_a : 0;
      int j = 0, pos = 0, cnt = 0;
      float x = 0.f, y = 0.f, r = 0.f;
      if (active) {
        j = list[t];
        pos = static_cast<int>(wpos[j]);
        x = rec_cx(recs[pos]); y = rec_cy(recs[pos]); r = recs[pos].r;
        if (since < kept)
          for_each_near(x, y, r, since, [&](int k) {
            if (cnt < kHitCap) hitbuf[cnt * kNmsThreads] = static_cast<Hit>(k);
            ++cnt;
          });
      }
      int total;
      const int off = base + block_exclusive_scan(cnt, s_warp, total);
      sub_end(2);
      const bool ok = active && (off + cnt <= kQ2Cap);
      const int m = __syncthreads_count(ok);   // the offsets are a prefix sum: the threads that fit are exactly tid < m
      if (m == 0) {
        if (base > 0) { flush(); continue; }   // retry the batch against an empty queue
        // the first candidate alone overflows the queue: thread 0 evaluates its pairs in place
        if (tid == 0) {
          const Rec rj = recs[pos];
          int fs = 0x7fffffff;
          for_each_near(x, y, r, since, [&](int k) {
            if (!kWeighted && fs != 0x7fffffff) return;
            bool above, above_m;
            classify_pair(recs[kept_position(k)], rj, above, above_m);
            if (above && k < fs) fs = k;
          });
          if (kWeighted) {
            for_each_near(x, y, r, since, [&](int k) {
              if (k > fs) return;
              bool above, above_m;
              classify_pair(recs[kept_position(k)], rj, above, above_m);
              if (above_m) accumulate(k, pos);
            });
          }
          if (fs != 0x7fffffff) { walive[j] = 0; ++st_hit; }
        }
        __syncthreads();
        bi += 1;
        continue;
      }
      if (ok) {
        if (kWeighted) wfs[j] = 0x7fffffff;
        const uint32_t tag = static_cast<uint32_t>(j) << 20;
        if (cnt <= kHitCap) {
          for (int c = 0; c < cnt; ++c) queue2[off + c] = tag | static_cast<uint32_t>(hitbuf[c * kNmsThreads]);
        } else {
          int w = off;
          for_each_near(x, y, r, since, [&](int k) { queue2[w++] = tag | static_cast<uint32_t>(k); });
        }
        if (tid == m - 1) s_qn = off + cnt;
      }
      __syncthreads();
      lap(1);
      base = s_qn;
      const bool cut = m < min(kNmsThreads, ns - bi);   // a candidate of this batch did not fit: evaluate, then go on with it
      bi += m;
      if (cut) flush();
    }
    if (base > 0) flush();
  };

  // order-preserving compaction of a survivor list by walive
  auto compact = [&](const uint16_t *in, int ns, uint16_t *out) -> int {
    const int per = (ns + kNmsThreads - 1) / kNmsThreads;
    const int lo = min(ns, tid * per), hi = min(ns, lo + per);
    int c = 0;
    for (int t = lo; t < hi; ++t) c += walive[in[t]] ? 1 : 0;
    int total;
    int off = block_exclusive_scan(c, s_warp, total);
    for (int t = lo; t < hi; ++t) {
      const uint16_t j = in[t];
      if (walive[j]) out[off++] = j;
    }
    __syncthreads();
    return total;
  };

  // ---- window assembly state (uniform)
  int rank_base = 0;                 // candidates consumed so far == rank of the next window's first candidate
  int bin_cur = 0;                   // next score bin
  int chunk_lo = 0, chunk_hi = 0;    // pending part of a bin larger than a window (sorted through global memory)
  bool done = false;                 // num_post_nms reached

  // next window -> wpos[0, wn) in rank order; returns wn (0: no candidates left)
  auto next_window = [&](bool need_order) -> int {
    int wn = 0;
    if (a.presorted) {
      wn = min(kWin, n_use - rank_base);
      for (int t = tid; t < wn; t += kNmsThreads) wpos[t] = static_cast<uint32_t>(rank_base + t);
      __syncthreads();
      return wn;
    }
    for (;;) {
      if (chunk_hi > chunk_lo) {
        wn = min(kWin, chunk_hi - chunk_lo);
        for (int t = tid; t < wn; t += kNmsThreads) wpos[t] = gorder[chunk_lo + t];
        chunk_lo += wn;
        __syncthreads();
        break;
      }
      if (bin_cur >= a.nb) return 0;
      const int b0 = bin_cur;
      const int base = s_bins[b0];
      int lo = b0, hi = a.nb;                // largest b with s_bins[b] - base <= kWin (every thread, same result)
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_bins[mid] - base <= kWin) lo = mid; else hi = mid - 1;
      }
      if (lo == b0) {
        // this bin alone exceeds a window (massively tied / saturated scores): sort it once through global memory
        const int m = s_bins[b0 + 1] - base;
        for (int t = tid; t < m; t += kNmsThreads) gorder[base + t] = static_cast<uint32_t>(base + t);
        __syncthreads();
        cta_bitonic_sort(skey + base, gorder + base, m);
        chunk_lo = base; chunk_hi = base + m;
        ++bin_cur;
        continue;
      }
      wn = s_bins[lo] - base;
      bin_cur = lo;
      if (wn == 0) continue;                 // only empty bins: go on (ends at bin_cur == nb)
      // the order inside the window only matters when candidates are consumed in rank order (need_order) or the
      // num_pre_nms cut falls inside it
      if (need_order || rank_base + wn > n_use) {
        for (int t = tid; t < wn; t += kNmsThreads) wkey[t] = skey[base + t];
        // The bins are already in rank order: only the order INSIDE each bin is missing.  Small bins (the rule): every
        // element counts the smaller keys of its own bin and goes straight to its place -- no dependent exchanges.
        // A bin above kRankSortMax sends the whole window to the bitonic network instead.
        bool big = false;
        for (int b = b0 + tid; b < lo; b += kNmsThreads) big |= (s_bins[b + 1] - s_bins[b]) > kRankSortMax;
        if (!__syncthreads_or(big)) {
          for (int t = tid; t < wn; t += kNmsThreads) {
            int bl = b0, bh = lo - 1;        // the bin of element t: largest b with s_bins[b] - base <= t
            while (bl < bh) {
              const int mid = (bl + bh + 1) >> 1;
              if (s_bins[mid] - base <= t) bl = mid; else bh = mid - 1;
            }
            const int e0 = s_bins[bl] - base, e1 = s_bins[bl + 1] - base;
            const unsigned long long mine = wkey[t];
            int rank = 0;
            for (int e = e0; e < e1; ++e) rank += wkey[e] < mine ? 1 : 0;
            wpos[e0 + rank] = static_cast<uint32_t>(base + t);
          }
        } else {
          for (int t = tid; t < wn; t += kNmsThreads) widx[t] = static_cast<uint16_t>(t);
          __syncthreads();
          cta_bitonic_sort(wkey, widx, wn);
          for (int t = tid; t < wn; t += kNmsThreads) wpos[t] = static_cast<uint32_t>(base + widx[t]);
        }
      } else {
        for (int t = tid; t < wn; t += kNmsThreads) wpos[t] = static_cast<uint32_t>(base + t);
      }
      __syncthreads();
      break;
    }
    return min(wn, n_use - rank_base);
  };

  while (rank_base < n_use && !done) {
    const int wn = next_window(true);
    if (wn <= 0) break;
    if (!kWeighted && a.lazy_boxes) {
      // records of this window, built here instead of by the bucketing pass for every candidate of the segment
      HardRec *recs_w = const_cast<HardRec *>(reinterpret_cast<const HardRec *>(recs));
      const RecFromBoxes8 build{a.lazy_boxes};
      for (int t = tid; t < wn; t += kNmsThreads) {
        const uint32_t pos = wpos[t];
        HardRec r;
        build.hard(static_cast<int>(a.src[beg + pos]), r);
        recs_w[pos] = r;
      }
      __syncthreads();
    }
    if (!geom_set) {
      // cell size of the grids from the padded radii of the first (top-scored) window
      float sr = 0.f, sc = 0.f;
      for (int t = tid; t < wn; t += kNmsThreads) {
        const float r = recs[wpos[t]].r;
        if (r > 0.f && r < 1.0e4f) { sr += r; sc += 1.f; }
      }
      for (int o = 16; o; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); sc += __shfl_xor_sync(0xffffffffu, sc, o); }
      if (lane == 0) { s_red[wid] = sr; s_red[kNmsWarps + wid] = sc; }
      __syncthreads();
      if (tid == 0) {
        float tr = 0.f, tc = 0.f;
        for (int w = 0; w < kNmsWarps; ++w) { tr += s_red[w]; tc += s_red[kNmsWarps + w]; }
        const float mr = tc > 0.f ? tr / tc : 1.f;
        s_cell[1] = a.rcap_mult * mr;                                      // r_cap: larger boxes go to the oversize lists
        s_cell[0] = 1.0f / fminf(fmaxf(a.cell_mult * mr, 1e-3f), 1.0e5f);  // cell = mean radius + r_cap: a typical query spans 3 x 3 cells
      }
      __syncthreads();
      inv_cell = s_cell[0]; r_cap = s_cell[1];
      geom_set = true;
    }
    for (int t = tid; t < wn; t += kNmsThreads) { surv_a[t] = static_cast<uint16_t>(t); walive[t] = 1; }
    __syncthreads();
    lap(0);
    uint16_t *list = surv_a, *other = surv_b;
    // list[0, ns): the window's candidates that are still alive, in rank order.  Zone A = [0, n_a) has met the kept
    // boxes below since_a, zone B = [n_a, ns) none yet.  Pulls are LAZY: a round only pulls what its frontier needs --
    // the leftovers of zone A and the next chunk of zone B -- and the frontier is only as large as the room under
    // num_post_nms can use.  A scan that stops at num_post_nms kept boxes (nms.py:53-56) never touches the rest of its
    // last window (at the bench workload: 2436 -> ~900 candidate pulls and 3 x 512 -> 512 + 512 + ~150 frontier boxes
    // per segment).  Every candidate still meets every higher-ranked kept box before it can be kept: same result.
    int ns = wn, n_a = 0, since_a = 0;

    while (true) {
      const int room = a.num_post - kept;
      const int nf_want = min(kF, room + (room >> 2) + 16);   // a later round tops up if fewer than `room` are kept
      // ================= a / c. pull against the kept boxes, drop the suppressed =================
      if (kept > 0) {
        // chunk: what the frontier wants, scaled by the share of the previous chunk that survived its pull (dense scenes
        // lose most of a chunk to the kept boxes: an under-filled frontier means more rounds) -- but not a second,
        // mostly idle batch of the pull (one candidate per thread) for a few more candidates
        int want = static_cast<int>(static_cast<float>(nf_want) * 1.1f / surv_rate) + 32;
        if (want <= kNmsThreads + kNmsThreads / 2) want = min(want, max(nf_want, kNmsThreads));
        want = min(want, kWin);
        const int np = min(ns, max(n_a, want));
        pull(list, np, n_a, since_a, false);
        const int ns2 = compact(list, np, other);
        if (np >= 64) surv_rate = fmaxf(static_cast<float>(ns2) / static_cast<float>(np), 0.05f);
        for (int t = np + tid; t < ns; t += kNmsThreads) other[ns2 + t - np] = list[t];   // zone B moves up behind the survivors
        __syncthreads();
        uint16_t *tmp = list; list = other; other = tmp;
        ns = ns2 + (ns - np);
        n_a = ns2;
        since_a = kept;
        if (ns == 0) break;
        if (n_a == 0) continue;   // the whole chunk was suppressed: next chunk
      } else {
        n_a = min(ns, nf_want);   // no kept box yet: nothing to meet; the frontier is zone A, the rest stays zone B
      }
      if (ns == 0) break;
      ++rounds;
      // ================= b. frontier = the first nf survivors of zone A; interacting pairs inside it =================
      const int nf = min(nf_want, n_a);
      for (int i = tid; i < kFrontBuckets; i += kNmsThreads) fheads[i] = -1;
      for (int i = tid; i < nf * kFW; i += kNmsThreads) {
        sup[i] = 0u;
        if (kWeighted) mrg[i] = 0u;
      }
      if (tid < kFW) { s_haspred[tid] = 0u; s_removed[tid] = 0u; }
      if (tid == 0) { s_qn = 0; s_nfos = 0; }
      float mx = 0.f, my = 0.f, mrad = 0.f;
      if (tid < nf) {
        const int j = list[tid];
        const Rec r = recs[wpos[j]];
        frec[tid] = r;
        mx = rec_cx(r); my = rec_cy(r); mrad = r.r;
        front_w[tid] = static_cast<uint16_t>(j);
        keptrank[tid] = -1;
      }
      __syncthreads();
      // frontier grid: same cells as the kept grid, chains of frontier slots
      if (tid < nf) {
        int next = -1;
        if (prune) {
          if (in_grid(mx, my, mrad)) next = atomicExch(&fheads[front_bucket(cell_of(mx), cell_of(my))], tid);
          else fos[atomicAdd(&s_nfos, 1)] = static_cast<uint16_t>(tid);
        }
        fq[tid] = make_float4(mx, my, mrad, __int_as_float(next));
      }
      __syncthreads();
      {
        auto mark = [&](int i, int j, bool above, bool above_m) {
          if (above) { ++st_hit; atomicOr(&sup[i * kFW + (j >> 5)], 1u << (j & 31)); }
          if (kWeighted && above_m) atomicOr(&mrg[i * kFW + (j >> 5)], 1u << (j & 31));
        };
        sub_begin();
        // Thread i collects the touching pairs it owns: the cells around it, plus the frontier's oversize list.  Count
        // first (the first kHitCap hits go to hitbuf), block scan, write the pairs at the scanned offsets: the queue is
        // ordered by the enumerating box, so the evaluation's record loads are mostly warp-uniform (with one shared
        // slot counter the pairs arrived in random order: evaluation 5.0 -> 3.5 Mcycles per step, greedy 1.26 -> 0.68).
        static_assert(kQ2Cap >= kF, "one frontier box's pairs always fit the queue");
        // Every touching pair is enumerated by exactly ONE of its two boxes, its owner: for two boxes that are both in
        // the grid, the one in the lexicographically smaller cell (row, column; same cell: the smaller slot) -- so a
        // box only walks its own cell and the cells AFTER it in its window (5 - 6 of 9 instead of all 9, and half the
        // entries); if either box is on the oversize list, the smaller slot.  A touching in-grid partner's cell always
        // lies inside the owner's window (reach = r + r_cap >= r + r_partner).
        auto walk_front = [&](int i, auto &&fn) {
          int ix0 = 0, ix1 = 0, iy0 = 0, iy1 = 0;
          bool skip;
          const bool gridded = prune && in_grid(mx, my, mrad);
          const int cix = cell_of(mx), ciy = cell_of(my);
          if (!query_cells(mx, my, mrad, ix0, ix1, iy0, iy1, skip)) {
            if (skip) return;
            if (!gridded) {
              for (int j = i + 1; j < nf; ++j) {
                const float4 q = fq[j];
                ++st_circle;
                if (!prune || touches(mx, my, mrad, q.x, q.y, q.z)) fn(j);
              }
            } else {   // (window larger than kMaxCellsPerQuery: only with an unusual grid geometry) same ownership rule, linear scan
              for (int j = 0; j < nf; ++j) {
                if (j == i) continue;
                const float4 q = fq[j];
                ++st_circle;
                if (!touches(mx, my, mrad, q.x, q.y, q.z)) continue;
                bool own = j > i;
                if (in_grid(q.x, q.y, q.z)) {
                  const int cjx = cell_of(q.x), cjy = cell_of(q.y);
                  own = cjy > ciy || (cjy == ciy && (cjx > cix || (cjx == cix && j > i)));
                }
                if (own) fn(j);
              }
            }
          } else {
            for (int iy = gridded ? ciy : iy0; iy <= iy1; ++iy)
              for (int ix = (gridded && iy == ciy) ? cix : ix0; ix <= ix1; ++ix) {
                const int jmin = (!gridded || (iy == ciy && ix == cix)) ? i : -1;   // own cell (or an oversize walker): larger slots only
                int j = fheads[front_bucket(ix, iy)];
                while (j >= 0) {
                  const float4 q = fq[j];
                  ++st_circle;
                  if (j > jmin && touches(mx, my, mrad, q.x, q.y, q.z)) fn(j);
                  j = __float_as_int(q.w);
                }
              }
            const int nfos = s_nfos;
            for (int o = 0; o < nfos; ++o) {
              const int j = fos[o];
              if (j > i) {
                const float4 q = fq[j];
                ++st_circle;
                if (touches(mx, my, mrad, q.x, q.y, q.z)) fn(j);
              }
            }
          }
        };
        int f0 = 0;   // frontier boxes below f0 have had their pairs evaluated (one pass unless the queue overflows)
        while (f0 < nf) {
          const bool active = tid >= f0 && tid < nf;
          int cnt = 0;
          if (active)
            walk_front(tid, [&](int j) {
              if (cnt < kHitCap) hitbuf[cnt * kNmsThreads] = static_cast<Hit>(j);
              ++cnt;
            });
          int total;
          const int off = block_exclusive_scan(cnt, s_warp, total);
          const bool ok = active && (off + cnt <= kQ2Cap);
          const int m = __syncthreads_count(ok);   // prefix sum: the threads that fit are exactly f0 <= tid < f0 + m, m >= 1
          if (ok) {
            auto entry = [&](int j) -> uint32_t {   // (higher-ranked slot << 10) | lower-ranked slot
              return static_cast<uint32_t>(j > tid ? ((tid << 10) | j) : ((j << 10) | tid));
            };
            if (cnt <= kHitCap) {
              for (int c = 0; c < cnt; ++c) queue2[off + c] = entry(static_cast<int>(hitbuf[c * kNmsThreads]));
            } else {
              int w = off;
              walk_front(tid, [&](int j) { queue2[w++] = entry(j); });
            }
            if (tid == f0 + m - 1) s_qn = off + cnt;
          }
          __syncthreads();
          sub_end(0);
          eval_queue(s_qn,
                     [&](int q, Rec &ra, Rec &rb) { ra = frec[queue2[q] >> 10]; rb = frec[queue2[q] & 1023]; },
                     [&](int q, bool above, bool above_m) {
                       mark(static_cast<int>(queue2[q] >> 10), static_cast<int>(queue2[q] & 1023), above, above_m);
                     });
          __syncthreads();
          sub_end(1);
          f0 += m;
        }
      }
      lap(3);
      // ================= greedy resolution of the frontier =================
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
        uint32_t valid = 0u, removed = 0u, pending = 0u, keptm = 0u;
        if (lane < kFW) {
          const int lo = lane << 5;
          valid = (nf >= lo + 32) ? 0xffffffffu : (nf <= lo ? 0u : ((1u << (nf - lo)) - 1u));
          removed = s_removed[lane];
          pending = valid & s_haspred[lane];       // must be visited in rank order
          keptm = valid & ~s_haspred[lane];
        }
        while (true) {
          const uint32_t cand = pending & ~removed;
          const uint32_t have = __ballot_sync(0xffffffffu, cand != 0u);
          if (!have) break;
          const int wsel = __ffs(have) - 1;
          const uint32_t cw = __shfl_sync(0xffffffffu, cand, wsel);
          const int bit = __ffs(cw) - 1;
          const int i = (wsel << 5) + bit;
          if (lane == wsel) keptm |= 1u << bit;
          if (lane < kFW) {
            // everything up to and including i is decided now
            const int lo = lane << 5;
            if (i >= lo + 31) pending = 0u;
            else if (i >= lo) pending &= ~((2u << (i - lo)) - 1u);
            removed |= sup[i * kFW + lane];
          }
        }
        // kept boxes in rank order, truncated to the room left under num_post_nms
        const int room = a.num_post - kept;
        const int cnt = lane < kFW ? __popc(keptm) : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < kFW; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        int rank = incl - cnt;
        if (lane < kFW) {
          uint32_t m = keptm;
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
      const int nk = s_nk;
      lap(4);

      // ================= publish the newly kept boxes: output list + kept-box grid =================
      if (tid < nk) {
        const int fi = keptf[tid];
        const int pos = static_cast<int>(wpos[front_w[fi]]);
        const int k = kept + tid;
        kept_pos[k] = pos;
        const float4 q = fq[fi];
        int next = -1;
        if (in_grid(q.x, q.y, q.z)) next = atomicExch(&heads[kept_bucket(cell_of(q.x), cell_of(q.y))], k);
        else kos[atomicAdd(&s_nos, 1)] = k;
        if (kSm) {
          kxyr[k] = make_float4(q.x, q.y, q.z, __uint_as_float((static_cast<uint32_t>(next + 1) << 20) | static_cast<uint32_t>(pos)));
        } else {
          kxyr[k] = make_float4(q.x, q.y, q.z, __int_as_float(pos));
          knext[k] = next;
        }
      }
      if (kWeighted) {
        // merge sets inside the frontier: candidate j joins every kept i < j with iou > merge_thr that comes no
        // later than its first suppressor; kept boxes join themselves.  (1) first suppressor of every frontier box
        // = min rank over the kept rows that contain it; (2) each kept row queues itself and the boxes of its merge
        // row that it reaches no later than their first suppressor; (3) the queue is accumulated one pair per lane.
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
            else accumulate(kept + tid, static_cast<int>(wpos[front_w[j]]));   // (<= kF * kF / 2 pairs: only if every pair merges)
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
          accumulate(kept + static_cast<int>(queue2[q] & 0xffffu), static_cast<int>(wpos[front_w[queue2[q] >> 16]]));
      }
      __syncthreads();
      lap(5);
      since_a = kept;             // the leftovers of zone A have met every kept box before this frontier
      kept += nk;
      list += nf;
      ns -= nf;
      n_a -= nf;
      // nms.py:53-56: only the first num_post_nms kept survive, so the scan can stop there
      if (kept >= a.num_post) { done = true; break; }
      if (ns == 0) break;
    }
    if (kWeighted && done) {
      // the candidates behind the last frontier still owe the kept boxes they have not met their merge contributions
      // (zone A: the last frontier's; zone B: all of them)
      pull(list, ns, n_a, since_a, true);
    }
    rank_base += wn;
  }
  if (kWeighted && done) {
    // ... and so do all candidates behind this window: one pull each against every kept box, merges only.
    auto tail_windows = [&](bool chunk_only) {
      while (rank_base < n_use && (!chunk_only || chunk_hi > chunk_lo)) {
        const int wn = next_window(false);
        if (wn <= 0) break;
        for (int t = tid; t < wn; t += kNmsThreads) surv_a[t] = static_cast<uint16_t>(t);
        __syncthreads();
        lap(0);
        pull(surv_a, wn, 0, 0, true);
        rank_base += wn;
      }
    };
    if (a.tail) {
      // One CTA doing this for a long tail is the whole call's critical path (200 k candidates: 39 ms), while the
      // candidates are independent of each other now: the keep-set is final.  What is left in whole score bins inside
      // the num_pre_nms cut goes to the grid-wide wnms_tail_kernel; this CTA only finishes a giant bin it is in the
      // middle of, and afterwards the one bin the cut falls into (whose members need their exact rank).
      tail_windows(true);
      if (rank_base < n_use) {
        int pos_lo, pos_hi, b_cut = 0;
        if (a.presorted) {
          pos_lo = rank_base; pos_hi = n_use;
        } else {
          int lo = bin_cur, hi = a.nb;             // largest b with s_bins[b] <= n_use (every thread, same result)
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_bins[mid] <= n_use) lo = mid; else hi = mid - 1;
          }
          b_cut = lo; pos_lo = s_bins[bin_cur]; pos_hi = s_bins[b_cut];
        }
        if (pos_hi > pos_lo) {
          if (kSm) {                                 // the kept-box grid leaves shared memory
            for (int i = tid; i < kept; i += kNmsThreads) a.kxyr_g[kbase + i] = kxyr[i];
            int *hg = a.heads_g + static_cast<size_t>(seg) * kBucketsSmem;
            for (int i = tid; i < kBucketsSmem; i += kNmsThreads) hg[i] = heads[i];
            const int nos = s_nos;
            for (int i = tid; i < nos; i += kNmsThreads) a.kos[kbase + i] = kos[i];
          }
          if (tid == 0) a.tail[seg] = TailInfo{pos_lo, pos_hi, kept, s_nos, inv_cell, r_cap};
          if (!a.presorted) bin_cur = b_cut;
          rank_base += pos_hi - pos_lo;
          __syncthreads();
        }
      }
    }
    tail_windows(false);
  }

  if (tid == 0) a.kept_count[seg] = kept;
#ifdef RV3D_NMS_DEBUG_PRINT   // per-segment figures for tools/dev_nms_segments.py (RV3D_NVCC_DEFS=-DRV3D_NMS_DEBUG_PRINT)
  if (tid == 0 && a.stats)
    printf("seg %d n %d consumed %d kept %d rounds %d cycles %lld ph %lld %lld %lld %lld %lld sub %lld %lld %lld\n", seg, n_seg, rank_base, kept, rounds,
           ph[0] + ph[1] + ph[2] + ph[3] + ph[4] + ph[5], ph[0], ph[1], ph[2], ph[3], ph[4], sub[0], sub[1], sub[2]);
#endif
  if (a.stats) {
    // warp-reduce (widened: the per-thread counters are 32-bit to keep the inner loops at one IADD) then one atomic per warp
    unsigned long long w_iou = st_iou, w_circle = st_circle, w_hit = st_hit, w_bound = st_bound;
    for (int o = 16; o; o >>= 1) {
      w_iou += __shfl_xor_sync(0xffffffffu, w_iou, o);
      w_circle += __shfl_xor_sync(0xffffffffu, w_circle, o);
      w_hit += __shfl_xor_sync(0xffffffffu, w_hit, o);
      w_bound += __shfl_xor_sync(0xffffffffu, w_bound, o);
    }
    if (lane == 0) {
      atomicAdd(a.stats + 0, w_iou);
      atomicAdd(a.stats + 3, w_circle);
      atomicAdd(a.stats + 18, w_hit);
      atomicAdd(a.stats + 19, w_bound);
    }
    if (tid == 0) {
      atomicAdd(a.stats + 1, static_cast<unsigned long long>(kept));
      atomicAdd(a.stats + 2, static_cast<unsigned long long>(rounds));
      unsigned long long tot = 0;
      for (int k = 0; k < 6; ++k) { atomicAdd(a.stats + 4 + k, static_cast<unsigned long long>(ph[k])); tot += ph[k]; }
      const unsigned long long prev = atomicMax(a.stats + 10, tot);  // slowest segment (cycles)
      if (tot > prev)                                                // (diagnostic, last writer wins) its phases
        for (int k = 0; k < 6; ++k) a.stats[12 + k] = static_cast<unsigned long long>(ph[k]);
      atomicMax(a.stats + 11, static_cast<unsigned long long>(n_seg));   // largest segment (candidates)
      atomicAdd(a.stats + 20, static_cast<unsigned long long>(min(rank_base, n_use)));   // candidates consumed by the scan
      for (int k = 0; k < 3; ++k) atomicAdd(a.stats + 21 + k, static_cast<unsigned long long>(sub[k]));
    }
  }
}

// ------------------------------------------------------------------------------------------
// weighted mode: merge contributions of the candidates behind the scan, grid-wide
// ------------------------------------------------------------------------------------------
// One thread per record of a published tail range (TailInfo).  The keep-set is final, so the candidates are independent:
// each walks its segment's kept-box grid (the same padded-circle test and upper bound as the scan kernel, the same
// bit-exact IoU with the kept box as box 1), finds its FIRST suppressor fs = the lowest kept index with iou > thr, and
// adds its score-weighted row to every kept box k <= fs with iou > merge_thr -- exactly what pull(.., merges_only) does
// for a window, without the window.  Accumulation is the same fp64 atomicAdd into NmsArgs::acc.
// kSm: the grid was copied out of the scan kernel's shared memory in its packed form ((next + 1) << 20 | position in w).
constexpr int kTailThreads = 128;
constexpr int kTailPairs = 768;      // per-warp buffer of (candidate, kept) pairs: circle hits, then the ones the bound lets through

template <bool kSm>
__global__ void __launch_bounds__(kTailThreads)
wnms_tail_kernel(NmsArgs a, int S) {
  // A warp works on 32 candidates at a time in three converged steps, so that the bit-exact IoU routine (a few thousand
  // instructions, called out of line) and the upper bound in front of it always run on dense lanes: (1) every lane walks
  // its candidate's grid window with the padded-circle test only and appends the hits to the warp's buffer; (1b) the
  // buffer is filtered by the upper bound on the IoU, one pair per lane, and compacted in place; (2) the survivors are
  // evaluated one pair per lane: flags + each candidate's first suppressor (shared-memory atomicMin); (3) the buffer is
  // walked again, one pair per lane, for the merges up to and including the first suppressor.  (ncu, first form with the
  // bound inside the walk: walk 45 %, bound 26 %, exact routine 26 % of 430 M warp instructions at ~20 % lane use.)
  __shared__ uint32_t s_pair[kTailThreads / 32][kTailPairs];     // lane << 20 | kept index
  __shared__ uint8_t s_flag[kTailThreads / 32][kTailPairs];
  __shared__ int s_fs[kTailThreads / 32][32];
  __shared__ int s_cnt[kTailThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t *pairs = s_pair[wid];
  uint8_t *flags = s_flag[wid];
  int *fsw = s_fs[wid];
  const int n_total = a.seg_begin[S - 1] + a.seg_count[S - 1];
  const float thr_any = fminf(a.thr, a.mthr);
  const bool prune = a.prune != 0;
  unsigned long long st_iou = 0, st_circle = 0, st_bound = 0;
  const int stride = gridDim.x * blockDim.x;
  for (int g0 = blockIdx.x * blockDim.x + wid * 32; g0 < n_total; g0 += stride) {   // warp-uniform trip count
    const int g = g0 + lane;
    bool live = g < n_total;
    int seg = 0, beg = 0, pos = 0, kbase = 0;
    TailInfo ti{};
    if (live) {
      int lo = 0, hi = S - 1;                     // largest seg with seg_begin[seg] <= g (empty segments share a begin)
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (a.seg_begin[mid] <= g) lo = mid; else hi = mid - 1;
      }
      seg = lo;
      beg = a.seg_begin[seg];
      pos = g - beg;
      ti = a.tail[seg];
      live = pos >= ti.pos_lo && pos < ti.pos_hi;
      kbase = a.kept_stride > 0 ? seg * a.kept_stride : beg;
    }
    if (!__any_sync(0xffffffffu, live)) continue;
    if (lane == 0) s_cnt[wid] = 0;
    fsw[lane] = 0x7fffffff;
    __syncwarp();
    const WRec *recs = static_cast<const WRec *>(a.recs) + beg;
    const float4 *kxyr = a.kxyr_g + kbase;
    const int *knext = kSm ? nullptr : a.knext_g + kbase;
    const int n_buckets = kSm ? kBucketsSmem : a.n_buckets_g;
    const int *heads = a.heads_g + static_cast<size_t>(seg) * n_buckets;
    const uint32_t bmask = static_cast<uint32_t>(n_buckets - 1);
    const int *kos = a.kos + kbase;
    auto position_of = [&](float w) -> int {
      const uint32_t u = __float_as_uint(w);
      return static_cast<int>(kSm ? (u & 0xfffffu) : u);
    };
    auto exact = [&](const WRec &rk, const WRec &rj, bool &above, bool &above_m) {
      const float iou = pair_iou(rk, rj);          // the kept box ranks higher: box 1 of the routine
      ++st_iou;
      above = iou > a.thr;
      above_m = iou > a.mthr;
    };
    auto accumulate = [&](int kb, int k, int row_index) {
      const float *row = a.data + static_cast<size_t>(row_index) * a.D;
      float v[kMaxD];
#pragma unroll
      for (int c = 0; c < kMaxD; ++c) v[c] = c < a.D ? row[c] : 0.f;
      const double sj = row[a.D - 1];
      double *acc = a.acc + static_cast<size_t>(kb + k) * a.D;
#pragma unroll
      for (int c = 0; c < kMaxD; ++c)
        if (c < a.D - 1) atomicAdd(acc + c, sj * static_cast<double>(v[c]));
      atomicAdd(acc + a.D - 1, sj);
      atomicAdd(a.merge_count + kb + k, 1);
    };
    // ---- (1) cheap tests; survivors go to the warp's buffer.  A pair that finds the buffer full is evaluated in place
    // and, because its candidate's first suppressor is not final yet, remembered in `late` for step (3).
    WRec rj{};
    int late_n = 0;
    if (live) {
      rj = recs[pos];
      const float x = rec_cx(rj), y = rec_cy(rj), r = rj.r;
      const float inv_cell = ti.inv_cell, r_cap = ti.r_cap;
      auto cell_of = [&](float v) -> int { return static_cast<int>(fminf(fmaxf(floorf(v * inv_cell), -32768.f), 32767.f)); };
      const int kbx = torus_bits_x(n_buckets);
      auto touches = [&](float qx, float qy, float qr) -> bool {
        const float dx = qx - x, dy = qy - y, rr = qr + r;
        return dx * dx + dy * dy <= rr * rr;
      };
      auto offer = [&](int k, int kpos) {
        const int slot = atomicAdd(&s_cnt[wid], 1);
        if (slot < kTailPairs) { pairs[slot] = (static_cast<uint32_t>(lane) << 20) | static_cast<uint32_t>(k); return; }
        ++st_bound;                                  // buffer full (32 candidates inside a very dense cluster of kept boxes)
        if (prune && !iou_may_exceed(recs[kpos], rj, thr_any)) return;
        bool above, above_m;
        exact(recs[kpos], rj, above, above_m);
        if (above) atomicMin(&fsw[lane], k);
        if (above_m) ++late_n;
      };
      bool scan_all = !prune, skip = false;
      int ix0 = 0, ix1 = 0, iy0 = 0, iy1 = 0;
      if (prune) {
        if (!(x == x) || !(y == y) || !(r == r)) skip = true;            // NaN boxes never interact
        const float reach = r + r_cap;
        if (!(reach <= kPosCap) || !(fabsf(x) <= kPosCap) || !(fabsf(y) <= kPosCap)) scan_all = true;
        else {
          ix0 = cell_of(x - reach); ix1 = cell_of(x + reach); iy0 = cell_of(y - reach); iy1 = cell_of(y + reach);
          scan_all = (ix1 - ix0 + 1) * (iy1 - iy0 + 1) > kMaxCellsPerQuery;
        }
      }
      if (skip) {
      } else if (scan_all) {
        for (int k = 0; k < ti.kept; ++k) {
          const float4 q = kxyr[k];
          ++st_circle;
          if (!prune || touches(q.x, q.y, q.z)) offer(k, position_of(q.w));
        }
      } else {
        for (int iy = iy0; iy <= iy1; ++iy)
          for (int ix = ix0; ix <= ix1; ++ix) {
            int k = heads[torus_bucket(ix, iy, kbx, bmask)];
            while (k >= 0) {
              const float4 q = kxyr[k];
              const int next = kSm ? static_cast<int>(__float_as_uint(q.w) >> 20) - 1 : knext[k];
              ++st_circle;
              if (touches(q.x, q.y, q.z)) offer(k, position_of(q.w));
              k = next;
            }
          }
        for (int o = 0; o < ti.nos; ++o) {
          const int k = kos[o];
          const float4 q = kxyr[k];
          ++st_circle;
          if (touches(q.x, q.y, q.z)) offer(k, position_of(q.w));
        }
      }
    }
    __syncwarp();
    // ---- (1b) the upper bound, one pair per lane; survivors move to the front of the buffer (order is irrelevant)
    int nq = min(s_cnt[wid], kTailPairs);
    if (prune) {
      int kept_pairs = 0;                            // warp-uniform
      for (int q0 = 0; q0 < nq; q0 += 32) {
        const int q = q0 + lane;
        const uint32_t e = q < nq ? pairs[q] : 0u;
        const int l = static_cast<int>(e >> 20), k = static_cast<int>(e & 0xfffffu);
        const int c_beg = __shfl_sync(0xffffffffu, beg, l), c_pos = __shfl_sync(0xffffffffu, pos, l);
        const int c_kbase = __shfl_sync(0xffffffffu, kbase, l);
        bool pass = false;
        if (q < nq) {
          const WRec *rs = static_cast<const WRec *>(a.recs) + c_beg;
          ++st_bound;
          pass = iou_may_exceed(rs[position_of((a.kxyr_g + c_kbase)[k].w)], rs[c_pos], thr_any);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, pass);
        __syncwarp();                                // every lane has read its entry before any slot of this round is reused
        if (pass) pairs[kept_pairs + __popc(m & ((1u << lane) - 1u))] = e;   // kept_pairs + rank <= q: never ahead of the reads
        kept_pairs += __popc(m);
      }
      nq = kept_pairs;
      __syncwarp();
    }
    // ---- (2) exact routine, one pair per lane.  The pair's candidate belongs to lane l: its segment data comes by shuffle.
    for (int q0 = 0; q0 < nq; q0 += 32) {
      const int q = q0 + lane;
      const uint32_t e = q < nq ? pairs[q] : 0u;
      const int l = static_cast<int>(e >> 20), k = static_cast<int>(e & 0xfffffu);
      const int c_beg = __shfl_sync(0xffffffffu, beg, l), c_pos = __shfl_sync(0xffffffffu, pos, l);
      const int c_kbase = __shfl_sync(0xffffffffu, kbase, l);
      if (q < nq) {
        const WRec *rs = static_cast<const WRec *>(a.recs) + c_beg;
        const int kpos = position_of((a.kxyr_g + c_kbase)[k].w);
        bool above, above_m;
        exact(rs[kpos], rs[c_pos], above, above_m);
        flags[q] = static_cast<uint8_t>((above ? 1u : 0u) | (above_m ? 2u : 0u));
        if (above) atomicMin(&fsw[l], k);
      }
    }
    __syncwarp();
    // ---- (3) merges up to and including the first suppressor
    for (int q0 = 0; q0 < nq; q0 += 32) {
      const int q = q0 + lane;
      const uint32_t e = q < nq ? pairs[q] : 0u;
      const int l = static_cast<int>(e >> 20), k = static_cast<int>(e & 0xfffffu);
      const int c_beg = __shfl_sync(0xffffffffu, beg, l), c_pos = __shfl_sync(0xffffffffu, pos, l);
      const int c_kbase = __shfl_sync(0xffffffffu, kbase, l);
      if (q < nq && (flags[q] & 2u) && k <= fsw[l]) accumulate(c_kbase, k, c_beg + c_pos);
    }
    if (late_n > 0) {
      // the pairs this lane evaluated in place: walk again, now that its first suppressor is final, skipping the buffered ones
      const int fs = fsw[lane];
      const float x = rec_cx(rj), y = rec_cy(rj), r = rj.r;
      for (int k = 0; k < ti.kept; ++k) {            // (rare: > 256 bound-passing pairs in one warp) plain scan over the kept boxes
        if (k > fs) break;
        const float4 q = kxyr[k];
        const float dx = q.x - x, dy = q.y - y, rr = q.z + r;
        if (prune && !(dx * dx + dy * dy <= rr * rr)) continue;
        bool buffered = false;
        for (int t = 0; t < nq; ++t) buffered |= pairs[t] == ((static_cast<uint32_t>(lane) << 20) | static_cast<uint32_t>(k));
        if (buffered) continue;
        const int kpos = position_of(q.w);
        if (prune && !iou_may_exceed(recs[kpos], rj, thr_any)) continue;
        bool above, above_m;
        exact(recs[kpos], rj, above, above_m);
        if (above_m) accumulate(kbase, k, beg + pos);
      }
    }
    __syncwarp();
  }
  if (a.stats) {
    for (int o = 16; o; o >>= 1) {
      st_iou += __shfl_xor_sync(0xffffffffu, st_iou, o);
      st_circle += __shfl_xor_sync(0xffffffffu, st_circle, o);
      st_bound += __shfl_xor_sync(0xffffffffu, st_bound, o);
    }
    if (lane == 0) {
      atomicAdd(a.stats + 0, st_iou);
      atomicAdd(a.stats + 3, st_circle);
      atomicAdd(a.stats + 19, st_bound);
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
  const int *seg_begin, *kept_count, *out_off, *kept_pos;
  int kept_stride;
  const uint32_t *src;
  const float *boxes;
  const double *acc;
  int total_classes, out_capacity, weighted, yaw_layout;
  float *out_params, *out_scores, *out_cats, *out_batch;
  const int *out_count;
  int *out_count_w;
  int *host_count;   // optional: mapped pinned host int the total is also stored to (no D2H copy needed to read it)
  int n_segments, fused_scan;
  // fused gather: every rank's buffer is (world, peer_capacity + 1, 16) f32; this rank writes slot `peer_rank` of each
  int n_peers, peer_rank, peer_capacity, sweep_offset;
  float *peer_rows[RV3D_MAX_PEERS];
  // sequence-flag protocol (peer_seq != null): *peer_seq = steps this rank has published.  Step k = *peer_seq + 1 goes
  // to slot k & 1 (peer_rows[q] + slot * peer_slot_stride); after its rows, the LAST block to finish stores k into the
  // header of slot `peer_rank` on every rank behind a system-scope fence, then sets *peer_seq = k.  Everything a
  // replayed CUDA graph needs (slot, sequence number) is read from device memory.
  uint32_t *peer_seq;
  long long peer_slot_stride;
  int *done_ticket;       // zeroed per call
};

__global__ void __launch_bounds__(128) pack_kernel(PackArgs a) {
  const int seg = blockIdx.x;
  const int cnt = a.kept_count[seg];
  const int beg = a.seg_begin[seg];
  const int kb = a.kept_stride > 0 ? seg * a.kept_stride : beg;
  // Output offset of the segment = kept boxes of the segments before it.  With few segments every block sums that
  // prefix itself (<= 4 loads per thread) instead of waiting for a separate one-block scan kernel; block (0, 0) also
  // publishes the total.  Many segments (n_segments > kPackFusedScan): kept_scan_kernel ran before, out_off is ready.
  __shared__ int s_part[8];
  __shared__ int s_last;
  int off, total = 0;
  const bool first = blockIdx.x == 0 && blockIdx.y == 0;
  const bool need_total = first || (a.n_peers > 0 && a.peer_seq);
  uint32_t step = 0;
  size_t slot_off = 0;
  if (a.n_peers > 0 && a.peer_seq) {
    step = *reinterpret_cast<volatile uint32_t *>(a.peer_seq) + 1u;   // only the last block advances it, after every block has read it
    slot_off = static_cast<size_t>(step & 1u) * static_cast<size_t>(a.peer_slot_stride);
  }
  if (a.fused_scan) {
    int part = 0, tot = 0;
    const int upto = need_total ? a.n_segments : seg;
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
  } else {
    off = a.out_off[seg];
    total = *a.out_count;
  }
  if (first && threadIdx.x == 0) {
    if (a.fused_scan) *a.out_count_w = total;
    if (a.host_count) *reinterpret_cast<volatile int *>(a.host_count) = total;
  }
  // grid = (segments, chunks of the kept list): every kept box has its own thread (the kernel is three dependent
  // loads and a sincos per row, i.e. pure latency).  The loop is warp-uniform (a warp owns 32 consecutive rows per
  // trip) so that the fused gather below can hand its rows over inside the warp.
  __shared__ float4 s_rows[4][32 * 4 + 4];                 // per warp: 32 rows x 64 B, staged for the peer stores
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  for (int t0 = blockIdx.y * blockDim.x + wrp * 32; t0 < cnt; t0 += gridDim.y * blockDim.x) {
    const int t = t0 + lane;
    const int row = off + t;
    const bool active = t < cnt && row < a.out_capacity;
    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0, r2 = r0, r3 = r0;
    if (active) {
      const float4 *src = reinterpret_cast<const float4 *>(a.boxes + static_cast<size_t>(a.src[beg + a.kept_pos[kb + t]]) * 8);
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
      double qs = 0.0, qc = 1.0;
      if (!a.yaw_layout || a.n_peers > 0) sincos(static_cast<double>(p[6] * 0.5f), &qs, &qc);  // SO3.py:122-134
      if (a.yaw_layout) {
        float *o = a.out_params + static_cast<size_t>(row) * 7;
#pragma unroll
        for (int c = 0; c < 7; ++c) o[c] = p[c];
      } else {
        float *o = a.out_params + static_cast<size_t>(row) * 10;
        o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; o[3] = p[3]; o[4] = p[4]; o[5] = p[5];
        o[6] = static_cast<float>(qc); o[7] = 0.f; o[8] = 0.f; o[9] = static_cast<float>(qs);
      }
      a.out_scores[row] = b1.w;
      a.out_cats[row] = static_cast<float>(seg % a.total_classes);   // nms.py:51: full_like(scores, j)
      a.out_batch[row] = static_cast<float>(seg / a.total_classes);  // nms.py:242
      r0 = make_float4(static_cast<float>(seg / a.total_classes + a.sweep_offset), static_cast<float>(seg % a.total_classes), b1.w, 0.f);
      r1 = make_float4(p[0], p[1], p[2], p[3]);
      r2 = make_float4(p[4], p[5], static_cast<float>(qc), 0.f);
      r3 = make_float4(0.f, static_cast<float>(qs), 0.f, 0.f);
    }
    if (a.n_peers > 0) {
      // The path's one exchange step, fused: the detections go straight into every rank's gather buffer through the
      // NVLink-mapped peer pointers (no staging copy, no collective call).  A warp's 32 rows are contiguous in the
      // destination (2 KB), so they are transposed through shared memory and stored as four fully coalesced 512-byte
      // stores per peer -- whole 128-byte lines on the wire instead of 16-byte fragments (measured with 16-byte stores
      // per thread: ~100 GB/s per GPU, 0.16 ms of the 8-GPU step).
      const unsigned valid = __ballot_sync(0xffffffffu, active && row < a.peer_capacity);   // a prefix of the warp
      const int nv = __popc(valid);
      float4 *tile = s_rows[wrp];
      tile[lane * 4 + 0] = r0; tile[lane * 4 + 1] = r1; tile[lane * 4 + 2] = r2; tile[lane * 4 + 3] = r3;
      __syncwarp();
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = tile[i * 32 + lane];
      __syncwarp();
      const size_t at = slot_off + (static_cast<size_t>(a.peer_rank) * (a.peer_capacity + 1) + 1 + (off + t0)) * 16;
      // consecutive warps start at different peers, so the stores fan out over all NVLink ports at once
      for (int qq = 0; qq < a.n_peers; ++qq) {
        const int q = (qq + wrp + blockIdx.x + blockIdx.y) % a.n_peers;
        float4 *dst = reinterpret_cast<float4 *>(a.peer_rows[q] + at);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i * 32 + lane < nv * 4) dst[i * 32 + lane] = v[i];
      }
    }
  }
  if (a.n_peers > 0) {
    const size_t hdr = slot_off + static_cast<size_t>(a.peer_rank) * (a.peer_capacity + 1) * 16;
    if (!a.peer_seq) {
      if (first && threadIdx.x == 0) {   // header row: [rows written, rows kept]; the caller orders readers (barrier)
        const float4 h = make_float4(static_cast<float>(total < a.peer_capacity ? total : a.peer_capacity),
                                     static_cast<float>(total), 0.f, 0.f);
        for (int q = 0; q < a.n_peers; ++q) *reinterpret_cast<float4 *>(a.peer_rows[q] + hdr) = h;
      }
    } else {
      // sequence-flag protocol: every block fences its peer stores system-wide, the last block to arrive publishes
      // [rows written, rows kept, seq] to every rank; a reader that sees seq in a header sees the rows behind it
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) s_last = (atomicAdd(a.done_ticket, 1) == static_cast<int>(gridDim.x * gridDim.y) - 1) ? 1 : 0;
      __syncthreads();
      if (s_last) {
        if (threadIdx.x < a.n_peers) {
          __threadfence_system();
          volatile float *h = a.peer_rows[threadIdx.x] + hdr;
          h[0] = static_cast<float>(total < a.peer_capacity ? total : a.peer_capacity);
          h[1] = static_cast<float>(total);
          __threadfence_system();
          reinterpret_cast<volatile uint32_t *>(h)[2] = step;
        }
        __syncthreads();
        if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t *>(a.peer_seq) = step;
      }
    }
  }
}

// consumer side of the sequence-flag protocol: k = *seq steps published by this rank so far; spin until every rank's
// header in slot k & 1 of THIS rank's buffer carries a sequence number >= k (k == 0: nothing to wait for)
__global__ void wait_peer_seq_kernel(const float *__restrict__ rows, long long slot_stride, int world, int peer_capacity,
                                     const uint32_t *__restrict__ seq) {
  const uint32_t k = *reinterpret_cast<const volatile uint32_t *>(seq);
  const int q = threadIdx.x;
  if (q < world && k != 0u) {
    const volatile uint32_t *h = reinterpret_cast<const volatile uint32_t *>(
        rows + static_cast<size_t>(k & 1u) * static_cast<size_t>(slot_stride) + static_cast<size_t>(q) * (peer_capacity + 1) * 16);
    while (static_cast<int32_t>(h[2] - k) < 0) __nanosleep(64);
  }
  __threadfence_system();
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

// geometry of one suppression call, all host-known: capacity (rows of keys / boxes), segments, kept capacity
struct NmsPlan {
  int cap, S, nb, kc, kept_stride, kept_rows, n_buckets_g;
  bool weighted, kept_in_smem;
  bool tail_pass;    // weighted: the candidates behind the scan are merged by the grid-wide tail kernel
  int D;
};

static int pow2_ceil(int64_t v) { int p = 1; while (p < v) p *= 2; return p; }

static NmsPlan make_plan(int cap, int S, int per_segment_max, int num_pre, int num_post, bool weighted, int D) {
  NmsPlan pl{};
  pl.cap = cap > 0 ? cap : 1; pl.S = S; pl.weighted = weighted; pl.D = D;
  int nb = kMaxBins;
  while (nb > 64 && static_cast<int64_t>(S) * nb > (int64_t(1) << 20)) nb >>= 1;
  pl.nb = nb;
  int kc = num_post < num_pre ? num_post : num_pre;
  if (kc > per_segment_max) kc = per_segment_max;
  if (kc > pl.cap) kc = pl.cap;
  if (kc < 1) kc = 1;
  pl.kc = kc;
  // a segment keeps at most as many boxes as it has candidates, so its slice [seg_begin, seg_end) of cap-sized
  // arrays is a private, sufficient region too: used when S * kc would be larger than that
  if (static_cast<int64_t>(S) * kc <= pl.cap) { pl.kept_stride = kc; pl.kept_rows = S * kc; }
  else { pl.kept_stride = 0; pl.kept_rows = pl.cap; }
  pl.kept_in_smem = kc <= kKeptSmem;
  // (a copy of every segment's 16 KB bucket array: not for calls with tens of thousands of segments)
  pl.tail_pass = weighted && (!pl.kept_in_smem || static_cast<int64_t>(S) * kBucketsSmem * 4 <= (int64_t(256) << 20));
  pl.n_buckets_g = 0;
  if (!pl.kept_in_smem) {
    int64_t nbk = pow2_ceil(2 * static_cast<int64_t>(kc));
    if (nbk < 4096) nbk = 4096;
    if (nbk > (1 << 18)) nbk = 1 << 18;
    while (nbk > 4096 && nbk * S > (int64_t(1) << 24)) nbk >>= 1;
    pl.n_buckets_g = static_cast<int>(nbk);
  }
  return pl;
}

struct NmsLayout {
  // zero block (one memset): hist | ticket + pack ticket | kept_count | acc | merge_count
  int *hist, *ticket, *kept_count;
  double *acc;
  int *merge_count;
  TailInfo *tail;
  size_t zero_bytes;
  uint32_t *minmax;      // [2] + BinMap
  BinMap *binmap;
  int *bin_start, *seg_count, *seg_begin, *out_off, *kept_pos, *kos, *n_dev;
  void *recs;
  unsigned long long *skey, *keys_own;
  uint32_t *src, *gorder;
  float *data;
  float4 *kxyr_g;
  int *knext_g, *heads_g;
  size_t heads_bytes, total;
};

static NmsLayout nms_layout(void *scratch, const NmsPlan &pl, bool own_keys, bool own_data) {
  Carver c(scratch);
  NmsLayout L{};
  const size_t cap = static_cast<size_t>(pl.cap), S = static_cast<size_t>(pl.S), kr = static_cast<size_t>(pl.kept_rows);
  L.hist = c.take<int>(S * pl.nb);
  L.ticket = c.take<int>(4);
  L.kept_count = c.take<int>(S);
  L.acc = nullptr; L.merge_count = nullptr;
  L.tail = nullptr;
  if (pl.weighted) {
    L.acc = c.take<double>(kr * pl.D);
    L.merge_count = c.take<int>(kr);
    if (pl.tail_pass) L.tail = c.take<TailInfo>(S);
  }
  L.zero_bytes = align_up(c.used, 256);
  L.minmax = c.take<uint32_t>(2);
  L.binmap = c.take<BinMap>(1);
  L.n_dev = c.take<int>(1);
  L.bin_start = c.take<int>(S * (pl.nb + 1));
  L.seg_count = c.take<int>(S);
  L.seg_begin = c.take<int>(S);
  L.out_off = c.take<int>(S);
  L.kept_pos = c.take<int>(kr);
  L.kos = c.take<int>(kr);
  L.recs = c.take<unsigned char>(cap * (pl.weighted ? sizeof(WRec) : sizeof(HardRec)));
  L.skey = c.take<unsigned long long>(cap);
  L.keys_own = own_keys ? c.take<unsigned long long>(cap) : nullptr;
  L.src = c.take<uint32_t>(cap);
  L.gorder = c.take<uint32_t>(cap);
  L.data = (pl.weighted && own_data) ? c.take<float>(cap * pl.D) : nullptr;
  L.kxyr_g = nullptr; L.knext_g = nullptr; L.heads_g = nullptr; L.heads_bytes = 0;
  if (!pl.kept_in_smem) {
    L.kxyr_g = c.take<float4>(kr);
    L.knext_g = c.take<int>(kr);
    L.heads_bytes = sizeof(int) * S * static_cast<size_t>(pl.n_buckets_g);
    L.heads_g = c.take<int>(S * static_cast<size_t>(pl.n_buckets_g));
  } else if (pl.tail_pass) {
    // the scan kernel's shared-memory grid is copied here (packed form) when it hands its tail to wnms_tail_kernel
    L.kxyr_g = c.take<float4>(kr);
    L.heads_g = c.take<int>(S * static_cast<size_t>(kBucketsSmem));   // written in full by the scan kernel: no memset
  }
  L.total = align_up(c.used, 256);
  return L;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
    else n = kNumSMs;
  }
  return n;
}
// grid of a grid-stride pass over at most `cap` rows (the live count is only known on the device)
static int stride_grid(int cap) {
  const int want = ceil_div(cap > 0 ? cap : 1, 256);
  const int most = sm_count() * 8;
  return want < most ? want : most;
}

template <typename Rec, bool kWeighted, bool kSm>
static int launch_nms_kernel(const NmsArgs &a, const NmsPlan &pl, cudaStream_t s) {
  constexpr int kF = Frontier<kWeighted>::kF;
  const size_t smem = nms_smem_bytes<Rec, kWeighted, kF>(kSm);
  if (smem > 220 * 1024) return RV3D_ERR_ARG;
  RV3D_CHECK_CUDA(cudaFuncSetAttribute(nms_pull_kernel<Rec, kWeighted, kF, kSm>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  nms_pull_kernel<Rec, kWeighted, kF, kSm><<<pl.S, kNmsThreads, smem, s>>>(a);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}
static int sm_count();
template <typename Rec, bool kWeighted>
static int launch_nms_segments(const NmsArgs &a, const NmsPlan &pl, cudaStream_t s) {
  const int rc = pl.kept_in_smem ? launch_nms_kernel<Rec, kWeighted, true>(a, pl, s) : launch_nms_kernel<Rec, kWeighted, false>(a, pl, s);
  if (rc != RV3D_OK || !kWeighted || !a.tail) return rc;
  const int want = ceil_div(pl.cap, kTailThreads), most = sm_count() * 8;
  const int grid = want < most ? want : most;
  if (pl.kept_in_smem) wnms_tail_kernel<true><<<grid, kTailThreads, 0, s>>>(a, pl.S);
  else wnms_tail_kernel<false><<<grid, kTailThreads, 0, s>>>(a, pl.S);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

// hist -> bin_scan -> (caller's scatter) : shared front of every sorted entry point
static int launch_bucketing(const unsigned long long *keys, const int32_t *n_ptr, const NmsPlan &pl, const NmsLayout &L,
                            KeyGeom g, bool range_known, uint32_t desc_lo, uint32_t desc_hi, BinMapArg &bma, cudaStream_t s) {
  bma.dev = nullptr;
  if (range_known) {
    bma.m = make_binmap(desc_lo, desc_hi, g.nb);
  } else {
    RV3D_CHECK_CUDA(cudaMemsetAsync(L.minmax, 0xFF, sizeof(uint32_t), s));
    RV3D_CHECK_CUDA(cudaMemsetAsync(L.minmax + 1, 0, sizeof(uint32_t), s));
    minmax_kernel<<<stride_grid(pl.cap), 256, 0, s>>>(keys, n_ptr, pl.cap, g, L.minmax);
    RV3D_CHECK_LAUNCH();
    binmap_from_minmax_kernel<<<1, 1, 0, s>>>(L.minmax, g.nb, L.binmap);
    RV3D_CHECK_LAUNCH();
    bma.m = BinMap{0, 0};
    bma.dev = L.binmap;
  }
  hist_kernel<<<stride_grid(pl.cap), 256, 0, s>>>(keys, n_ptr, pl.cap, g, bma, pl.S, L.hist);
  RV3D_CHECK_LAUNCH();
  bin_scan_kernel<<<pl.S, 256, 0, s>>>(L.hist, g.nb, pl.S, L.bin_start, L.seg_count, L.seg_begin, L.ticket);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

static uint32_t desc_code(float score, int score_bits) {
  uint32_t bits;
  memcpy(&bits, &score, sizeof(bits));
  const uint32_t d = ~orderable_f32(bits);
  return score_bits >= 32 ? d : (d & ((1u << score_bits) - 1u));
}

static NmsArgs base_args(const NmsPlan &pl, const NmsLayout &L, int num_pre, int num_post, float thr, float mthr,
                         bool weighted, int flags, int64_t *stats) {
  NmsArgs a{};
  a.recs = L.recs; a.skey = L.skey; a.gorder = L.gorder; a.seg_begin = L.seg_begin; a.seg_count = L.seg_count;
  a.bin_start = L.bin_start; a.nb = pl.nb; a.presorted = 0;
  a.num_pre = num_pre; a.num_post = num_post; a.thr = thr; a.mthr = mthr;
  a.prune = (thr >= 0.f && (!weighted || mthr >= 0.f)) ? 1 : 0;
  (void)flags;
  a.kept_stride = pl.kept_stride; a.kept_pos = L.kept_pos; a.kept_count = L.kept_count;
  a.kxyr_g = L.kxyr_g; a.knext_g = L.knext_g; a.heads_g = L.heads_g; a.n_buckets_g = pl.n_buckets_g; a.kos = L.kos;
  a.tail = L.tail;
  if (const char *e = getenv("RV3D_NMS_NO_TAIL")) { if (atoi(e)) a.tail = nullptr; }   // experiments / A-B tests only
  a.data = L.data; a.D = pl.D; a.acc = L.acc; a.merge_count = L.merge_count;
  a.stats = reinterpret_cast<unsigned long long *>(stats);
  a.rcap_mult = 2.0f; a.cell_mult = 3.0f;   // measured sweep (profiles/r02_nms_geometry.md): r_cap 2, cell 3 mean radii
  if (const char *e = getenv("RV3D_NMS_RCAP")) { const float v = static_cast<float>(atof(e)); if (v > 0.f) a.rcap_mult = v; }   // experiments only
  if (const char *e = getenv("RV3D_NMS_CELL")) { const float v = static_cast<float>(atof(e)); if (v > 0.f) a.cell_mult = v; }
  return a;
}

}  // namespace rv3d

using namespace rv3d;

static bool nms_params_ok(const rv3d_nms_params *p) {
  return p && p->batch > 0 && p->total_classes > 0 && p->total_candidates > 0 && p->capacity > 0 && p->num_pre_nms > 0 &&
         p->num_post_nms > 0 && p->out_capacity >= 0 && (p->mode == RV3D_NMS_HARD || p->mode == RV3D_NMS_WEIGHTED) &&
         (p->out_layout == RV3D_OUT_QUAT || p->out_layout == RV3D_OUT_YAW) &&
         static_cast<int64_t>(p->batch) * p->total_classes < (int64_t(1) << 24);
}

static NmsPlan plan_of(const rv3d_nms_params *p) {
  return make_plan(p->capacity, p->batch * p->total_classes, p->total_candidates, p->num_pre_nms, p->num_post_nms,
                   p->mode == RV3D_NMS_WEIGHTED, 9);
}

extern "C" size_t rv3d_nms_scratch_bytes(const rv3d_nms_params *p) {
  if (!nms_params_ok(p)) return 0;
  return nms_layout(nullptr, plan_of(p), false, true).total;
}

extern "C" int rv3d_nms(const rv3d_nms_params *p, const uint64_t *keys_in, const float *boxes,
                        const int32_t *n_candidates, float *out_params, float *out_scores, float *out_categories,
                        float *out_batch, int32_t *out_count, int64_t *stats, void *scratch, size_t scratch_bytes,
                        rv3d_stream_t stream) {
  RV3D_CHECK_ARG(nms_params_ok(p) && out_count && scratch && n_candidates);
  RV3D_CHECK_ARG(p->peer_world >= 0 && p->peer_world <= RV3D_MAX_PEERS);
  if (p->peer_world > 0) {
    RV3D_CHECK_ARG(p->out_layout == RV3D_OUT_QUAT && p->peer_rank >= 0 && p->peer_rank < p->peer_world && p->peer_capacity > 0);
    for (int q = 0; q < p->peer_world; ++q) RV3D_CHECK_ARG(p->peer_rows[q] && aligned(p->peer_rows[q], 16));
  }
  RV3D_CHECK_ARG(keys_in && boxes && out_params && out_scores && out_categories && out_batch);
  if (!aligned(boxes, 16) || !aligned(keys_in, 8) || !aligned(scratch, 256)) return RV3D_ERR_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int S = p->batch * p->total_classes;
  const int idx_bits = bits_for(p->total_candidates);
  const int score_bits = p->score_bits ? p->score_bits : 32;
  RV3D_CHECK_ARG(score_bits == 31 || score_bits == 32);
  if (bits_for(S) + score_bits + idx_bits > 64) return RV3D_ERR_KEYBITS;
  const bool weighted = p->mode == RV3D_NMS_WEIGHTED;
  const NmsPlan pl = plan_of(p);
  RV3D_CHECK_ARG(pl.kc < (1 << 20));   // kept indices travel in 20 bits of a queue entry
  const NmsLayout L = nms_layout(scratch, pl, false, true);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  const auto *keys = reinterpret_cast<const unsigned long long *>(keys_in);

  RV3D_CHECK_CUDA(cudaMemsetAsync(L.hist, 0, L.zero_bytes, s));
  if (L.heads_g) RV3D_CHECK_CUDA(cudaMemsetAsync(L.heads_g, 0xFF, L.heads_bytes, s));

  // K3: counting sort by (segment, score bin)
  const KeyGeom g{idx_bits, score_bits, pl.nb};
  const bool range_known = p->score_hi > p->score_lo;
  BinMapArg bma{};
  int rc = launch_bucketing(keys, n_candidates, pl, L, g, range_known, desc_code(p->score_hi, score_bits),
                            desc_code(p->score_lo, score_bits), bma, s);
  if (rc != RV3D_OK) return rc;
  const RecFromBoxes8 build{boxes};
  if (weighted)
    scatter_records_kernel<true><<<stride_grid(pl.cap), 256, 0, s>>>(keys, build, n_candidates, pl.cap, g, bma, S, L.hist, L.bin_start,
                                                                    L.seg_begin, L.recs, L.skey, L.src, L.data);
  else
    scatter_records_kernel<false><<<stride_grid(pl.cap), 256, 0, s>>>(keys, build, n_candidates, pl.cap, g, bma, S, L.hist, L.bin_start,
                                                                     L.seg_begin, L.recs, L.skey, L.src, nullptr, 1);
  RV3D_CHECK_LAUNCH();

  // K4 / K5
  NmsArgs a = base_args(pl, L, p->num_pre_nms, p->num_post_nms, p->iou_threshold, p->merge_threshold, weighted, p->flags, stats);
  if (!weighted) { a.lazy_boxes = boxes; a.src = L.src; }
  rc = weighted ? launch_nms_segments<WRec, true>(a, pl, s) : launch_nms_segments<HardRec, false>(a, pl, s);
  if (rc != RV3D_OK) return rc;

  constexpr int kPackFusedScan = 512;   // up to this many segments the pack blocks sum their own output offset
  const bool fused_scan = S <= kPackFusedScan;
  if (!fused_scan) {
    kept_scan_kernel<<<1, 1024, 0, s>>>(L.kept_count, S, L.out_off, out_count, p->out_capacity);
    RV3D_CHECK_LAUNCH();
  }
  PackArgs pa{};
  pa.out_count_w = out_count; pa.n_segments = S; pa.fused_scan = fused_scan ? 1 : 0;
  pa.host_count = p->host_count;
  pa.seg_begin = L.seg_begin; pa.kept_stride = pl.kept_stride; pa.kept_count = L.kept_count; pa.out_off = L.out_off;
  pa.kept_pos = L.kept_pos; pa.src = L.src; pa.boxes = boxes; pa.acc = L.acc;
  pa.total_classes = p->total_classes; pa.out_capacity = p->out_capacity; pa.weighted = weighted ? 1 : 0; pa.yaw_layout = p->out_layout == RV3D_OUT_YAW;
  pa.out_params = out_params; pa.out_scores = out_scores; pa.out_cats = out_categories; pa.out_batch = out_batch;
  pa.out_count = out_count;
  pa.n_peers = p->peer_world; pa.peer_rank = p->peer_rank; pa.peer_capacity = p->peer_capacity; pa.sweep_offset = p->sweep_offset;
  pa.peer_seq = p->peer_seq; pa.peer_slot_stride = p->peer_slot_stride; pa.done_ticket = L.ticket + 1;
  for (int q = 0; q < p->peer_world; ++q) pa.peer_rows[q] = p->peer_rows[q];
  {
    const int per_seg = p->num_post_nms < pl.kc ? p->num_post_nms : pl.kc;
    const int chunks = ceil_div(per_seg < (1 << 20) ? per_seg : (1 << 20), 128);
    pack_kernel<<<dim3(S, chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks)), 128, 0, s>>>(pa);
  }
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_peer_wait(const float *rows, int64_t slot_stride, int32_t world, int32_t peer_capacity,
                              const uint32_t *seq, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(rows && world > 0 && world <= RV3D_MAX_PEERS && peer_capacity > 0 && seq && slot_stride >= 0);
  wait_peer_seq_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(rows, slot_stride, world, peer_capacity, seq);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

// ==========================================================================================
// free-standing operator forms
// ==========================================================================================
namespace rv3d {

__global__ void score_keys_kernel(const float *__restrict__ scores, int n, int idx_bits,
                                  unsigned long long *__restrict__ keys, int *__restrict__ n_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_dev = n;
  if (i >= n) return;
  const uint32_t desc = ~orderable_f32(__float_as_uint(scores[i]));
  keys[i] = (static_cast<unsigned long long>(desc) << idx_bits) | static_cast<uint32_t>(i);
}

__global__ void single_segment_kernel(int n, int *seg_begin, int *seg_count) {
  seg_begin[0] = 0; seg_count[0] = n;
}

__global__ void w_recs_from_boxes5_kernel(const float *__restrict__ boxes5, int n, WRec *__restrict__ recs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *b = boxes5 + static_cast<size_t>(i) * 5;
  recs[i] = make_w_rec(b[0], b[1], b[2], b[3], b[4]);
}

__global__ void keep_indices_kernel(const int *__restrict__ kept_pos, const int *__restrict__ kept_count,
                                    const uint32_t *__restrict__ src, int64_t *__restrict__ keep,
                                    int32_t *__restrict__ n_keep) {
  const int cnt = kept_count[0];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_keep = cnt;
  if (i < cnt) keep[i] = src ? src[kept_pos[i]] : kept_pos[i];
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
  const HardRec ra = make_hard_rec(a[0], a[1], a[3], a[4], a[6], 1.0, true);  // mmcv: angle in radians, its rotation direction
  const HardRec rb = make_hard_rec(b[0], b[1], b[3], b[4], b[6], 1.0, true);
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
  out[t] = rot_iou(make_hard_rec(a[0], a[1], a[2], a[3], a[4], 1.0, true), make_hard_rec(b[0], b[1], b[2], b[3], b[4], 1.0, true));
}

// The pruning decision of the NMS kernels, exposed for tests/test_gpu_iou_decisions.py: aligned pairs.
// decision[i]: 2 = stopped by the upper bound (the kernels treat the pair as "not above the threshold" without running
// the exact routine), 0 = sent to the exact routine;  bound[i]: the upper bound on the IoU the kernels compare with the
// threshold (NaN when it is not trusted);  exact[i]: the bit-exact routine's value.
template <typename Rec>
__device__ __forceinline__ void decide_pair(const Rec &ra, const Rec &rb, float thr, int8_t &decision, float &bound, float &ex) {
  const Obb oa = obb_of(ra), ob = obb_of(rb);
  const float inter = fminf(overlap_in_frame(oa, ob), overlap_in_frame(ob, oa));
  const float uni = fabsf(oa.w * oa.h) + fabsf(ob.w * ob.h) - inter;
  bound = (!(uni > 0.f) || !(inter == inter)) ? CUDART_NAN_F : (inter / uni) * 1.02f + 1e-5f;
  decision = static_cast<int8_t>(iou_may_exceed(ra, rb, thr) ? 0 : 2);
  ex = pair_iou(ra, rb);
}

__global__ void __launch_bounds__(128)
pair_decisions_kernel(const float *__restrict__ A, const float *__restrict__ B, int64_t n, float thr, int routine,
                      int8_t *__restrict__ decision, float *__restrict__ approx, float *__restrict__ exact) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *a = A + i * 5, *b = B + i * 5;
  if (routine == 0)
    decide_pair(make_hard_rec(a[0], a[1], a[2], a[3], a[4], 0.01745329251), make_hard_rec(b[0], b[1], b[2], b[3], b[4], 0.01745329251),
                thr, decision[i], approx[i], exact[i]);
  else
    decide_pair(make_w_rec(a[0], a[1], a[2], a[3], a[4]), make_w_rec(b[0], b[1], b[2], b[3], b[4]), thr, decision[i], approx[i], exact[i]);
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

struct SortLayout {
  unsigned long long *keys, *keys_alt;
  uint32_t *order, *order_alt, *segs;
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

static SortLayout sort_layout(void *scratch, int n) {
  Carver c(scratch);
  SortLayout L{};
  const int nn = n > 0 ? n : 1;
  L.keys = c.take<unsigned long long>(nn);
  L.keys_alt = c.take<unsigned long long>(nn);
  L.order = c.take<uint32_t>(nn);
  L.order_alt = c.take<uint32_t>(nn);
  L.segs = c.take<uint32_t>(nn);
  L.cub_bytes = cub_sort_bytes(nn, 64);
  L.cub_tmp = c.take<unsigned char>(L.cub_bytes ? L.cub_bytes : 1);
  L.total = align_up(c.used, 256);
  return L;
}

}  // namespace rv3d

extern "C" size_t rv3d_nms_rotated_scratch_bytes(int32_t n) {
  const int nn = n > 0 ? n : 1;
  return nms_layout(nullptr, make_plan(nn, 1, nn, nn, nn, false, 0), true, false).total;
}

extern "C" int rv3d_nms_rotated(const float *boxes, const float *scores, int32_t n, float iou_threshold,
                                int64_t *keep, int32_t *n_keep, void *scratch, size_t scratch_bytes,
                                rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && n < (1 << 20) && n_keep && scratch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    RV3D_CHECK_CUDA(cudaMemsetAsync(n_keep, 0, sizeof(int32_t), s));
    return RV3D_OK;
  }
  RV3D_CHECK_ARG(boxes && scores && keep);
  if (!aligned(scratch, 256)) return RV3D_ERR_ALIGN;
  const NmsPlan pl = make_plan(n, 1, n, n, n, false, 0);
  const NmsLayout L = nms_layout(scratch, pl, true, false);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  RV3D_CHECK_CUDA(cudaMemsetAsync(L.hist, 0, L.zero_bytes, s));
  if (L.heads_g) RV3D_CHECK_CUDA(cudaMemsetAsync(L.heads_g, 0xFF, L.heads_bytes, s));
  const int idx_bits = bits_for(n);
  score_keys_kernel<<<ceil_div(n, 256), 256, 0, s>>>(scores, n, idx_bits, L.keys_own, L.n_dev);
  RV3D_CHECK_LAUNCH();
  const KeyGeom g{idx_bits, 32, pl.nb};
  BinMapArg bma{};
  int rc = launch_bucketing(L.keys_own, L.n_dev, pl, L, g, false, 0, 0, bma, s);
  if (rc != RV3D_OK) return rc;
  scatter_records_kernel<false><<<stride_grid(n), 256, 0, s>>>(L.keys_own, RecFromBoxes5{boxes}, L.n_dev, n, g, bma, 1, L.hist,
                                                               L.bin_start, L.seg_begin, L.recs, L.skey, L.src, nullptr);
  RV3D_CHECK_LAUNCH();
  const NmsArgs a = base_args(pl, L, n, n, iou_threshold, 0.f, false, 0, nullptr);
  rc = launch_nms_segments<HardRec, false>(a, pl, s);
  if (rc != RV3D_OK) return rc;
  keep_indices_kernel<<<ceil_div(n, 256), 256, 0, s>>>(L.kept_pos, L.kept_count, L.src, keep, n_keep);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" size_t rv3d_wnms_scratch_bytes(int32_t n, int32_t d) {
  const int nn = n > 0 ? n : 1;
  return nms_layout(nullptr, make_plan(nn, 1, nn, nn, nn, true, d > 0 ? d : 1), false, false).total;
}

extern "C" int rv3d_wnms(const float *boxes, const float *data, int32_t n, int32_t d, float nms_threshold,
                         float merge_threshold, float *output, int64_t *keep, int64_t *count, int32_t *n_out,
                         void *scratch, size_t scratch_bytes, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && n < (1 << 20) && d >= 1 && d <= kMaxD && n_out && scratch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    RV3D_CHECK_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), s));
    return RV3D_OK;
  }
  RV3D_CHECK_ARG(boxes && data && output && keep && count);
  if (!aligned(scratch, 256)) return RV3D_ERR_ALIGN;
  const NmsPlan pl = make_plan(n, 1, n, n, n, true, d);
  const NmsLayout L = nms_layout(scratch, pl, false, false);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  RV3D_CHECK_CUDA(cudaMemsetAsync(L.hist, 0, L.zero_bytes, s));
  if (L.heads_g) RV3D_CHECK_CUDA(cudaMemsetAsync(L.heads_g, 0xFF, L.heads_bytes, s));
  w_recs_from_boxes5_kernel<<<ceil_div(n, 256), 256, 0, s>>>(boxes, n, static_cast<WRec *>(L.recs));
  RV3D_CHECK_LAUNCH();
  single_segment_kernel<<<1, 1, 0, s>>>(n, L.seg_begin, L.seg_count);
  RV3D_CHECK_LAUNCH();
  RV3D_CHECK_CUDA(cudaMemsetAsync(output, 0, sizeof(float) * static_cast<size_t>(n) * d, s));  // nms.py:155,173
  RV3D_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int64_t) * static_cast<size_t>(n), s));
  NmsArgs a = base_args(pl, L, n, n, nms_threshold, merge_threshold, true, 0, nullptr);
  a.presorted = 1; a.bin_start = nullptr; a.skey = nullptr; a.gorder = nullptr;   // the caller sorted (nms.py:148-154)
  a.data = data;
  const int rc = launch_nms_segments<WRec, true>(a, pl, s);
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

extern "C" int rv3d_pair_decisions(const float *boxes_a, const float *boxes_b, int64_t n, float iou_threshold,
                                   int32_t routine, int8_t *decision, float *approx, float *exact, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && (routine == 0 || routine == 1));
  if (n == 0) return RV3D_OK;
  RV3D_CHECK_ARG(boxes_a && boxes_b && decision && approx && exact && n < (int64_t(1) << 37));
  pair_decisions_kernel<<<ceil_div(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(boxes_a, boxes_b, n, iou_threshold,
                                                                                       routine, decision, approx, exact);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" size_t rv3d_pack_candidates_scratch_bytes(int32_t n) { return sort_layout(nullptr, n).total; }

extern "C" int rv3d_pack_candidates(const uint64_t *keys, const float *boxes, int32_t n, int32_t batch,
                                    int32_t total_classes, int32_t total_candidates, int32_t score_bits, float *out_params,
                                    float *out_scores, int64_t *out_categories, int64_t *out_batch, void *scratch,
                                    size_t scratch_bytes, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && batch > 0 && total_classes > 0 && total_candidates > 0 && scratch);
  RV3D_CHECK_ARG(score_bits == 31 || score_bits == 32);
  if (n == 0) return RV3D_OK;
  RV3D_CHECK_ARG(keys && boxes && out_params && out_scores && out_categories && out_batch);
  if (!aligned(scratch, 256) || !aligned(boxes, 16)) return RV3D_ERR_ALIGN;
  const SortLayout L = sort_layout(scratch, n);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int idx_bits = bits_for(total_candidates);
  remap_keys_kernel<<<ceil_div(n, 256), 256, 0, s>>>(reinterpret_cast<const unsigned long long *>(keys), n, idx_bits, score_bits,
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
