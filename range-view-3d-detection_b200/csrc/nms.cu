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
#include <cub/device/device_radix_sort.cuh>
#include <math_constants.h>

#include "common.cuh"
#include "iou.cuh"

namespace rv3d {

constexpr int kNmsThreads = 512;
constexpr int kF = 256;                 // frontier size
constexpr int kFW = kF / 32;            // words per bit-matrix row
constexpr int kPairCap = 12288;         // frontier pair queue entries (u16 each)
constexpr int kTile = 4096;             // candidates per kill-phase tile
constexpr int kKillCap = 8192;          // kill-phase queue entries
constexpr int kMaxD = 16;               // max data columns of the weighted merge

// ------------------------------------------------------------------------------------------
// small kernels around the sort
// ------------------------------------------------------------------------------------------
__global__ void iota_kernel(uint32_t *v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

// seg_begin / seg_end from sorted keys (both zero-initialised: empty segments stay [0,0))
__global__ void segment_bounds_kernel(const unsigned long long *__restrict__ keys, int n, int shift,
                                      int *__restrict__ seg_begin, int *__restrict__ seg_end) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = static_cast<uint32_t>(keys[i] >> shift);
  if (i == 0 || static_cast<uint32_t>(keys[i - 1] >> shift) != s) seg_begin[s] = i;
  if (i == n - 1 || static_cast<uint32_t>(keys[i + 1] >> shift) != s) seg_end[s] = i + 1;
}

// exclusive scan of min(seg size, cap) over S segments (S is small: one block, serial chunks)
__global__ void segment_capacity_scan_kernel(const int *__restrict__ seg_begin, const int *__restrict__ seg_end,
                                             int S, int cap, int *__restrict__ base) {
  __shared__ int s_part[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (S + nt - 1) / nt;
  const int lo = min(S, tid * per), hi = min(S, lo + per);
  int sum = 0;
  for (int s = lo; s < hi; ++s) sum += min(seg_end[s] - seg_begin[s], cap);
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < nt; ++t) { const int v = s_part[t]; s_part[t] = run; run += v; }
  }
  __syncthreads();
  int run = s_part[tid];
  for (int s = lo; s < hi; ++s) { base[s] = run; run += min(seg_end[s] - seg_begin[s], cap); }
}

// sorted position -> suppression record (+ merge row for the weighted mode)
template <bool kWeighted>
__global__ void __launch_bounds__(256)
prepare_records_kernel(const uint32_t *__restrict__ order, const float *__restrict__ boxes, int n,
                       void *__restrict__ recs, float *__restrict__ data) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
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
  int *kept_pos;           // [kept_base[s] + i] = position inside the segment
  int *kept_count;         // [s]
  // weighted merge
  const float *data;       // (n, D) rows in sorted order, score last
  int D;
  double *acc;             // [(kept_base[s] + i) * D + c]: c < D-1 weighted sums, c = D-1 weight sum
  int *merge_count;        // [kept_base[s] + i]
  unsigned long long *stats;
};

template <typename Rec>
struct Smem {
  uint32_t *alive;                 // nwords
  int *front_pos;                  // kF
  Rec *frec;                       // kF
  uint32_t *sup;                   // kF * kFW
  uint32_t *mrg;                   // kF * kFW (weighted)
  uint16_t *keptf;                 // kF
  float *kx, *ky, *kr;             // kF each
  uint32_t *queue;                 // max(kPairCap/2, kKillCap) words
  float *qiou;                     // kKillCap (weighted)
  uint8_t *firstsup;               // kTile (weighted)
};

template <typename Rec, bool kWeighted>
__host__ __device__ inline size_t nms_smem_bytes(int nwords) {
  size_t b = 0;
  b += align_up_c(sizeof(uint32_t) * nwords, 16);
  b += sizeof(int) * kF;
  b += sizeof(Rec) * kF;
  b += sizeof(uint32_t) * kF * kFW * (kWeighted ? 2 : 1);
  b += align_up_c(sizeof(uint16_t) * kF, 16);
  b += sizeof(float) * kF * 3;
  b += sizeof(uint32_t) * (kKillCap > kPairCap / 2 ? kKillCap : kPairCap / 2);
  if (kWeighted) b += sizeof(float) * kKillCap + kTile;
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

template <typename Rec, bool kWeighted>
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_segment_kernel(NmsArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_warp[kNmsThreads / 32 + 1];
  __shared__ int s_qn, s_nk;

  const int seg = blockIdx.x;
  const int beg = a.seg_begin[seg];
  const int n = min(a.seg_end[seg] - beg, a.num_pre);  // top num_pre_nms by score (nms.py:29-32)
  const int tid = threadIdx.x;
  if (n <= 0) {
    if (tid == 0) a.kept_count[seg] = 0;
    return;
  }
  const int nwords = (n + 31) >> 5;
  const Rec *recs = static_cast<const Rec *>(a.recs) + beg;
  const int kbase = a.kept_base[seg];

  // ---- carve shared memory
  Smem<Rec> S;
  {
    unsigned char *p = smem_raw;
    S.alive = reinterpret_cast<uint32_t *>(p); p += align_up_c(sizeof(uint32_t) * nwords, 16);
    S.frec = reinterpret_cast<Rec *>(p); p += sizeof(Rec) * kF;
    S.front_pos = reinterpret_cast<int *>(p); p += sizeof(int) * kF;
    S.sup = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * kF * kFW;
    S.mrg = S.sup;
    if (kWeighted) { S.mrg = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * kF * kFW; }
    S.kx = reinterpret_cast<float *>(p); p += sizeof(float) * kF;
    S.ky = reinterpret_cast<float *>(p); p += sizeof(float) * kF;
    S.kr = reinterpret_cast<float *>(p); p += sizeof(float) * kF;
    S.queue = reinterpret_cast<uint32_t *>(p); p += sizeof(uint32_t) * (kKillCap > kPairCap / 2 ? kKillCap : kPairCap / 2);
    S.keptf = reinterpret_cast<uint16_t *>(p); p += align_up_c(sizeof(uint16_t) * kF, 16);
    S.qiou = nullptr; S.firstsup = nullptr;
    if (kWeighted) {
      S.qiou = reinterpret_cast<float *>(p); p += sizeof(float) * kKillCap;
      S.firstsup = p;
    }
  }
  for (int w = tid; w < nwords; w += kNmsThreads)
    S.alive[w] = (w == nwords - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;

  unsigned long long st_iou = 0, st_circle = 0;
  int kept_total = 0;
  int cursor_word = 0;
  int rounds = 0;
  __syncthreads();

  while (true) {
    // ================= 1. frontier: first kF alive candidates =================
    int nf = 0;
    for (int base = cursor_word; base < nwords && nf < kF; base += kNmsThreads) {
      const int wi = base + tid;
      const uint32_t word = wi < nwords ? S.alive[wi] : 0u;
      int total;
      const int off = block_exclusive_scan(__popc(word), s_warp, total);
      if (word && nf + off < kF) {
        uint32_t m = word;
        int r = nf + off;
        while (m && r < kF) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          S.front_pos[r++] = (wi << 5) + bit;
        }
      }
      nf = min(kF, nf + total);
    }
    __syncthreads();
    if (nf == 0) break;
    ++rounds;
    // the frontier is decided this round: clear its bits, load its records, zero the bit-matrices
    if (tid < nf) {
      const int pos = S.front_pos[tid];
      atomicAnd(&S.alive[pos >> 5], ~(1u << (pos & 31)));
      S.frec[tid] = recs[pos];
    }
    for (int i = tid; i < kF * kFW; i += kNmsThreads) {
      S.sup[i] = 0u;
      if (kWeighted) S.mrg[i] = 0u;
    }
    if (tid == 0) s_qn = 0;
    __syncthreads();
    cursor_word = S.front_pos[0] >> 5;

    // ================= 2. pairs inside the frontier =================
    {
      uint16_t *pq = reinterpret_cast<uint16_t *>(S.queue);
      for (int idx = tid; idx < nf * nf; idx += kNmsThreads) {
        const int i = idx / nf, j = idx - i * nf;
        if (i >= j) continue;
        const Rec &ri = S.frec[i];
        const Rec &rj = S.frec[j];
        const float dx = rec_cx(ri) - rec_cx(rj), dy = rec_cy(ri) - rec_cy(rj), rr = ri.r + rj.r;
        ++st_circle;
        if (!(dx * dx + dy * dy <= rr * rr)) continue;
        const int slot = atomicAdd(&s_qn, 1);
        if (slot < kPairCap) {
          pq[slot] = static_cast<uint16_t>((i << 8) | j);
        } else {  // queue full: evaluate in place (rare)
          const float iou = pair_iou(ri, rj);
          ++st_iou;
          if (iou > a.thr) atomicOr(&S.sup[i * kFW + (j >> 5)], 1u << (j & 31));
          if (kWeighted && iou > a.mthr) atomicOr(&S.mrg[i * kFW + (j >> 5)], 1u << (j & 31));
        }
      }
      __syncthreads();
      const int qn = min(s_qn, kPairCap);
      for (int q = tid; q < qn; q += kNmsThreads) {
        const int i = pq[q] >> 8, j = pq[q] & 255;
        const float iou = pair_iou(S.frec[i], S.frec[j]);
        ++st_iou;
        if (iou > a.thr) atomicOr(&S.sup[i * kFW + (j >> 5)], 1u << (j & 31));
        if (kWeighted && iou > a.mthr) atomicOr(&S.mrg[i * kFW + (j >> 5)], 1u << (j & 31));
      }
      __syncthreads();
    }

    // ================= 3. greedy resolution of the frontier (one warp) =================
    if (tid < 32) {
      const int lane = tid;
      // lane w < kFW owns word w of the `removed` set; bits >= nf are pre-removed
      uint32_t removed = 0u;
      if (lane < kFW) {
        const int lo = lane << 5;
        removed = (nf >= lo + 32) ? 0u : (nf <= lo ? 0xffffffffu : ~((1u << (nf - lo)) - 1u));
      }
      int nk = 0;
      const int room = a.num_post - kept_total;
      int next = 0;
      while (nk < room) {
        uint32_t avail = 0u;
        if (lane < kFW) {
          avail = ~removed;
          const int lo = lane << 5;
          if (next >= lo + 32) avail = 0u;
          else if (next > lo) avail &= ~((1u << (next - lo)) - 1u);
        }
        const uint32_t have = __ballot_sync(0xffffffffu, avail != 0u);
        if (!have) break;
        const int wsel = __ffs(have) - 1;
        const uint32_t aw = __shfl_sync(0xffffffffu, avail, wsel);
        const int i = (wsel << 5) + __ffs(aw) - 1;
        if (lane == 0) S.keptf[nk] = static_cast<uint16_t>(i);
        ++nk;
        if (lane < kFW) removed |= S.sup[i * kFW + lane];
        next = i + 1;
      }
      if (lane == 0) s_nk = nk;
    }
    __syncthreads();
    const int nk = s_nk;

    // ================= 4. publish the newly kept boxes =================
    if (tid < nk) {
      const int fi = S.keptf[tid];
      a.kept_pos[kbase + kept_total + tid] = S.front_pos[fi];
      S.kx[tid] = rec_cx(S.frec[fi]);
      S.ky[tid] = rec_cy(S.frec[fi]);
      S.kr[tid] = S.frec[fi].r;
    }
    if (kWeighted) {
      // merge sets inside the frontier: candidate j joins every kept i < j with iou > merge_thr
      // that comes no later than its first suppressor; kept boxes join themselves.
      if (tid < nf) {
        const int j = tid;
        const size_t rowj = static_cast<size_t>(beg + S.front_pos[j]) * a.D;
        const double sj = a.data[rowj + a.D - 1];
        for (int t = 0; t < nk; ++t) {
          const int i = S.keptf[t];
          if (i > j) break;
          const bool self = (i == j);
          const bool m = self || (S.mrg[i * kFW + (j >> 5)] >> (j & 31)) & 1u;
          if (m) {
            double *acc = a.acc + static_cast<size_t>(kbase + kept_total + t) * a.D;
            for (int c = 0; c < a.D - 1; ++c) atomicAdd(acc + c, sj * static_cast<double>(a.data[rowj + c]));
            atomicAdd(acc + a.D - 1, sj);
            atomicAdd(a.merge_count + kbase + kept_total + t, 1);
          }
          if (self || ((S.sup[i * kFW + (j >> 5)] >> (j & 31)) & 1u)) break;
        }
      }
    }
    kept_total += nk;
    __syncthreads();
    // nms.py:53-56: only the first num_post_nms kept survive, so the scan can stop there.  The
    // weighted mode still owes the last kept boxes their merge sets from the candidates behind
    // the frontier, so it runs one more kill phase before leaving.
    const bool last_round = kept_total >= a.num_post;
    if (last_round && !kWeighted) break;

    // ================= 5. kill phase: alive candidates vs the newly kept boxes =================
    for (int tile0 = cursor_word << 5; tile0 < n; tile0 += kTile) {
      if (tid == 0) s_qn = 0;
      if (kWeighted)
        for (int i = tid; i < kTile; i += kNmsThreads) S.firstsup[i] = 255;
      __syncthreads();
      // 4 threads per bitmap word, 8 candidates each
      for (int sub = tid; sub < (kTile >> 3); sub += kNmsThreads) {
        const int wi = (tile0 >> 5) + (sub >> 2);
        if (wi >= nwords) continue;
        uint32_t bits = (S.alive[wi] >> ((sub & 3) << 3)) & 0xffu;
        while (bits) {
          const int bit = __ffs(bits) - 1;
          bits &= bits - 1;
          const int j = (wi << 5) + ((sub & 3) << 3) + bit;
          const Rec &rj = recs[j];
          const float jx = rec_cx(rj), jy = rec_cy(rj), jr = rj.r;
          for (int t = 0; t < nk; ++t) {
            const float dx = S.kx[t] - jx, dy = S.ky[t] - jy, rr = S.kr[t] + jr;
            if (!(dx * dx + dy * dy <= rr * rr)) continue;
            const int slot = atomicAdd(&s_qn, 1);
            if (slot < kKillCap) {
              S.queue[slot] = (static_cast<uint32_t>(j - tile0) << 8) | t;
            } else if (!kWeighted) {  // queue full: evaluate in place (hard mode only needs ANY suppressor)
              ++st_iou;
              if (pair_iou(S.frec[S.keptf[t]], rj) > a.thr) {
                atomicAnd(&S.alive[j >> 5], ~(1u << (j & 31)));
                break;
              }
            }
          }
          st_circle += nk;
        }
      }
      __syncthreads();
      int qn = s_qn;
      if (kWeighted && qn > kKillCap) {
        // (weighted) overflow: the dropped pairs are redone serially per candidate by thread 0..;
        // keep it simple and exact: re-run this tile candidate-by-candidate without the queue
        qn = -1;
      }
      if (!kWeighted) {
        qn = min(qn, kKillCap);
        for (int q = tid; q < qn; q += kNmsThreads) {
          const uint32_t e = S.queue[q];
          const int j = tile0 + (e >> 8), t = e & 255;
          if (!((S.alive[j >> 5] >> (j & 31)) & 1u)) continue;  // already suppressed by another pair
          ++st_iou;
          if (pair_iou(S.frec[S.keptf[t]], recs[j]) > a.thr) atomicAnd(&S.alive[j >> 5], ~(1u << (j & 31)));
        }
      } else if (qn >= 0) {
        // pass 1: IoU of every queued pair; remember each candidate's FIRST suppressor (rank order)
        for (int q = tid; q < qn; q += kNmsThreads) {
          const uint32_t e = S.queue[q];
          const int jl = e >> 8, t = e & 255;
          const float iou = pair_iou(S.frec[S.keptf[t]], recs[tile0 + jl]);
          ++st_iou;
          S.qiou[q] = iou;
          if (iou > a.thr) {
            // atomicMin on a byte: CAS on the containing word
            uint32_t *wp = reinterpret_cast<uint32_t *>(S.firstsup + (jl & ~3));
            const int sh = (jl & 3) << 3;
            uint32_t old = *wp, assumed;
            do {
              assumed = old;
              const uint32_t cur = (assumed >> sh) & 255u;
              if (cur <= static_cast<uint32_t>(t)) break;
              old = atomicCAS(wp, assumed, (assumed & ~(255u << sh)) | (static_cast<uint32_t>(t) << sh));
            } while (old != assumed);
          }
        }
        __syncthreads();
        // pass 2: merges up to and including the first suppressor; the suppressor clears the bit
        for (int q = tid; q < qn; q += kNmsThreads) {
          const uint32_t e = S.queue[q];
          const int jl = e >> 8, t = e & 255;
          const int fs = S.firstsup[jl];
          if (t > fs) continue;
          const int j = tile0 + jl;
          if (S.qiou[q] > a.mthr) {
            const size_t rowj = static_cast<size_t>(beg + j) * a.D;
            const double sj = a.data[rowj + a.D - 1];
            double *acc = a.acc + static_cast<size_t>(kbase + kept_total - nk + t) * a.D;
            for (int c = 0; c < a.D - 1; ++c) atomicAdd(acc + c, sj * static_cast<double>(a.data[rowj + c]));
            atomicAdd(acc + a.D - 1, sj);
            atomicAdd(a.merge_count + kbase + kept_total - nk + t, 1);
          }
          if (t == fs) atomicAnd(&S.alive[j >> 5], ~(1u << (j & 31)));
        }
      } else {
        // exact serial fallback for an overflowing weighted tile: one thread per candidate
        for (int jl = tid; jl < kTile; jl += kNmsThreads) {
          const int j = tile0 + jl;
          if (j >= n || !((S.alive[j >> 5] >> (j & 31)) & 1u)) continue;
          const Rec rj = recs[j];
          const float jx = rec_cx(rj), jy = rec_cy(rj);
          for (int t = 0; t < nk; ++t) {
            const float dx = S.kx[t] - jx, dy = S.ky[t] - jy, rr = S.kr[t] + rj.r;
            if (!(dx * dx + dy * dy <= rr * rr)) continue;
            const float iou = pair_iou(S.frec[S.keptf[t]], rj);
            ++st_iou;
            if (iou > a.mthr) {
              const size_t rowj = static_cast<size_t>(beg + j) * a.D;
              const double sj = a.data[rowj + a.D - 1];
              double *acc = a.acc + static_cast<size_t>(kbase + kept_total - nk + t) * a.D;
              for (int c = 0; c < a.D - 1; ++c) atomicAdd(acc + c, sj * static_cast<double>(a.data[rowj + c]));
              atomicAdd(acc + a.D - 1, sj);
              atomicAdd(a.merge_count + kbase + kept_total - nk + t, 1);
            }
            if (iou > a.thr) { atomicAnd(&S.alive[j >> 5], ~(1u << (j & 31))); break; }
          }
        }
      }
      __syncthreads();
    }
    if (last_round) break;
  }

  if (tid == 0) a.kept_count[seg] = kept_total;
  if (a.stats) {
    // warp-reduce then one atomic per warp
    for (int o = 16; o; o >>= 1) {
      st_iou += __shfl_xor_sync(0xffffffffu, st_iou, o);
      st_circle += __shfl_xor_sync(0xffffffffu, st_circle, o);
    }
    if ((tid & 31) == 0) {
      atomicAdd(a.stats + 0, st_iou);
      atomicAdd(a.stats + 3, st_circle);
    }
    if (tid == 0) {
      atomicAdd(a.stats + 1, static_cast<unsigned long long>(kept_total));
      atomicAdd(a.stats + 2, static_cast<unsigned long long>(rounds));
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
};

__global__ void __launch_bounds__(128) pack_kernel(PackArgs a) {
  const int seg = blockIdx.x;
  const int cnt = a.kept_count[seg];
  const int off = a.out_off[seg], kb = a.kept_base[seg], beg = a.seg_begin[seg];
  for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
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
  L.kept_base = c.take<int>(S);
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
  L.cub_bytes = cub_sort_bytes(nn, end_bit);
  L.cub_tmp = c.take<unsigned char>(L.cub_bytes);
  L.total = align_up(c.used, 256);
  return L;
}

template <typename Rec, bool kWeighted>
static int launch_nms_segments(const NmsArgs &a, int S, int max_seg_n, cudaStream_t s) {
  const int nwords = (max_seg_n + 31) / 32;
  const size_t smem = nms_smem_bytes<Rec, kWeighted>(nwords);
  if (smem > 200 * 1024) return RV3D_ERR_ARG;  // num_pre_nms too large for the shared-memory bitmap
  RV3D_CHECK_CUDA(cudaFuncSetAttribute(nms_segment_kernel<Rec, kWeighted>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  nms_segment_kernel<Rec, kWeighted><<<S, kNmsThreads, smem, s>>>(a);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

}  // namespace rv3d

using namespace rv3d;

extern "C" size_t rv3d_nms_scratch_bytes(const rv3d_nms_params *p) {
  if (!p || p->batch <= 0 || p->total_classes <= 0 || p->n_candidates < 0) return 0;
  const int S = p->batch * p->total_classes;
  const int end_bit = bits_for(S) + 32 + bits_for(p->total_candidates);
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
  const int end_bit = bits_for(S) + 32 + idx_bits;
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
  segment_bounds_kernel<<<ceil_div(n, 256), 256, 0, s>>>(skeys, n, 32 + idx_bits, L.seg_begin, L.seg_end);
  RV3D_CHECK_LAUNCH();
  segment_capacity_scan_kernel<<<1, 1024, 0, s>>>(L.seg_begin, L.seg_end, S, p->num_post_nms, L.kept_base);
  RV3D_CHECK_LAUNCH();

  NmsArgs a{};
  a.recs = L.recs; a.seg_begin = L.seg_begin; a.seg_end = L.seg_end; a.kept_base = L.kept_base;
  a.num_pre = p->num_pre_nms; a.num_post = p->num_post_nms; a.thr = p->iou_threshold; a.mthr = p->merge_threshold;
  a.kept_pos = L.kept_pos; a.kept_count = L.kept_count; a.data = L.data; a.D = 9; a.acc = L.acc;
  a.merge_count = L.merge_count; a.stats = reinterpret_cast<unsigned long long *>(stats);
  const int max_seg_n = n < p->num_pre_nms ? n : p->num_pre_nms;
  int rc;
  if (weighted) {
    RV3D_CHECK_CUDA(cudaMemsetAsync(L.acc, 0, sizeof(double) * static_cast<size_t>(n) * 9, s));
    RV3D_CHECK_CUDA(cudaMemsetAsync(L.merge_count, 0, sizeof(int) * static_cast<size_t>(n), s));
    prepare_records_kernel<true><<<ceil_div(n, 256), 256, 0, s>>>(order, boxes, n, L.recs, L.data);
    RV3D_CHECK_LAUNCH();
    rc = launch_nms_segments<WRec, true>(a, S, max_seg_n, s);
  } else {
    prepare_records_kernel<false><<<ceil_div(n, 256), 256, 0, s>>>(order, boxes, n, L.recs, nullptr);
    RV3D_CHECK_LAUNCH();
    rc = launch_nms_segments<HardRec, false>(a, S, max_seg_n, s);
  }
  if (rc != RV3D_OK) return rc;

  kept_scan_kernel<<<1, 1024, 0, s>>>(L.kept_count, S, L.out_off, out_count, p->out_capacity);
  RV3D_CHECK_LAUNCH();
  PackArgs pa{};
  pa.seg_begin = L.seg_begin; pa.kept_base = L.kept_base; pa.kept_count = L.kept_count; pa.out_off = L.out_off;
  pa.kept_pos = L.kept_pos; pa.order = order; pa.boxes = boxes; pa.acc = L.acc;
  pa.total_classes = p->total_classes; pa.out_capacity = p->out_capacity; pa.weighted = weighted ? 1 : 0; pa.yaw_layout = p->out_layout == RV3D_OUT_YAW;
  pa.out_params = out_params; pa.out_scores = out_scores; pa.out_cats = out_categories; pa.out_batch = out_batch;
  pack_kernel<<<S, 128, 0, s>>>(pa);
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

// threshold-only branch: rows ordered by (sweep, candidate index)
__global__ void remap_keys_kernel(const unsigned long long *__restrict__ keys, int n, int idx_bits,
                                  int total_classes, unsigned long long *__restrict__ out, uint32_t *__restrict__ order,
                                  uint32_t *__restrict__ seg_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = keys[i];
  const uint32_t seg = static_cast<uint32_t>(k >> (32 + idx_bits));
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
  a.num_pre = n; a.num_post = n; a.thr = iou_threshold; a.mthr = 0.f;
  a.kept_pos = L.kept_pos; a.kept_count = L.kept_count;
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
  a.kept_pos = L.kept_pos; a.kept_count = L.kept_count; a.data = data; a.D = d; a.acc = L.acc;
  a.merge_count = L.merge_count;
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

extern "C" size_t rv3d_pack_candidates_scratch_bytes(int32_t n) { return op_layout(nullptr, n, false, 0, true).total; }

extern "C" int rv3d_pack_candidates(uint64_t *keys, const float *boxes, int32_t n, int32_t batch,
                                    int32_t total_classes, int32_t total_candidates, float *out_params,
                                    float *out_scores, int64_t *out_categories, int64_t *out_batch, void *scratch,
                                    size_t scratch_bytes, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && batch > 0 && total_classes > 0 && total_candidates > 0 && scratch);
  if (n == 0) return RV3D_OK;
  RV3D_CHECK_ARG(keys && boxes && out_params && out_scores && out_categories && out_batch);
  if (!aligned(scratch, 256) || !aligned(boxes, 16)) return RV3D_ERR_ALIGN;
  const OpLayout L = op_layout(scratch, n, false, 0, true);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int idx_bits = bits_for(total_candidates);
  remap_keys_kernel<<<ceil_div(n, 256), 256, 0, s>>>(reinterpret_cast<unsigned long long *>(keys), n, idx_bits,
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
