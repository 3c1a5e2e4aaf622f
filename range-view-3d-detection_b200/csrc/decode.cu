// decode.cu -- dense per-pixel box decoding (K2) and its free-standing operator forms.
//
// Replaces (paths relative to /root/reference):
//   math/ops/coding.py:110-144          decode_range_view  (+ :79-107 egovehicle_from_azimuth)
//   nn/decoders/range_decoder.py:49-77  sigmoid*mask, max over classes, decode, sample_by_range
//   nn/decoders/range_decoder.py:127-156 sample_by_range
//   math/ops/nms.py:212-215             min-confidence gather
//   math/linalg/lie/SO3.py:122-134      yaw_to_quat
//
// K2 (decode_compact) is one pass over the head outputs, one CTA per tile of 512 pixels (phases described at the
// kernel): TMA-staged logits / cart / mask -> class max + conservative logit bound + sample_by_range mask on the
// 4-pixels-per-thread lanes -> exact float32 sigmoid / threshold / first-index check on dense lanes -> block scan,
// ONE global atomicAdd per tile -> fp64 box decode (csrc/fastmath.cuh) of the emitting pixels on dense lanes,
// 1..n_partitions (key, box) rows each.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include <cmath>

#include "common.cuh"
#include "decode_math.cuh"
#include "fastmath.cuh"

namespace rv3d {

// ------------------------------------------------------------------------------------------
// decode_range_view as a dense operator: (B,8,H,W) + (B,3,H,W) -> (B,7,H,W)
// ------------------------------------------------------------------------------------------
// T = dtype of the regressands and of the result (coding.py:126,144); TC = dtype of cart, widened on its own
// (coding.py:128) -- under autocast the heads are half precision while cart stays float32
template <typename T, typename TC>
__global__ void __launch_bounds__(256)
decode_dense_kernel(const T *__restrict__ reg, const TC *__restrict__ cart, T *__restrict__ out, int HW,
                    int az_inv) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  float r[8], c[3];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = Ld<T>::one(reg + (static_cast<size_t>(b) * 8 + k) * HW + p);
#pragma unroll
  for (int k = 0; k < 3; ++k) c[k] = Ld<TC>::one(cart + (static_cast<size_t>(b) * 3 + k) * HW + p);
  double o[7];
  decode_box(r, c, az_inv != 0, o);
#pragma unroll
  for (int k = 0; k < 7; ++k) out[(static_cast<size_t>(b) * 7 + k) * HW + p] = Ld<T>::cast(o[k]);
}

// ------------------------------------------------------------------------------------------
// sample_by_range as a dense operator
// ------------------------------------------------------------------------------------------
struct PartArgs {
  int n;
  float lower[RV3D_MAX_PARTITIONS], upper[RV3D_MAX_PARTITIONS];
  int rate[RV3D_MAX_PARTITIONS], wsub[RV3D_MAX_PARTITIONS], off[RV3D_MAX_PARTITIONS + 1];
  // w / rate without an integer division: shift when rate is a power of two, else
  // umulhi(w, magic) with magic = floor(2^32 / rate) + 1 (exact for w * rate < 2^32, i.e. any image width)
  int shift[RV3D_MAX_PARTITIONS];
  uint32_t magic[RV3D_MAX_PARTITIONS];
};
__device__ __forceinline__ int fast_div(int w, int shift, uint32_t magic) {
  return shift >= 0 ? (w >> shift) : static_cast<int>(__umulhi(static_cast<uint32_t>(w), magic));
}

static PartArgs make_parts(const rv3d_partitions *p, int H, int W) {
  PartArgs a{};
  a.n = p ? p->n_partitions : 0;
  int off = 0;
  for (int i = 0; i < a.n; ++i) {
    a.lower[i] = p->lower[i]; a.upper[i] = p->upper[i]; a.rate[i] = p->rate[i];
    a.wsub[i] = (W + p->rate[i] - 1) / p->rate[i];
    a.shift[i] = -1;
    for (int sh = 0; sh < 31; ++sh)
      if (p->rate[i] == (1 << sh)) a.shift[i] = sh;
    a.magic[i] = static_cast<uint32_t>((1ull << 32) / static_cast<unsigned long long>(p->rate[i])) + 1u;
    a.off[i] = off;
    off += H * a.wsub[i];
  }
  a.off[a.n] = off;
  if (a.n == 0) a.off[0] = H * W;
  return a;
}

// The first kRegParts partitions in the form the decode kernel's filter phase wants them (registers, no branches):
//   * range test on the SUM OF SQUARES: d = fl(sqrtf(s)) is monotone in s, so `d > lower` and `d <= upper` are
//     `s > S(lower)` and `s <= S(upper)` with S(b) = the largest float whose correctly rounded square root is <= b
//     (found on the host by stepping through neighbouring floats; CUDA's sqrtf under --prec-sqrt=true and glibc's are
//     both correctly rounded).  No square root per pixel, bit-identical decisions.  Float32 cart only.
//   * column stride for 4 consecutive columns w0 .. w0+3 at once: with n = (-w0) mod rate the multiples of `rate` among
//     them are the set bits of (pattern << n) & 0xF, pattern = {bit k*rate : k*rate < 4}.
constexpr int kRegParts = 4;
struct FilterParts {
  float s_lo[kRegParts], s_hi[kRegParts];
  float lower[kRegParts], upper[kRegParts];
  int rate[kRegParts], shift[kRegParts];
  uint32_t magic[kRegParts], pattern[kRegParts];
};

static float sumsq_bound(float b) {
  if (b != b) return b;                         // NaN: every comparison fails, like the reference's
  if (b < 0.0f) return -1.0f;                   // no norm is <= b
  if (std::isinf(b)) return b;
  double sq = static_cast<double>(b) * static_cast<double>(b);
  float s = sq >= 3.4028234663852886e38 ? 3.4028234663852886e38f : static_cast<float>(sq);
  while (s > 0.0f && std::sqrt(s) > b) s = std::nextafterf(s, -INFINITY);
  while (s < 3.4028234663852886e38f && std::sqrt(std::nextafterf(s, INFINITY)) <= b) s = std::nextafterf(s, INFINITY);
  return s;
}

static FilterParts make_filter_parts(const PartArgs &pa) {
  FilterParts f{};
  for (int i = 0; i < kRegParts; ++i) {
    const bool on = i < pa.n;
    f.lower[i] = on ? pa.lower[i] : INFINITY; f.upper[i] = on ? pa.upper[i] : -INFINITY;   // unused slots never match
    f.s_lo[i] = on ? sumsq_bound(pa.lower[i]) : INFINITY; f.s_hi[i] = on ? sumsq_bound(pa.upper[i]) : -INFINITY;
    f.rate[i] = on ? pa.rate[i] : 1; f.shift[i] = on ? pa.shift[i] : 0; f.magic[i] = on ? pa.magic[i] : 0u;
    uint32_t pat = 0u;
    for (int k = 0; on && k * pa.rate[i] < 4; ++k) pat |= 1u << (k * pa.rate[i]);
    f.pattern[i] = pat;
  }
  return f;
}

// ||cart||_2 in float32, the way torch's CPU / CUDA vector_norm reduces three floats
__device__ __forceinline__ float norm3(float x, float y, float z) { return sqrtf((x * x + y * y) + z * z); }

__global__ void __launch_bounds__(256)
sample_by_range_kernel(const float *__restrict__ scores, const int64_t *__restrict__ cats,
                       const float *__restrict__ cub, const float *__restrict__ cart, PartArgs pa, int H, int W,
                       float *__restrict__ o_scores, int64_t *__restrict__ o_cats, float *__restrict__ o_cub) {
  const int b = blockIdx.y;
  const int K = pa.off[pa.n];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int i = 0;
  while (i + 1 < pa.n && k >= pa.off[i + 1]) ++i;
  const int local = k - pa.off[i];
  const int h = local / pa.wsub[i];
  const int w = (local - h * pa.wsub[i]) * pa.rate[i];
  const int HW = H * W, p = h * W + w;
  const float *c = cart + static_cast<size_t>(b) * 3 * HW + p;
  const float d = norm3(c[0], c[HW], c[2 * HW]);
  const bool in = (d > pa.lower[i]) && (d <= pa.upper[i]);
  const float s = scores[static_cast<size_t>(b) * HW + p];
  o_scores[static_cast<size_t>(b) * K + k] = in ? s : s * 0.0f;  // scores * partition (NaN/inf semantics kept)
  o_cats[static_cast<size_t>(b) * K + k] = cats[static_cast<size_t>(b) * HW + p];
  float *oc = o_cub + (static_cast<size_t>(b) * K + k) * 7;
#pragma unroll
  for (int j = 0; j < 7; ++j) oc[j] = cub[(static_cast<size_t>(b) * 7 + j) * HW + p];
}

// ------------------------------------------------------------------------------------------
// K2: fused decode + threshold + compaction
// ------------------------------------------------------------------------------------------
// 64-bit sort key: [ segment = sweep * total_classes + class | ~orderable(score) | candidate ].
// Ascending key order == segment asc, score desc, candidate index asc: the total order the
// suppression step works in (DESIGN.md "K3").
struct KeyPack {
  int idx_bits;    // width of the low field (candidate index)
  int score_bits;  // 32, or 31 when every score is >= 0 (then the top bit of the order code is constant):
                   // one radix pass less at the Waymo shape (6 + 31 + 19 = 56 key bits instead of 57)
  __device__ __forceinline__ unsigned long long make(uint32_t seg, float score, uint32_t cand) const {
    uint32_t desc = ~orderable_f32(__float_as_uint(score));
    if (score_bits < 32) desc &= (1u << score_bits) - 1u;
    return (static_cast<unsigned long long>(seg) << (score_bits + idx_bits)) |
           (static_cast<unsigned long long>(desc) << idx_bits) | cand;
  }
  // the decode kernel's form: score_bits is the compile-time RV3D_SCORE_BITS_DECODE and the score is >= 0
  __device__ __forceinline__ unsigned long long make_nonneg(uint32_t seg, float score, uint32_t cand) const {
    const uint32_t desc = ~__float_as_uint(score) & ((1u << RV3D_SCORE_BITS_DECODE) - 1u);   // ~(bits | 0x80000000), 31 bits
    return (static_cast<unsigned long long>(seg) << (RV3D_SCORE_BITS_DECODE + idx_bits)) |
           (static_cast<unsigned long long>(desc) << idx_bits) | cand;
  }
};

struct DecodeArgs {
  int B, C, H, W, az_inv, cat_off, cand_off, total_classes, capacity;
  float thr;
  float inv_w;  // 1 / W (row of a tile's first pixel; corrected by one step on the device)
  float x_lo;   // conservative logit bound: sigmoid_T(x) >= thr implies x >= x_lo (logit_lower_bound)
  KeyPack kp;
  PartArgs pa;
  FilterParts fp;
};

// ---- TMA 1-D bulk copy + mbarrier (inline PTX; SASS: UBLKCP / SYNCS) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA prefetch of a contiguous span into L2 (no shared memory, no completion to wait for; SASS: UBLKPF)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// shared-memory reads of 4 consecutive elements, widened to float
template <typename T> struct Sm;
template <> struct Sm<float> {
  static __device__ __forceinline__ void four(const float *p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ float one(const float *p) { return *p; }
};
template <> struct Sm<__half> {
  static __device__ __forceinline__ void four(const __half *p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2 *>(p);
    const __half2 a = *reinterpret_cast<const __half2 *>(&t.x), b = *reinterpret_cast<const __half2 *>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ float one(const __half *p) { return __half2float(*p); }
};
template <> struct Sm<__nv_bfloat16> {
  static __device__ __forceinline__ void four(const __nv_bfloat16 *p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2 *>(p);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
  static __device__ __forceinline__ float one(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
};

constexpr int kPxPerThread = 4;

// bytes of dynamic shared memory for a tile of kTile pixels
static size_t decode_smem_bytes(int C, int tile, size_t elem, size_t cart_elem) {
  size_t b = static_cast<size_t>(C) * tile * elem + static_cast<size_t>(3) * tile * cart_elem;   // logit and cart planes (regressands are gathered)
  b += tile;                                              // mask
  b = align_up(b, 16);
  b += static_cast<size_t>(tile) * (4 + 2 + 1 + 1);   // q_score, q_pix, q_cls, q_emit
  return align_up(b, 16) + 64;
}

// Running max over the classes for 4 consecutive pixels (first index wins ties; NaN never becomes the max).
//   runner   max logit among the classes BEFORE the winner (for the first-index tie check of the score phase)
//   nansum   sum of the logits: NaN iff some logit is NaN (or +inf and -inf meet) -- one FADD per (pixel, class)
//            instead of a comparison + predicate merge; the rare NaN sum is re-examined exactly by the caller
// kC > 0: class count known at compile time (fully unrolled, everything stays in registers / predicates).
template <typename T, int kC, int kTile>
__device__ __forceinline__ void class_max(const T *s_logits, int lp0, int C, float (&best)[4], float (&runner)[4],
                                          float (&nansum)[4], int (&cls)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) { best[j] = -CUDART_INF_F; runner[j] = -CUDART_INF_F; nansum[j] = 0.f; cls[j] = 0; }
  const int n = kC > 0 ? kC : C;
#pragma unroll
  for (int c = 0; c < n; ++c) {
    float v[4];
    Sm<T>::four(s_logits + c * kTile + lp0, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      nansum[j] += v[j];
      const bool gt = v[j] > best[j];
      runner[j] = gt ? best[j] : runner[j];
      cls[j] = gt ? c : cls[j];
      best[j] = gt ? v[j] : best[j];
    }
  }
}


// One CTA per tile of kTile consecutive pixels of one sweep, kTile / 4 threads.
//   stage    the logit, cart and mask planes' slices of the tile are brought into shared memory: with kBulk,
//            C + 4 TMA 1-D bulk copies issued by one thread, completion on an mbarrier (no registers, no
//            per-thread loads; several resident CTAs overlap each other's copies and math); without
//            (unaligned shapes) plain cooperative loads into the same layout.  The 8 regressand planes are
//            NOT staged: only emitting pixels need them, the last phase gathers them (8 independent loads per pixel).
//   filter   every thread owns 4 consecutive pixels: running max over the classes (first index wins, NaN kills the
//            pixel like torch.max) and ONE comparison against a conservative logit bound x_lo (sigmoid_T(x) >= thr
//            implies x >= x_lo, computed on the host with slack for the float32 / half rounding); for threads with a
//            survivor, the sample_by_range mask of the 4 pixels (FilterParts: range test on the sum of squares, stride
//            test for 4 columns at once).  Pixels that may pass AND can emit go to the queue (warp-aggregated
//            shared-memory reservation).  No sigmoid, no square root, no division on these sparse lanes: with one pixel
//            in three alive a per-pixel branch costs the whole warp its full price (ncu r02a: sigmoid + threshold +
//            partition + stride tests were 36 % of the kernel's instructions at 30 % lane use).
//   score    dense lanes over the queue: float32 sigmoid the way the reference's CUDA path rounds it, exact threshold,
//            first-index tie check (only where the filter flagged a close runner-up).
//   scan     block scan of the emitted rows; ONE global atomicAdd per tile reserves the output rows.
//   decode   every thread decodes the queue entries it scored: fp64 decode (3 exp, 1-2 atan2, rotation), 1..n_partitions
//            (key, box) rows each.
template <typename T, typename TC, int kTile, bool kBulk>
__global__ void __launch_bounds__(kTile / kPxPerThread, kTile == 512 ? 9 : 12)
decode_compact_kernel(DecodeArgs a, const T *__restrict__ logits, const T *__restrict__ reg,
                      const TC *__restrict__ cart, const uint8_t *__restrict__ mask,
                      unsigned long long *__restrict__ out_keys, float *__restrict__ out_boxes,
                      int32_t *__restrict__ counter) {
  constexpr int kThreads = kTile / kPxPerThread;
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ uint32_t s_scan[kWarps];
  __shared__ uint32_t s_base, s_q1n;
  __shared__ __align__(8) uint64_t s_bar;
  // partition constants out of the kernel-parameter bank (dynamic indexing there costs uniform moves)
  __shared__ float s_lower[RV3D_MAX_PARTITIONS], s_upper[RV3D_MAX_PARTITIONS];
  __shared__ int s_rate[RV3D_MAX_PARTITIONS], s_shift[RV3D_MAX_PARTITIONS], s_off[RV3D_MAX_PARTITIONS], s_wsub[RV3D_MAX_PARTITIONS];
  __shared__ uint32_t s_magic[RV3D_MAX_PARTITIONS];

  const int b = blockIdx.y;
  const int HW = a.H * a.W;
  const int tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5;
  const int blk0 = blockIdx.x * kTile;
  const int npx = min(kTile, HW - blk0);

  // ---- carve shared memory: [C][kTile] logits | [3][kTile] cart | mask | queues
  T *s_logits = reinterpret_cast<T *>(dsm);
  TC *s_cart = reinterpret_cast<TC *>(s_logits + static_cast<size_t>(a.C) * kTile);   // C * kTile * sizeof(T) is a multiple of 16
  uint8_t *s_mask = reinterpret_cast<uint8_t *>(s_cart + 3 * kTile);
  unsigned char *qp = dsm + align_up_c(static_cast<size_t>(a.C) * kTile * sizeof(T) + 3 * kTile * sizeof(TC) + kTile, 16);
  float *q_score = reinterpret_cast<float *>(qp);                    // queue: score of entry t
  uint16_t *q_pix = reinterpret_cast<uint16_t *>(q_score + kTile);   // queue: local pixel id
  uint8_t *q_cls = reinterpret_cast<uint8_t *>(q_pix + kTile);       // queue: class index within the task (C <= 128 checked on the host)
  uint8_t *q_emit = q_cls + kTile;                                   // queue: partition bit-mask

  if (tid < RV3D_MAX_PARTITIONS) {
    s_lower[tid] = a.pa.lower[tid]; s_upper[tid] = a.pa.upper[tid]; s_rate[tid] = a.pa.rate[tid];
    s_shift[tid] = a.pa.shift[tid]; s_magic[tid] = a.pa.magic[tid]; s_off[tid] = a.pa.off[tid]; s_wsub[tid] = a.pa.wsub[tid];
  }
  if (tid == 0) s_q1n = 0u;
  const int n_parts = a.pa.n;

  const T *rg = reg + static_cast<size_t>(b) * 8 * HW + blk0;   // regressands: gathered in the decode phase for the emitting pixels only

  // ---------------- stage the tile ----------------
  {
    const T *lg = logits + static_cast<size_t>(b) * a.C * HW + blk0;
    const TC *ct = cart + static_cast<size_t>(b) * 3 * HW + blk0;
    const uint8_t *mk = mask + static_cast<size_t>(b) * HW + blk0;
    if (kBulk) {
      if (tid == 0) mbar_init(&s_bar, 1);
      __syncthreads();
      if (tid == 0) {
        const uint32_t plane = static_cast<uint32_t>(npx) * sizeof(T);
        const uint32_t cplane = static_cast<uint32_t>(npx) * sizeof(TC);
        mbar_expect_tx(&s_bar, plane * a.C + cplane * 3 + npx);
        for (int c = 0; c < a.C; ++c) bulk_g2s(s_logits + c * kTile, lg + static_cast<size_t>(c) * HW, plane, &s_bar);
        for (int k = 0; k < 3; ++k) bulk_g2s(s_cart + k * kTile, ct + static_cast<size_t>(k) * HW, cplane, &s_bar);
        bulk_g2s(s_mask, mk, npx, &s_bar);
      }
      mbar_wait(&s_bar, 0);
    } else {
      __syncthreads();
      for (int i = tid; i < npx; i += kThreads) {
        for (int c = 0; c < a.C; ++c) s_logits[static_cast<size_t>(c) * kTile + i] = lg[static_cast<size_t>(c) * HW + i];
        for (int k = 0; k < 3; ++k) s_cart[k * kTile + i] = ct[static_cast<size_t>(k) * HW + i];
        s_mask[i] = mk[i];
      }
      __syncthreads();
    }
  }

  // (row, col) of the tile's first pixel: float reciprocal + one correction step instead of an integer division
  int row0b = __float2int_rz(__int2float_rn(blk0) * a.inv_w);
  int col0b = blk0 - row0b * a.W;
  if (col0b < 0) { col0b += a.W; --row0b; }
  if (col0b >= a.W) { col0b -= a.W; ++row0b; }

  // ---------------- filter: class max, conservative logit bound, partition / stride mask ----------------
  const bool zero_passes = 0.0f >= a.thr;
  {
    const int lp0 = tid * kPxPerThread;          // first local pixel of this thread
    float best[kPxPerThread], runner[kPxPerThread], nansum[kPxPerThread];
    int cls[kPxPerThread];
    if (a.C == 3) class_max<T, 3, kTile>(s_logits, lp0, 3, best, runner, nansum, cls);
    else class_max<T, 0, kTile>(s_logits, lp0, a.C, best, runner, nansum, cls);
    const uchar4 m4 = *reinterpret_cast<const uchar4 *>(s_mask + lp0);
    const uint32_t mbits = (m4.x ? 1u : 0u) | (m4.y ? 2u : 0u) | (m4.z ? 4u : 0u) | (m4.w ? 8u : 0u);
    uint32_t maybe = 0u;
#pragma unroll
    for (int j = 0; j < kPxPerThread; ++j) {
      bool bad = false;
      if (nansum[j] != nansum[j]) {             // rare: a NaN logit, or infinities of both signs
        for (int c = 0; c < a.C; ++c) {
          const float x = Sm<T>::one(s_logits + c * kTile + lp0 + j);
          bad |= (x != x);
        }
      }
      // a masked pixel scores sigmoid * 0 == 0 for every class: it survives only when 0 >= thr
      const bool pass = zero_passes || (((mbits >> j) & 1u) && best[j] >= a.x_lo);
      if ((lp0 + j < npx) && !bad && pass) maybe |= 1u << j;
      // bit 7 of the queued class: an earlier class may round to the same float32 sigmoid (checked in the score phase)
      if (cls[j] > 0 && (best[j] >= 15.f || runner[j] > best[j] - 1.0f)) cls[j] |= 0x80;
    }
    // sample_by_range for the pixels that may pass: bit i of emit = the pixel's range lies in partition i (or 0 >= thr,
    // where out-of-partition copies survive with score 0) AND its column is a multiple of rate_i.  A pixel whose mask
    // comes out empty can never produce a row, whatever its exact score: it leaves here, on the 4-pixels-per-thread
    // lanes, before the sigmoid (at the bench workload 95 % of the pixels pass the threshold but 2 in 3 are strided out).
    uint32_t emit[kPxPerThread];
#pragma unroll
    for (int j = 0; j < kPxPerThread; ++j) emit[j] = n_parts == 0 ? 1u : 0u;
    if (maybe && n_parts > 0) {
      float cxv[4], cyv[4], czv[4];
      Sm<TC>::four(s_cart + lp0, cxv);
      Sm<TC>::four(s_cart + kTile + lp0, cyv);
      Sm<TC>::four(s_cart + 2 * kTile + lp0, czv);
      int w0 = col0b + lp0;
      while (w0 >= a.W) w0 -= a.W;
      // stride masks of the 4 columns, 4 bits per partition (FilterParts); a thread whose pixels straddle the end of
      // an image row (one in W / 4) tests its columns one by one instead
      const bool straddle = w0 + (kPxPerThread - 1) >= a.W;
      uint32_t M = 0u;
#pragma unroll
      for (int i = 0; i < kRegParts; ++i) {
        const int R = a.fp.rate[i];
        const int r0 = a.fp.shift[i] >= 0 ? (w0 & (R - 1)) : (w0 - static_cast<int>(__umulhi(static_cast<uint32_t>(w0), a.fp.magic[i])) * R);
        const int n = r0 ? R - r0 : 0;
        M |= (n < 4 ? ((a.fp.pattern[i] << n) & 0xFu) : 0u) << (4 * i);
      }
#pragma unroll
      for (int j = 0; j < kPxPerThread; ++j) {
        uint32_t part_in = 0u, stride_ok;
        if (sizeof(TC) == 4) {
          const float ss = (cxv[j] * cxv[j] + cyv[j] * cyv[j]) + czv[j] * czv[j];   // norm3 without its square root (FilterParts)
#pragma unroll
          for (int i = 0; i < kRegParts; ++i) part_in |= ((ss > a.fp.s_lo[i]) && (ss <= a.fp.s_hi[i])) ? (1u << i) : 0u;
        } else {
          // cart.norm(dim=1) accumulates in float32 and rounds to cart's dtype; the bounds are a float32 tensor
          const float d = Ld<TC>::round_f32(norm3(cxv[j], cyv[j], czv[j]));
#pragma unroll
          for (int i = 0; i < kRegParts; ++i) part_in |= ((d > a.fp.lower[i]) && (d <= a.fp.upper[i])) ? (1u << i) : 0u;
        }
        if (!straddle) {
          const uint32_t x = (M >> j) & 0x1111u;
          stride_ok = (x | (x >> 3) | (x >> 6) | (x >> 9)) & 0xFu;
        } else {
          int w = w0 + j;
          if (w >= a.W) w -= a.W;
          stride_ok = 0u;
          for (int i = 0; i < n_parts && i < kRegParts; ++i)
            stride_ok |= (w - fast_div(w, s_shift[i], s_magic[i]) * s_rate[i]) ? 0u : (1u << i);
        }
        if (n_parts > kRegParts) {                 // partitions beyond the register set: the plain form
          const float d = Ld<TC>::round_f32(norm3(cxv[j], cyv[j], czv[j]));
          int w = w0 + j;
          if (w >= a.W) w -= a.W;
          for (int i = kRegParts; i < n_parts; ++i) {
            part_in |= ((d > s_lower[i]) && (d <= s_upper[i])) ? (1u << i) : 0u;
            stride_ok |= (w - fast_div(w, s_shift[i], s_magic[i]) * s_rate[i]) ? 0u : (1u << i);
          }
        }
        emit[j] = (zero_passes ? ((1u << n_parts) - 1u) : part_in) & stride_ok;
        if (!emit[j]) maybe &= ~(1u << j);
      }
    }
    // warp-aggregated reservation in the queue (its order is irrelevant: every row carries its own key)
    const uint32_t cnt = __popc(maybe);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    uint32_t wbase = 0u;
    if (lane == 31 && incl) wbase = atomicAdd(&s_q1n, incl);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    uint32_t pos = wbase + incl - cnt;
#pragma unroll
    for (int j = 0; j < kPxPerThread; ++j) {
      if (maybe & (1u << j)) {
        q_pix[pos] = static_cast<uint16_t>(lp0 + j);
        q_cls[pos] = static_cast<uint8_t>(cls[j]);
        q_emit[pos] = static_cast<uint8_t>(emit[j]);
        ++pos;
      }
    }
  }
  __syncthreads();
  const uint32_t n1 = s_q1n;
  // The decode phase gathers the 8 regressand planes of the emitting pixels; when a good part of the tile is still
  // alive, one thread asks the TMA unit to pull the tile's slices of those planes into L2 now, so that the gathers
  // (a dependent DRAM round trip per warp otherwise: ncu r02f long_scoreboard 3.4 per issue) find them there.
  if (kBulk && tid == 0 && n1 * 8u >= static_cast<uint32_t>(kTile)) {
    const uint32_t plane = static_cast<uint32_t>(npx) * sizeof(T);
    for (int k = 0; k < 8; ++k) bulk_prefetch_l2(rg + static_cast<size_t>(k) * HW, plane);
  }

  // ---------------- score: exact sigmoid / threshold / first index, dense lanes ----------------
  uint32_t n_emit = 0;
  for (uint32_t t = tid; t < n1; t += kThreads) {
    const int lp = q_pix[t];
    const int qc = q_cls[t];
    int cls = qc & 0x7f;
    float score = 0.f;
    if (s_mask[lp]) {
      const float best = Sm<T>::one(s_logits + cls * kTile + lp);
      score = sigmoid_t<T>(best);
      // torch.max returns the FIRST index attaining the max of the float32 scores: an earlier class with a smaller
      // logit can round to the same float32 sigmoid (always when saturated)
      if ((qc & 0x80) && score >= a.thr) {
        for (int c = 0; c < cls; ++c) {
          const float x = Sm<T>::one(s_logits + c * kTile + lp);
          if ((best >= 15.f || x > best - 1.0f) && sigmoid_t<T>(x) == score) { cls = c; break; }
        }
      }
    } else {
      cls = 0;  // sigmoid * 0 == 0 for every class -> argmax 0
    }
    uint32_t emit = q_emit[t];
    if (!(score >= a.thr)) emit = 0u;             // 0 >= thr: every score (>= 0) passes
    q_cls[t] = static_cast<uint8_t>(cls);
    q_emit[t] = static_cast<uint8_t>(emit);
    q_score[t] = score;
    n_emit += __popc(emit);
  }

  // ---------------- block scan of the emitted rows ----------------
  // Every warp adds up the warp totals itself: no serial warp-0 step, one barrier.  The ONE global atomicAdd per tile
  // that reserves the block's output rows is issued by thread 0 right after that barrier; a thread decodes the queue
  // entries it scored, so no shared-memory hand-over is needed in between.
  uint32_t incl = n_emit;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_scan[wid] = incl;
  __syncthreads();
  uint32_t before = 0u, total = 0u;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    const uint32_t v = s_scan[w];
    before += w < wid ? v : 0u;
    total += v;
  }
  if (tid == 0) s_base = total ? static_cast<uint32_t>(atomicAdd(counter, static_cast<int>(total))) : 0u;
  __syncthreads();
  uint32_t row = s_base + before + (incl - n_emit);

  // ---------------- decode: dense-lane fp64 decode of the emitting pixels ----------------
  for (uint32_t t = tid; t < n1; t += kThreads) {
    uint32_t e = q_emit[t];
    if (!e) continue;                              // failed the exact threshold (only pixels inside the bound's slack)
    const int lp = q_pix[t];
    const int p = blk0 + lp;
    float r[8], c[3];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = Ld<T>::one(rg + static_cast<size_t>(k) * HW + lp);   // 8 independent loads in flight
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = Sm<TC>::one(s_cart + k * kTile + lp);
    double o[7];
    decode_box(r, c, a.az_inv != 0, o);
    // decode_range_view casts back to the input dtype (coding.py:144); widen that to f32
    float box[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) box[k] = static_cast<float>(Ld<T>::cast(o[k]));
    const uint32_t seg = static_cast<uint32_t>(b) * a.total_classes + q_cls[t] + a.cat_off;
    int h = row0b, w = col0b + lp;
    while (w >= a.W) { w -= a.W; ++h; }
    const float sc = q_score[t];
    while (e) {
      const int i = __ffs(e) - 1;
      e &= e - 1;
      float s_out = sc;
      uint32_t cand;
      if (n_parts == 0) {
        cand = p;
      } else {
        cand = s_off[i] + h * s_wsub[i] + fast_div(w, s_shift[i], s_magic[i]);
        // the partition that does not contain the pixel contributes score 0 (only when 0 >= thr: otherwise a
        // pixel only emits into partitions that contain it)
        if (zero_passes) {
          const float d = Ld<TC>::round_f32(norm3(c[0], c[1], c[2]));
          if (!((d > s_lower[i]) && (d <= s_upper[i]))) s_out = 0.0f;
        }
      }
      if (row < static_cast<uint32_t>(a.capacity)) {
        out_keys[row] = a.kp.make_nonneg(seg, s_out, cand + a.cand_off);
        float4 *ob = reinterpret_cast<float4 *>(out_boxes + static_cast<size_t>(row) * 8);
        ob[0] = make_float4(box[0], box[1], box[2], box[3]);
        ob[1] = make_float4(box[4], box[5], box[6], s_out);
      }
      ++row;
    }
  }
}

// dense candidates -> (key, box) rows
__global__ void __launch_bounds__(256)
compact_candidates_kernel(const float *__restrict__ cub, const float *__restrict__ scores,
                          const int64_t *__restrict__ cats, int K, int total_classes, float thr, int apply_thr,
                          int capacity, KeyPack kp, unsigned long long *__restrict__ out_keys,
                          float *__restrict__ out_boxes, int32_t *__restrict__ counter) {
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_base;
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  bool live = false;
  float s = 0.f;
  if (k < K) {
    s = scores[static_cast<size_t>(b) * K + k];
    live = apply_thr ? (s >= thr) : true;
  }
  const uint32_t bal = __ballot_sync(0xffffffffu, live);
  if (lane == 0) s_warp[wid] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
    for (int i = 0; i < 8; ++i) { const uint32_t c = s_warp[i]; s_warp[i] = tot; tot += c; }
    s_base = tot ? static_cast<uint32_t>(atomicAdd(counter, static_cast<int>(tot))) : 0u;
  }
  __syncthreads();
  if (!live) return;
  const uint32_t row = s_base + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
  if (row >= static_cast<uint32_t>(capacity)) return;
  const float *c = cub + (static_cast<size_t>(b) * K + k) * 7;
  const uint32_t seg = static_cast<uint32_t>(b) * total_classes + static_cast<uint32_t>(cats[static_cast<size_t>(b) * K + k]);
  out_keys[row] = kp.make(seg, s, static_cast<uint32_t>(k));
  float4 *ob = reinterpret_cast<float4 *>(out_boxes + static_cast<size_t>(row) * 8);
  ob[0] = make_float4(c[0], c[1], c[2], c[3]);
  ob[1] = make_float4(c[4], c[5], c[6], s);
}

__global__ void yaw_to_quat_kernel(const float *__restrict__ yaw, float *__restrict__ quat, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s, c;
  sincos(static_cast<double>(yaw[i] * 0.5f), &s, &c);
  reinterpret_cast<float4 *>(quat)[i] = make_float4(static_cast<float>(c), 0.f, 0.f, static_cast<float>(s));
}

template <typename T, typename TC, int kTile, bool kBulk>
static int launch_decode_tile(const DecodeArgs &a, const void *logits, const void *reg, const void *cart,
                              const uint8_t *mask, unsigned long long *keys, float *boxes, int32_t *counter,
                              cudaStream_t s) {
  const int HW = a.H * a.W;
  const size_t smem = decode_smem_bytes(a.C, kTile, sizeof(T), sizeof(TC));
  if (smem > 200 * 1024) return RV3D_ERR_ARG;   // too many classes for one tile
  RV3D_CHECK_CUDA(cudaFuncSetAttribute(decode_compact_kernel<T, TC, kTile, kBulk>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(ceil_div(HW, kTile), a.B);
  decode_compact_kernel<T, TC, kTile, kBulk><<<grid, kTile / kPxPerThread, smem, s>>>(
      a, static_cast<const T *>(logits), static_cast<const T *>(reg), static_cast<const TC *>(cart), mask, keys, boxes,
      counter);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

template <typename T, typename TC>
static int launch_decode_compact(const DecodeArgs &a, const void *logits, const void *reg, const void *cart,
                                 const uint8_t *mask, unsigned long long *keys, float *boxes, int32_t *counter,
                                 cudaStream_t s) {
  const int HW = a.H * a.W;
  // TMA bulk copies need 16-byte aligned sources and sizes: every plane slice of a tile must start on a
  // 16-byte boundary (HW % 16 == 0 covers the 1-byte mask plane too)
  const bool bulk = (HW % 16 == 0) && aligned(logits, 16) && aligned(cart, 16) && aligned(mask, 16);
  // tile size: keep a CTA's stage under ~48 KB so 4-6 CTAs are resident per SM
  const bool small_tile = (static_cast<size_t>(a.C) * sizeof(T) + 3 * sizeof(TC)) * 512 > 40 * 1024;
  if (bulk) {
    if (small_tile) return launch_decode_tile<T, TC, 256, true>(a, logits, reg, cart, mask, keys, boxes, counter, s);
    return launch_decode_tile<T, TC, 512, true>(a, logits, reg, cart, mask, keys, boxes, counter, s);
  }
  if (small_tile) return launch_decode_tile<T, TC, 256, false>(a, logits, reg, cart, mask, keys, boxes, counter, s);
  return launch_decode_tile<T, TC, 512, false>(a, logits, reg, cart, mask, keys, boxes, counter, s);
}

}  // namespace rv3d

using namespace rv3d;

template <typename T, typename TC>
static void launch_decode_dense(const void *reg, const void *cart, void *out, int HW, int batch, int az_inv, cudaStream_t s) {
  dim3 grid(ceil_div(HW, 256), batch);
  decode_dense_kernel<T, TC><<<grid, 256, 0, s>>>(static_cast<const T *>(reg), static_cast<const TC *>(cart),
                                                  static_cast<T *>(out), HW, az_inv);
}

// the dtype pairs the path meets: everything in one dtype, or half-precision heads with float32 cart (autocast)
#define RV3D_DISPATCH_DTYPES(dtype, cart_dtype, CALL)                                             \
  do {                                                                                            \
    if ((dtype) == RV3D_F32 && (cart_dtype) == RV3D_F32) { CALL(float, float); }                  \
    else if ((dtype) == RV3D_F16 && (cart_dtype) == RV3D_F16) { CALL(__half, __half); }           \
    else if ((dtype) == RV3D_BF16 && (cart_dtype) == RV3D_BF16) { CALL(__nv_bfloat16, __nv_bfloat16); } \
    else if ((dtype) == RV3D_F16 && (cart_dtype) == RV3D_F32) { CALL(__half, float); }            \
    else if ((dtype) == RV3D_BF16 && (cart_dtype) == RV3D_F32) { CALL(__nv_bfloat16, float); }    \
    else return RV3D_ERR_ARG;                                                                     \
  } while (0)

extern "C" int rv3d_decode_range_view(const void *regressands, const void *cart, void *out, int32_t dtype,
                                      int32_t cart_dtype, int32_t batch, int32_t height, int32_t width,
                                      int32_t azimuth_invariant, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(regressands && cart && out && batch > 0 && height > 0 && width > 0);
  const int HW = height * width;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define RV3D_CALL(T, TC) launch_decode_dense<T, TC>(regressands, cart, out, HW, batch, azimuth_invariant, s)
  RV3D_DISPATCH_DTYPES(dtype, cart_dtype, RV3D_CALL);
#undef RV3D_CALL
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

static bool parts_ok(const rv3d_partitions *p) {
  if (!p) return true;
  if (p->n_partitions < 0 || p->n_partitions > RV3D_MAX_PARTITIONS) return false;
  for (int i = 0; i < p->n_partitions; ++i)
    if (p->rate[i] <= 0) return false;
  return true;
}

extern "C" int64_t rv3d_num_candidates(const rv3d_partitions *parts, int32_t height, int32_t width) {
  if (!parts_ok(parts) || height <= 0 || width <= 0) return -1;
  const PartArgs a = make_parts(parts, height, width);
  return a.off[a.n];
}

extern "C" int rv3d_sample_by_range(const float *scores, const int64_t *categories, const float *cuboids,
                                    const float *cart, const rv3d_partitions *parts, int32_t batch,
                                    int32_t height, int32_t width, float *out_scores, int64_t *out_categories,
                                    float *out_cuboids, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(scores && categories && cuboids && cart && parts && out_scores && out_categories && out_cuboids);
  RV3D_CHECK_ARG(batch > 0 && height > 0 && width > 0 && parts_ok(parts) && parts->n_partitions > 0);
  const PartArgs a = make_parts(parts, height, width);
  dim3 grid(ceil_div(a.off[a.n], 256), batch);
  sample_by_range_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, categories, cuboids, cart, a, height, width, out_scores, out_categories, out_cuboids);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

// Smallest logit that can still reach the confidence threshold: sigmoid_T(x) >= thr  =>  x >= logit_lower_bound.
// sigmoid_T is the float32 evaluation rounded to the tensor dtype (relative error < 1e-6 + half an ulp of T), so the
// bound is the exact logit of a threshold lowered by that much and then some; it only has to be conservative -- every
// pixel that passes it gets the exact test.
static float logit_lower_bound(float thr, int32_t dtype) {
  if (!(thr > 0.0f)) return -INFINITY;            // 0 >= thr (or NaN): nothing is filtered here
  const double rel = dtype == RV3D_F32 ? 1.0e-5 : (dtype == RV3D_F16 ? 2.0e-3 : 1.6e-2);
  const double t = static_cast<double>(thr) * (1.0 - rel) - 1.0e-6;
  if (t <= 0.0) return -INFINITY;
  if (t >= 1.0) return INFINITY;                  // thr > 1: no sigmoid reaches it
  const double x = std::log(t / (1.0 - t));
  return static_cast<float>(x - 1.0e-4 * (1.0 + std::fabs(x)));
}

extern "C" int rv3d_decode_compact(const rv3d_decode_params *p, const void *logits, const void *regressands,
                                   const void *cart, const uint8_t *mask, uint64_t *out_keys, float *out_boxes,
                                   int32_t *counter, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(p && logits && regressands && cart && mask && out_keys && out_boxes && counter);
  RV3D_CHECK_ARG(p->batch > 0 && p->n_classes > 0 && p->height > 0 && p->width > 0 && p->capacity >= 0);
  RV3D_CHECK_ARG(parts_ok(&p->parts) && p->total_classes >= p->category_offset + p->n_classes);
  RV3D_CHECK_ARG(static_cast<int64_t>(p->height) * p->width < (int64_t(1) << 29));   // 32-bit element offsets over the 8 regressand planes
  if (!aligned(out_boxes, 16) || !aligned(out_keys, 8)) return RV3D_ERR_ALIGN;
  DecodeArgs a;
  a.B = p->batch; a.C = p->n_classes; a.H = p->height; a.W = p->width; a.az_inv = p->azimuth_invariant;
  a.cat_off = p->category_offset; a.cand_off = p->candidate_offset; a.total_classes = p->total_classes;
  a.capacity = p->capacity; a.thr = p->min_confidence;
  a.x_lo = logit_lower_bound(p->min_confidence, p->dtype);
  a.inv_w = 1.0f / static_cast<float>(p->width);
  RV3D_CHECK_ARG(p->height < (1 << 22));   // row estimate from a float32 quotient is within one of the true row
  RV3D_CHECK_ARG(p->n_classes <= 128);   // queue 1 keeps the class in 7 bits
  a.pa = make_parts(&p->parts, p->height, p->width);
  a.fp = make_filter_parts(a.pa);
  RV3D_CHECK_ARG(p->total_candidates >= p->candidate_offset + a.pa.off[a.pa.n]);
  a.kp.idx_bits = bits_for(p->total_candidates);
  a.kp.score_bits = RV3D_SCORE_BITS_DECODE;   // sigmoid * mask is never negative
  if (bits_for(static_cast<int64_t>(p->batch) * p->total_classes) + a.kp.score_bits + a.kp.idx_bits > 64) return RV3D_ERR_KEYBITS;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto *keys = reinterpret_cast<unsigned long long *>(out_keys);
#define RV3D_CALL(T, TC) return launch_decode_compact<T, TC>(a, logits, regressands, cart, mask, keys, out_boxes, counter, s)
  RV3D_DISPATCH_DTYPES(p->dtype, p->cart_dtype, RV3D_CALL);
#undef RV3D_CALL
  return RV3D_ERR_ARG;
}

extern "C" int rv3d_compact_candidates(const float *cuboids, const float *scores, const int64_t *categories,
                                       int32_t batch, int32_t k, int32_t total_classes, float min_confidence,
                                       int32_t apply_threshold, int32_t capacity, uint64_t *out_keys,
                                       float *out_boxes, int32_t *counter, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(batch > 0 && k >= 0 && total_classes > 0 && capacity >= 0 && out_keys && out_boxes && counter);
  if (k == 0) return RV3D_OK;
  RV3D_CHECK_ARG(cuboids && scores && categories);
  if (!aligned(out_boxes, 16) || !aligned(out_keys, 8)) return RV3D_ERR_ALIGN;
  KeyPack kp;
  kp.idx_bits = bits_for(k);
  kp.score_bits = 32;                         // caller-supplied scores may be negative
  if (bits_for(static_cast<int64_t>(batch) * total_classes) + 32 + kp.idx_bits > 64) return RV3D_ERR_KEYBITS;
  dim3 grid(ceil_div(k, 256), batch);
  compact_candidates_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      cuboids, scores, categories, k, total_classes, min_confidence, apply_threshold, capacity, kp,
      reinterpret_cast<unsigned long long *>(out_keys), out_boxes, counter);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" int rv3d_yaw_to_quat(const float *yaw, float *quat, int64_t n, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && (n == 0 || (yaw && quat)));
  if (n == 0) return RV3D_OK;
  if (!aligned(quat, 16)) return RV3D_ERR_ALIGN;
  yaw_to_quat_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(yaw, quat, n);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}
