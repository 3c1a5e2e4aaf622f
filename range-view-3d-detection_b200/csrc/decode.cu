// decode.cu -- dense per-pixel box decoding (K2) and its free-standing operator forms.
//
// Replaces (paths relative to /root/reference):
//   math/ops/coding.py:110-144          decode_range_view  (+ :79-107 egovehicle_from_azimuth)
//   nn/decoders/range_decoder.py:49-77  sigmoid*mask, max over classes, decode, sample_by_range
//   nn/decoders/range_decoder.py:127-156 sample_by_range
//   math/ops/nms.py:212-215             min-confidence gather
//   math/linalg/lie/SO3.py:122-134      yaw_to_quat
//
// K2 (decode_compact) is one pass over the head outputs:
//   phase A  every thread owns 4 consecutive pixels: float4 loads of the C logit planes + mask,
//            running max (first index wins ties, NaN kills the pixel like torch.max would),
//            correctly-rounded float32 sigmoid of the winner, threshold; for pixels that pass,
//            the range-partition / column-stride test of sample_by_range on ||cart||.
//   scan     block-wide exclusive scan of (live pixels, emitted candidates); ONE global atomicAdd
//            per 1024 pixels reserves the output rows.
//   phase B  live pixels are compacted into a shared-memory queue so the fp64 decode
//            (3 exp, 2 atan2, sincos) runs with dense lanes instead of ~20 %-full warps;
//            each live pixel writes 1..n_partitions (key, box) rows.
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
  double yaw = fast_atan2(static_cast<double>(reg[6]), static_cast<double>(reg[7]));  // :136 (fastmath.cuh: <= 1 ulp)
  const double cx = cart[0], cy = cart[1], cz = cart[2];
  if (az_inv) {                                                                    // :79-107
    const double phi = fast_atan2(cy, cx);
    // cos / sin of atan2(cy, cx) are cx / h and cy / h (one reciprocal instead of a sincos); degenerate or
    // non-finite rays take the library path
    const double h2 = cx * cx + cy * cy;
    double s, c;
    if (h2 > 1e-60 && h2 < 1e60) {
      const double rh = 1.0 / sqrt(h2);
      c = cx * rh; s = cy * rh;
    } else {
      sincos(phi, &s, &c);
    }
    const double x = c * ox - s * oy;
    const double y = s * ox + c * oy;
    ox = x; oy = y;
    yaw += phi;
  }
  out[0] = cx + ox; out[1] = cy + oy; out[2] = cz + oz;                            // :142
  out[3] = exp(static_cast<double>(reg[3]));                                       // :132
  out[4] = exp(static_cast<double>(reg[4]));
  out[5] = exp(static_cast<double>(reg[5]));
  out[6] = yaw;
}

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
  KeyPack kp;
  PartArgs pa;
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
  b += static_cast<size_t>(tile) * (2 + 2 + 1 + 2 + 2 + 4 + 4);   // q_pix, q_meta, q_emit, q_h, q_w, q_score, q_off
  return align_up(b, 16) + 64;
}

// One CTA per tile of kTile consecutive pixels of one sweep, kTile / 4 threads.
//   stage   the logit, cart and mask planes' slices of the tile are brought into shared memory: with kBulk,
//           C + 4 TMA 1-D bulk copies issued by one thread, completion on an mbarrier (no registers, no
//           per-thread loads; several resident CTAs overlap each other's copies and math); without
//           (unaligned shapes) plain cooperative loads into the same layout.  The 8 regressand planes are
//           NOT staged: only live pixels need them, phase B gathers them (8 independent loads per pixel),
//           which halves the tile's shared memory (more resident CTAs) and, at realistic candidate
//           densities, skips most of their bytes.
//   phase A / scan / phase B   as described at the top of this file, reading shared memory only.
template <typename T, typename TC, int kTile, bool kBulk>
__global__ void __launch_bounds__(kTile / kPxPerThread)
decode_compact_kernel(DecodeArgs a, const T *__restrict__ logits, const T *__restrict__ reg,
                      const TC *__restrict__ cart, const uint8_t *__restrict__ mask,
                      unsigned long long *__restrict__ out_keys, float *__restrict__ out_boxes,
                      int32_t *__restrict__ counter) {
  constexpr int kThreads = kTile / kPxPerThread;
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(128) unsigned char dsm[];
  __shared__ uint32_t s_scan[kWarps];
  __shared__ uint32_t s_base, s_total_live;
  __shared__ __align__(8) uint64_t s_bar;
  // partition constants out of the kernel-parameter bank (dynamic indexing there costs uniform moves)
  __shared__ float s_lower[RV3D_MAX_PARTITIONS], s_upper[RV3D_MAX_PARTITIONS];
  __shared__ int s_rate[RV3D_MAX_PARTITIONS], s_shift[RV3D_MAX_PARTITIONS], s_off[RV3D_MAX_PARTITIONS], s_wsub[RV3D_MAX_PARTITIONS];
  __shared__ uint32_t s_magic[RV3D_MAX_PARTITIONS];

  const int b = blockIdx.y;
  const int HW = a.H * a.W;
  const int tid = threadIdx.x;
  const int blk0 = blockIdx.x * kTile;
  const int npx = min(kTile, HW - blk0);

  // ---- carve shared memory: [C][kTile] logits | [3][kTile] cart | mask | queues
  T *s_logits = reinterpret_cast<T *>(dsm);
  TC *s_cart = reinterpret_cast<TC *>(s_logits + static_cast<size_t>(a.C) * kTile);   // C * kTile * sizeof(T) is a multiple of 16
  uint8_t *s_mask = reinterpret_cast<uint8_t *>(s_cart + 3 * kTile);
  unsigned char *qp = dsm + align_up_c(static_cast<size_t>(a.C) * kTile * sizeof(T) + 3 * kTile * sizeof(TC) + kTile, 16);
  float *q_score = reinterpret_cast<float *>(qp);                    // score of live pixel t
  uint32_t *q_off = reinterpret_cast<uint32_t *>(q_score + kTile);   // exclusive emit offset inside the block
  uint16_t *q_pix = reinterpret_cast<uint16_t *>(q_off + kTile);     // local pixel id of live pixel t
  uint16_t *q_meta = q_pix + kTile;                                  // class index within the task
  uint8_t *q_emit = reinterpret_cast<uint8_t *>(q_meta + kTile);     // partition bit-mask
  uint16_t *q_h = reinterpret_cast<uint16_t *>(q_emit + kTile);      // image row / column of live pixel t
  uint16_t *q_w = q_h + kTile;

  if (tid < RV3D_MAX_PARTITIONS) {
    s_lower[tid] = a.pa.lower[tid]; s_upper[tid] = a.pa.upper[tid]; s_rate[tid] = a.pa.rate[tid];
    s_shift[tid] = a.pa.shift[tid]; s_magic[tid] = a.pa.magic[tid]; s_off[tid] = a.pa.off[tid]; s_wsub[tid] = a.pa.wsub[tid];
  }
  const int n_parts = a.pa.n;

  const T *rg = reg + static_cast<size_t>(b) * 8 * HW + blk0;   // regressands: gathered in phase B for the live pixels only

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
        for (int c = 0; c < a.C; ++c) bulk_g2s(s_logits + static_cast<size_t>(c) * kTile, lg + static_cast<size_t>(c) * HW, plane, &s_bar);
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

  // ---------------- phase A: class max over the staged logits ----------------
  const int lp0 = tid * kPxPerThread;          // first local pixel of this thread
  const int row0 = (blk0 + lp0) / a.W;         // one division per thread; (row, col) advance incrementally
  const int col0 = (blk0 + lp0) - row0 * a.W;
  float best[kPxPerThread], runner[kPxPerThread];   // runner = max logit among the classes before `cls`
  int cls[kPxPerThread];
  bool bad[kPxPerThread];
#pragma unroll
  for (int j = 0; j < kPxPerThread; ++j) { best[j] = -CUDART_INF_F; runner[j] = -CUDART_INF_F; cls[j] = 0; bad[j] = false; }
#pragma unroll 4
  for (int c = 0; c < a.C; ++c) {
    float v[4];
    Sm<T>::four(s_logits + static_cast<size_t>(c) * kTile + lp0, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bad[j] |= (v[j] != v[j]);
      if (v[j] > best[j]) { runner[j] = best[j]; best[j] = v[j]; cls[j] = c; }
    }
  }
  const uchar4 m4 = *reinterpret_cast<const uchar4 *>(s_mask + lp0);
  const uint32_t mbits = (m4.x ? 1u : 0u) | (m4.y ? 2u : 0u) | (m4.z ? 4u : 0u) | (m4.w ? 8u : 0u);

  // score + threshold + partition test
  float score[kPxPerThread];
  uint32_t emit[kPxPerThread];
  uint32_t n_live = 0, n_emit = 0;
  const bool zero_passes = 0.0f >= a.thr;
  float cx[4], cy[4], cz[4];
  Sm<TC>::four(s_cart + lp0, cx);
  Sm<TC>::four(s_cart + kTile + lp0, cy);
  Sm<TC>::four(s_cart + 2 * kTile + lp0, cz);
#pragma unroll
  for (int j = 0; j < kPxPerThread; ++j) {
    emit[j] = 0;
    score[j] = 0.f;
    if (lp0 + j >= npx || bad[j]) continue;
    if (mbits & (1u << j)) {
      score[j] = sigmoid_t<T>(best[j]);
      // torch.max returns the FIRST index attaining the max of the float32 scores: an earlier
      // class with a smaller logit can round to the same float32 sigmoid (always when saturated)
      if (cls[j] > 0 && score[j] >= a.thr && (best[j] >= 15.f || runner[j] > best[j] - 1.0f)) {
        for (int c = 0; c < cls[j]; ++c) {
          const float x = Sm<T>::one(s_logits + static_cast<size_t>(c) * kTile + lp0 + j);
          if ((best[j] >= 15.f || x > best[j] - 1.0f) && sigmoid_t<T>(x) == score[j]) { cls[j] = c; break; }
        }
      }
    } else {
      cls[j] = 0;  // sigmoid * 0 == 0 for every class -> argmax 0
    }
    if (!(score[j] >= a.thr || zero_passes)) continue;
    if (n_parts == 0) {
      if (score[j] >= a.thr) emit[j] = 1u;
    } else {
      // bit i of part_in: the pixel's range lies in partition i; the column-stride test comes below
      // cart.norm(dim=1) accumulates in float32 and rounds to cart's dtype; the bounds are a float32 tensor
      const float d = Ld<TC>::round_f32(norm3(cx[j], cy[j], cz[j]));
      uint32_t part_in = 0u;
      for (int i = 0; i < n_parts; ++i) part_in |= ((d > s_lower[i]) && (d <= s_upper[i])) ? (1u << i) : 0u;
      emit[j] = score[j] >= a.thr ? part_in : 0u;
      if (zero_passes) emit[j] = (1u << n_parts) - 1u;     // 0 >= thr: out-of-partition copies survive with score 0
    }
  }
  if (n_parts > 0) {
    // column stride: partition i keeps the columns w with w % rate_i == 0
    int wj[kPxPerThread];
#pragma unroll
    for (int j = 0; j < kPxPerThread; ++j) { wj[j] = col0 + j; if (wj[j] >= a.W) wj[j] -= a.W; }   // W >= 4: one wrap at most
    for (int i = 0; i < n_parts; ++i) {
      const int sh = s_shift[i], rate = s_rate[i];
      const uint32_t mg = s_magic[i];
#pragma unroll
      for (int j = 0; j < kPxPerThread; ++j)
        if (wj[j] - fast_div(wj[j], sh, mg) * rate) emit[j] &= ~(1u << i);
    }
  }
#pragma unroll
  for (int j = 0; j < kPxPerThread; ++j) {
    n_live += emit[j] ? 1u : 0u;
    n_emit += __popc(emit[j]);
  }

  // ---------------- block scan of (live, emit) packed as hi16 | lo16 ----------------
  // per block: live <= 512, emit <= 512 * 8 -> both fit 16 bits, no carry between the halves
  const uint32_t mine = (n_live << 16) | n_emit;
  uint32_t incl = mine;
  const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_scan[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const uint32_t v = lane < kWarps ? s_scan[lane] : 0u;
    uint32_t inc2 = v;
#pragma unroll
    for (int o = 1; o < kWarps; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc2, o);
      if (lane >= o) inc2 += t;
    }
    if (lane < kWarps) s_scan[lane] = inc2 - v;  // exclusive warp offsets
    if (lane == kWarps - 1) {
      const uint32_t total_emit = inc2 & 0xffffu;
      s_total_live = inc2 >> 16;
      // ONE global atomic per tile reserves the block's output rows
      s_base = total_emit ? static_cast<uint32_t>(atomicAdd(counter, static_cast<int>(total_emit))) : 0u;
    }
  }
  __syncthreads();
  const uint32_t excl = s_scan[wid] + (incl - mine);
  uint32_t live_pos = excl >> 16, emit_pos = excl & 0xffffu;
#pragma unroll
  for (int j = 0; j < kPxPerThread; ++j) {
    if (!emit[j]) continue;
    q_pix[live_pos] = static_cast<uint16_t>(lp0 + j);
    q_meta[live_pos] = static_cast<uint16_t>(cls[j]);
    {
      const bool wrap = col0 + j >= a.W;
      q_h[live_pos] = static_cast<uint16_t>(row0 + (wrap ? 1 : 0));
      q_w[live_pos] = static_cast<uint16_t>(col0 + j - (wrap ? a.W : 0));
    }
    q_emit[live_pos] = static_cast<uint8_t>(emit[j]);
    q_score[live_pos] = score[j];
    q_off[live_pos] = emit_pos;
    ++live_pos;
    emit_pos += __popc(emit[j]);
  }
  __syncthreads();
  const uint32_t total_live = s_total_live;
  const uint32_t base = s_base;

  // ---------------- phase B: dense-lane fp64 decode of the live pixels ----------------
  for (uint32_t t = tid; t < total_live; t += kThreads) {
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
    const uint32_t seg = static_cast<uint32_t>(b) * a.total_classes + q_meta[t] + a.cat_off;
    const int h = q_h[t], w = q_w[t];   // (row, col) computed incrementally in phase A
    uint32_t e = q_emit[t];
    uint32_t row = base + q_off[t];
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
        // the partition that does not contain the pixel contributes score 0 (only when 0 >= thr)
        const float d = Ld<TC>::round_f32(norm3(c[0], c[1], c[2]));
        if (!((d > s_lower[i]) && (d <= s_upper[i]))) s_out = 0.0f;
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

extern "C" int rv3d_decode_compact(const rv3d_decode_params *p, const void *logits, const void *regressands,
                                   const void *cart, const uint8_t *mask, uint64_t *out_keys, float *out_boxes,
                                   int32_t *counter, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(p && logits && regressands && cart && mask && out_keys && out_boxes && counter);
  RV3D_CHECK_ARG(p->batch > 0 && p->n_classes > 0 && p->height > 0 && p->width > 0 && p->capacity >= 0);
  RV3D_CHECK_ARG(parts_ok(&p->parts) && p->total_classes >= p->category_offset + p->n_classes);
  RV3D_CHECK_ARG(static_cast<int64_t>(p->height) * p->width < (int64_t(1) << 30));
  if (!aligned(out_boxes, 16) || !aligned(out_keys, 8)) return RV3D_ERR_ALIGN;
  DecodeArgs a;
  a.B = p->batch; a.C = p->n_classes; a.H = p->height; a.W = p->width; a.az_inv = p->azimuth_invariant;
  a.cat_off = p->category_offset; a.cand_off = p->candidate_offset; a.total_classes = p->total_classes;
  a.capacity = p->capacity; a.thr = p->min_confidence;
  a.pa = make_parts(&p->parts, p->height, p->width);
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
