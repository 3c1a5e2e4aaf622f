// assign.cu -- training-time callers of the path's operators.
//
// Replaces (paths relative to /root/reference):
//   math/ops/assignment.py:121-139  the per-instance loop of compute_classification_targets:
//       for every panoptic instance: affinities_i.topk(min(k, n)) -> zeros.scatter -> masked_scatter_
//   (the decode calls at :105-114 are decode.cu's dense operator, the affinity at :20-73 is
//    rv3d_iou3d_aligned / rv3d_box_iou_rotated)
//
// Fused form (rv3d_classification_targets): two passes over the pixels, nothing dense in between.
//   pass 1 (targets_affinity_kernel): a FOREGROUND pixel (panoptic id > 0) decodes its predicted and its target box (the
//     decoder of decode_math.cuh, rounded to the tensor dtype exactly like the dense operator), evaluates its affinity
//     (bit-exact rotated BEV IoU, or the Gaussian of the centre distance) and enters it into its instance's top-k list:
//     k 64-bit slots per (sweep, instance), filled by a chain of atomicMax -- slot j keeps the maximum of everything that
//     ever reached it and passes the smaller value on, so whatever the interleaving slot j ends up holding the
//     (j + 1)-th largest key.  Keys are (affinity, lowest pixel first), the order torch.topk's ties are given here.
//   pass 2 (targets_finish_kernel): a pixel is kept iff its key reaches its instance's k-th slot; it then writes all
//     four results of the function (affinities x one-hot labels, foreground / background masks, regression weights).
// The two dense decodes (2 x 7 planes written and read back), the one-hot tensors, the boolean gathers and the global
// sort of the unfused form are gone; the only dense traffic is the function's own inputs and outputs.
//
// The reference walks the instances in Python (one_hot mask, masked_select, topk, two masked_scatter_ per instance,
// a host sync each).  Here every foreground pixel carries its instance's segment id, ONE stable radix sort orders
// (segment asc, affinity desc, pixel asc) and a pixel is in its instance's top-k iff the entry k places before it
// belongs to another segment.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "decode_math.cuh"
#include "iou.cuh"

namespace rv3d {

__global__ void __launch_bounds__(256)
topk_keys_kernel(const float *__restrict__ aff, const int32_t *__restrict__ seg, int64_t n,
                 unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // torch.topk ranks NaN above everything; orderable_f32 already places (positive) NaN above +inf
  const uint32_t desc = ~orderable_f32(__float_as_uint(aff[i]));
  keys[i] = (static_cast<unsigned long long>(static_cast<uint32_t>(seg[i])) << 32) | desc;
  vals[i] = static_cast<uint32_t>(i);
}

__global__ void __launch_bounds__(256)
topk_mark_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n, int k,
                 const float *__restrict__ aff, float *__restrict__ likelihood) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t s = static_cast<uint32_t>(keys[p] >> 32);
  const bool kept = k > 0 && (p < k || static_cast<uint32_t>(keys[p - k] >> 32) != s);
  const uint32_t i = vals[p];
  likelihood[i] = kept ? aff[i] : 0.0f;   // zeros_like(affinities_i).scatter(0, indices, likelihoods) (:131-133)
}


struct TargetsArgs {
  int B, C, HW;          // C = background_index (number of foreground classes)
  int az_inv_targets;    // targets_config.enable_azimuth_invariant_targets (the predictions always decode with it on, :105-109)
  int gaussian;          // 0: BEV IoU (:64-73), 1: exp(-|dc| / sigma^2) (:151-161)
  int k;                 // slots per instance; 0 = topk(0); kKeepAll = every pixel of an instance is in its top-k
  int id_cap;            // instance ids per sweep the slot table holds
  float sigma2;          // float32(sigma ** 2)
};
constexpr int kKeepAll = 0x7fffffff;

__device__ __forceinline__ unsigned long long target_key(float aff, uint32_t pix_in_sweep) {
  // torch.topk ranks NaN above everything; orderable_f32 places (positive) NaN above +inf.  Larger key = better:
  // higher affinity first, then the lower pixel index.
  return (static_cast<unsigned long long>(orderable_f32(__float_as_uint(aff))) << 32) | (0xffffffffu - pix_in_sweep);
}

template <typename T>
__global__ void __launch_bounds__(128)
targets_affinity_kernel(TargetsArgs a, const T *__restrict__ input, const T *__restrict__ target,
                        const T *__restrict__ cart, const long long *__restrict__ panoptics,
                        float *__restrict__ aff_px, unsigned long long *__restrict__ slots, int *__restrict__ status) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.HW) return;
  const size_t gp = static_cast<size_t>(b) * a.HW + p;
  const long long id = panoptics[gp];
  if (id <= 0) return;                                   // one_hot(panoptics)[:, 1:]: id 0 is background (:118)
  if (id >= a.id_cap) {                                  // the caller's capacity promise is broken: flag it, or fail loudly
    if (status) atomicExch(status, 1); else __trap();
    return;
  }
  float ri[8], rt[8], c[3];
#pragma unroll
  for (int k = 0; k < 8; ++k) ri[k] = Ld<T>::one(input + (static_cast<size_t>(b) * 8 + k) * a.HW + p);
#pragma unroll
  for (int k = 0; k < 8; ++k) rt[k] = Ld<T>::one(target + (static_cast<size_t>(b) * 8 + k) * a.HW + p);
#pragma unroll
  for (int k = 0; k < 3; ++k) c[k] = Ld<T>::one(cart + (static_cast<size_t>(b) * 3 + k) * a.HW + p);
  double dp[7], dg[7];
  decode_box(ri, c, true, dp);                            // :105-109
  decode_box(rt, c, a.az_inv_targets != 0, dg);           // :110-114
  float pd[7], gt[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {                           // decode_range_view returns the tensor dtype (coding.py:144)
    pd[k] = static_cast<float>(Ld<T>::cast(dp[k]));
    gt[k] = static_cast<float>(Ld<T>::cast(dg[k]));
  }
  float aff;
  if (a.gaussian) {
    const float dx = pd[0] - gt[0], dy = pd[1] - gt[1], dz = pd[2] - gt[2];
    const float d = sqrtf((dx * dx + dy * dy) + dz * dz);
    aff = expf(-d / a.sigma2);
  } else {
    // XYLWA_INDICES = (0, 1, 3, 4, 6); mmcv box_iou_rotated: angle in radians; .clamp(0, 1) keeps NaN
    const float v = rot_iou(make_hard_rec(pd[0], pd[1], pd[3], pd[4], pd[6], 1.0, true), make_hard_rec(gt[0], gt[1], gt[3], gt[4], gt[6], 1.0, true));
    aff = v != v ? v : fminf(fmaxf(v, 0.0f), 1.0f);
  }
  aff_px[gp] = aff;
  if (a.k <= 0 || a.k == kKeepAll) return;
  unsigned long long key = target_key(aff, static_cast<uint32_t>(p));
  unsigned long long *sl = slots + (static_cast<size_t>(b) * a.id_cap + static_cast<size_t>(id)) * a.k;
  // the k-th slot only ever grows: a key that does not beat it now cannot be among the final k
  if (key <= *reinterpret_cast<volatile unsigned long long *>(sl + a.k - 1)) return;
  for (int j = 0; j < a.k && key != 0ull; ++j) {
    const unsigned long long old = atomicMax(sl + j, key);
    if (old < key) key = old;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
targets_finish_kernel(TargetsArgs a, const long long *__restrict__ panoptics, const long long *__restrict__ labels,
                      const uint8_t *__restrict__ mask, const float *__restrict__ aff_px,
                      const unsigned long long *__restrict__ slots, float *__restrict__ affinities, T *__restrict__ foreground,
                      uint8_t *__restrict__ background, uint8_t *__restrict__ reg_weights) {
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.HW) return;
  const size_t gp = static_cast<size_t>(b) * a.HW + p;
  const long long id = panoptics[gp];
  const long long label = labels[gp];
  float like = 0.0f;
  if (id > 0 && id < a.id_cap && a.k > 0) {
    const float aff = aff_px[gp];
    bool kept = true;
    if (a.k != kKeepAll)
      kept = target_key(aff, static_cast<uint32_t>(p)) >= slots[(static_cast<size_t>(b) * a.id_cap + static_cast<size_t>(id)) * a.k + a.k - 1];
    like = kept ? aff : 0.0f;                              // zeros_like(affinities_i).scatter(0, indices, likelihoods) (:131-133)
  }
  const bool fg = like != 0.0f;                            // likelihoods.bool(): NaN counts as foreground (:137-139)
  for (int c = 0; c < a.C; ++c)                            // affinities * all_foreground (:142): a literal product (NaN * 0 = NaN)
    affinities[(static_cast<size_t>(b) * a.C + c) * a.HW + p] = like * (label == c ? 1.0f : 0.0f);
  foreground[gp] = Ld<T>::cast(fg ? 1.0 : 0.0);
  background[gp] = (!fg && mask[gp] != 0) ? 1 : 0;         // :141
  reg_weights[gp] = (label >= 0 && label < a.C) ? 1 : 0;   // all_foreground.any(dim=1) (:143)
}

struct TopkLayout {
  unsigned long long *keys, *keys_alt;
  uint32_t *vals, *vals_alt;
  unsigned char *cub_tmp;
  size_t cub_bytes, total;
};

static TopkLayout topk_layout(void *base, int64_t n) {
  TopkLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { void *p = base ? static_cast<unsigned char *>(base) + off : nullptr; off += align_up(bytes, 256); return p; };
  const size_t nn = static_cast<size_t>(n > 0 ? n : 1);
  L.keys = static_cast<unsigned long long *>(take(nn * 8));
  L.keys_alt = static_cast<unsigned long long *>(take(nn * 8));
  L.vals = static_cast<uint32_t *>(take(nn * 4));
  L.vals_alt = static_cast<uint32_t *>(take(nn * 4));
  cub::DoubleBuffer<unsigned long long> kb(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> vb(nullptr, nullptr);
  L.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, L.cub_bytes, kb, vb, static_cast<int>(nn), 0, 64);
  L.cub_tmp = static_cast<unsigned char *>(take(L.cub_bytes));
  L.total = off;
  return L;
}

}  // namespace rv3d

using namespace rv3d;

extern "C" size_t rv3d_instance_topk_scratch_bytes(int64_t n) { return topk_layout(nullptr, n).total; }

extern "C" int rv3d_instance_topk(const float *affinity, const int32_t *segment, int64_t n, int32_t n_segments,
                                  int32_t k, float *likelihood, void *scratch, size_t scratch_bytes,
                                  rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && n < (int64_t(1) << 31) && n_segments > 0 && k >= 0);
  if (n == 0) return RV3D_OK;
  RV3D_CHECK_ARG(affinity && segment && likelihood && scratch);
  if (!aligned(scratch, 256)) return RV3D_ERR_ALIGN;
  const TopkLayout L = topk_layout(scratch, n);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  topk_keys_kernel<<<ceil_div(n, 256), 256, 0, s>>>(affinity, segment, n, L.keys, L.vals);
  RV3D_CHECK_LAUNCH();
  cub::DoubleBuffer<unsigned long long> kb(L.keys, L.keys_alt);
  cub::DoubleBuffer<uint32_t> vb(L.vals, L.vals_alt);
  size_t cub_bytes = L.cub_bytes;
  // stable: equal (segment, affinity) keep pixel order
  RV3D_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, cub_bytes, kb, vb, static_cast<int>(n), 0,
                                                  32 + bits_for(n_segments), s));
  topk_mark_kernel<<<ceil_div(n, 256), 256, 0, s>>>(kb.Current(), vb.Current(), n, k, affinity, likelihood);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}

extern "C" size_t rv3d_classification_targets_scratch_bytes(int32_t batch, int32_t height, int32_t width, int32_t k, int32_t id_capacity) {
  const size_t px = static_cast<size_t>(batch) * height * width;
  const size_t kk = (k > 0 && k != kKeepAll) ? static_cast<size_t>(k) : 0;
  return align_up(px * 4, 256) + align_up(static_cast<size_t>(batch) * id_capacity * kk * 8 + 8, 256) + 256;
}

extern "C" int rv3d_classification_targets(const float *input, const float *target, const int64_t *labels, const float *cart,
                                           const uint8_t *mask, const int64_t *panoptics, int32_t batch, int32_t n_classes,
                                           int32_t height, int32_t width, int32_t affinity_fn, int32_t az_inv_targets,
                                           int32_t k, float sigma2, int32_t id_capacity, float *affinities,
                                           float *foreground, uint8_t *background, uint8_t *reg_weights, int32_t *status,
                                           void *scratch, size_t scratch_bytes, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(batch >= 0 && n_classes >= 0 && height > 0 && width > 0 && id_capacity > 0 && k >= 0);
  RV3D_CHECK_ARG(affinity_fn == 0 || affinity_fn == 1);
  RV3D_CHECK_ARG(k == kKeepAll || k <= 64);
  RV3D_CHECK_ARG(static_cast<int64_t>(height) * width < (int64_t(1) << 31));
  if (batch == 0) return RV3D_OK;
  RV3D_CHECK_ARG(input && target && labels && cart && mask && panoptics && affinities && foreground && background && reg_weights && scratch);
  if (!aligned(scratch, 256)) return RV3D_ERR_ALIGN;
  if (scratch_bytes < rv3d_classification_targets_scratch_bytes(batch, height, width, k, id_capacity)) return RV3D_ERR_SCRATCH;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  TargetsArgs a;
  a.B = batch; a.C = n_classes; a.HW = height * width; a.az_inv_targets = az_inv_targets; a.gaussian = affinity_fn;
  a.k = k; a.id_cap = id_capacity; a.sigma2 = sigma2;
  const size_t px = static_cast<size_t>(batch) * a.HW;
  float *aff_px = static_cast<float *>(scratch);
  unsigned long long *slots = reinterpret_cast<unsigned long long *>(static_cast<unsigned char *>(scratch) + align_up(px * 4, 256));
  const size_t kk = (k > 0 && k != kKeepAll) ? static_cast<size_t>(k) : 0;
  if (kk) RV3D_CHECK_CUDA(cudaMemsetAsync(slots, 0, static_cast<size_t>(batch) * id_capacity * kk * 8, s));
  targets_affinity_kernel<float><<<dim3(ceil_div(a.HW, 128), batch), 128, 0, s>>>(a, input, target, cart,
      reinterpret_cast<const long long *>(panoptics), aff_px, slots, status);
  RV3D_CHECK_LAUNCH();
  targets_finish_kernel<float><<<dim3(ceil_div(a.HW, 256), batch), 256, 0, s>>>(a, reinterpret_cast<const long long *>(panoptics),
      reinterpret_cast<const long long *>(labels), mask, aff_px, slots, affinities, foreground, background, reg_weights);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}
