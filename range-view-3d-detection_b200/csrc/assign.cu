// assign.cu -- training-time callers of the path's operators.
//
// Replaces (paths relative to /root/reference):
//   math/ops/assignment.py:121-139  the per-instance loop of compute_classification_targets:
//       for every panoptic instance: affinities_i.topk(min(k, n)) -> zeros.scatter -> masked_scatter_
//   (the decode calls at :105-114 are decode.cu's dense operator, the affinity at :20-73 is
//    rv3d_iou3d_aligned / rv3d_box_iou_rotated)
//
// The reference walks the instances in Python (one_hot mask, masked_select, topk, two masked_scatter_ per instance,
// a host sync each).  Here every foreground pixel carries its instance's segment id, ONE stable radix sort orders
// (segment asc, affinity desc, pixel asc) and a pixel is in its instance's top-k iff the entry k places before it
// belongs to another segment.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

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
