// records.cu -- detection wire format.
//
// Replaces (paths relative to /root/reference):
//   math/ops/coding.py:31-58     build_dataframe's thirteen per-column `.tolist()` device->host reads
//   nn/arch/detector.py:45-60    SERIALIZED_SCHEMA (the row the evaluation pipeline serialises)
//   nn/arch/detector.py:573-584  prepare_for_evaluation's range filter: ||(tx,ty,tz)||_2 <= max_range_m
//
// One fixed 64-byte record per detection, built and (optionally) range-filtered on the device with the order
// of the decoder's output preserved (sweep asc, class asc, score desc), so the host needs ONE copy.
#include <cub/device/device_select.cuh>

#include "common.cuh"

namespace rv3d {

static_assert(sizeof(rv3d_detection_record) == 64, "wire record is 64 bytes");

__global__ void __launch_bounds__(256)
records_kernel(const float *__restrict__ params, const float *__restrict__ scores, const float *__restrict__ cats,
               const float *__restrict__ batch, long long n, const long long *__restrict__ stamp, int B, float max_range,
               int filter, rv3d_detection_record *__restrict__ rec, uint8_t *__restrict__ flag) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rv3d_detection_record r;
#pragma unroll
  for (int k = 0; k < 10; ++k) r.params[k] = params[i * 10 + k];
  r.score = scores[i];
  r.category_index = static_cast<int32_t>(cats[i]);      // categories.int() (coding.py:54)
  const int b = static_cast<int32_t>(batch[i]);
  r.batch_index = b;
  r.timestamp_ns = (stamp && b >= 0 && b < B) ? stamp[b] : 0;
  // np.linalg.norm of the float32 columns: float32 sqrt((x*x + y*y) + z*z), compared with the float64 limit
  const float nx = r.params[0], ny = r.params[1], nz = r.params[2];
  r.range_m = sqrtf((nx * nx + ny * ny) + nz * nz);
  rec[i] = r;
  flag[i] = (!filter || static_cast<double>(r.range_m) <= static_cast<double>(max_range)) ? 1 : 0;
}

struct RecLayout {
  rv3d_detection_record *tmp;
  uint8_t *flag;
  unsigned char *cub_tmp;
  size_t cub_bytes, total;
};

static RecLayout rec_layout(void *base, long long n) {
  RecLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { void *p = base ? static_cast<unsigned char *>(base) + off : nullptr; off += align_up(bytes, 256); return p; };
  const size_t nn = static_cast<size_t>(n > 0 ? n : 1);
  L.tmp = static_cast<rv3d_detection_record *>(take(nn * sizeof(rv3d_detection_record)));
  L.flag = static_cast<uint8_t *>(take(nn));
  L.cub_bytes = 0;
  cub::DeviceSelect::Flagged(nullptr, L.cub_bytes, static_cast<rv3d_detection_record *>(nullptr), static_cast<uint8_t *>(nullptr),
                             static_cast<rv3d_detection_record *>(nullptr), static_cast<int32_t *>(nullptr), static_cast<int>(nn));
  L.cub_tmp = static_cast<unsigned char *>(take(L.cub_bytes));
  L.total = off;
  return L;
}

}  // namespace rv3d

using namespace rv3d;

extern "C" size_t rv3d_detection_records_scratch_bytes(int64_t n) { return rec_layout(nullptr, n).total; }

extern "C" int rv3d_detection_records(const float *params, const float *scores, const float *categories,
                                      const float *batch_index, int64_t n, const int64_t *sweep_timestamp_ns,
                                      int32_t batch, float max_range_m, int32_t apply_range_filter,
                                      rv3d_detection_record *out, int32_t *out_count, void *scratch,
                                      size_t scratch_bytes, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(n >= 0 && n < (int64_t(1) << 31) && batch >= 0 && out_count);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    RV3D_CHECK_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t), s));
    return RV3D_OK;
  }
  RV3D_CHECK_ARG(params && scores && categories && batch_index && out && scratch);
  if (!aligned(scratch, 256) || !aligned(out, 8)) return RV3D_ERR_ALIGN;
  const RecLayout L = rec_layout(scratch, n);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  records_kernel<<<ceil_div(n, 256), 256, 0, s>>>(params, scores, categories, batch_index, n,
                                                  reinterpret_cast<const long long *>(sweep_timestamp_ns), batch, max_range_m,
                                                  apply_range_filter, L.tmp, L.flag);
  RV3D_CHECK_LAUNCH();
  size_t cub_bytes = L.cub_bytes;
  RV3D_CHECK_CUDA(cub::DeviceSelect::Flagged(L.cub_tmp, cub_bytes, L.tmp, L.flag, out, out_count, static_cast<int>(n), s));
  return RV3D_OK;
}
