// records.cu -- detection wire format.
//
// Replaces (paths relative to /root/reference):
//   math/ops/coding.py:31-58     build_dataframe's thirteen per-column `.tolist()` device->host reads
//   nn/arch/detector.py:45-60    SERIALIZED_SCHEMA (the row the evaluation pipeline serialises)
//   nn/arch/detector.py:573-584  prepare_for_evaluation: range filter ||(tx,ty,tz)||_2 <= max_range_m, sort by score
//                                (descending), unique rows
//   nn/arch/detector.py:366-380  validation_step's group_by (log_id, timestamp_ns): per-sweep slices of the record stream
//
// One fixed 64-byte record per detection, built and (optionally) range-filtered on the device with the order
// of the decoder's output preserved (sweep asc, class asc, score desc), so the host needs ONE copy.
#include <cub/device/device_radix_sort.cuh>
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

// ---- prepare_for_evaluation's `.sort(score, descending=True).unique()` on the record stream -------------------------
// key = [ ~orderable(score) | 32-bit hash of the 64 record bytes ]: one radix sort brings equal rows next to each
// other inside their score group; a row is a duplicate when an earlier row of its key run has the same 64 bytes
// (the backward scan makes hash collisions harmless).  Rows beyond the device-side count get the all-ones key.
// polars' unique() leaves the order of the survivors unspecified; here it is (score desc, hash asc, input order).
__device__ __forceinline__ uint32_t hash_record(const rv3d_detection_record &r) {
  const uint32_t *w = reinterpret_cast<const uint32_t *>(&r);
  uint32_t h = 0x811C9DC5u;
#pragma unroll
  for (int k = 0; k < 16; ++k) { h ^= w[k]; h *= 0x01000193u; h ^= h >> 15; }
  return h;
}
__device__ __forceinline__ bool same_record(const rv3d_detection_record &a, const rv3d_detection_record &b) {
  const uint4 *x = reinterpret_cast<const uint4 *>(&a), *y = reinterpret_cast<const uint4 *>(&b);
  bool eq = true;
#pragma unroll
  for (int k = 0; k < 4; ++k) eq &= x[k].x == y[k].x && x[k].y == y[k].y && x[k].z == y[k].z && x[k].w == y[k].w;
  return eq;
}

__global__ void __launch_bounds__(256)
record_keys_kernel(const rv3d_detection_record *__restrict__ rec, const int32_t *__restrict__ count, long long cap,
                   unsigned long long *__restrict__ keys, uint32_t *__restrict__ idx) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  idx[i] = static_cast<uint32_t>(i);
  if (i >= *count) { keys[i] = ~0ull; return; }
  const rv3d_detection_record r = rec[i];
  const uint32_t desc = ~orderable_f32(__float_as_uint(r.score));
  keys[i] = (static_cast<unsigned long long>(desc) << 32) | hash_record(r);
}

__global__ void __launch_bounds__(256)
record_unique_kernel(const rv3d_detection_record *__restrict__ rec, const int32_t *__restrict__ count, long long cap,
                     const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ idx,
                     rv3d_detection_record *__restrict__ sorted, uint8_t *__restrict__ flag) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  if (i >= *count) { flag[i] = 0; return; }          // the live rows sort in front of the padding
  const rv3d_detection_record r = rec[idx[i]];
  const unsigned long long k = keys[i];
  bool dup = false;
  for (long long j = i - 1; j >= 0 && keys[j] == k && !dup; --j) dup = same_record(r, rec[idx[j]]);
  sorted[i] = r;
  flag[i] = dup ? 0 : 1;
}

// first record of every sweep in a stream ordered by batch_index (the decoder's order): offsets[b] = #records with
// batch_index < b, b = 0 .. B  (one binary search per thread)
__global__ void record_group_offsets_kernel(const rv3d_detection_record *__restrict__ rec, const int32_t *__restrict__ count,
                                            int B, int32_t *__restrict__ offsets) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  int lo = 0, hi = *count;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rec[mid].batch_index < b) lo = mid + 1; else hi = mid;
  }
  offsets[b] = lo;
}

struct UniqLayout {
  unsigned long long *keys, *keys_alt;
  uint32_t *idx, *idx_alt;
  rv3d_detection_record *sorted;
  uint8_t *flag;
  unsigned char *cub_tmp;
  size_t cub_bytes, total;
};

static UniqLayout uniq_layout(void *base, long long n) {
  UniqLayout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { void *p = base ? static_cast<unsigned char *>(base) + off : nullptr; off += align_up(bytes, 256); return p; };
  const size_t nn = static_cast<size_t>(n > 0 ? n : 1);
  L.keys = static_cast<unsigned long long *>(take(nn * 8)); L.keys_alt = static_cast<unsigned long long *>(take(nn * 8));
  L.idx = static_cast<uint32_t *>(take(nn * 4)); L.idx_alt = static_cast<uint32_t *>(take(nn * 4));
  L.sorted = static_cast<rv3d_detection_record *>(take(nn * sizeof(rv3d_detection_record)));
  L.flag = static_cast<uint8_t *>(take(nn));
  size_t a = 0, b = 0;
  cub::DoubleBuffer<unsigned long long> kb(nullptr, nullptr);
  cub::DoubleBuffer<uint32_t> vb(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, a, kb, vb, static_cast<int>(nn), 0, 64);
  cub::DeviceSelect::Flagged(nullptr, b, static_cast<rv3d_detection_record *>(nullptr), static_cast<uint8_t *>(nullptr),
                             static_cast<rv3d_detection_record *>(nullptr), static_cast<int32_t *>(nullptr), static_cast<int>(nn));
  L.cub_bytes = a > b ? a : b;
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

extern "C" size_t rv3d_records_sort_unique_scratch_bytes(int64_t capacity) { return uniq_layout(nullptr, capacity).total; }

extern "C" int rv3d_records_sort_unique(const rv3d_detection_record *records, const int32_t *count, int64_t capacity,
                                        rv3d_detection_record *out, int32_t *out_count, void *scratch, size_t scratch_bytes,
                                        rv3d_stream_t stream) {
  RV3D_CHECK_ARG(capacity >= 0 && capacity < (int64_t(1) << 31) && count && out_count);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (capacity == 0) {
    RV3D_CHECK_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int32_t), s));
    return RV3D_OK;
  }
  RV3D_CHECK_ARG(records && out && scratch);
  if (!aligned(scratch, 256) || !aligned(out, 16) || !aligned(records, 16)) return RV3D_ERR_ALIGN;
  const UniqLayout L = uniq_layout(scratch, capacity);
  if (scratch_bytes < L.total) return RV3D_ERR_SCRATCH;
  const int g = ceil_div(capacity, 256);
  record_keys_kernel<<<g, 256, 0, s>>>(records, count, capacity, L.keys, L.idx);
  RV3D_CHECK_LAUNCH();
  cub::DoubleBuffer<unsigned long long> kb(L.keys, L.keys_alt);
  cub::DoubleBuffer<uint32_t> vb(L.idx, L.idx_alt);
  size_t cub_bytes = L.cub_bytes;
  RV3D_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(L.cub_tmp, cub_bytes, kb, vb, static_cast<int>(capacity), 0, 64, s));
  record_unique_kernel<<<g, 256, 0, s>>>(records, count, capacity, kb.Current(), vb.Current(), L.sorted, L.flag);
  RV3D_CHECK_LAUNCH();
  cub_bytes = L.cub_bytes;
  RV3D_CHECK_CUDA(cub::DeviceSelect::Flagged(L.cub_tmp, cub_bytes, L.sorted, L.flag, out, out_count, static_cast<int>(capacity), s));
  return RV3D_OK;
}

extern "C" int rv3d_records_group_offsets(const rv3d_detection_record *records, const int32_t *count, int32_t batch,
                                          int32_t *offsets, rv3d_stream_t stream) {
  RV3D_CHECK_ARG(records && count && offsets && batch >= 0);
  record_group_offsets_kernel<<<ceil_div(batch + 1, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(records, count, batch, offsets);
  RV3D_CHECK_LAUNCH();
  return RV3D_OK;
}
