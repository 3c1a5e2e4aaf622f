// common.cuh -- shared helpers for librv3d (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rv3d.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "librv3d is written for sm_100a (B200) only"
#endif

#define RV3D_CHECK_ARG(cond) \
  do {                       \
    if (!(cond)) return RV3D_ERR_ARG; \
  } while (0)

#define RV3D_CHECK_CUDA(expr)                      \
  do {                                             \
    cudaError_t e__ = (expr);                      \
    if (e__ != cudaSuccess) return RV3D_ERR_CUDA;  \
  } while (0)

#define RV3D_CHECK_LAUNCH() RV3D_CHECK_CUDA(cudaGetLastError())

namespace rv3d {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74

static inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ constexpr size_t align_up_c(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ceil_div(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }
static inline int bits_for(int64_t n) {  // bits needed to hold values 0..n-1
  int b = 0;
  while ((int64_t(1) << b) < n) ++b;
  return b;
}

// float <-> order-preserving uint32 (ascending)
__host__ __device__ __forceinline__ uint32_t orderable_f32(uint32_t bits) {
  return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t unorderable_f32(uint32_t u) {
  return (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
}

// streaming (read-once) loads: keep them out of L1
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream_f(const float *p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// L2 residency hints (createpolicy + .L2::cache_hint): data that a later kernel re-reads is marked evict_last,
// write-once streams evict_first, so the 126 MB L2 keeps the former while the latter pass through.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float4 *ptr, uint64_t policy) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(ptr), "l"(policy));
  return r;
}
__device__ __forceinline__ void st_f_hint(float *ptr, float v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(ptr), "f"(v), "l"(policy) : "memory");
}

}  // namespace rv3d
