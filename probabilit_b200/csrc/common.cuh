// Shared helpers for the probabilit_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

namespace pbl {

// status codes returned through the C ABI (include/probabilit_b200.h)
enum Status : int {
  kOk = 0,
  kNotPositiveDefinite = 1,  // rank correlation not PD  (reference correlation.py:399-403)
  kNonFinite = 2,            // NaN in X / non-finite scores (scipy check_finite in :409)
  kBadShape = 3,
  kCudaError = 4,
  kInternal = 5,             // look-back watchdog fired, etc.
  kRetry = 6,                // stage API only: repeat from stage_begin (the plan now sorts on 64 bits)
};

void set_last_error(const std::string& msg);
const char* get_last_error();

#define PBL_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      char _buf[512];                                                                     \
      snprintf(_buf, sizeof(_buf), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),    \
               __FILE__, __LINE__, cudaGetErrorString(_e));                               \
      ::pbl::set_last_error(_buf);                                                        \
      return ::pbl::kCudaError;                                                           \
    }                                                                                     \
  } while (0)

// every kernel launch is counted (bench.py reports it as gpu_launches) and checked
extern std::atomic<long long> g_kernel_launches;
#define PBL_LAUNCH_CHECK()                                                   \
  do {                                                                       \
    ::pbl::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);        \
    PBL_CUDA_CHECK(cudaGetLastError());                                      \
  } while (0)

#define PBL_RETURN_IF(expr)                \
  do {                                     \
    int _s = (expr);                       \
    if (_s != ::pbl::kOk) return _s;       \
  } while (0)

// Order-preserving map fp64 bits -> u64 (and back).  -0.0 < +0.0 as keys; tie detection
// is done on the decoded doubles so they still tie (scipy/stats/_stats_py.py:10408).
__host__ __device__ __forceinline__ uint64_t flip_f64(uint64_t u) {
  return u ^ ((u >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ uint64_t unflip_f64(uint64_t k) {
  return k ^ ((k >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
}
__device__ __forceinline__ double key_to_double(uint64_t k) {
  return __longlong_as_double((long long)unflip_f64(k));
}

// streaming (read-once / write-once) accesses: keep them out of L1
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream_f64(const double* p) {
  return __longlong_as_double((long long)ld_stream_u64((const uint64_t*)p));
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t lane_id() {
  uint32_t r;
  asm("mov.u32 %0, %%laneid;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t r;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r));
  return r;
}

static inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace pbl
