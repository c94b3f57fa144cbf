// extern "C" boundary (include/probabilit_b200.h): plain pointers and sizes only.
#include <atomic>
#include <cstring>
#include <string>

#include "../../include/probabilit_b200.h"
#include "common.cuh"
#include "ic.cuh"

namespace pbl {
std::atomic<long long> g_kernel_launches{0};
static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* get_last_error() { return g_last_error.c_str(); }
}  // namespace pbl

struct pbl_ic_plan {
  pbl::IcPlan* impl;
};

using pbl::kBadShape;
using pbl::kOk;

extern "C" {

int pbl_version(void) { return 100; }
const char* pbl_last_error(void) { return pbl::get_last_error(); }

int64_t pbl_kernel_launches(void) { return (int64_t)pbl::g_kernel_launches.load(); }

int pbl_sort_profile_enable(int on) {
  pbl::sort_profile_enable(on != 0);
  return kOk;
}
int pbl_sort_profile_read(int64_t* launches, double* total_ms, int64_t* keys) {
  pbl::sort_profile_read(launches, total_ms, keys);
  return kOk;
}

int pbl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int pbl_set_device(int device) {
  PBL_CUDA_CHECK(cudaSetDevice(device));
  return kOk;
}

int pbl_get_device(int* device) {
  if (!device) return kBadShape;
  PBL_CUDA_CHECK(cudaGetDevice(device));
  return kOk;
}

int pbl_device_malloc(void** ptr, uint64_t bytes) {
  PBL_CUDA_CHECK(cudaMalloc(ptr, bytes ? bytes : 16));
  return kOk;
}
int pbl_device_free(void* ptr) {
  PBL_CUDA_CHECK(cudaFree(ptr));
  return kOk;
}
int pbl_host_malloc_pinned(void** ptr, uint64_t bytes) {
  PBL_CUDA_CHECK(cudaMallocHost(ptr, bytes ? bytes : 16));
  return kOk;
}
int pbl_host_free_pinned(void* ptr) {
  PBL_CUDA_CHECK(cudaFreeHost(ptr));
  return kOk;
}
int pbl_memcpy_h2d(void* dst, const void* src, uint64_t bytes, void* stream) {
  PBL_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return kOk;
}
int pbl_memcpy_d2h(void* dst, const void* src, uint64_t bytes, void* stream) {
  PBL_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return kOk;
}
// ---- peer memory over NVLink (multi-GPU transposes, probabilit_b200/distributed.py) ----
int pbl_ipc_export(const void* ptr_dev, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
  cudaIpcMemHandle_t h;
  PBL_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr_dev)));
  memcpy(handle64, &h, sizeof(h));
  return kOk;
}
int pbl_ipc_open(const void* handle64, void** ptr_dev) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  PBL_CUDA_CHECK(cudaIpcOpenMemHandle(ptr_dev, h, cudaIpcMemLazyEnablePeerAccess));
  return kOk;
}
int pbl_ipc_close(void* ptr_dev) {
  PBL_CUDA_CHECK(cudaIpcCloseMemHandle(ptr_dev));
  return kOk;
}

namespace {
constexpr int kCopyStreams = 8;  // pool size; g_copy_streams of them are used
int g_copy_streams = 4;
struct CopyPool {
  bool ready = false;
  cudaStream_t s[kCopyStreams];
  cudaEvent_t fork, join[kCopyStreams];
};
CopyPool g_copy_pool[64];
}  // namespace

int pbl_peer_copy_streams(int32_t n) {
  if (n < 1 || n > kCopyStreams) return kBadShape;
  g_copy_streams = n;
  return kOk;
}

// `count` device-to-device copies (local or peer-mapped pointers, any mix), spread over a small
// pool of side streams so that several copy engines / NVLink ports work at once; ordered after
// everything already in `stream`, and `stream` continues after all of them.
int pbl_peer_copy_many(int32_t count, void* const* dst, const void* const* src, const uint64_t* bytes,
                       void* stream) {
  int dev = 0;
  PBL_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return kBadShape;
  CopyPool& cp = g_copy_pool[dev];
  if (!cp.ready) {
    PBL_CUDA_CHECK(cudaEventCreateWithFlags(&cp.fork, cudaEventDisableTiming));
    for (int i = 0; i < kCopyStreams; ++i) {
      PBL_CUDA_CHECK(cudaStreamCreateWithFlags(&cp.s[i], cudaStreamNonBlocking));
      PBL_CUDA_CHECK(cudaEventCreateWithFlags(&cp.join[i], cudaEventDisableTiming));
    }
    cp.ready = true;
  }
  cudaStream_t main = (cudaStream_t)stream;
  const int width = g_copy_streams;
  const int used = count < width ? count : width;
  PBL_CUDA_CHECK(cudaEventRecord(cp.fork, main));
  for (int i = 0; i < used; ++i) PBL_CUDA_CHECK(cudaStreamWaitEvent(cp.s[i], cp.fork, 0));
  for (int i = 0; i < count; ++i)
    if (bytes[i])
      PBL_CUDA_CHECK(cudaMemcpyAsync(dst[i], src[i], bytes[i], cudaMemcpyDefault, cp.s[i % width]));
  for (int i = 0; i < used; ++i) {
    PBL_CUDA_CHECK(cudaEventRecord(cp.join[i], cp.s[i]));
    PBL_CUDA_CHECK(cudaStreamWaitEvent(main, cp.join[i], 0));
  }
  return kOk;
}

int pbl_stream_synchronize(void* stream) {
  PBL_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return kOk;
}

// ---------------------------------------------------------------------------- Iman-Conover
int pbl_ic_plan_create_ex(int64_t n, int32_t k, int32_t col_batch, int32_t flags, pbl_ic_plan** plan) {
  if (!plan) return kBadShape;
  *plan = nullptr;
  pbl::IcPlan* impl = nullptr;
  PBL_RETURN_IF(pbl::ic_plan_create(n, k, col_batch, flags, &impl));
  *plan = new pbl_ic_plan{impl};
  return kOk;
}

int pbl_ic_plan_create(int64_t n, int32_t k, int32_t col_batch, pbl_ic_plan** plan) {
  return pbl_ic_plan_create_ex(n, k, col_batch, 0, plan);
}

int pbl_ic_plan_destroy(pbl_ic_plan* plan) {
  if (!plan) return kOk;
  pbl::ic_plan_destroy(plan->impl);
  delete plan;
  return kOk;
}

uint64_t pbl_ic_plan_bytes(const pbl_ic_plan* plan) { return plan ? plan->impl->bytes : 0; }

int pbl_ic_plan_set_target(pbl_ic_plan* plan, const double* P_lower) {
  if (!plan || !P_lower) return kBadShape;
  return pbl::ic_plan_set_target(plan->impl, P_lower);
}

int pbl_ic_plan_run(pbl_ic_plan* plan, const double* X, int64_t xrs, int64_t xcs, double* Y,
                    int64_t yrs, int64_t ycs, void* stream) {
  if (!plan || !X || !Y) return kBadShape;
  if (!plan->impl->has_target) {
    pbl::set_last_error("User must call `set_target` first.");
    return kBadShape;
  }
  return pbl::ic_plan_run(plan->impl, X, xrs, xcs, Y, yrs, ycs, (cudaStream_t)stream);
}

int pbl_cholesky_plan_run(pbl_ic_plan* plan, const double* X, int64_t xrs, int64_t xcs, double* Y,
                          int64_t yrs, int64_t ycs, void* stream) {
  if (!plan || !X || !Y) return kBadShape;
  return pbl::cholesky_correlator_run(plan->impl, X, xrs, xcs, Y, yrs, ycs, (cudaStream_t)stream);
}

int pbl_permcorr_begin(pbl_ic_plan* plan, const double* X, int64_t xrs, int64_t xcs, double* Y, int32_t spearman,
                       const double* target, const double* weights, void* stream) {
  if (!plan || !X || !Y || !target || !weights) return kBadShape;
  return pbl::permcorr_begin(plan->impl, X, xrs, xcs, Y, spearman, target, weights, (cudaStream_t)stream);
}
int pbl_permcorr_steps(pbl_ic_plan* plan, double* Y, const int32_t* step_col, const int32_t* step_off,
                       const int32_t* step_cnt, const int64_t* swaps, int64_t n_swaps_total, int64_t n_steps,
                       double tol, int64_t* steps_done, int32_t* converged, double* errors, int64_t errors_cap,
                       int64_t* n_errors, void* stream) {
  if (!plan || !Y || (n_steps > 0 && (!step_col || !step_off || !step_cnt || !swaps))) return kBadShape;
  return pbl::permcorr_steps(plan->impl, Y, step_col, step_off, step_cnt, swaps, n_swaps_total, n_steps, tol,
                             steps_done, converged, errors, errors_cap, n_errors, (cudaStream_t)stream);
}
int pbl_permcorr_corr(pbl_ic_plan* plan, double* corr) {
  if (!plan || !corr) return kBadShape;
  return pbl::permcorr_read_corr(plan->impl, corr);
}

int pbl_copy_strided_f64(const double* src, int64_t srs, int64_t scs, double* dst, int64_t drs, int64_t dcs, int64_t n,
                         int32_t k, void* stream) {
  if (n < 0 || k < 0 || (n > 0 && k > 0 && (!src || !dst))) return kBadShape;
  return pbl::copy_strided(src, srs, scs, dst, drs, dcs, n, k, (cudaStream_t)stream);
}

int pbl_corrcoef_f64(pbl_ic_plan* plan, const double* X, int64_t xrs, int64_t xcs, int32_t spearman, double* out,
                     void* stream) {
  if (!plan || !X || !out) return kBadShape;
  return pbl::corrcoef_run(plan->impl, X, xrs, xcs, spearman, out, (cudaStream_t)stream);
}

static bool contiguous_layout(int64_t n, int32_t k, int64_t rs, int64_t cs) {
  return (rs == 1 && cs == n) || (cs == 1 && rs == k) || (n == 1 && cs == 1) || (k == 1 && rs == 1);
}

int pbl_ic_plan_run_host(pbl_ic_plan* plan, const double* X, int64_t xrs, int64_t xcs, double* Y, int64_t yrs,
                         int64_t ycs, double* X_staging_dev, double* Y_staging_dev, void* stream) {
  if (!plan || !X || !Y || !X_staging_dev || !Y_staging_dev) return kBadShape;
  if (!plan->impl->has_target) {
    pbl::set_last_error("User must call `set_target` first.");
    return kBadShape;
  }
  if (!contiguous_layout(plan->impl->n, plan->impl->k, xrs, xcs) ||
      !contiguous_layout(plan->impl->n, plan->impl->k, yrs, ycs)) {
    pbl::set_last_error("pbl_ic_plan_run_host: X and Y must be contiguous in C or F order");
    return kBadShape;
  }
  return pbl::ic_plan_run_host(plan->impl, X, xrs, xcs, Y, yrs, ycs, X_staging_dev, Y_staging_dev,
                               (cudaStream_t)stream);
}

int pbl_iman_conover_f64(const double* X, int64_t n, int32_t k, int64_t xrs, int64_t xcs,
                         const double* P_lower, double* Y, int64_t yrs, int64_t ycs) {
  if (!X || !Y || !P_lower) return kBadShape;
  if (!contiguous_layout(n, k, xrs, xcs) || !contiguous_layout(n, k, yrs, ycs)) {
    pbl::set_last_error("pbl_iman_conover_f64: X and Y must be contiguous in C or F order");
    return kBadShape;
  }
  pbl_ic_plan* plan = nullptr;
  double *dX = nullptr, *dY = nullptr;
  const size_t bytes = (size_t)n * k * sizeof(double);
  int rc = pbl_ic_plan_create(n, k, 0, &plan);
  if (rc == kOk) rc = pbl_ic_plan_set_target(plan, P_lower);
  if (rc == kOk && cudaMalloc((void**)&dX, bytes) != cudaSuccess) rc = pbl::kCudaError;
  if (rc == kOk && cudaMalloc((void**)&dY, bytes) != cudaSuccess) rc = pbl::kCudaError;
  if (rc == pbl::kCudaError && !*pbl::get_last_error()) pbl::set_last_error("cudaMalloc failed");
  if (rc == kOk) rc = pbl_ic_plan_run_host(plan, X, xrs, xcs, Y, yrs, ycs, dX, dY, nullptr);
  cudaFree(dX);
  cudaFree(dY);
  pbl_ic_plan_destroy(plan);
  return rc;
}

int pbl_ic_stage_begin(pbl_ic_plan* plan, void* stream) {
  if (!plan) return kBadShape;
  PBL_CUDA_CHECK(cudaMemsetAsync(plan->impl->flags, 0, 8 * sizeof(uint32_t), (cudaStream_t)stream));
  return kOk;
}
int pbl_ic_stage_rank_scores(pbl_ic_plan* plan, const double* X, int64_t rs, int64_t cs,
                             int32_t col0, int32_t ncols, void* stream) {
  if (!plan || !X || col0 < 0 || ncols < 0 || col0 + ncols > plan->impl->k) return kBadShape;
  return pbl::ic_stage_rank_scores(plan->impl, X, rs, cs, col0, ncols, (cudaStream_t)stream);
}
int pbl_ic_stage_gram(pbl_ic_plan* plan, void* stream) {
  if (!plan) return kBadShape;
  return pbl::ic_stage_gram(plan->impl, (cudaStream_t)stream);
}
int pbl_ic_stage_solve(pbl_ic_plan* plan, int64_t n_total, void* stream) {
  if (!plan) return kBadShape;
  return pbl::ic_stage_solve(plan->impl, n_total, (cudaStream_t)stream);
}
int pbl_ic_stage_transform(pbl_ic_plan* plan, void* stream) {
  if (!plan) return kBadShape;
  return pbl::ic_stage_transform(plan->impl, (cudaStream_t)stream);
}
int pbl_ic_stage_rank_gather(pbl_ic_plan* plan, double* Y, int64_t rs, int64_t cs, int32_t col0,
                             int32_t ncols, void* stream) {
  if (!plan || !Y || col0 < 0 || ncols < 0 || col0 + ncols > plan->impl->k) return kBadShape;
  return pbl::ic_stage_rank_gather(plan->impl, Y, rs, cs, col0, ncols, (cudaStream_t)stream);
}
int pbl_ic_stage_status(pbl_ic_plan* plan, void* stream) {
  if (!plan) return kBadShape;
  return pbl::ic_read_status(plan->impl, (cudaStream_t)stream);
}

int pbl_ic_plan_set_chunk_hook(pbl_ic_plan* plan, int64_t chunk_rows, int32_t first_chunk, pbl_chunk_fn fn,
                               void* user) {
  if (!plan) return kBadShape;
  pbl::IcPlan* p = plan->impl;
  p->chunk_rows = fn ? chunk_rows : 0;
  p->chunk_first = first_chunk;
  p->chunk_fn = fn;
  p->chunk_user = user;
  return kOk;
}

int pbl_ic_plan_buffer(pbl_ic_plan* plan, int32_t what, void** ptr, uint64_t* bytes) {
  if (!plan || !ptr) return kBadShape;
  pbl::IcPlan* p = plan->impl;
  const uint64_t nk = (uint64_t)p->n * p->k * 8, kk = (uint64_t)p->k * p->k * 8;
  uint64_t b = 0;
  switch (what) {
    case 0: *ptr = p->scores; b = nk; break;
    case 1: *ptr = p->sortedX; b = nk; break;
    case 2: *ptr = p->gram; b = kk; break;
    case 3: *ptr = p->colsum; b = (uint64_t)p->k * 8; break;
    case 4: *ptr = p->T; b = kk; break;
    case 5: *ptr = p->work; b = kk; break;
    default: return kBadShape;
  }
  if (bytes) *bytes = b;
  return kOk;
}

}  // extern "C"
