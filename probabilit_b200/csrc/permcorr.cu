// PermutationCorrelator on the device (reference src/probabilit/correlation.py:473-703) with the
// incremental correlation update of CorrelationMatrix (:757-921).
//
// The algorithm is a serial randomised hill climb: step (iteration, column k) proposes swapping
// rows i[] <-> j[] inside column k, computes the change of row/column k of the correlation
// matrix from the 2*s touched rows only (O(s K), :882-907) and keeps the swap iff the weighted
// squared error of that row drops (:679-686).  It is latency bound, not bandwidth bound, so it
// runs as ONE persistent thread block that loops over the steps (no launch per step, state in
// shared/global memory, one thread per variable for the O(s K) part).  The swap indices come
// from the host: they are the reference's own NumPy stream (SwapIndexGenerator, :428-470), which
// is what makes the accept/reject sequence -- and therefore the output -- reproduce the
// reference's exactly.
//
// Reductions whose rounding decides an accept/reject use NumPy's summation order (pairwise_sum of
// numpy/_core/src/umath/loops_utils.h.src: 8 accumulators, blocks of 128) so that the decision
// variable is computed like np.average / np.sum compute it.
#include <vector>

#include "../../include/probabilit_b200.h"
#include "common.cuh"
#include "ic.cuh"

namespace pbl {

struct PermCorrState {
  int k = 0;
  int64_t n = 0;
  int spearman = 0;
  double* corr = nullptr;    // [k][k]
  double* numer = nullptr;   // [k][k]
  double* denom = nullptr;   // [k]
  double* target = nullptr;  // [k][k]
  double* weights = nullptr; // [k][k], normalised to sum 1 (correlation.py:592-593)
  double* scratch = nullptr; // [k*k] work array for the error sums, [3k] step vectors
  double* mean = nullptr;    // [k]
  double* errors = nullptr;  // [cap] current_error after every k == 0 step
  int64_t errors_cap = 0;
  int64_t* idx = nullptr;    // swap indices of one chunk
  int64_t idx_cap = 0;
  int32_t* meta = nullptr;   // per step: column, offset, count
  int64_t meta_cap = 0;
  int64_t* result = nullptr; // [2]: steps done, converged flag
  const double* xs = nullptr;  // the matrix correlations are induced on (ranks for spearman, else Y)
};

namespace {

// numpy pairwise_sum for doubles (single thread)
__device__ double np_pairwise_sum(const double* a, int64_t n) {
  if (n < 8) {
    double res = 0.0;
    for (int64_t i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int64_t i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// corr = (numer / denom[None, :]) / denom[:, None]   (correlation.py:850-852)
__global__ void permcorr_init_kernel(const double* __restrict__ gram, double m, const double* __restrict__ denom,
                                     int k, double* __restrict__ numer, double* __restrict__ corr,
                                     uint32_t* __restrict__ flags) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < k * k; e += gridDim.x * blockDim.x) {
    const int i = e / k, j = e % k;
    const double num = gram[e] / m;
    numer[e] = num;
    corr[e] = (num / denom[j]) / denom[i];
    // np.isclose(denominator, 0): |d| <= 1e-8  -> "X has one or several constant columns"
    if (i == j && !(fabs(denom[i]) > 1e-8)) flags[kFlagReserved] = 1u;
  }
}

// _error (:597-601): sqrt(sum_{i<j} w_ij (obs_ij - target_ij)^2), np.sum order
__device__ double permcorr_error(const PermCorrState& s, int tid, int nth) {
  const int k = s.k;
  // row-major upper-triangular order = np.triu_indices(k, 1)
  for (int i = tid; i < k; i += nth) {
    int64_t base = (int64_t)i * (2 * k - i - 1) / 2;
    for (int j = i + 1; j < k; ++j) {
      const double d = s.corr[(size_t)i * k + j] - s.target[(size_t)i * k + j];
      s.scratch[base + (j - i - 1)] = s.weights[(size_t)i * k + j] * (d * d);
    }
  }
  __syncthreads();
  __shared__ double err;
  if (tid == 0) err = sqrt(np_pairwise_sum(s.scratch, (int64_t)k * (k - 1) / 2));
  __syncthreads();
  return err;
}

__global__ void __launch_bounds__(256)
permcorr_steps_kernel(const PermCorrState s, double* __restrict__ Y, double* __restrict__ ranks, int64_t n_steps,
                      double tol, int check_every_cycle) {
  const int k = s.k, tid = threadIdx.x, nth = blockDim.x;
  const int64_t n = s.n;
  const double m = (double)n;
  double* X_ = s.spearman ? ranks : Y;  // [k][n]
  double* vec_delta = s.scratch + (size_t)k * k;  // [k] delta of row/column `col`
  double* vec_old = vec_delta + k;                // [k] w * (target - old)^2
  double* vec_new = vec_old + k;                  // [k] w * (target - new)^2
  __shared__ int accept;
  int64_t step = 0;
  int64_t n_checks = 1;
  bool converged = false;
  {  // errors[0] = the error before the first step of this chunk (what verbose mode prints, :655-660)
    const double e0 = permcorr_error(s, tid, nth);
    if (tid == 0 && s.errors_cap > 0) s.errors[0] = e0;
  }
  for (; step < n_steps; ++step) {
    const int col = s.meta[3 * step], cnt = s.meta[3 * step + 2];
    const int64_t* si = s.idx + 2 * (int64_t)s.meta[3 * step + 1];
    const int64_t* sj = si + cnt;
    const double dcol = s.denom[col];
    for (int c = tid; c < k; c += nth) {
      // _delta_numerator (:882-907): sum over the swaps, in order
      double dn = 0.0;
      for (int t = 0; t < cnt; ++t) {
        const int64_t i = si[t], j = sj[t];
        const double ric = X_[(size_t)col * n + i], rjc = X_[(size_t)col * n + j];
        const double a = X_[(size_t)c * n + i] - X_[(size_t)c * n + j];
        dn += __dmul_rn(a, rjc - ric);
      }
      if (c == col) dn = 0.0;
      const double delta = dn / ((m * s.denom[c]) * dcol);  // :911-912
      vec_delta[c] = delta;
      const double oldv = s.corr[(size_t)col * k + c];
      const double newv = s.corr[(size_t)c * k + col] + delta;  // update_column (:914-919)
      const double w = s.weights[(size_t)col * k + c], t = s.target[(size_t)col * k + c];
      const double eo = t - oldv, en = t - newv;
      vec_old[c] = __dmul_rn(__dmul_rn(eo, eo), w);  // np.average: (a * w).sum() / w.sum()
      vec_new[c] = __dmul_rn(__dmul_rn(en, en), w);
    }
    __syncthreads();
    if (tid == 0) {
      const double scl = np_pairwise_sum(s.weights + (size_t)col * k, k);
      const double old_error = np_pairwise_sum(vec_old, k) / scl;
      const double new_error = np_pairwise_sum(vec_new, k) / scl;
      accept = new_error < old_error ? 1 : 0;
    }
    __syncthreads();
    if (accept) {  // commit (:861-880)
      for (int c = tid; c < k; c += nth) {
        const double d = vec_delta[c];  // 0 on the diagonal, where the reference adds it twice
        s.corr[(size_t)c * k + col] += d;
        s.corr[(size_t)col * k + c] += d;
      }
      for (int t = tid; t < cnt; t += nth) {
        const int64_t i = si[t], j = sj[t];
        const double vi = X_[(size_t)col * n + i], vj = X_[(size_t)col * n + j];
        X_[(size_t)col * n + i] = vj;
        X_[(size_t)col * n + j] = vi;
        if (s.spearman) {
          const double yi = Y[(size_t)col * n + i], yj = Y[(size_t)col * n + j];
          Y[(size_t)col * n + i] = yj;
          Y[(size_t)col * n + j] = yi;
        }
      }
      __threadfence_block();
    }
    __syncthreads();
    if (col == 0 && check_every_cycle) {
      const double e = permcorr_error(s, tid, nth);
      if (tid == 0 && n_checks < s.errors_cap) s.errors[n_checks] = e;
      ++n_checks;
      if (e < tol) {
        converged = true;
        ++step;
        break;
      }
    }
  }
  if (tid == 0) {
    s.result[0] = step;
    s.result[1] = converged ? 1 : 0;
  }
}

template <typename T>
int ensure(T** p, int64_t* cap, int64_t want) {
  if (*cap >= want) return kOk;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  PBL_CUDA_CHECK(cudaMalloc((void**)p, (size_t)std::max<int64_t>(want, 16) * sizeof(T)));
  *cap = want;
  return kOk;
}

}  // namespace

void permcorr_free(void* st) {
  PermCorrState* s = static_cast<PermCorrState*>(st);
  if (!s) return;
  cudaFree(s->corr);
  cudaFree(s->errors);
  cudaFree(s->idx);
  cudaFree(s->meta);
  cudaFree(s->result);
  delete s;
}

// CorrelationMatrix.__init__ (:819-853) on Y = copy of X (column-major [k][n], owned by the caller)
int permcorr_begin(IcPlan* p, const double* X, int64_t xrs, int64_t xcs, double* Y, int spearman,
                   const double* target_host, const double* weights_host, cudaStream_t stream) {
  const int k = p->k;
  const int64_t n = p->n;
  if (spearman && p->rows_only) {
    set_last_error("permcorr: the Spearman mode needs a plan with a sort workspace");
    return kBadShape;
  }
  PermCorrState* s = static_cast<PermCorrState*>(p->permcorr);
  if (!s) {
    s = new PermCorrState();
    s->k = k;
    s->n = n;
    const size_t kk = (size_t)k * k;
    PBL_CUDA_CHECK(cudaMalloc((void**)&s->corr, (5 * kk + 5 * (size_t)k + 16) * 8));
    s->numer = s->corr + kk;
    s->target = s->numer + kk;
    s->weights = s->target + kk;
    s->scratch = s->weights + kk;  // kk + 3k
    s->denom = s->scratch + kk + 3 * (size_t)k;
    s->mean = s->denom + k;
    PBL_CUDA_CHECK(cudaMalloc((void**)&s->result, 2 * sizeof(int64_t)));
    p->permcorr = s;
  }
  s->spearman = spearman;
  const size_t kk = (size_t)k * k;
  PBL_CUDA_CHECK(cudaMemcpyAsync(s->target, target_host, kk * 8, cudaMemcpyHostToDevice, stream));
  PBL_CUDA_CHECK(cudaMemcpyAsync(s->weights, weights_host, kk * 8, cudaMemcpyHostToDevice, stream));
  PBL_CUDA_CHECK(cudaMemsetAsync(p->flags, 0, 8 * sizeof(uint32_t), stream));
  // Y <- X (self.X = X.copy(), :831), any input layout -> column-major
  PBL_CUDA_CHECK(cudaMemsetAsync(s->mean, 0, (size_t)k * 8, stream));
  PBL_RETURN_IF(centre_columns(p, X, xrs, xcs, s->mean, Y, stream));
  const double* basis = Y;  // what correlations are induced on
  if (spearman) {
    PBL_RETURN_IF(ic_stage_rank_scores(p, X, xrs, xcs, 0, k, stream, /*ranks_only=*/true));
    // keep the ranks in sortedX's storage: scores is needed for the centred copy below
    PBL_CUDA_CHECK(cudaMemcpyAsync(p->sortedX, p->scores, (size_t)n * k * 8, cudaMemcpyDeviceToDevice, stream));
    basis = p->sortedX;
  }
  s->xs = basis;
  // X_centered = X_ - mean; numerator = X_c^T X_c / m; denominator = std(X_c)   (:843-846)
  PBL_RETURN_IF(column_moments(p, basis, 1, n, nullptr, s->mean, 0, stream));
  PBL_RETURN_IF(centre_columns(p, basis, 1, n, s->mean, p->scores, stream));
  PBL_RETURN_IF(ic_stage_gram(p, stream));
  PBL_RETURN_IF(column_moments(p, p->scores, 1, n, nullptr, s->denom, 0, stream));     // mean of the centred data
  PBL_RETURN_IF(column_moments(p, p->scores, 1, n, s->denom, s->denom, 1, stream));    // np.std about it
  permcorr_init_kernel<<<std::min(64, (k * k + 255) / 256), 256, 0, stream>>>(p->gram, (double)n, s->denom, k,
                                                                              s->numer, s->corr, p->flags);
  PBL_LAUNCH_CHECK();
  uint32_t h[8];
  PBL_CUDA_CHECK(cudaMemcpyAsync(h, p->flags, sizeof(h), cudaMemcpyDeviceToHost, stream));
  PBL_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (h[kFlagNaN]) {
    set_last_error("array must not contain infs or NaNs");
    return kNonFinite;
  }
  if (h[kFlagReserved]) {
    set_last_error("X has one or several constant columns");
    return kNotPositiveDefinite;
  }
  return kOk;
}

int permcorr_steps(IcPlan* p, double* Y, const int32_t* step_col, const int32_t* step_off, const int32_t* step_cnt,
                   const int64_t* swaps, int64_t n_swaps_total, int64_t n_steps, double tol, int64_t* steps_done,
                   int32_t* converged, double* errors_host, int64_t errors_cap, int64_t* n_errors,
                   cudaStream_t stream) {
  PermCorrState* s = static_cast<PermCorrState*>(p->permcorr);
  if (!s) {
    set_last_error("permcorr_steps: call permcorr_begin first");
    return kBadShape;
  }
  if (n_steps <= 0) {
    if (steps_done) *steps_done = 0;
    if (converged) *converged = 0;
    if (n_errors) *n_errors = 0;
    return kOk;
  }
  std::vector<int32_t> meta((size_t)3 * n_steps);
  for (int64_t t = 0; t < n_steps; ++t) {
    if (step_col[t] < 0 || step_col[t] >= p->k || step_cnt[t] < 0 || step_off[t] < 0 ||
        (int64_t)step_off[t] + step_cnt[t] > n_swaps_total) {
      set_last_error("permcorr_steps: step table out of range");
      return kBadShape;
    }
    meta[3 * t] = step_col[t];
    meta[3 * t + 1] = step_off[t];
    meta[3 * t + 2] = step_cnt[t];
  }
  for (int64_t t = 0; t < 2 * n_swaps_total; ++t) {
    if (swaps[t] < 0 || swaps[t] >= p->n) {
      set_last_error("permcorr_steps: swap index out of range");
      return kBadShape;
    }
  }
  PBL_RETURN_IF(ensure(&s->meta, &s->meta_cap, 3 * n_steps));
  PBL_RETURN_IF(ensure(&s->idx, &s->idx_cap, 2 * n_swaps_total));
  const int64_t want_err = n_steps / std::max(1, p->k) + 3;
  PBL_RETURN_IF(ensure(&s->errors, &s->errors_cap, want_err));
  PBL_CUDA_CHECK(cudaMemcpyAsync(s->meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, stream));
  PBL_CUDA_CHECK(cudaMemcpyAsync(s->idx, swaps, (size_t)2 * n_swaps_total * 8, cudaMemcpyHostToDevice, stream));
  double* ranks = s->spearman ? p->sortedX : nullptr;
  permcorr_steps_kernel<<<1, 256, 0, stream>>>(*s, Y, ranks, n_steps, tol, 1);
  PBL_LAUNCH_CHECK();
  int64_t res[2] = {0, 0};
  PBL_CUDA_CHECK(cudaMemcpyAsync(res, s->result, sizeof(res), cudaMemcpyDeviceToHost, stream));
  PBL_CUDA_CHECK(cudaStreamSynchronize(stream));
  if (steps_done) *steps_done = res[0];
  if (converged) *converged = (int32_t)res[1];
  int64_t ne = 1;  // errors[0] = error before the chunk, then one per completed k == 0 step
  for (int64_t t = 0; t < res[0]; ++t) ne += step_col[t] == 0 ? 1 : 0;
  if (n_errors) *n_errors = ne;
  if (errors_host && ne > 0) {
    const int64_t take = std::min(ne, std::min(errors_cap, s->errors_cap));
    PBL_CUDA_CHECK(cudaMemcpy(errors_host, s->errors, (size_t)take * 8, cudaMemcpyDeviceToHost));
  }
  return kOk;
}

int permcorr_read_corr(IcPlan* p, double* corr_host) {
  PermCorrState* s = static_cast<PermCorrState*>(p->permcorr);
  if (!s) return kBadShape;
  PBL_CUDA_CHECK(cudaMemcpy(corr_host, s->corr, (size_t)p->k * p->k * 8, cudaMemcpyDeviceToHost));
  return kOk;
}

// Layout conversion of an (n, k) matrix: a block moves a tile of 32 rows x 32 columns through shared memory,
// reading along the source's unit-stride dimension and writing along the destination's, so that both sides are
// coalesced when one layout is row-major and the other column-major.
__global__ void __launch_bounds__(256)
copy_strided_kernel(const double* __restrict__ src, int64_t srs, int64_t scs, double* __restrict__ dst, int64_t drs,
                    int64_t dcs, int64_t n, int k) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = (int)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const bool src_rows_fast = srs <= scs;  // consecutive rows are adjacent in the source
  const bool dst_rows_fast = drs <= dcs;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + (src_rows_fast ? tx : i);
    const int c = c0 + (src_rows_fast ? i : tx);
    if (r < n && c < k) tile[(int)(r - r0)][c - c0] = src[r * srs + (int64_t)c * scs];
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + (dst_rows_fast ? tx : i);
    const int c = c0 + (dst_rows_fast ? i : tx);
    if (r < n && c < k) dst[r * drs + (int64_t)c * dcs] = tile[(int)(r - r0)][c - c0];
  }
}

int copy_strided(const double* src, int64_t srs, int64_t scs, double* dst, int64_t drs, int64_t dcs, int64_t n, int k,
                 cudaStream_t stream) {
  if (n == 0 || k == 0) return kOk;
  const int64_t row_tiles = (n + 31) / 32;
  if (row_tiles > 0x7FFFFFFF) {
    set_last_error("copy_strided: too many rows");
    return kBadShape;
  }
  copy_strided_kernel<<<dim3((unsigned)row_tiles, (unsigned)((k + 31) / 32)), 256, 0, stream>>>(src, srs, scs, dst, drs,
                                                                                                 dcs, n, k);
  PBL_LAUNCH_CHECK();
  return kOk;
}

}  // namespace pbl
