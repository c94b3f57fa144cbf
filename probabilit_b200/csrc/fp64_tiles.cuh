// FP64 Gram and transform for WIDE problems (k > 64; BASELINE.json configs[3]: N = 1e6, d = 1024), where both are
// FP64-compute-bound (2 N k^2 flop each against 16 N k bytes: 128 flop/B at k = 1024) and the denominator is the
// device's DFMA rate (tools/micro/fp64_peak.cu: 34.2 TFLOP/s measured), not HBM.  Included by ic.cu.
//
// Both kernels are a 128 x 128 output tile per block of 256 threads, 8 x 8 outputs per thread (64 independent
// DFMA chains, 16 shared-memory loads per 64 DFMAs: 2 B of shared-memory traffic per DFMA against the 4 B of
// the 4 x 4 register tiles they replace, which were shared-memory-bound at 25-34 % of the DFMA rate), the
// long dimension consumed 16 deep per step from a double buffer: the global loads of step s+1 are issued
// before the DFMAs of step s and parked in registers, one block barrier per step.
//   reference: np.corrcoef(normal_scores, rowvar=False) (correlation.py:398) and
//              solve_triangular(...) @ P.T == scores @ T (correlation.py:409-414)
#pragma once

constexpr int kBT = 128;   // tile edge
constexpr int kBD = 16;    // depth per step
constexpr int kBLD = 129;  // padded row of a TRANSPOSED stage (odd: the transposing stores spread over the banks)
constexpr size_t kGramBigSmem = (size_t)2 * 2 * kBD * kBLD * sizeof(double);
constexpr size_t kTransformBigSmem = (size_t)2 * 2 * kBD * kBT * sizeof(double);

// acc[i][j] += a[i] * b[j] over the kBD rows of a stage; a from sa[r][ia + 16 i], b from sb[r][ib + 16 j]
template <int LDA, int LDB>
__device__ __forceinline__ void tile_fma(double (&acc)[8][8], const double* __restrict__ sa,
                                         const double* __restrict__ sb, const int ia, const int ib) {
#pragma unroll 4
  for (int r = 0; r < kBD; ++r) {
    double a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = sa[r * LDA + ia + 16 * i];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = sb[r * LDB + ib + 16 * j];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
}

// Partial Gram of the rows [blockIdx.x * rows_per_block, ...) for the pair of 128-column tiles blockIdx.y;
// output layout of gram_kernel<TG> with CT = 128 (consumed by gram_reduce_kernel).
__global__ void __launch_bounds__(256, 1)
gram_big_kernel(const double* __restrict__ S, int64_t n, int k, double* __restrict__ partials, int nrb,
                int64_t rows_per_block) {
  extern __shared__ __align__(16) double bsm[];
  double* sA = bsm;                       // [2][kBD][kBLD]  (row r, column c of tile I)
  double* sB = bsm + 2 * kBD * kBLD;      // [2][kBD][kBLD]  tile J
  const int nt = (k + kBT - 1) / kBT;
  int I = 0, J = 0;
  {
    int p = blockIdx.y;
    for (I = 0; I < nt; ++I) {
      const int cnt = nt - I;
      if (p < cnt) { J = I + p; break; }
      p -= cnt;
    }
  }
  const bool diag = I == J;
  const int tid = threadIdx.x;
  const int ti = tid >> 4, tj = tid & 15;
  const int lc = tid >> 1, lr = (tid & 1) * 8;  // loader: column lc of the tile, rows lr .. lr + 7 of the step
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(n, r_begin + rows_per_block);
  const int colA = I * kBT + lc, colB = J * kBT + lc;
  const double* pA = S + (int64_t)min(colA, k - 1) * n;
  const double* pB = S + (int64_t)min(colB, k - 1) * n;
  const bool okA = colA < k, okB = colB < k;

  double acc[8][8], sum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sum[i] = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  }
  double ra[8], rb[8];
  auto fetch = [&](int64_t r0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int64_t row = r0 + lr + e;
      const bool in = row < r_end;
      ra[e] = (okA && in) ? __ldg(pA + row) : 0.0;
      rb[e] = (!diag && okB && in) ? __ldg(pB + row) : 0.0;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sA[(buf * kBD + lr + e) * kBLD + lc] = ra[e];
      if (!diag) sB[(buf * kBD + lr + e) * kBLD + lc] = rb[e];
    }
  };
  fetch(r_begin);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += kBD, buf ^= 1) {
    const bool more = r0 + kBD < r_end;
    if (more) fetch(r0 + kBD);
    const double* a = sA + buf * kBD * kBLD;
    const double* b = diag ? a : sB + buf * kBD * kBLD;
    tile_fma<kBLD, kBLD>(acc, a, b, ti, tj);
    if (diag && tj == 0) {
#pragma unroll 4
      for (int r = 0; r < kBD; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) sum[i] += a[r * kBLD + ti + 16 * i];
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
  }
  double* out = partials + ((size_t)blockIdx.y * nrb + blockIdx.x) * (kBT * kBT + kBT);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) out[(ti + 16 * i) * kBT + tj + 16 * j] = acc[i][j];
    if (tj == 0) out[kBT * kBT + ti + 16 * i] = sum[i];
  }
}

// S <- S @ T in place (T upper triangular, row-major k x k) for the 128 rows of the block: column tiles in
// DESCENDING order, so that a tile is overwritten only after every tile that reads it as input is done.
__global__ void __launch_bounds__(256, 1)
transform_big_kernel(double* __restrict__ S, int64_t n, int k, const double* __restrict__ T) {
  extern __shared__ __align__(16) double bsm[];
  double* sS = bsm;                     // [2][kBD][kBT]  (depth j, row r)
  double* sT = bsm + 2 * kBD * kBT;     // [2][kBD][kBT]  (depth j, column c)
  const int tid = threadIdx.x;
  const int tr = tid & 15, tc = tid >> 4;
  const int lj = tid >> 4, lx = (tid & 15) * 8;  // loader: depth lj of the step, 8 consecutive rows / columns
  const int64_t r0 = (int64_t)blockIdx.x * kBT;
  const int nt = (k + kBT - 1) / kBT;
  double rs[8], rt[8];
  for (int kt = nt - 1; kt >= 0; --kt) {
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
    const int jmax = min(k, (kt + 1) * kBT);  // T[j][c] = 0 for j > c
    auto fetch = [&](int j0) {
      const int gj = j0 + lj;
      const bool okj = gj < jmax;
      const double* ps = S + (int64_t)min(gj, k - 1) * n + r0 + lx;
      const double* pt = T + (size_t)min(gj, k - 1) * k + kt * kBT + lx;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        rs[e] = (okj && r0 + lx + e < n) ? ps[e] : 0.0;
        rt[e] = (okj && kt * kBT + lx + e < k) ? __ldg(pt + e) : 0.0;
      }
    };
    auto stash = [&](int buf) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sS[(buf * kBD + lj) * kBT + lx + e] = rs[e];
        sT[(buf * kBD + lj) * kBT + lx + e] = rt[e];
      }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int j0 = 0; j0 < jmax; j0 += kBD, buf ^= 1) {
      const bool more = j0 + kBD < jmax;
      if (more) fetch(j0 + kBD);
      tile_fma<kBT, kBT>(acc, sS + buf * kBD * kBT, sT + buf * kBD * kBT, tr, tc);
      if (more) stash(buf ^ 1);
      __syncthreads();
    }
    // every read of this block's rows of column tile kt is complete (barrier above); lower column tiles never
    // read tile kt again, so it can be overwritten now
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gc = kt * kBT + tc + 16 * j;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t row = r0 + tr + 16 * i;
        if (gc < k && row < n) S[(int64_t)gc * n + row] = acc[i][j];
      }
    }
    __syncthreads();  // (the stage buffers are refilled by the next column tile)
  }
}
