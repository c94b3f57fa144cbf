// Microbenchmark: throughput of MATCH.ANY vs the 8-ballot emulation, per SM, on random 8-bit digits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o match_bench match_bench.cu && ./match_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, int iters, int distinct_mask) {
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x = x * 1664525u + 1013904223u;
      uint32_t d = (x >> 24) & distinct_mask;
      uint32_t m;
      if (MODE == 0) {
        m = __match_any_sync(0xFFFFFFFFu, d);
      } else if (MODE == 1) {
        m = 0xFFFFFFFFu;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const bool bit = (d >> b) & 1u;
          const uint32_t bal = __ballot_sync(0xFFFFFFFFu, bit);
          m &= bit ? bal : ~bal;
        }
      } else {
        m = d;  // baseline: LCG only
      }
      acc += __popc(m) + (m & 1);
    }
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int MODE>
float run(uint32_t* out, int iters, int mask) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<MODE><<<148 * 8, 256>>>(out, 10, mask);
  cudaEventRecord(a);
  k<MODE><<<148 * 8, 256>>>(out, iters, mask);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  uint32_t* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 2000;
  const double matches_per_sm = 8.0 * 8 /*warps per block*/ * iters * 16;  // warp-level ops per SM
  for (int mask : {255, 15, 1}) {
    float t0 = run<0>(out, iters, mask), t1 = run<1>(out, iters, mask), t2 = run<2>(out, iters, mask);
    printf("distinct<=%3d  match.any %.3f ms (%.1f clk/warp-op/SM @1.965GHz)  ballot8 %.3f ms (%.1f)  baseline %.3f ms (%.1f)\n",
           mask + 1, t0, t0 * 1e-3 * 1.965e9 / matches_per_sm, t1, t1 * 1e-3 * 1.965e9 / matches_per_sm, t2,
           t2 * 1e-3 * 1.965e9 / matches_per_sm);
  }
  return 0;
}
