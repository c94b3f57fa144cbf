// FP64 FMA peak of the device (the roofline denominator of the d = 1024 Gram / transform kernels, which are
// FP64-compute-bound: SURVEY.md section 7, hard part 8).  Independent DFMA chains per thread, no memory traffic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/fp64_peak tools/micro/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  constexpr int ILP = 8;
  const int blocks = sms * 8, threads = 256, iters = 1 << 16;
  double* out;
  cudaMalloc(&out, (size_t)blocks * threads * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * (double)blocks * threads * ILP * iters;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  printf("{\"what\": \"DFMA chains, %d blocks x %d threads x ILP %d\", \"sms\": %d, \"fp64_tflops\": %.2f, "
         "\"dfma_per_clk_per_sm_at_1965MHz\": %.1f}\n",
         blocks, threads, ILP, sms, best, best * 1e12 / 2.0 / sms / 1.965e9);
  return 0;
}
