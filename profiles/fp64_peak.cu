// fp64_peak.cu — measures the DFMA issue ceiling of the B200 (denominator of the FP64-pipe roofline quoted in DESIGN.md).
// build + run on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a profiles/fp64_peak.cu -o /tmp/fp64_peak && /tmp/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
  double* out;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 64 * iters * (double)blocks * threads;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dfma_per_clk_per_sm\": %.1f, \"clock_mhz\": %d}\n", p.name, p.multiProcessorCount, best,
         best * 1e12 / 2 / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000);
  return 0;
}
