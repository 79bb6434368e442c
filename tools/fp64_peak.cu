// fp64_peak.cu -- measured FP64 CUDA-core peak of this GPU (the denominator of bench.py's "fp64" roofline object;
// "confirm with an FP64-FMA microbenchmark and use the measured peak").
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_prof/fp64_peak tools/fp64_peak.cu
//   run  : tools/_prof/fp64_peak > profiles/fp64_peak.json        (one JSON object)
// Three rates: DFMA (2 flop / instruction, the datasheet convention), DMUL+DADD pairs (what a -fmad=false build like
// this engine's can issue at best: 1 flop / instruction), and dependent-chain latency of DFMA.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: fma, 1: mul+add (not contracted)
__global__ void __launch_bounds__(256) burn(double* out, int iters, double a, double b) {
  double x0 = a + threadIdx.x, x1 = a - threadIdx.x, x2 = b + threadIdx.x, x3 = b - threadIdx.x;
  double x4 = a * 0.5, x5 = b * 0.5, x6 = a * 0.25, x7 = b * 0.25;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
      } else {
        x0 = __dadd_rn(__dmul_rn(x0, a), b); x1 = __dadd_rn(__dmul_rn(x1, a), b);
        x2 = __dadd_rn(__dmul_rn(x2, a), b); x3 = __dadd_rn(__dmul_rn(x3, a), b);
        x4 = __dadd_rn(__dmul_rn(x4, a), b); x5 = __dadd_rn(__dmul_rn(x5, a), b);
        x6 = __dadd_rn(__dmul_rn(x6, a), b); x7 = __dadd_rn(__dmul_rn(x7, a), b);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
  double x = a;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 32; ++u) x = fma(x, a, b);
  }
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
double rate(int grid, int iters, double* out) {   // instructions of the measured kind per second, best of 5
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int r = 0; r < 6; ++r) {
    cudaEventRecord(e0);
    burn<MODE><<<grid, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * 256.0 * iters * 16.0 * 8.0;   // FMA (or MUL+ADD pairs) executed
    if (r > 0 && ops / (ms * 1e-3) > best) best = ops / (ms * 1e-3);
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
  double* out; long long* cyc;
  const int grid = p.multiProcessorCount * 8;
  cudaMalloc(&out, sizeof(double) * grid * 256);
  cudaMallocManaged(&cyc, sizeof(long long));
  const double fma_rate = rate<0>(grid, 4096, out);
  const double pair_rate = rate<1>(grid, 4096, out);
  chain<<<1, 32>>>(out, cyc, 2000, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz_max\": %d, "
         "\"dfma_tflops\": %.3f, \"dfma_per_clk_per_sm\": %.2f, "
         "\"dmul_dadd_tflops\": %.3f, \"dmul_dadd_instr_per_clk_per_sm\": %.2f, "
         "\"dfma_dependent_latency_cycles\": %.2f, "
         "\"how\": \"8 independent chains per thread, 256 threads x 8 CTAs per SM, 16x unrolled, best of 5 (CUDA events); "
         "dfma_tflops counts 2 flop per DFMA, dmul_dadd_tflops counts 2 flop per DMUL+DADD pair (the -fmad=false rate)\"}\n",
         p.name, p.multiProcessorCount, clk, fma_rate * 2 / 1e12, fma_rate / (clk * 1e3) / p.multiProcessorCount,
         pair_rate * 2 / 1e12, pair_rate * 2 / (clk * 1e3) / p.multiProcessorCount, (double)*cyc / (2000.0 * 32.0));
  return 0;
}
