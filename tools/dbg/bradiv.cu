// cost of the convergence check (BRA.DIV) ptxas puts in front of shfl.sync when it cannot prove the warp converged
#include <cstdio>
#include <cuda_runtime.h>
__device__ __noinline__ double chain_noinline(double x, int n) {
  for (int i = 0; i < n; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  return x;
}
__global__ void k_top(double* out, long long* cyc, int n) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x;
}
__global__ void k_fn(double* out, long long* cyc, int n) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
  x = chain_noinline(x, n);
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  out[threadIdx.x] = x;
}
__global__ void k_branch(double* out, long long* cyc, int n) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
  if ((threadIdx.x >> 5) == (unsigned)(n & 0)) {   // warp-uniform in fact, unknown to the compiler
    for (int i = 0; i < n; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  out[threadIdx.x] = x;
}
int main() {
  double* out; long long* cyc; cudaMallocManaged(&out, 4096); cudaMallocManaged(&cyc, 64);
  for (int i = 0; i < 256; ++i) out[i] = i;
  const int n = 1024;
  for (int rep = 0; rep < 2; ++rep) {
    k_top<<<1, 32>>>(out, cyc, n); k_fn<<<1, 32>>>(out, cyc, n); k_branch<<<1, 64>>>(out, cyc, n); cudaDeviceSynchronize();
  }
  printf("shfl.f64+dadd per step: top-level %.1f  noinline fn %.1f  under tid-branch %.1f cycles\n", double(cyc[0]) / n, double(cyc[1]) / n, double(cyc[2]) / n);
  return 0;
}
