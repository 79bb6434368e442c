// cycles per call of the warp-level LA kernels on real matrices (corpus dumped by the host emulation)
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../../pydegensac_b200/csrc/warpla.h"
using namespace dg;
__global__ void bench(const double* mats, int n, double* out, long long* cyc, int mode) {
  __shared__ WarpScratch ws;
  const int lane = threadIdx.x & 31;
  long long total = 0;
  if (threadIdx.x < 32) {
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
      for (int t = lane; t < 81; t += 32) ws.A[t] = mats[(size_t)i * 81 + t];
      __syncwarp();
      const long long t0 = clock64();
      if (mode == 0) warp_smallest_eigvec9(&ws, lane, 32);
      else if (mode == 1) { warp_jacobi_eig9(&ws, lane, 32); }
      else if (mode == 2) { warp_null_8x9(&ws, lane, 32); }
      else if (mode == 3) { if (lane == 0) { double q[9]; for (int j = 0; j < 9; ++j) q[j] = ws.A[j]; enforce_rank2(q); for (int j = 0; j < 9; ++j) ws.cs[j] = q[j]; } }
      __syncwarp();
      total += clock64() - t0;
      if (lane < 9) out[(size_t)i * 9 + lane] = ws.cs[lane];
      __syncwarp();
    }
    if (lane == 0) atomicAdd((unsigned long long*)cyc, (unsigned long long)total);
  }
}
int main(int argc, char** argv) {
  FILE* fp = fopen(argc > 1 ? argv[1] : "tools/_prof/eig_dump.bin", "rb");
  if (!fp) { printf("no corpus\n"); return 1; }
  std::vector<double> h; double buf[81];
  while (fread(buf, sizeof(double), 81, fp) == 81) h.insert(h.end(), buf, buf + 81);
  fclose(fp);
  const int n = (int)(h.size() / 81);
  double *d, *o; long long* c;
  cudaMalloc(&d, h.size() * 8); cudaMalloc(&o, (size_t)n * 9 * 8); cudaMallocManaged(&c, 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  const char* names[] = {"smallest_eigvec9", "jacobi_eig9", "null_8x9 (first 72 entries as system)", "enforce_rank2 (first 9 entries)"};
  for (int mode = 0; mode < 4; ++mode)
    for (int cfg = 0; cfg < 2; ++cfg) {
      const int grid = cfg ? 296 : 1, threads = cfg ? 256 : 32;
      *c = 0;
      bench<<<grid, threads>>>(d, n, o, c, mode); cudaDeviceSynchronize();
      *c = 0;
      bench<<<grid, threads>>>(d, n, o, c, mode); cudaDeviceSynchronize();
      printf("%-42s grid %3d x %3d : %.0f cycles/call  (%s)\n", names[mode], grid, threads, double(*c) / n, cudaGetErrorString(cudaGetLastError()));
    }
  // accuracy of mode 0 against the residual
  bench<<<296, 256>>>(d, n, o, c, 0); cudaDeviceSynchronize();
  std::vector<double> x((size_t)n * 9); cudaMemcpy(x.data(), o, x.size() * 8, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int i = 0; i < n; ++i) {
    const double* A = &h[(size_t)i * 81]; const double* v = &x[(size_t)i * 9];
    double r[9], rho = 0, fro = 0, r2 = 0;
    for (int a = 0; a < 9; ++a) { r[a] = 0; for (int b = 0; b < 9; ++b) { r[a] += A[a * 9 + b] * v[b]; fro += A[a * 9 + b] * A[a * 9 + b]; } rho += r[a] * v[a]; }
    for (int a = 0; a < 9; ++a) r2 += (r[a] - rho * v[a]) * (r[a] - rho * v[a]);
    worst = fmax(worst, sqrt(r2 / fro));
  }
  printf("worst relative residual %.3g over %d matrices\n", worst, n);
  return 0;
}
