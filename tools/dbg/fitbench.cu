// stand-alone cycles per call of the block-level fit routines (one CTA of 256 threads, warp 0 works, 7 warps wait)
#include <cstdio>
#include <vector>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../pydegensac_b200/csrc/ffit.h"
using namespace dg;
__global__ void bench(const double* pts, int N, int* lists, double* wts, double* out, long long* cyc, int iters, int mode) {
  __shared__ BlockScratch sc;
  Ctx c;
  c.tid = threadIdx.x; c.nt = blockDim.x; c.lane = threadIdx.x & 31; c.wid = threadIdx.x >> 5; c.nw = blockDim.x >> 5;
  c.N = N; c.x1 = pts; c.y1 = pts + N; c.x2 = pts + 2 * N; c.y2 = pts + 3 * N; c.sc = &sc; c.t32 = nullptr;
  int* list = lists + (size_t)blockIdx.x * 1024;
  for (int i = threadIdx.x; i < 600; i += blockDim.x) list[i] = (i * 3 + blockIdx.x) % N;
  __syncthreads();
  DrawCursor cur; cur.seed = 1234 + blockIdx.x; cur.k = 7; cur.j = 8;
  double f[9], acc = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) blk_sample8_fit_F(c, list, 600, wts, cur, f);
    else if (mode == 1) { blk_randsubset(c, list, 600, 14, cur); blk_fit_F(c, list + 600 - 14, 14, nullptr, f); }
    else if (mode == 2) { blk_fit_F(c, list, 600, nullptr, f); }
    acc += f[0];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[blockIdx.x] = (t1 - t0) / iters; out[blockIdx.x] = acc; }
}
int main() {
  const int N = 2000;
  std::vector<double> h(4 * N), w(N);
  srand(3);
  for (int i = 0; i < N; ++i) { double x = rand() % 640, y = rand() % 480; h[i] = x; h[N + i] = y; h[2 * N + i] = x + 5 + (rand() % 100) * 0.01; h[3 * N + i] = y - 3 + (rand() % 100) * 0.01; w[i] = 0.5 + (rand() % 100) * 0.01; }
  double *d, *dw, *out; int* lists; long long* cyc;
  cudaMalloc(&d, h.size() * 8); cudaMalloc(&dw, N * 8); cudaMallocManaged(&out, 296 * 8); cudaMalloc(&lists, 296 * 1024 * 4); cudaMallocManaged(&cyc, 296 * 8);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), N * 8, cudaMemcpyHostToDevice);
  const char* names[] = {"sample8+fit8 (weighted)", "randsubset14 + 14-pt eig fit", "600-pt big fit"};
  for (int mode = 0; mode < 3; ++mode)
    for (int grid : {1, 148, 296}) {
      bench<<<grid, 256>>>(d, N, lists, dw, out, cyc, 50, mode); cudaDeviceSynchronize();
      bench<<<grid, 256>>>(d, N, lists, dw, out, cyc, 50, mode); cudaDeviceSynchronize();
      printf("%-30s grid %3d x 256: %lld cycles/call (%s)\n", names[mode], grid, cyc[0], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
