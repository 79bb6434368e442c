// latency microbenchmarks (cycles) on the target GPU: dependent chains of the primitive ops the replay path is made of
#include <cstdio>
#include <cuda_runtime.h>
__device__ double g_sink;
__global__ void lat(double* out, const double* gbuf, int* idx, double x0) {
  __shared__ double sm[512];
  const int t = threadIdx.x;
  sm[t] = x0 + t; sm[t + 256] = x0;
  __syncthreads();
  double x = x0, y = x0 * 0.5 + 1.0;
  long long t0, t1; int n = 256;
  auto rec = [&](int id) { if (t == 0) out[id] = double(t1 - t0) / n; };
  t0 = clock64(); for (int i = 0; i < n; ++i) x = x + y; t1 = clock64(); rec(0);
  t0 = clock64(); for (int i = 0; i < n; ++i) x = x * y; t1 = clock64(); rec(1);
  t0 = clock64(); for (int i = 0; i < n; ++i) x = fma(x, y, y); t1 = clock64(); rec(2);
  x = 1.7 + x0;
  t0 = clock64(); for (int i = 0; i < n; ++i) x = y / x; t1 = clock64(); rec(3);
  t0 = clock64(); for (int i = 0; i < n; ++i) x = sqrt(x + 2.0); t1 = clock64(); rec(4);
  t0 = clock64(); for (int i = 0; i < n; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1); t1 = clock64(); rec(5);
  int j = t & 7;
  t0 = clock64(); for (int i = 0; i < n; ++i) j = idx[j]; t1 = clock64(); rec(6);     // dependent global (L1 hit)
  x += j;
  t0 = clock64(); for (int i = 0; i < n; ++i) { __syncthreads(); } t1 = clock64(); rec(7);
  int k = t & 7;
  t0 = clock64(); for (int i = 0; i < n; ++i) k = (int)sm[k & 255] & 255; t1 = clock64(); rec(8);   // smem dependent
  x += k;
  float f = (float)x0;
  t0 = clock64(); for (int i = 0; i < n; ++i) f = f * f + 1.0f; t1 = clock64(); rec(9);
  unsigned h = (unsigned)x0;
  t0 = clock64(); for (int i = 0; i < n; ++i) { h += i; h = (h << 16) ^ ((i << 11) ^ h); h += h >> 11; } t1 = clock64(); rec(10);
  // independent FP64 throughput: 8 chains
  double a0 = x0, a1 = x0 + 1, a2 = x0 + 2, a3 = x0 + 3, a4 = x0 + 4, a5 = x0 + 5, a6 = x0 + 6, a7 = x0 + 7;
  t0 = clock64(); for (int i = 0; i < n; ++i) { a0 = a0 * y + y; a1 = a1 * y + y; a2 = a2 * y + y; a3 = a3 * y + y; a4 = a4 * y + y; a5 = a5 * y + y; a6 = a6 * y + y; a7 = a7 * y + y; } t1 = clock64(); rec(11);
  // ld.global.cg dependent (L2 hit)
  int m = t & 7;
  t0 = clock64(); for (int i = 0; i < n; ++i) m = __ldcg(idx + m); t1 = clock64(); rec(12);
  double r = 1.0 / (x0 + 3.0);
  t0 = clock64(); for (int i = 0; i < n; ++i) r = rsqrt(r + 2.0); t1 = clock64(); rec(13);
  g_sink = x + f + h + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + m + r;
}
int main() {
  double* out; int* idx; cudaMallocManaged(&out, 256); cudaMallocManaged(&idx, 1024);
  for (int i = 0; i < 256; ++i) idx[i] = (i * 5 + 3) & 7;
  const char* names[] = {"dadd dep", "dmul dep", "dfma dep", "ddiv dep", "dsqrt dep", "shfl f64 dep", "ld.global dep (L1)", "__syncthreads", "ld.shared dep (+cvt)", "ffma dep", "hash step", "8 indep dfma (per 8)", "ld.cg dep (L2)", "drsqrt dep"};
  for (int threads : {32, 256}) {
    lat<<<1, threads>>>(out, nullptr, idx, 1.25); cudaDeviceSynchronize();
    lat<<<1, threads>>>(out, nullptr, idx, 1.25); cudaDeviceSynchronize();
    printf("threads=%d\n", threads);
    for (int i = 0; i < 14; ++i) printf("  %-24s %.1f cycles\n", names[i], out[i]);
  }
  // all SMs busy x 2 CTAs, 256 threads
  lat<<<296, 256>>>(out, nullptr, idx, 1.25); cudaDeviceSynchronize();
  printf("grid=296x256\n");
  for (int i = 0; i < 14; ++i) printf("  %-24s %.1f cycles\n", names[i], out[i]);
  return 0;
}
