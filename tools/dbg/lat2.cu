// dependent-chain latencies with runtime operands (no constant folding)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, const double* in, int n) {
  double x = in[0], y = in[1], z = in[2];
  const int t = threadIdx.x;
  long long t0, t1;
  auto rec = [&](int id) { if (t == 0) out[id] = double(t1 - t0) / n; };
  t0 = clock64(); for (int i = 0; i < n; ++i) x = x + y; t1 = clock64(); rec(0);
  t0 = clock64(); for (int i = 0; i < n; ++i) x = x * z; t1 = clock64(); rec(1);
  t0 = clock64(); for (int i = 0; i < n; ++i) x = fma(x, z, y); t1 = clock64(); rec(2);
  t0 = clock64(); for (int i = 0; i < n; ++i) x = x * z + y; t1 = clock64(); rec(3);   // mul+add (fmad=false)
  t0 = clock64(); for (int i = 0; i < n; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + y; t1 = clock64(); rec(4);
  float f = (float)x, g = (float)y;
  t0 = clock64(); for (int i = 0; i < n; ++i) f = f + g; t1 = clock64(); rec(5);
  int k = (int)y + t;
  t0 = clock64(); for (int i = 0; i < n; ++i) k = __shfl_xor_sync(0xffffffffu, k, 1) + 1; t1 = clock64(); rec(6);
  t0 = clock64(); for (int i = 0; i < n; ++i) x = 1.0 / x + y; t1 = clock64(); rec(7);
  t0 = clock64(); for (int i = 0; i < n; ++i) x = rsqrt(x) + y; t1 = clock64(); rec(8);
  t0 = clock64(); for (int i = 0; i < n; ++i) x = (x > y) ? x - z : x + z; t1 = clock64(); rec(9);   // compare+select chain
  unsigned long long u = (unsigned long long)k;
  t0 = clock64(); for (int i = 0; i < n; ++i) u = (u * 0xD2511F53ull) >> 11; t1 = clock64(); rec(10);
  out[32 + t] = x + f + k + (double)u;
}
int main() {
  double *out, *in; cudaMallocManaged(&out, 4096); cudaMallocManaged(&in, 64);
  in[0] = 1.000001; in[1] = 1e-9; in[2] = 0.9999999;
  const char* names[] = {"dadd dep", "dmul dep", "dfma dep", "dmul+dadd dep", "shfl.f64 + dadd dep", "fadd dep", "shfl.i32 + iadd dep", "1/x + dadd dep", "rsqrt + dadd dep", "dsetp+sel chain", "imad.wide+shift"};
  for (int threads : {32, 128, 256}) {
    lat<<<1, threads>>>(out, in, 512); cudaDeviceSynchronize();
    lat<<<1, threads>>>(out, in, 512); cudaDeviceSynchronize();
    printf("threads=%d:", threads);
    for (int i = 0; i < 11; ++i) printf("  %s %.1f |", names[i], out[i]);
    printf("\n");
  }
  lat<<<296, 256>>>(out, in, 512); cudaDeviceSynchronize();
  printf("grid 296x256:");
  for (int i = 0; i < 11; ++i) printf("  %s %.1f |", names[i], out[i]);
  printf("\n");
}
