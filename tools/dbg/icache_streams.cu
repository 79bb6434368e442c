// instruction-cache behaviour with SEVERAL independent instruction streams per SM: warp w of the CTA loops over its own
// straight-line region of S bytes.  Prints cycles per instruction per warp for streams x region size.
#include <cstdio>
#include <cuda_runtime.h>
#define F1 asm volatile("fma.rn.f32 %0, %0, %4, %5;\n\tfma.rn.f32 %1, %1, %4, %5;\n\tfma.rn.f32 %2, %2, %4, %5;\n\tfma.rn.f32 %3, %3, %4, %5;" : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3) : "f"(a), "f"(b));
#define F4 F1 F1 F1 F1
#define F16 F4 F4 F4 F4
#define F32 F16 F16            /* 128 instructions = 2 KB */
template <int ID, int REP>
__device__ __noinline__ float body(float a, float b, int iters, long long* cyc) {
  float x0 = a + ID, x1 = b, x2 = a + b, x3 = a - b;
  long long t0 = 0;
  for (int it = 0; it < iters + 1; ++it) {
    if (it == 1) t0 = clock64();
#pragma unroll
    for (int r = 0; r < REP; ++r) { F32 }
  }
  *cyc = clock64() - t0;
  return x0 + x1 + x2 + x3;
}
template <int REP>
__global__ void k(float* out, long long* cyc, int iters, float a, float b, int same) {
  const int w = same ? 0 : threadIdx.x >> 5;
  float r = 0; long long c = 0;
  switch (w) {
#define C(i) case i: r = body<i, REP>(a, b, iters, &c); break;
    C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
  }
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 16 + (threadIdx.x >> 5)] = c;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int REP>
void run(float* out, long long* cyc) {
  for (int same = 0; same < 2; ++same)
    for (int streams : {1, 2, 4, 8, 16}) {
      const int iters = 4096 / REP;
      for (int rep = 0; rep < 2; ++rep) { k<REP><<<148, 32 * streams>>>(out, cyc, iters, 1.0001f, 0.5f, same); cudaDeviceSynchronize(); }
      double s = 0;
      for (int w = 0; w < streams; ++w) s += double(cyc[w]);
      printf("region %3d KB x %2d warps (%s): total %4d KB, %.2f cycles/instr per warp, %.2f IPC per SM\n", 2 * REP, streams,
             same ? "same code" : "own code ", same ? 2 * REP : 2 * REP * streams, s / streams / iters / (128.0 * REP),
             streams / (s / streams / iters / (128.0 * REP)));
    }
}
int main() {
  float* out; long long* cyc; cudaMallocManaged(&out, 148 * 512 * 4); cudaMallocManaged(&cyc, 148 * 16 * 8);
  run<1>(out, cyc); run<2>(out, cyc); run<4>(out, cyc); run<8>(out, cyc); run<16>(out, cyc);
  return 0;
}
