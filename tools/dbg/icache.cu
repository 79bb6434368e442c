// instruction-cache capacity probe: straight-line code of K independent FFMAs executed in a loop by one warp
#include <cstdio>
#include <cuda_runtime.h>
#define F1 asm volatile("fma.rn.f32 %0, %0, %4, %5;\n\tfma.rn.f32 %1, %1, %4, %5;\n\tfma.rn.f32 %2, %2, %4, %5;\n\tfma.rn.f32 %3, %3, %4, %5;" : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3) : "f"(a), "f"(b));
#define F4 F1 F1 F1 F1
#define F16 F4 F4 F4 F4
#define F64 F16 F16 F16 F16
#define F256 F64 F64 F64 F64
#define F1K F256 F256 F256 F256
#define BODY(NAME, CODE) \
__global__ void NAME(float* out, long long* cyc, int iters, float a, float b) { \
  float x0 = a, x1 = b, x2 = a + b, x3 = a - b; \
  long long t0 = 0; \
  for (int it = 0; it < iters + 1; ++it) { if (it == 1) t0 = clock64(); CODE } \
  long long t1 = clock64(); \
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0; \
  out[threadIdx.x] = x0 + x1 + x2 + x3; }
BODY(k1, F256)                       // 1k instr = 16 KB
BODY(k2, F256 F256)                  // 32 KB
BODY(k4, F1K)                        // 64 KB
BODY(k6, F1K F256 F256)              // 96 KB
BODY(k8, F1K F1K)                    // 128 KB
BODY(k12, F1K F1K F1K)               // 192 KB
BODY(k16, F1K F1K F1K F1K)           // 256 KB
BODY(k32, F1K F1K F1K F1K F1K F1K F1K F1K)   // 512 KB
int main() {
  float* out; long long* cyc; cudaMallocManaged(&out, 4096); cudaMallocManaged(&cyc, 8 * 1024);
  struct { const char* n; void (*k)(float*, long long*, int, float, float); int ninstr; } ks[] = {
    {"16 KB", k1, 1024}, {"32 KB", k2, 2048}, {"64 KB", k4, 4096}, {"96 KB", k6, 6144}, {"128 KB", k8, 8192}, {"192 KB", k12, 12288}, {"256 KB", k16, 16384}, {"512 KB", k32, 32768}};
  for (auto& e : ks) {
    for (int grid : {1, 148}) {
      e.k<<<grid, 32>>>(out, cyc, 20, 1.0001f, 0.5f); cudaDeviceSynchronize();
      e.k<<<grid, 32>>>(out, cyc, 20, 1.0001f, 0.5f); cudaDeviceSynchronize();
      printf("%-7s grid %3d x 32 thr: %.2f cycles/instr\n", e.n, grid, double(cyc[0]) / 20 / e.ninstr);
    }
  }
  // two different kernels' worth of code on one SM: 2 warps in one CTA running different halves is not possible here; instead 4 warps same code
  for (auto& e : ks) { e.k<<<148, 128>>>(out, cyc, 20, 1.0001f, 0.5f); cudaDeviceSynchronize(); printf("%-7s grid 148 x 128 thr: %.2f cycles/instr per warp\n", e.n, double(cyc[0]) / 20 / e.ninstr); }
  return 0;
}
