// frontend.cu -- the steps either side of the RANSAC path, device resident:
//
//   BEFORE  descriptor matching: brute-force L2 two-nearest-neighbour search + SNN ratio test (+ optional mutual
//           check) producing the [N,2] correspondence arrays the RANSAC kernel consumes -- what the reference's
//           pipeline does on the host with cv2.BFMatcher.knnMatch(k=2) and `m.distance < 0.9 * n.distance`
//           (examples/simple-example.py:46-53), so that src_pts/dst_pts never visit the host.
//           The train descriptors are streamed through shared memory with Blackwell/Hopper BULK ASYNC COPIES
//           (cp.async.bulk global->shared, completion on an mbarrier; UBLKCP in SASS), double buffered behind the math.
//   AFTER   pose from a fundamental matrix: E = K2^T F K1, the four (R, t) candidates, cheirality vote over the
//           inliers (one warp per pair);  the reference stops at F (the survey lists E/pose as the step after).
//   ALSO    the reference's alternative 7-point null-space solver nullspace_qr7x9 (Ftools.c:594, USE_QR) as a batch
//           kernel over many samples (la.h restates LAPACK's dgeqp3 kernel).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/degensac_b200.h"
#include "common.h"
#include "la.h"
#include "hgeom.h"
#include "rng.h"

namespace {

thread_local char g_err2[256] = "";
int fail2(int code, const char* what, cudaError_t e = cudaSuccess) {
  if (e != cudaSuccess) snprintf(g_err2, sizeof(g_err2), "%s: %s", what, cudaGetErrorString(e));
  else snprintf(g_err2, sizeof(g_err2), "%s", what);
  return code;
}
#define CU2(call)                                                     \
  do {                                                                \
    cudaError_t e__ = (call);                                         \
    if (e__ != cudaSuccess) return fail2(DGB200_E_CUDA, #call, e__);  \
  } while (0)

// ------------------------------------------------------------------------------------------------ matcher
constexpr int kTQ = 32;          // query descriptors per CTA
constexpr int kTT = 64;          // train descriptors per shared-memory stage
constexpr int kMatchThreads = 128;   // 32 queries x 4 train groups

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
// 1-D bulk async copy global -> shared (TMA engine, no tensor map needed): size and both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct MatchArgs {
  const float* d1; const float* d2;   // [n1][D], [n2][D]
  int n1, n2, D;
  int use_bulk;                       // train rows are 16-byte aligned: stage them with cp.async.bulk
  int mutual;                         // also track, per train descriptor, its nearest query (col_best)
  int* nn_idx; float* nn_d1; float* nn_d2;          // per query: nearest train index, squared distances of the two nearest
  unsigned long long* col_best;       // per train descriptor: (squared distance bits << 32 | query index), atomicMin
};

// Two nearest train descriptors of every query descriptor (squared L2 in FP32, accumulated in descriptor order).
__global__ void __launch_bounds__(kMatchThreads) nn2_kernel(MatchArgs a) {
  extern __shared__ __align__(128) unsigned char msm[];
  __shared__ __align__(8) uint64_t bar[2];
  const int D = a.D, qs = D + 4;                        // padded query row: conflict-free 128-bit reads across queries
  float* tq = reinterpret_cast<float*>(msm);            // [kTQ][D + 4]
  float* tt = tq + kTQ * qs;                             // 2 stages x [kTT][D]
  const size_t stage_floats = (size_t)kTT * D;
  const int tid = threadIdx.x, q = tid & 31, g = tid >> 5;
  const int q0 = blockIdx.x * kTQ;
  for (int i = tid; i < kTQ * D; i += kMatchThreads) {
    const int r = i / D, c = i % D;
    tq[r * qs + c] = (q0 + r < a.n1) ? a.d1[(size_t)(q0 + r) * D + c] : 0.f;
  }
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int ntiles = (a.n2 + kTT - 1) / kTT;
  auto issue = [&](int t) {          // stage tile t (one elected thread; or cooperative loads when unaligned)
    const int rows = min(kTT, a.n2 - t * kTT);
    float* dst = tt + (size_t)(t & 1) * stage_floats;
    if (a.use_bulk) {
      if (tid == 0) {
        const uint32_t bytes = (uint32_t)(rows * D * sizeof(float));
        mbar_expect_tx(&bar[t & 1], bytes);
        bulk_g2s(dst, a.d2 + (size_t)t * kTT * D, bytes, &bar[t & 1]);
      }
    } else {
      const float* src = a.d2 + (size_t)t * kTT * D;
      for (int i = tid; i < rows * D; i += kMatchThreads) dst[i] = src[i];
    }
  };
  float b1 = INFINITY, b2 = INFINITY;
  int i1 = -1;
  if (ntiles > 0) issue(0);
  for (int t = 0; t < ntiles; ++t) {
    if (t + 1 < ntiles) issue(t + 1);                 // next tile in flight behind this one's math
    if (a.use_bulk) mbar_wait(&bar[t & 1], (uint32_t)((t >> 1) & 1)); else __syncthreads();
    const int rows = min(kTT, a.n2 - t * kTT);
    const float* B = tt + (size_t)(t & 1) * stage_floats;
    const float* A = tq + q * qs;
    // this thread: train rows g*16 .. g*16+15 of the tile, four at a time (query chunk loaded once per four rows)
    for (int jb = g * 16; jb < g * 16 + 16 && jb < rows; jb += 4) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      const float* B0 = B + (size_t)jb * D;
      const int r1 = (jb + 1 < rows) ? 1 : 0, r2 = (jb + 2 < rows) ? 2 : 0, r3 = (jb + 3 < rows) ? 3 : 0;
      for (int k = 0; k < D; k += 4) {
        const float4 av = *reinterpret_cast<const float4*>(A + k);
        const float4 v0 = *reinterpret_cast<const float4*>(B0 + k);
        const float4 v1 = *reinterpret_cast<const float4*>(B0 + (size_t)r1 * D + k);
        const float4 v2 = *reinterpret_cast<const float4*>(B0 + (size_t)r2 * D + k);
        const float4 v3 = *reinterpret_cast<const float4*>(B0 + (size_t)r3 * D + k);
        float d;
        d = av.x - v0.x; s0 = fmaf(d, d, s0); d = av.y - v0.y; s0 = fmaf(d, d, s0); d = av.z - v0.z; s0 = fmaf(d, d, s0); d = av.w - v0.w; s0 = fmaf(d, d, s0);
        d = av.x - v1.x; s1 = fmaf(d, d, s1); d = av.y - v1.y; s1 = fmaf(d, d, s1); d = av.z - v1.z; s1 = fmaf(d, d, s1); d = av.w - v1.w; s1 = fmaf(d, d, s1);
        d = av.x - v2.x; s2 = fmaf(d, d, s2); d = av.y - v2.y; s2 = fmaf(d, d, s2); d = av.z - v2.z; s2 = fmaf(d, d, s2); d = av.w - v2.w; s2 = fmaf(d, d, s2);
        d = av.x - v3.x; s3 = fmaf(d, d, s3); d = av.y - v3.y; s3 = fmaf(d, d, s3); d = av.z - v3.z; s3 = fmaf(d, d, s3); d = av.w - v3.w; s3 = fmaf(d, d, s3);
      }
      const float sv[4] = {s0, s1, s2, s3};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = jb + u;
        if (j < rows) {                                // (warp-uniform: g and rows are)
          const float s = sv[u];
          const int gj = t * kTT + j;
          if (s < b1 || (s == b1 && gj < i1)) { b2 = b1; b1 = s; i1 = gj; }
          else if (s < b2) b2 = s;
          if (a.mutual) {
            // nearest QUERY of train descriptor gj among this warp's 32 queries, then one atomic per warp and row
            unsigned long long key = (q0 + q < a.n1) ? (((unsigned long long)__float_as_uint(s) << 32) | (unsigned)(q0 + q))
                                                      : 0xffffffffffffffffull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
              key = other < key ? other : key;
            }
            if (q == 0) atomicMin(a.col_best + gj, key);
          }
        }
      }
    }
    __syncthreads();    // everybody is done with stage t&1 before tile t+2 overwrites it
  }
  // merge the four train groups of each query (shared memory reuse: the query tile is dead)
  float* m1 = reinterpret_cast<float*>(msm);
  float* m2 = m1 + kMatchThreads;
  int* mi = reinterpret_cast<int*>(m2 + kMatchThreads);
  m1[tid] = b1; m2[tid] = b2; mi[tid] = i1;
  __syncthreads();
  if (g == 0 && q0 + q < a.n1) {
    float c1 = INFINITY, c2 = INFINITY;
    int ci = -1;
    for (int gg = 0; gg < 4; ++gg) {
      const float x1 = m1[gg * 32 + q], x2 = m2[gg * 32 + q];
      const int xi = mi[gg * 32 + q];
      if (xi >= 0 && (x1 < c1 || (x1 == c1 && xi < ci))) { c2 = c1; c1 = x1; ci = xi; }
      else if (x1 < c2) c2 = x1;
      if (x2 < c2) c2 = x2;
    }
    a.nn_idx[q0 + q] = ci; a.nn_d1[q0 + q] = c1; a.nn_d2[q0 + q] = c2;
  }
}

// SNN ratio test (+ mutual check) and ORDERED compaction of the accepted matches; gathers the keypoint coordinates
// into the [N,2] float64 arrays the RANSAC entry points take.  One CTA (n1 is a few thousand).
__global__ void __launch_bounds__(1024) select_kernel(int n1, const int* nn_idx, const float* nn_d1, const float* nn_d2,
                                                      const unsigned long long* col_best, float ratio2, int mutual,
                                                      const double* kp1, const double* kp2, int kp_dim, int* match_q,
                                                      int* match_t, double* x1y1, double* x2y2, int out_dim, int cap,
                                                      int* count) {
  __shared__ int wsum[32];
  __shared__ int base;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  for (int q0 = 0; q0 < n1; q0 += 1024) {
    const int q = q0 + tid;
    bool ok = false;
    int j = -1;
    if (q < n1) {
      j = nn_idx[q];
      ok = j >= 0 && nn_d1[q] < ratio2 * nn_d2[q];          // d1 < ratio * d2 on L2 distances (simple-example.py:50)
      if (ok && mutual) ok = (unsigned)(col_best[j] & 0xffffffffu) == (unsigned)q;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) wsum[wid] = __popc(m);
    __syncthreads();
    int off = base;
    for (int w = 0; w < wid; ++w) off += wsum[w];
    off += __popc(m & ((1u << lane) - 1u));
    if (ok && off < cap) {
      match_q[off] = q; match_t[off] = j;
      if (x1y1) {
        for (int c = 0; c < out_dim; ++c) {
          x1y1[(size_t)off * out_dim + c] = kp1[(size_t)q * kp_dim + c];
          x2y2[(size_t)off * out_dim + c] = kp2[(size_t)j * kp_dim + c];
        }
      }
    }
    __syncthreads();
    if (tid == 0) { int tot = 0; for (int w = 0; w < 32; ++w) tot += wsum[w]; base += tot; }
    __syncthreads();
  }
  if (tid == 0) *count = base < cap ? base : cap;
}

// ------------------------------------------------------------------------------------------------ pose from F
struct PoseArgs {
  const double* F; const double* K1; const double* K2;   // [P][9] row-major; K: [9] shared or [P][9]
  int k_per_pair;
  const double* x1y1; const double* x2y2; const unsigned char* mask;   // [P][n][dim], [P][n]
  int n_pairs, n, dim;
  double* R; double* t; int* good;   // [P][9], [P][3], [P]
};

__device__ inline void mat3_mul(const double* A, const double* B, double* C) {   // row-major C = A B
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
__device__ inline double det3r(const double* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}

// One warp per pair.  Lane 0 factors E (one-sided Jacobi SVD, la.h) and builds the four candidates; all lanes vote:
// a correspondence supports (R, t) when its two-view triangulation lambda1 R x1 + t = lambda2 x2 (least squares in
// lambda) has both depths positive.  Result: the candidate with the most supporters (first in the order
// (R1,+t), (R1,-t), (R2,+t), (R2,-t) on ties), t of unit length.
__global__ void pose_kernel(PoseArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.n_pairs) return;
  const int p = warp;
  __shared__ double sh[8][4 * 12];        // per warp: 4 candidates x (R[9], t[3])
  double* cand = sh[(threadIdx.x >> 5) & 7];
  const double* K1 = a.K1 + (a.k_per_pair ? (size_t)p * 9 : 0);
  const double* K2 = a.K2 + (a.k_per_pair ? (size_t)p * 9 : 0);
  double k1i[9], k2i[9];
  {   // inverse of an upper-triangular-ish calibration matrix: general 3x3 inverse
    double m[9];
    for (int i = 0; i < 9; ++i) m[i] = K1[i];
    dg::inv3(m);
    for (int i = 0; i < 9; ++i) k1i[i] = m[i];
    for (int i = 0; i < 9; ++i) m[i] = K2[i];
    dg::inv3(m);
    for (int i = 0; i < 9; ++i) k2i[i] = m[i];
  }
  if (lane == 0) {
    const double* F = a.F + (size_t)p * 9;
    double K2t[9], tmp[9], E[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) K2t[3 * i + j] = K2[3 * j + i];
    mat3_mul(K2t, F, tmp);
    mat3_mul(tmp, K1, E);
    // E = U S V^T with G = E V (orthogonal columns, norms = singular values)
    double G[9], V[9], sv[3];
    dg::svd3_onesided(E, G, V, sv);
    int o[3] = {0, 1, 2};                   // sort singular values, largest first
    for (int i = 0; i < 3; ++i) for (int j = i + 1; j < 3; ++j) if (sv[o[j]] > sv[o[i]]) { const int tt = o[i]; o[i] = o[j]; o[j] = tt; }
    double U[9], Vs[9];
    for (int c = 0; c < 2; ++c) {
      const int k = o[c];
      const double inv = sv[k] > 0 ? 1.0 / sv[k] : 0.0;
      for (int r = 0; r < 3; ++r) { U[3 * r + c] = G[3 * r + k] * inv; Vs[3 * r + c] = V[3 * r + k]; }
    }
    {  // third columns: cross products (unit, orthogonal), so that det(U) = det(V) = +1
      const double u0[3] = {U[0], U[3], U[6]}, u1[3] = {U[1], U[4], U[7]}, v0[3] = {Vs[0], Vs[3], Vs[6]}, v1[3] = {Vs[1], Vs[4], Vs[7]};
      double cu[3], cv[3];
      dg::cross3(cu, u0, u1); dg::cross3(cv, v0, v1);
      for (int r = 0; r < 3; ++r) { U[3 * r + 2] = cu[r]; Vs[3 * r + 2] = cv[r]; }
    }
    const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1}, Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    double Vt[9], UW[9], R1[9], R2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Vt[3 * i + j] = Vs[3 * j + i];
    mat3_mul(U, W, UW); mat3_mul(UW, Vt, R1);
    mat3_mul(U, Wt, UW); mat3_mul(UW, Vt, R2);
    for (int c = 0; c < 4; ++c) {
      const double* Rc = (c < 2) ? R1 : R2;
      const double sgn = (c & 1) ? -1.0 : 1.0;
      for (int i = 0; i < 9; ++i) cand[12 * c + i] = Rc[i];
      for (int r = 0; r < 3; ++r) cand[12 * c + 9 + r] = sgn * U[3 * r + 2];
    }
  }
  __syncwarp();
  int votes[4] = {0, 0, 0, 0};
  const double* g1 = a.x1y1 + (size_t)p * a.n * a.dim;
  const double* g2 = a.x2y2 + (size_t)p * a.n * a.dim;
  const unsigned char* mk = a.mask ? a.mask + (size_t)p * a.n : nullptr;
  for (int i = lane; i < a.n; i += 32) {
    if (mk && !mk[i]) continue;
    const double px1 = g1[(size_t)i * a.dim], py1 = g1[(size_t)i * a.dim + 1], px2 = g2[(size_t)i * a.dim], py2 = g2[(size_t)i * a.dim + 1];
    const double x1[3] = {k1i[0] * px1 + k1i[1] * py1 + k1i[2], k1i[3] * px1 + k1i[4] * py1 + k1i[5], k1i[6] * px1 + k1i[7] * py1 + k1i[8]};
    const double x2[3] = {k2i[0] * px2 + k2i[1] * py2 + k2i[2], k2i[3] * px2 + k2i[4] * py2 + k2i[5], k2i[6] * px2 + k2i[7] * py2 + k2i[8]};
    for (int c = 0; c < 4; ++c) {
      const double* Rc = cand + 12 * c;
      const double* tc = Rc + 9;
      const double a1[3] = {Rc[0] * x1[0] + Rc[1] * x1[1] + Rc[2] * x1[2], Rc[3] * x1[0] + Rc[4] * x1[1] + Rc[5] * x1[2],
                            Rc[6] * x1[0] + Rc[7] * x1[1] + Rc[8] * x1[2]};
      // minimise | l1 a1 - l2 x2 + t |: normal equations of the 3 x 2 system [a1, -x2] (l1, l2)^T = -t
      const double aa = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2];
      const double bb = x2[0] * x2[0] + x2[1] * x2[1] + x2[2] * x2[2];
      const double ab = a1[0] * x2[0] + a1[1] * x2[1] + a1[2] * x2[2];
      const double at = a1[0] * tc[0] + a1[1] * tc[1] + a1[2] * tc[2];
      const double bt = x2[0] * tc[0] + x2[1] * tc[1] + x2[2] * tc[2];
      const double det = aa * bb - ab * ab;
      const double l1 = (-at * bb + ab * bt) / det;
      const double l2 = (aa * bt - ab * at) / det;
      if (l1 > 0 && l2 > 0) ++votes[c];
    }
  }
  for (int c = 0; c < 4; ++c)
    for (int o = 16; o > 0; o >>= 1) votes[c] += __shfl_xor_sync(0xffffffffu, votes[c], o);
  if (lane == 0) {
    int best = 0;
    for (int c = 1; c < 4; ++c) if (votes[c] > votes[best]) best = c;
    for (int i = 0; i < 9; ++i) a.R[(size_t)p * 9 + i] = cand[12 * best + i];
    for (int i = 0; i < 3; ++i) a.t[(size_t)p * 3 + i] = cand[12 * best + 9 + i];
    if (a.good) a.good[p] = votes[best];
  }
}

// ------------------------------------------------------------------------------------------------ QR null space
__global__ void qr7x9_kernel(const double* A, double* N, int* rc, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double a[63], n[18];
  for (int k = 0; k < 63; ++k) a[k] = A[(size_t)i * 63 + k];
  for (int k = 0; k < 18; ++k) n[k] = 0.0;
  const int r = dg::nullspace_qr7x9(a, n);
  for (int k = 0; k < 18; ++k) N[(size_t)i * 18 + k] = n[k];
  if (rc) rc[i] = r;
}

// ------------------------------------------------------------------------------------------------ division check
// Bit-for-bit comparison of the shared-reciprocal quotients used by h_pinvJ (hgeom.h) with IEEE division on random
// operands spanning 60 binary orders of magnitude; returns the number of mismatches in *bad.
__global__ void div_check_kernel(unsigned long long seed, long long count, unsigned long long* bad) {
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long local = 0;
  for (long long i = i0; i < count; i += (long long)gridDim.x * blockDim.x) {
    uint32_t o[4];
    dg::philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 7u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    // mantissas from 52 random bits, exponents in [-30, 30], random signs
    const unsigned long long mx = ((unsigned long long)o[0] << 20) ^ o[1], mn = ((unsigned long long)o[2] << 20) ^ o[3];
    const int ex = (int)(o[1] % 61u) - 30, en = (int)(o[3] % 61u) - 30;
    double x = __longlong_as_double((long long)(((unsigned long long)(1023 + ex) << 52) | (mx & 0xfffffffffffffull)));
    double n = __longlong_as_double((long long)(((unsigned long long)(1023 + en) << 52) | (mn & 0xfffffffffffffull)));
    if (o[0] & 1u) x = -x;
    if (o[2] & 1u) n = -n;
    const double q = dg::div_by_shared_rcp(x, n, __drcp_rn(n));
    if (__double_as_longlong(q) != __double_as_longlong(x / n)) ++local;
  }
  if (local) atomicAdd(bad, local);
}

}  // namespace

extern "C" {

int dgb200_debug_div_check(unsigned long long seed, long long count, unsigned long long* h_bad) {
  unsigned long long* d = nullptr;
  CU2(cudaMalloc(&d, sizeof(unsigned long long)));
  CU2(cudaMemset(d, 0, sizeof(unsigned long long)));
  div_check_kernel<<<1184, 256>>>(seed, count, d);
  CU2(cudaGetLastError());
  CU2(cudaMemcpy(h_bad, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return 0;
}

const char* dgb200_frontend_last_error(void) { return g_err2; }

size_t dgb200_match_workspace_bytes(int n1, int n2) {
  return (size_t)n1 * (sizeof(int) + 2 * sizeof(float)) + (size_t)n2 * sizeof(unsigned long long) + 256;
}

int dgb200_match_descriptors_dev(const float* d_desc1, int n1, const float* d_desc2, int n2, int D, float ratio, int mutual,
                                 const double* d_kp1, const double* d_kp2, int kp_dim, int* d_match_q, int* d_match_t,
                                 double* d_x1y1, double* d_x2y2, int out_dim, int capacity, int* d_count, void* d_workspace,
                                 void* stream) {
  if (!d_desc1 || !d_desc2 || !d_match_q || !d_match_t || !d_count || !d_workspace) return fail2(DGB200_E_ARG, "null buffer");
  if (n1 < 1 || n2 < 2) return fail2(DGB200_E_ARG, "need n1 >= 1 query and n2 >= 2 train descriptors");
  if (D < 4 || D > 256 || (D & 3)) return fail2(DGB200_E_ARG, "descriptor length must be a multiple of 4 in [4, 256]");
  if (d_x1y1 && (!d_kp1 || !d_kp2 || !d_x2y2 || out_dim < 2 || out_dim > kp_dim)) return fail2(DGB200_E_ARG, "keypoint arrays / out_dim inconsistent");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* w = (unsigned char*)d_workspace;
  MatchArgs a;
  a.d1 = d_desc1; a.d2 = d_desc2; a.n1 = n1; a.n2 = n2; a.D = D;
  a.col_best = (unsigned long long*)w; w += (size_t)n2 * sizeof(unsigned long long);
  a.nn_idx = (int*)w; w += (size_t)n1 * sizeof(int);
  a.nn_d1 = (float*)w; w += (size_t)n1 * sizeof(float);
  a.nn_d2 = (float*)w;
  a.use_bulk = (((uintptr_t)d_desc2) & 15) == 0 ? 1 : 0;       // rows are multiples of 16 bytes (D % 4 == 0)
  a.mutual = mutual;
  CU2(cudaMemsetAsync(a.col_best, 0xff, (size_t)n2 * sizeof(unsigned long long), st));
  const size_t smem = sizeof(float) * ((size_t)kTQ * (D + 4) + 2 * (size_t)kTT * D);
  CU2(cudaFuncSetAttribute(nn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  nn2_kernel<<<(n1 + kTQ - 1) / kTQ, kMatchThreads, smem, st>>>(a);
  CU2(cudaGetLastError());
  select_kernel<<<1, 1024, 0, st>>>(n1, a.nn_idx, a.nn_d1, a.nn_d2, a.col_best, ratio * ratio, mutual, d_kp1, d_kp2, kp_dim,
                                    d_match_q, d_match_t, d_x1y1, d_x2y2, out_dim, capacity, d_count);
  CU2(cudaGetLastError());
  return 0;
}

int dgb200_pose_from_fundamental_batch_dev(const double* d_F, const double* d_K1, const double* d_K2, int k_per_pair,
                                           const double* d_x1y1, const double* d_x2y2, const uint8_t* d_mask, int n_pairs,
                                           int n, int dim, double* d_R_out, double* d_t_out, int32_t* d_good_out,
                                           void* stream) {
  if (!d_F || !d_K1 || !d_K2 || !d_x1y1 || !d_x2y2 || !d_R_out || !d_t_out) return fail2(DGB200_E_ARG, "null buffer");
  if (n_pairs < 1 || n < 1 || dim < 2) return fail2(DGB200_E_ARG, "bad sizes");
  PoseArgs a;
  a.F = d_F; a.K1 = d_K1; a.K2 = d_K2; a.k_per_pair = k_per_pair; a.x1y1 = d_x1y1; a.x2y2 = d_x2y2; a.mask = d_mask;
  a.n_pairs = n_pairs; a.n = n; a.dim = dim; a.R = d_R_out; a.t = d_t_out; a.good = d_good_out;
  const int threads = 256;
  pose_kernel<<<(n_pairs * 32 + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(a);
  CU2(cudaGetLastError());
  return 0;
}

int dgb200_nullspace_qr7x9_batch_dev(const double* d_A, double* d_N, int32_t* d_rc, int count, void* stream) {
  if (!d_A || !d_N || count < 1) return fail2(DGB200_E_ARG, "bad arguments");
  qr7x9_kernel<<<(count + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_A, d_N, d_rc, count);
  CU2(cudaGetLastError());
  return 0;
}

}  // extern "C"
