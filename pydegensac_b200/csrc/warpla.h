// warpla.h -- warp-cooperative small dense solves (one warp, operands in a per-warp shared-memory tile).
//
// The LO and DEGENSAC steps are latency chains of 9x9 eigen-problems and 9x8 null-vector problems; run by a
// single thread they dominated the pair time (profiles/r01: 64 % of all cycles).  Here one warp does them:
//   * symmetric 9x9 eigen-decomposition by PARALLEL-ORDER Jacobi: in each of the 9 rounds of a sweep the
//     4 disjoint index pairs {(r+k) mod 9, (r-k) mod 9} are rotated at once (36+36 two-element updates
//     spread over the lanes), replacing LAPACK dsyev_ on the reference side (lapwrap.c:67);
//   * Householder QR of the 9 x len (len <= 8) system with one lane per column, giving the vector
//     orthogonal to all columns, replacing CCMATH svduv's "last column of U" (Ftools.c:373,383).
// The code is SPMD over `lane` in [0, W): W = 32 on the device, W = 1 in the host emulation, where every
// strided loop degenerates to the sequential order -- same arithmetic, same results.
#pragma once
#include <stddef.h>
#include "common.h"
#include "la.h"

namespace dg {

struct WarpScratch {
  double A[81];
  double V[81];
  double aux[168];   // second Jacobi buffer (A and V) / DLT rows / Householder vectors
  double cs[16];
};

static_assert(offsetof(WarpScratch, V) == 81 * sizeof(double), "[A|V] must be contiguous");

#if DG_DEVICE_PASS
#define DG_WSYNC() __syncwarp()
DG_ENG inline double wl_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#else
#define DG_WSYNC() ((void)0)
inline double wl_sum(double v) { return v; }
#endif

#if DG_DEVICE_PASS
DG_ENG inline double dg_rsqrt(double x) { return rsqrt(x); }
#else
inline double dg_rsqrt(double x) { return 1.0 / sqrt(x); }
#endif

// Eigen-decomposition of the symmetric matrix in ws->A (full 9x9 row-major).  On return ws->A holds the
// eigenvalues on its diagonal and column k of ws->V the eigenvector of A[k][k].
// Per round: (1) four lanes compute the rotations of the four disjoint pairs (one sqrt, one division and
// one rsqrt each); (2) every lane rebuilds its entries of J^T A J directly from the four old entries they
// depend on (A is double-buffered between ws->A and ws->aux) and rotates its entries of V.
DG_ENGN void warp_jacobi_eig9(WarpScratch* ws, int lane, int W) {
  // [A | V] lives in ws->A..ws->V (162 contiguous doubles) and ping-pongs with ws->aux
  double* cur = ws->A;
  double* nxt = ws->aux;
  #pragma unroll 1
  for (int t = lane; t < 81; t += W) cur[81 + t] = (t / 9 == t % 9) ? 1.0 : 0.0;
  DG_WSYNC();
  #pragma unroll 1
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dia = 0.0;
    #pragma unroll 1
    for (int t = lane; t < 81; t += W) {
      const int r = t / 9, cc = t % 9;
      const double v = cur[t];
      if (r < cc) off += v * v;
      else if (r == cc) dia += v * v;
    }
    off = wl_sum(off);
    dia = wl_sum(dia);
    if (!(off > 4e-30 * dia) || off == 0.0) break;  // off-norm at rounding level: converged
    #pragma unroll 1
    for (int r = 0; r < 9; ++r) {
      #pragma unroll 1
      for (int k = lane; k < 4; k += W) {
        const int i = (r + k + 1) % 9, j = (r + 9 - k - 1) % 9;
        const int p = i < j ? i : j, q = i < j ? j : i;
        const double apq = cur[p * 9 + q];
        double c = 1.0, s = 0.0;
        if (apq != 0.0) {
          const double zeta = cur[q * 10] - cur[p * 10];
          const double h = sqrt(zeta * zeta + 4.0 * apq * apq);
          const double tt = (zeta >= 0.0 ? 2.0 * apq : -2.0 * apq) / (fabs(zeta) + h);
          c = dg_rsqrt(tt * tt + 1.0);
          s = tt * c;
        }
        ws->cs[2 * k] = c;
        ws->cs[2 * k + 1] = s;
      }
      DG_WSYNC();
      // Index x != r is rotated with partner 2r-x (mod 9): x' = c*x -/+ s*partner (- for the smaller index).
      // new A[i][j] = sum over {i,pi} x {j,pj} of J[a][i] * A[a][b] * J[b][j];  new V[:,j] = c*V[:,j] -/+ s*V[:,pj]
      #pragma unroll 1
      for (int t = lane; t < 162; t += W) {
        const int u = (t < 81) ? t : t - 81;
        const int i = u / 9, j = u % 9;
        int pj = j;
        double cj = 1.0, sj = 0.0;
        if (j != r) {
          const int d = (j - r + 9) % 9;
          const int k = (d <= 4) ? d - 1 : 8 - d;
          pj = (2 * r - j + 18) % 9;
          const double c = ws->cs[2 * k], s = ws->cs[2 * k + 1];
          cj = c;
          sj = (j < pj) ? -s : s;
        }
        if (t < 81) {
          int pi = i;
          double ci = 1.0, si = 0.0;
          if (i != r) {
            const int d = (i - r + 9) % 9;
            const int k = (d <= 4) ? d - 1 : 8 - d;
            pi = (2 * r - i + 18) % 9;
            const double c = ws->cs[2 * k], s = ws->cs[2 * k + 1];
            ci = c;
            si = (i < pi) ? -s : s;
          }
          const double u0 = ci * cur[i * 9 + j] + si * cur[pi * 9 + j];
          const double u1 = ci * cur[i * 9 + pj] + si * cur[pi * 9 + pj];
          nxt[t] = cj * u0 + sj * u1;
        } else {
          nxt[t] = cj * cur[81 + i * 9 + j] + sj * cur[81 + i * 9 + pj];
        }
      }
      DG_WSYNC();
      double* tmp = cur; cur = nxt; nxt = tmp;
    }
  }
  if (cur != ws->A) {
    #pragma unroll 1
    for (int t = lane; t < 162; t += W) ws->A[t] = cur[t];
    DG_WSYNC();
  }
}

// Eigenvector of the smallest eigenvalue of the symmetric matrix whose lower triangle (row-major packed,
// 45 entries: (0,0),(1,0),(1,1),(2,0)...) is in `packed`; result in ws->cs[0..8] (visible after DG_WSYNC).
DG_ENGN void warp_min_eigvec9_packed(WarpScratch* ws, const double* packed, int lane, int W) {
  #pragma unroll 1
  for (int t = lane; t < 45; t += W) {
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    const int j = t - i * (i + 1) / 2;
    const double v = packed[t];
    ws->A[9 * i + j] = v;
    ws->A[9 * j + i] = v;
  }
  DG_WSYNC();
  warp_jacobi_eig9(ws, lane, W);
  DG_WSYNC();
  if (lane == 0) {
    int m = 0;
    #pragma unroll 1
    for (int i = 1; i < 9; ++i)
      if (ws->A[i * 10] < ws->A[m * 10]) m = i;
    #pragma unroll 1
    for (int i = 0; i < 9; ++i) ws->cs[i] = ws->V[i * 9 + m];
  }
  DG_WSYNC();
}

// Unit vector orthogonal to the `len` (1..8) columns of the 9 x len row-major matrix in ws->A[0 .. 9*len)
// (destroyed); result in ws->cs[0..8].  Same Householder arithmetic as left_null_9xk (la.h), with the
// reflector applied to the trailing columns one lane per column.
DG_ENGN void warp_left_null_9xk(WarpScratch* ws, int len, int lane, int W) {
  double* Z = ws->A;
  double* vs = ws->aux;        // len x 9 reflectors
  double* beta = ws->aux + 80;
  #pragma unroll 1
  for (int c = 0; c < len; ++c) {
    if (lane == 0) {
      double nrm = 0.0;
      #pragma unroll 1
      for (int r = c; r < 9; ++r) nrm += Z[r * len + c] * Z[r * len + c];
      nrm = sqrt(nrm);
      #pragma unroll 1
      for (int r = 0; r < 9; ++r) vs[c * 9 + r] = 0.0;
      if (nrm == 0.0) {
        beta[c] = 0.0;
      } else {
        const double x0 = Z[c * len + c];
        const double alpha = (x0 >= 0.0) ? -nrm : nrm;
        #pragma unroll 1
        for (int r = c; r < 9; ++r) vs[c * 9 + r] = Z[r * len + c];
        vs[c * 9 + c] = x0 - alpha;
        double vn = 0.0;
        #pragma unroll 1
        for (int r = c; r < 9; ++r) vn += vs[c * 9 + r] * vs[c * 9 + r];
        beta[c] = (vn > 0.0) ? 2.0 / vn : 0.0;
      }
    }
    DG_WSYNC();
    if (beta[c] != 0.0) {
      #pragma unroll 1
      for (int cc = c + lane; cc < len; cc += W) {
        double dot = 0.0;
        #pragma unroll 1
        for (int r = c; r < 9; ++r) dot += vs[c * 9 + r] * Z[r * len + cc];
        dot *= beta[c];
        #pragma unroll 1
        for (int r = c; r < 9; ++r) Z[r * len + cc] -= dot * vs[c * 9 + r];
      }
    }
    DG_WSYNC();
  }
  if (lane == 0) {
    double q[9];
    #pragma unroll 1
    for (int r = 0; r < 9; ++r) q[r] = 0.0;
    q[8] = 1.0;
    #pragma unroll 1
    for (int c = len - 1; c >= 0; --c) {
      double dot = 0.0;
      #pragma unroll 1
      for (int r = c; r < 9; ++r) dot += vs[c * 9 + r] * q[r];
      dot *= beta[c];
      #pragma unroll 1
      for (int r = c; r < 9; ++r) q[r] -= dot * vs[c * 9 + r];
    }
    #pragma unroll 1
    for (int r = 0; r < 9; ++r) ws->cs[r] = q[r];
  }
  DG_WSYNC();
}

}  // namespace dg
