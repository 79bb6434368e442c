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
#include "common.h"
#include "la.h"

namespace dg {

struct WarpScratch {
  double A[81];
  double V[81];
  double aux[112];   // rows / Householder vectors
  double cs[16];
};

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

// Eigen-decomposition of the symmetric matrix in ws->A (full 9x9 row-major).  On return ws->A holds the
// eigenvalues on its diagonal and column k of ws->V the eigenvector of A[k][k].
DG_ENGN void warp_jacobi_eig9(WarpScratch* ws, int lane, int W) {
  double* A = ws->A;
  double* V = ws->V;
  for (int t = lane; t < 81; t += W) V[t] = (t / 9 == t % 9) ? 1.0 : 0.0;
  DG_WSYNC();
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int t = lane; t < 81; t += W) {
      const int r = t / 9, cc = t % 9;
      const double v = A[t];
      if (r < cc) off += v * v;
      else if (r == cc) dia += v * v;
    }
    off = wl_sum(off);
    dia = wl_sum(dia);
    if (!(off > 4e-30 * dia) || off == 0.0) break;  // off-norm at rounding level: converged
    for (int r = 0; r < 9; ++r) {
      for (int k = lane; k < 4; k += W) {
        const int i = (r + k + 1) % 9, j = (r + 9 - k - 1) % 9;
        const int p = i < j ? i : j, q = i < j ? j : i;
        const double apq = A[p * 9 + q];
        double c = 1.0, s = 0.0;
        if (apq != 0.0) {
          const double app = A[p * 10], aqq = A[q * 10];
          const double theta = (aqq - app) / (2.0 * apq);
          const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          c = 1.0 / sqrt(tt * tt + 1.0);
          s = tt * c;
        }
        ws->cs[2 * k] = c;
        ws->cs[2 * k + 1] = s;
      }
      DG_WSYNC();
      for (int t = lane; t < 72; t += W) {   // columns p,q of A and of V
        double* M = (t < 36) ? A : V;
        const int u = t % 36, k = u / 9, row = u % 9;
        const int i = (r + k + 1) % 9, j = (r + 9 - k - 1) % 9;
        const int p = i < j ? i : j, q = i < j ? j : i;
        const double c = ws->cs[2 * k], s = ws->cs[2 * k + 1];
        const double mp = M[row * 9 + p], mq = M[row * 9 + q];
        M[row * 9 + p] = c * mp - s * mq;
        M[row * 9 + q] = s * mp + c * mq;
      }
      DG_WSYNC();
      for (int t = lane; t < 36; t += W) {   // rows p,q of A
        const int k = t / 9, col = t % 9;
        const int i = (r + k + 1) % 9, j = (r + 9 - k - 1) % 9;
        const int p = i < j ? i : j, q = i < j ? j : i;
        const double c = ws->cs[2 * k], s = ws->cs[2 * k + 1];
        const double ap = A[p * 9 + col], aq = A[q * 9 + col];
        A[p * 9 + col] = c * ap - s * aq;
        A[q * 9 + col] = s * ap + c * aq;
      }
      DG_WSYNC();
    }
  }
}

// Eigenvector of the smallest eigenvalue of the symmetric matrix whose lower triangle (row-major packed,
// 45 entries: (0,0),(1,0),(1,1),(2,0)...) is in `packed`; result in ws->cs[0..8] (visible after DG_WSYNC).
DG_ENGN void warp_min_eigvec9_packed(WarpScratch* ws, const double* packed, int lane, int W) {
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
    for (int i = 1; i < 9; ++i)
      if (ws->A[i * 10] < ws->A[m * 10]) m = i;
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
  for (int c = 0; c < len; ++c) {
    if (lane == 0) {
      double nrm = 0.0;
      for (int r = c; r < 9; ++r) nrm += Z[r * len + c] * Z[r * len + c];
      nrm = sqrt(nrm);
      for (int r = 0; r < 9; ++r) vs[c * 9 + r] = 0.0;
      if (nrm == 0.0) {
        beta[c] = 0.0;
      } else {
        const double x0 = Z[c * len + c];
        const double alpha = (x0 >= 0.0) ? -nrm : nrm;
        for (int r = c; r < 9; ++r) vs[c * 9 + r] = Z[r * len + c];
        vs[c * 9 + c] = x0 - alpha;
        double vn = 0.0;
        for (int r = c; r < 9; ++r) vn += vs[c * 9 + r] * vs[c * 9 + r];
        beta[c] = (vn > 0.0) ? 2.0 / vn : 0.0;
      }
    }
    DG_WSYNC();
    if (beta[c] != 0.0) {
      for (int cc = c + lane; cc < len; cc += W) {
        double dot = 0.0;
        for (int r = c; r < 9; ++r) dot += vs[c * 9 + r] * Z[r * len + cc];
        dot *= beta[c];
        for (int r = c; r < 9; ++r) Z[r * len + cc] -= dot * vs[c * 9 + r];
      }
    }
    DG_WSYNC();
  }
  if (lane == 0) {
    double q[9];
    for (int r = 0; r < 9; ++r) q[r] = 0.0;
    q[8] = 1.0;
    for (int c = len - 1; c >= 0; --c) {
      double dot = 0.0;
      for (int r = c; r < 9; ++r) dot += vs[c * 9 + r] * q[r];
      dot *= beta[c];
      for (int r = c; r < 9; ++r) q[r] -= dot * vs[c * 9 + r];
    }
    for (int r = 0; r < 9; ++r) ws->cs[r] = q[r];
  }
  DG_WSYNC();
}

}  // namespace dg
