// la.h -- tiny dense linear algebra done by one thread in registers / local memory.
// No cuSOLVER, no LAPACK: every solve on the hot path is at most 9x9.
#pragma once
#include "common.h"

namespace dg {

// ---------------------------------------------------------------------------------------------
// Null space of a 9x9 row-major system by Gauss-Jordan elimination with partial pivoting.
// Semantics follow the reference's `nullspace` (utools.c:97-167): column sweep, pivot searched
// from the diagonal row downwards with strict improvement, pivots below 1e-12 declare a free
// column, the k-th null vector is (-column of the reduced matrix ; unit on the free column).
// Returns the number of null vectors (written to ns[k*9 + ...]).
// ---------------------------------------------------------------------------------------------
template <int NN>
DG_HD int nullspaceN(double* M, double* ns) {
  const double tol = 1e-12;
  int freec[NN], pivc[NN];
  int nfree = 0, npiv = 0, row = 0;
  #pragma unroll 1
  for (int col = 0; col < NN; ++col) {
    int best = row;
    double mag = fabs(M[NN * row + col]);
    #pragma unroll 1
    for (int r = row + 1; r < NN; ++r) {
      const double t = fabs(M[NN * r + col]);
      if (mag < t) { mag = t; best = r; }
    }
    if (mag < tol) {
      freec[nfree++] = col;
      #pragma unroll 1
      for (int r = row; r < NN; ++r) M[NN * r + col] = 0.0;
      continue;
    }
    pivc[npiv++] = col;
    if (best != row) {
      #pragma unroll 1
      for (int c = col; c < NN; ++c) {
        const double t = M[NN * row + c];
        M[NN * row + c] = M[NN * best + c];
        M[NN * best + c] = t;
      }
    }
    const double p = M[NN * row + col];
    #pragma unroll 1
    for (int c = col; c < NN; ++c) M[NN * row + c] /= p;
    #pragma unroll 1
    for (int r = 0; r < NN; ++r) {
      if (r == row) continue;
      const double a = M[NN * r + col];
      #pragma unroll 1
      for (int c = col; c < NN; ++c) M[NN * r + c] -= a * M[NN * row + c];
    }
    ++row;
  }
  #pragma unroll 1
  for (int k = 0; k < nfree; ++k) {
    const int j = freec[k];
    #pragma unroll 1
    for (int l = 0; l < npiv; ++l) ns[k * NN + pivc[l]] = -M[l * NN + j];
    #pragma unroll 1
    for (int l = 0; l < nfree; ++l) ns[k * NN + freec[l]] = (j == freec[l]) ? 1.0 : 0.0;
  }
  return nfree;
}
DG_HDN int nullspace9(double* M, double* ns) { return nullspaceN<9>(M, ns); }
DG_HDN int nullspace15(double* M, double* ns) { return nullspaceN<15>(M, ns); }   // two-ellipse H solver (h2el.h)

// ---------------------------------------------------------------------------------------------
// Symmetric 9x9 eigen-decomposition by cyclic Jacobi rotations (replaces LAPACK dsyev_, which the
// reference reaches through lapwrap.c:67 from u2f/u2fw/u2h).  A is destroyed; on return column k of
// V (V[r*9+k]) is the eigenvector of d[k].  Eigenvalues are NOT sorted; callers pick the minimum.
// ---------------------------------------------------------------------------------------------
DG_HDN void jacobi_eig9(double* A, double* V, double* d) {
  #pragma unroll 1
  for (int i = 0; i < 81; ++i) V[i] = 0.0;
  for (int i = 0; i < 9; ++i) V[i * 10] = 1.0;
  #pragma unroll 1
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int p = 0; p < 9; ++p) {
      dia += A[p * 10] * A[p * 10];
      #pragma unroll 1
      for (int q = p + 1; q < 9; ++q) off += A[p * 9 + q] * A[p * 9 + q];
    }
    if (!(off > 4e-30 * dia) || off == 0.0) break;  // off-norm at rounding level: converged
    for (int p = 0; p < 8; ++p) {
      #pragma unroll 1
      for (int q = p + 1; q < 9; ++q) {
        const double apq = A[p * 9 + q];
        if (apq == 0.0) continue;
        const double app = A[p * 10], aqq = A[q * 10];
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 9; ++k) {  // columns p,q
          const double akp = A[k * 9 + p], akq = A[k * 9 + q];
          A[k * 9 + p] = c * akp - s * akq;
          A[k * 9 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 9; ++k) {  // rows p,q
          const double apk = A[p * 9 + k], aqk = A[q * 9 + k];
          A[p * 9 + k] = c * apk - s * aqk;
          A[q * 9 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 9; ++k) {
          const double vkp = V[k * 9 + p], vkq = V[k * 9 + q];
          V[k * 9 + p] = c * vkp - s * vkq;
          V[k * 9 + q] = s * vkp + c * vkq;
        }
      }
    }
  }
  for (int i = 0; i < 9; ++i) d[i] = A[i * 10];
}

// Eigenvector of the smallest eigenvalue of the symmetric 9x9 matrix C (destroyed) -> v[9].
DG_HDN void min_eigvec9(double* C, double* v) {
  double V[81], d[9];
  jacobi_eig9(C, V, d);
  int j = 0;
  for (int i = 1; i < 9; ++i)
    if (d[i] < d[j]) j = i;
  for (int i = 0; i < 9; ++i) v[i] = V[i * 9 + j];
}

// ---------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi SVD of a 3x3 row-major matrix: G = A*V has orthogonal columns whose
// norms are the singular values.  Used for the rank-2 projection of F (reference singulF,
// Ftools.c:330-347, LAPACK dgesvd_) and for the epipole in Hdetect (DegUtils.c:109, CCMATH svduv).
// ---------------------------------------------------------------------------------------------
#if DG_DEVICE_PASS
#define DG_RSQRT(x) rsqrt(x)
#else
#define DG_RSQRT(x) (1.0 / sqrt(x))
#endif
// One Hestenes rotation of columns (p,q) with compile-time indices: everything stays in registers.
#define DG_SVD3_ROT(p, q)                                                                         \
  {                                                                                               \
    const double al = G[p] * G[p] + G[3 + p] * G[3 + p] + G[6 + p] * G[6 + p];                    \
    const double be = G[q] * G[q] + G[3 + q] * G[3 + q] + G[6 + q] * G[6 + q];                    \
    const double ga = G[p] * G[q] + G[3 + p] * G[3 + q] + G[6 + p] * G[6 + q];                    \
    if (ga != 0.0 && ga * ga > 1e-30 * (al * be)) {                                               \
      rotated = true;                                                                             \
      const double zeta = be - al;                                                                \
      const double hh = sqrt(zeta * zeta + 4.0 * ga * ga);                                        \
      const double t = (zeta >= 0.0 ? 2.0 * ga : -2.0 * ga) / (fabs(zeta) + hh);                  \
      const double c = DG_RSQRT(1.0 + t * t), s = c * t;                                          \
      double gp, gq;                                                                              \
      gp = G[p]; gq = G[q]; G[p] = c * gp - s * gq; G[q] = s * gp + c * gq;                        \
      gp = G[3 + p]; gq = G[3 + q]; G[3 + p] = c * gp - s * gq; G[3 + q] = s * gp + c * gq;        \
      gp = G[6 + p]; gq = G[6 + q]; G[6 + p] = c * gp - s * gq; G[6 + q] = s * gp + c * gq;        \
      gp = V[p]; gq = V[q]; V[p] = c * gp - s * gq; V[q] = s * gp + c * gq;                        \
      gp = V[3 + p]; gq = V[3 + q]; V[3 + p] = c * gp - s * gq; V[3 + q] = s * gp + c * gq;        \
      gp = V[6 + p]; gq = V[6 + q]; V[6 + p] = c * gp - s * gq; V[6 + q] = s * gp + c * gq;        \
    }                                                                                             \
  }
DG_HDN void svd3_onesided(const double* A, double* Gout, double* Vout, double* sv) {
  double G[9], V[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { G[i] = A[i]; V[i] = 0.0; }
  V[0] = V[4] = V[8] = 1.0;
#pragma unroll 1
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
    DG_SVD3_ROT(0, 1)
    DG_SVD3_ROT(0, 2)
    DG_SVD3_ROT(1, 2)
    if (!rotated) break;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) sv[c] = sqrt(G[c] * G[c] + G[3 + c] * G[3 + c] + G[6 + c] * G[6 + c]);
#pragma unroll
  for (int i = 0; i < 9; ++i) { Gout[i] = G[i]; Vout[i] = V[i]; }
}

// 2^-e for the binary exponent e of a positive normal double (exact power of two; 0 when t is not usable).
DG_HD double pow2_inv_exponent(double t) {
  if (!(t > 1e-290) || !(t < 1e290)) return 0.0;
#if DG_DEVICE_PASS
  const int hi = __double2hiint(t);
  return __hiloint2double((2046 - ((hi >> 20) & 0x7ff)) << 20, 0);
#else
  int e;
  frexp(t, &e);
  return ldexp(1.0, 1 - e);
#endif
}

// Right singular vector v of the SMALLEST singular value of the 3x3 row-major F, without an SVD:
// the cofactor matrix C of F = U S V^T is U diag(s1 s2, s0 s2, s0 s1) V^T (up to sign), so v is the dominant
// eigenvector of K = C^T C, whose eigenvalue ratio is (s2/s1)^2.  Seven squarings of K (power 128) leave a
// numerically rank-one matrix whenever s2/s1 < ~0.85; its largest column is v.  Working on the cofactors
// instead of F^T F keeps the accuracy at eps * s1/(s1 - s2) even for badly scaled F (un-normalised 8-point
// fits).  Returns false (caller falls back to the Jacobi SVD) when the residual test fails.
DG_HD bool smallest_right_sv3_fast(const double* F, double* v) {
  double C[9];
  C[0] = F[4] * F[8] - F[5] * F[7]; C[1] = F[5] * F[6] - F[3] * F[8]; C[2] = F[3] * F[7] - F[4] * F[6];
  C[3] = F[2] * F[7] - F[1] * F[8]; C[4] = F[0] * F[8] - F[2] * F[6]; C[5] = F[1] * F[6] - F[0] * F[7];
  C[6] = F[1] * F[5] - F[2] * F[4]; C[7] = F[2] * F[3] - F[0] * F[5]; C[8] = F[0] * F[4] - F[1] * F[3];
  // K = C^T C, symmetric: k00 k01 k02 k11 k12 k22
  double k00 = C[0] * C[0] + C[3] * C[3] + C[6] * C[6];
  double k01 = C[0] * C[1] + C[3] * C[4] + C[6] * C[7];
  double k02 = C[0] * C[2] + C[3] * C[5] + C[6] * C[8];
  double k11 = C[1] * C[1] + C[4] * C[4] + C[7] * C[7];
  double k12 = C[1] * C[2] + C[4] * C[5] + C[7] * C[8];
  double k22 = C[2] * C[2] + C[5] * C[5] + C[8] * C[8];
  const double sc = pow2_inv_exponent(k00 + k11 + k22);
  if (sc == 0.0) return false;
  k00 *= sc; k01 *= sc; k02 *= sc; k11 *= sc; k12 *= sc; k22 *= sc;   // trace in [1,2): 7 squarings stay in range
  double m00 = k00, m01 = k01, m02 = k02, m11 = k11, m12 = k12, m22 = k22;
  // rolled on purpose: straight-line code beyond the ~32 KB instruction cache streams at ~6 cycles per instruction
  // per warp on B200 (tools/dbg/icache.cu), a 36-instruction loop body runs at dependency latency
#pragma unroll 1
  for (int it = 0; it < 7; ++it) {
    const double n00 = m00 * m00 + m01 * m01 + m02 * m02;
    const double n01 = m00 * m01 + m01 * m11 + m02 * m12;
    const double n02 = m00 * m02 + m01 * m12 + m02 * m22;
    const double n11 = m01 * m01 + m11 * m11 + m12 * m12;
    const double n12 = m01 * m02 + m11 * m12 + m12 * m22;
    const double n22 = m02 * m02 + m12 * m12 + m22 * m22;
    m00 = n00; m01 = n01; m02 = n02; m11 = n11; m12 = n12; m22 = n22;
  }
  double x, y, z;
  if (m00 >= m11 && m00 >= m22) { x = m00; y = m01; z = m02; }
  else if (m11 >= m22)          { x = m01; y = m11; z = m12; }
  else                          { x = m02; y = m12; z = m22; }
  const double n2 = x * x + y * y + z * z;
  if (!(n2 > 0.0) || !(n2 < 1e300)) return false;
  const double r = DG_RSQRT(n2);
  x *= r; y *= r; z *= r;
  // residual against the (scaled) K itself
  const double w0 = k00 * x + k01 * y + k02 * z;
  const double w1 = k01 * x + k11 * y + k12 * z;
  const double w2 = k02 * x + k12 * y + k22 * z;
  const double mu = w0 * x + w1 * y + w2 * z;
  const double r0 = w0 - mu * x, r1 = w1 - mu * y, r2 = w2 - mu * z;
  if (!(r0 * r0 + r1 * r1 + r2 * r2 <= 1e-29 * (mu * mu))) return false;
  v[0] = x; v[1] = y; v[2] = z;
  return true;
}

// Rank-2 projection F <- U diag(s0,s1,0) V^T == F - (F v_min) v_min^T   (reference: singulF, Ftools.c:330-347)
DG_HDN void enforce_rank2_slow(double* F) {
  double G[9], V[9], sv[3];
  svd3_onesided(F, G, V, sv);
  int m = 0;
  if (sv[1] < sv[m]) m = 1;
  if (sv[2] < sv[m]) m = 2;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) F[3 * i + j] -= G[3 * i + m] * V[3 * j + m];
}
// inlinable flavour: the fast path keeps F in the caller's registers
DG_HD void enforce_rank2_inl(double (&F)[9]) {
  double v[3];
  if (smallest_right_sv3_fast(F, v)) {
    const double g0 = F[0] * v[0] + F[1] * v[1] + F[2] * v[2];
    const double g1 = F[3] * v[0] + F[4] * v[1] + F[5] * v[2];
    const double g2 = F[6] * v[0] + F[7] * v[1] + F[8] * v[2];
    F[0] -= g0 * v[0]; F[1] -= g0 * v[1]; F[2] -= g0 * v[2];
    F[3] -= g1 * v[0]; F[4] -= g1 * v[1]; F[5] -= g1 * v[2];
    F[6] -= g2 * v[0]; F[7] -= g2 * v[1]; F[8] -= g2 * v[2];
    return;
  }
  double T[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) T[i] = F[i];
  enforce_rank2_slow(T);
#pragma unroll
  for (int i = 0; i < 9; ++i) F[i] = T[i];
}
DG_HDN void enforce_rank2(double* F) {
  double v[3];
  if (smallest_right_sv3_fast(F, v)) {
    const double g0 = F[0] * v[0] + F[1] * v[1] + F[2] * v[2];
    const double g1 = F[3] * v[0] + F[4] * v[1] + F[5] * v[2];
    const double g2 = F[6] * v[0] + F[7] * v[1] + F[8] * v[2];
    F[0] -= g0 * v[0]; F[1] -= g0 * v[1]; F[2] -= g0 * v[2];
    F[3] -= g1 * v[0]; F[4] -= g1 * v[1]; F[5] -= g1 * v[2];
    F[6] -= g2 * v[0]; F[7] -= g2 * v[1]; F[8] -= g2 * v[2];
    return;
  }
  double G[9], V[9], sv[3];
  svd3_onesided(F, G, V, sv);
  int m = 0;
  if (sv[1] < sv[m]) m = 1;
  if (sv[2] < sv[m]) m = 2;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) F[3 * i + j] -= G[3 * i + m] * V[3 * j + m];
}

// Right singular vector of the smallest singular value of a 3x3 row-major matrix (A v ~ 0).
DG_HDN void right_null3(const double* A, double* v) {
  double G[9], V[9], sv[3];
  svd3_onesided(A, G, V, sv);
  int m = 0;
  if (sv[1] < sv[m]) m = 1;
  if (sv[2] < sv[m]) m = 2;
  v[0] = V[m]; v[1] = V[3 + m]; v[2] = V[6 + m];
}

// ---------------------------------------------------------------------------------------------
// Third right singular vector of a 3x3 row-major matrix AS CCMATH's svduv RETURNS IT
// (matutls/svduv.c + ldvmat.c + qrbdv.c): Householder bidiagonalisation followed by Golub-Kahan-Reinsch
// implicit-shift QR sweeps, singular values left UNSORTED.  The reference's Hdetect takes column 2 of V
// as the epipole (DegUtils.c:109-110) whether or not the vanishing singular value ended up in slot 2
// (it lands in slot 1 for roughly a quarter of rank-2 inputs), so the DEGENSAC test only reproduces if
// the same sweep order is followed.  Left rotations are not accumulated (U is never used).
// ---------------------------------------------------------------------------------------------
DG_HDN void gkr_third_right_vector3(const double* Ain, double* vout) {
  double a[9], d[3], e[2], V[9];
  for (int i = 0; i < 9; ++i) a[i] = Ain[i];
  e[0] = e[1] = 0.0;
  // --- column 0 reflector
  {
    double w0 = a[0], w1 = a[3], w2 = a[6];
    double s = w0 * w0 + w1 * w1 + w2 * w2, h = 0.0;
    if (s > 0.) {
      h = sqrt(s);
      if (a[0] < 0.) h = -h;
      s += a[0] * h;
      s = 1. / s;
      w0 += h;
      for (int k = 1; k < 3; ++k) {
        double r = w0 * a[k] + w1 * a[3 + k] + w2 * a[6 + k];
        r *= s;
        a[k] -= r * w0; a[3 + k] -= r * w1; a[6 + k] -= r * w2;
      }
    }
    d[0] = -h;
  }
  // --- row 0 reflector over columns 1,2 (the only one that shapes V for n = 3)
  double hb = 0.0, u2 = 0.0;  // V = I - hb * (1,u2)(1,u2)^T on the trailing 2x2 block
  {
    double s = a[1] * a[1] + a[2] * a[2], h = 0.0;
    if (s > 0.) {
      h = sqrt(s);
      if (a[1] < 0.) h = -h;
      hb = 1. + fabs(a[1] / h);
      s += a[1] * h;
      s = 1. / s;
      const double p0 = a[1] + h;
      const double t = 1. / p0;
      for (int row = 1; row < 3; ++row) {
        double r = p0 * a[3 * row + 1] + a[2] * a[3 * row + 2];
        r *= s;
        a[3 * row + 1] -= r * p0;
        a[3 * row + 2] -= r * a[2];
      }
      u2 = a[2] * t;
    }
    e[0] = -h;
  }
  // --- column 1 reflector (rows 1,2)
  {
    double w0 = a[4], w1 = a[7];
    double s = w0 * w0 + w1 * w1, h = 0.0;
    if (s > 0.) {
      h = sqrt(s);
      if (a[4] < 0.) h = -h;
      s += a[4] * h;
      s = 1. / s;
      w0 += h;
      double r = w0 * a[5] + w1 * a[8];
      r *= s;
      a[5] -= r * w0; a[8] -= r * w1;
    }
    d[1] = -h;
  }
  e[1] = a[5];
  d[2] = a[8];
  for (int i = 0; i < 9; ++i) V[i] = 0.0;
  V[0] = 1.0; V[4] = 1.0; V[8] = 1.0;
  if (hb != 0.) {
    V[4] = 1. - hb;
    V[7] = -hb * u2;
    const double sdot = hb * u2;      // (V[8]=1) * u2 * hb
    V[8] = 1. - sdot * u2;
    V[5] = -sdot;
  }
  // --- implicit-shift QR sweeps on the bidiagonal (d,e)
  int m = 3;
  double t = fabs(d[0]);
  for (int j = 1; j < 3; ++j) {
    const double s = fabs(d[j]) + fabs(e[j - 1]);
    if (s > t) t = s;
  }
  t *= 1.e-15;
  #pragma unroll 1
  for (int it = 0; m > 1 && it < 300; ++it) {
    int k;
    #pragma unroll 1
    for (k = m - 1; k > 0; --k) {
      if (fabs(e[k - 1]) < t) break;
      if (fabs(d[k - 1]) < t) {
        double s = 1., c = 0.;
        #pragma unroll 1
        for (int i = k; i < m; ++i) {
          const double aa = s * e[i - 1], bb = d[i];
          e[i - 1] *= c;
          const double u = sqrt(aa * aa + bb * bb);
          d[i] = u;
          s = -aa / u;
          c = bb / u;
        }
        break;
      }
    }
    double y = d[k], x = d[m - 1], u = e[m - 2];
    double aa = (y + x) * (y - x) - u * u, s = y * e[k], bb = s + s;
    u = sqrt(aa * aa + bb * bb);
    if (u != 0.) {
      double c = sqrt((u + aa) / (u + u));
      if (c != 0.) s /= (c * u); else s = 1.;
      #pragma unroll 1
      for (int i = k; i < m - 1; ++i) {
        bb = e[i];
        if (i > k) {
          aa = s * e[i];
          bb *= c;
          e[i - 1] = u = sqrt(x * x + aa * aa);
          c = x / u;
          s = aa / u;
        }
        aa = c * y + s * bb;
        bb = c * bb - s * y;
        for (int r = 0; r < 3; ++r) {
          const double w = c * V[3 * r + i] + s * V[3 * r + i + 1];
          V[3 * r + i + 1] = c * V[3 * r + i + 1] - s * V[3 * r + i];
          V[3 * r + i] = w;
        }
        s *= d[i + 1];
        d[i] = u = sqrt(aa * aa + s * s);
        y = c * d[i + 1];
        c = aa / u;
        s /= u;
        x = c * bb + s * y;
        y = c * y - s * bb;
      }
    }
    e[m - 2] = x;
    d[m - 1] = y;
    if (fabs(x) < t) --m;
    if (m == k + 1) --m;
  }
  vout[0] = V[2]; vout[1] = V[5]; vout[2] = V[8];  // the sign flip svduv applies for d[2] < 0 does not matter
}

// ---------------------------------------------------------------------------------------------
// Unit vector orthogonal to the `len` (<= 8) columns of the 9 x len row-major matrix Z (destroyed):
// the last column of the full Q of a Householder QR.  This is what the reference takes from CCMATH
// svduv as "last column of U" in the len <= 8 branch of u2f/u2fw (Ftools.c:373,383,433,443).
// ---------------------------------------------------------------------------------------------
DG_HDN void left_null_9xk(double* Z, int len, double* q) {
  double vs[8][9];
  double beta[8];
  #pragma unroll 1
  for (int c = 0; c < len; ++c) {
    double nrm = 0.0;
    #pragma unroll 1
    for (int r = c; r < 9; ++r) nrm += Z[r * len + c] * Z[r * len + c];
    nrm = sqrt(nrm);
    for (int r = 0; r < 9; ++r) vs[c][r] = 0.0;
    if (nrm == 0.0) { beta[c] = 0.0; continue; }
    const double x0 = Z[c * len + c];
    const double alpha = (x0 >= 0.0) ? -nrm : nrm;
    #pragma unroll 1
    for (int r = c; r < 9; ++r) vs[c][r] = Z[r * len + c];
    vs[c][c] = x0 - alpha;
    double vn = 0.0;
    #pragma unroll 1
    for (int r = c; r < 9; ++r) vn += vs[c][r] * vs[c][r];
    beta[c] = (vn > 0.0) ? 2.0 / vn : 0.0;
    #pragma unroll 1
    for (int cc = c; cc < len; ++cc) {
      double dot = 0.0;
      #pragma unroll 1
      for (int r = c; r < 9; ++r) dot += vs[c][r] * Z[r * len + cc];
      dot *= beta[c];
      #pragma unroll 1
      for (int r = c; r < 9; ++r) Z[r * len + cc] -= dot * vs[c][r];
    }
  }
  for (int r = 0; r < 9; ++r) q[r] = 0.0;
  q[8] = 1.0;
  #pragma unroll 1
  for (int c = len - 1; c >= 0; --c) {
    double dot = 0.0;
    #pragma unroll 1
    for (int r = c; r < 9; ++r) dot += vs[c][r] * q[r];
    dot *= beta[c];
    #pragma unroll 1
    for (int r = c; r < 9; ++r) q[r] -= dot * vs[c][r];
  }
}

// ---------------------------------------------------------------------------------------------
// Null space of the 7 x 9 system of a 7-point sample by column-pivoted Householder QR and back substitution: the
// reference's alternative minimal solver nullspace_qr7x9 (Ftools.c:594-668, selected at compile time by USE_QR,
// exp_ranF.c:1346-1349), which calls LAPACK dgeqp3_.  dgeqp3 on a 7 x 9 matrix runs its unblocked kernel dlaqp2
// (pivot = first column of largest partial norm, dlarfg reflectors, norm downdating with the sqrt(eps) safeguard);
// that kernel is restated here so that the pivot order -- which fixes WHICH two coordinates of the basis vectors are
// the unit/zero pair -- follows LAPACK's.  A: 7 x 9 row-major; N: two null vectors of 9 (x_1 has 1 at the last pivoted
// column, x_2 at the one before).  Returns 0, or -1 on a zero diagonal of R (as the reference).
// ---------------------------------------------------------------------------------------------
DG_HDN int nullspace_qr7x9(const double* A, double* N) {
  const int rows = 7, cols = 9;
  double T[63], vn1[9], vn2[9];
  int p[9];
  #pragma unroll 1
  for (int i = 0; i < rows; ++i)
    #pragma unroll 1
    for (int j = 0; j < cols; ++j) T[i + rows * j] = A[cols * i + j];
  #pragma unroll 1
  for (int j = 0; j < cols; ++j) {
    double ss = 0.0;
    #pragma unroll 1
    for (int i = 0; i < rows; ++i) ss += T[i + rows * j] * T[i + rows * j];
    vn1[j] = sqrt(ss); vn2[j] = vn1[j]; p[j] = j;
  }
  const double tol3z = 1.4901161193847656e-08;   // sqrt(dlamch('Epsilon')) = sqrt(2^-53)... LAPACK: sqrt(eps), eps = 2^-53
  #pragma unroll 1
  for (int i = 0; i < rows; ++i) {
    int pvt = i;
    #pragma unroll 1
    for (int j = i + 1; j < cols; ++j) if (vn1[j] > vn1[pvt]) pvt = j;
    if (pvt != i) {
      #pragma unroll 1
      for (int r = 0; r < rows; ++r) { const double t = T[r + rows * pvt]; T[r + rows * pvt] = T[r + rows * i]; T[r + rows * i] = t; }
      const int tp = p[pvt]; p[pvt] = p[i]; p[i] = tp;
      vn1[pvt] = vn1[i]; vn2[pvt] = vn2[i];
    }
    // reflector H(i) (dlarfg)
    double tau = 0.0;
    if (i < rows - 1) {
      double xn = 0.0;
      #pragma unroll 1
      for (int r = i + 1; r < rows; ++r) xn += T[r + rows * i] * T[r + rows * i];
      xn = sqrt(xn);
      if (xn != 0.0) {
        const double alpha = T[i + rows * i];
        double beta = sqrt(alpha * alpha + xn * xn);
        if (alpha >= 0.0) beta = -beta;
        tau = (beta - alpha) / beta;
        const double sc = 1.0 / (alpha - beta);
        #pragma unroll 1
        for (int r = i + 1; r < rows; ++r) T[r + rows * i] *= sc;
        T[i + rows * i] = beta;
      }
    }
    // apply H(i)^T to the trailing columns (dlarf, side = left), v = (1, T[i+1.., i])
    if (tau != 0.0) {
      #pragma unroll 1
      for (int j = i + 1; j < cols; ++j) {
        double w = T[i + rows * j];
        #pragma unroll 1
        for (int r = i + 1; r < rows; ++r) w += T[r + rows * i] * T[r + rows * j];
        w *= tau;
        T[i + rows * j] -= w;
        #pragma unroll 1
        for (int r = i + 1; r < rows; ++r) T[r + rows * j] -= w * T[r + rows * i];
      }
    }
    // partial column norms (dlaqp2)
    #pragma unroll 1
    for (int j = i + 1; j < cols; ++j) {
      if (vn1[j] != 0.0) {
        double temp = fabs(T[i + rows * j]) / vn1[j];
        temp = 1.0 - temp * temp;
        if (temp < 0.0) temp = 0.0;
        const double q = vn1[j] / vn2[j];
        const double temp2 = temp * (q * q);
        if (temp2 <= tol3z) {
          if (i < rows - 1) {
            double ss = 0.0;
            #pragma unroll 1
            for (int r = i + 1; r < rows; ++r) ss += T[r + rows * j] * T[r + rows * j];
            vn1[j] = sqrt(ss); vn2[j] = vn1[j];
          } else { vn1[j] = 0.0; vn2[j] = 0.0; }
        } else {
          vn1[j] *= sqrt(temp);
        }
      }
    }
  }
  // back substitution (Ftools.c:646-666)
  double* sol = N;
  #pragma unroll 1
  for (int k = 1; k <= cols - rows; ++k) {
    #pragma unroll 1
    for (int c = rows; c < cols; ++c) sol[p[c]] = 0.0;
    sol[p[cols - k]] = 1.0;
    #pragma unroll 1
    for (int r = rows - 1; r >= 0; --r) {
      if (T[r * rows + r] == 0.0) return -1;
      double a = 0.0;
      #pragma unroll 1
      for (int c = r + 1; c < cols; ++c) a += T[c * rows + r] * sol[p[c]];
      sol[p[r]] = -a / T[r * rows + r];
    }
    sol += cols;
  }
  return 0;
}

// 3x3 inverse in place (row-major) by Gauss-Jordan with partial pivoting.  Returns nonzero when a
// pivot falls below 1e-15 x the largest pivot seen so far (the singularity rule of CCMATH minv,
// matutls/minv.c:11,27), in which case the matrix content is unspecified.
DG_HDN int inv3(double* a) {
  double m[3][6];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { m[i][j] = a[3 * i + j]; m[i][3 + j] = (i == j) ? 1.0 : 0.0; }
  double tq = 0.0;
  for (int c = 0; c < 3; ++c) {
    int best = c;
    double s = fabs(m[c][c]);
    #pragma unroll 1
    for (int r = c + 1; r < 3; ++r) {
      const double t = fabs(m[r][c]);
      if (t > s) { s = t; best = r; }
    }
    tq = tq > s ? tq : s;
    if (s < 1e-15 * tq || s == 0.0) return -1;
    if (best != c)
      for (int k = 0; k < 6; ++k) { const double t = m[c][k]; m[c][k] = m[best][k]; m[best][k] = t; }
    const double inv = 1.0 / m[c][c];
    for (int k = 0; k < 6; ++k) m[c][k] *= inv;
    for (int r = 0; r < 3; ++r) {
      if (r == c) continue;
      const double f = m[r][c];
      for (int k = 0; k < 6; ++k) m[r][k] -= f * m[c][k];
    }
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[3 * i + j] = m[i][3 + j];
  return 0;
}

// In-place inverse of a 3x3 row-major matrix with EXACTLY the arithmetic of CCMATH's minv (matutls/minv.c:10-71):
// column-wise Crout LU with row pivoting (pivot compared with 1e-15 x the largest pivot so far), the two triangular
// factors inverted in place, multiplied, and the row interchanges undone as column interchanges.  Needed bit for bit:
// the symmetric-transfer metrics of the H driver push every correspondence through this inverse, and on scenes whose
// best sample is supported by nothing but its own four points the MSAC scores 4 - O(1e-13) of two exact fits differ
// only by the rounding noise of this routine -- which decides whether the reference schedules one more LO.  Returns -1
// on a vanishing pivot, with the matrix left half-processed exactly as the reference leaves it (its callers ignore the
// return value).
DG_HDN int minv3(double* a) {
  const int n = 3;
  int le[3];
  double q0[3];
  double tq = 0.0;
  const double zr = 1.e-15;
  #pragma unroll 1
  for (int j = 0; j < n; ++j) {
    if (j > 0) {
      for (int i = 0; i < n; ++i) q0[i] = a[i * n + j];
      #pragma unroll 1
      for (int i = 1; i < n; ++i) {
        const int lc = i < j ? i : j;
        double t = 0.0;
        #pragma unroll 1
        for (int k = 0; k < lc; ++k) t += a[i * n + k] * q0[k];
        q0[i] -= t;
      }
      for (int i = 0; i < n; ++i) a[i * n + j] = q0[i];
    }
    double s = fabs(a[j * n + j]);
    int lc = j;
    #pragma unroll 1
    for (int k = j + 1; k < n; ++k) {
      const double t = fabs(a[k * n + j]);
      if (t > s) { s = t; lc = k; }
    }
    tq = tq > s ? tq : s;
    if (s < zr * tq) return -1;
    le[j] = lc;
    if (lc != j) {
      for (int k = 0; k < n; ++k) { const double t = a[j * n + k]; a[j * n + k] = a[lc * n + k]; a[lc * n + k] = t; }
    }
    const double t = 1. / a[j * n + j];
    #pragma unroll 1
    for (int k = j + 1; k < n; ++k) a[k * n + j] *= t;
    a[j * n + j] = t;
  }
  #pragma unroll 1
  for (int j = 1; j < n; ++j)
    #pragma unroll 1
    for (int k = 0; k < j; ++k) a[k * n + j] *= a[j * n + j];
  #pragma unroll 1
  for (int j = 1; j < n; ++j) {
    #pragma unroll 1
    for (int i = 0; i < j; ++i) q0[i] = a[i * n + j];
    #pragma unroll 1
    for (int k = 0; k < j; ++k) {
      double t = 0.0;
      #pragma unroll 1
      for (int i = k; i < j; ++i) t -= a[k * n + i] * q0[i];
      q0[k] = t;
    }
    #pragma unroll 1
    for (int i = 0; i < j; ++i) a[i * n + j] = q0[i];
  }
  #pragma unroll 1
  for (int j = n - 2; j >= 0; --j) {
    int m = n - j - 1;
    #pragma unroll 1
    for (int i = 0; i < m; ++i) q0[i] = a[(j + 1 + i) * n + j];
    #pragma unroll 1
    for (int k = n - 1; k > j; --k) {
      double t = -a[k * n + j];
      #pragma unroll 1
      for (int i = j + 1; i < k; ++i) t -= a[k * n + i] * q0[i - j - 1];
      q0[--m] = t;
    }
    m = n - j - 1;
    #pragma unroll 1
    for (int i = 0; i < m; ++i) a[(j + 1 + i) * n + j] = q0[i];
  }
  #pragma unroll 1
  for (int k = 0; k < n - 1; ++k) {
    for (int i = 0; i < n; ++i) q0[i] = a[i * n + k];
    #pragma unroll 1
    for (int j = 0; j < n; ++j) {
      double t;
      int i;
      if (j > k) { t = 0.0; i = j; } else { t = q0[j]; i = k + 1; }
      #pragma unroll 1
      for (; i < n; ++i) t += a[j * n + i] * q0[i];
      q0[j] = t;
    }
    for (int i = 0; i < n; ++i) a[i * n + k] = q0[i];
  }
  #pragma unroll 1
  for (int j = n - 2; j >= 0; --j) {
    const int lc = le[j];
    for (int k = 0; k < n; ++k) { const double t = a[k * n + j]; a[k * n + j] = a[k * n + lc]; a[k * n + lc] = t; }
  }
  return 0;
}

DG_HD double det3(const double* A) {  // utools.c:196-202
  double r = (A[0] * A[4] * A[8] + A[2] * A[3] * A[7] + A[1] * A[5] * A[6]);
  r -= (A[2] * A[4] * A[6] + A[0] * A[5] * A[7] + A[1] * A[3] * A[8]);
  return r;
}

DG_HD void cross3(double* o, const double* a, const double* b) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// ---------------------------------------------------------------------------------------------
// SuperFastHash over an ordered list of int32 indices, bit-exact with hash.c:4-47 for a byte string
// of 4*n bytes (little endian): used to de-duplicate LO inlier sets (exp_ranF.c:675-686).
// ---------------------------------------------------------------------------------------------
DG_HD uint32_t sfh_init(int n) { return (uint32_t)(4 * n); }
DG_HD uint32_t sfh_word(uint32_t hash, uint32_t w) {
  hash += (w & 0xffffu);
  const uint32_t tmp = ((w >> 16) << 11) ^ hash;
  hash = (hash << 16) ^ tmp;
  hash += hash >> 11;
  return hash;
}
DG_HD uint32_t sfh_final(uint32_t hash) {
  hash ^= hash << 3;
  hash += hash >> 5;
  hash ^= hash << 4;
  hash += hash >> 17;
  hash ^= hash << 25;
  hash += hash >> 6;
  return hash;
}
DG_HD uint32_t superfasthash_i32(const int* idx, int n) {
  if (n <= 0) return 0u;
  uint32_t h = sfh_init(n);
  #pragma unroll 1
  for (int i = 0; i < n; ++i) h = sfh_word(h, (uint32_t)idx[i]);
  return sfh_final(h);
}

// Number of samples for a confidence level (reference nsamples, rtools.c:202-225).
DG_HD int nsamples(int ninl, int ptNum, int samsiz, double conf) {
  double a = 1.0, b = 1.0;
  #pragma unroll 1
  for (int i = 0; i < samsiz; ++i) {
    a *= ninl - i;
    b *= ptNum - i;
  }
  a = a / b;
  if (a < kEps) return kMaxSamples;
  a = 1.0 - a;
  if (a < kEps) return 1;
  b = log(1.0 - conf) / log(a);
  if (b > kMaxSamples) return kMaxSamples;
  return (int)ceil(b);
}

// MSAC gain (reference truncQuad, rtools.c:228-236): width (thr*9)/4.
DG_HD double trunc_quad(double e, double thr) {
  if (thr == 0) return 0.0;
  const double w = thr * 9 / 4;
  if (e >= w) return 0.0;
  return 1 - (e / w);
}

}  // namespace dg
