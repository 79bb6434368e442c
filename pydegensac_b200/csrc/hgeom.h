// hgeom.h -- per-thread homography geometry: 4-point DLT rows, orientation / singularity tests,
// Sampson and symmetric transfer residuals, normalised DLT pieces.
// Convention: h[9] is COLUMN-major and maps image 2 -> image 1:
//   x1 = (h0 x2 + h3 y2 + h6) / (h2 x2 + h5 y2 + h8),  y1 = (h1 x2 + h4 y2 + h7) / (same).
#pragma once
#include "common.h"
#include "la.h"

namespace dg {

// The two DLT rows of one correspondence (reference lin_hg, Htools.c:20-58).
DG_HD void h_lin_rows(double x1, double y1, double x2, double y2, double* r0, double* r1) {
  r0[0] = x2;  r0[1] = 0.0; r0[2] = -x1 * x2;
  r0[3] = y2;  r0[4] = 0.0; r0[5] = -x1 * y2;
  r0[6] = 1.0; r0[7] = 0.0; r0[8] = -x1 * 1.0;
  r1[0] = 0.0; r1[1] = x2;  r1[2] = -y1 * x2;
  r1[3] = 0.0; r1[4] = y2;  r1[5] = -y1 * y2;
  r1[6] = 0.0; r1[7] = 1.0; r1[8] = -y1 * 1.0;
}

// x / n for several x with ONE reciprocal: rn = RN(1/n) (__drcp_rn), q0 = RN(x rn), then two exact-residual
// corrections q <- RN(q + (x - q n) rn) (FMA).  With a correctly rounded reciprocal the first correction makes q
// faithful and the second one correctly rounded (Markstein's theorem), i.e. bit-identical to IEEE x / n -- which is
// what the reference computes eight times per correspondence in pinvJ (Htools.c:157-158).  6 instructions per quotient
// instead of ~22.  Valid away from overflow / underflow: the caller falls back to plain division outside the guarded
// range.  (tools: dgb200_debug_div_check compares 10^8 random quotients bit for bit on the device.)
#if defined(__CUDACC__)
__device__ __forceinline__ double div_by_shared_rcp(double x, double n, double rn) {
  double q = x * rn;
  q = fma(fma(-q, n, x), rn, q);
  q = fma(fma(-q, n, x), rn, q);
  return q;
}
#endif

// Closed-form pseudo-inverse of the 2x4 Sampson Jacobian (reference pinvJ, Htools.c:135-159).
DG_HD void h_pinvJ(double a, double b, double c, double d, double e, double* pJ) {
  const double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d, e2 = e * e;
  const double c2pd2 = c2 + d2, ab = a * b, de = d * e;
  const double Q = c * (c2pd2 + e2);
  pJ[0] = -b * de + a * (c2 + e2);
  pJ[1] = b * c2pd2 - a * de;
  pJ[2] = Q;
  pJ[3] = -c * (a * d + b * e);
  pJ[4] = d * (b2 + c2) - ab * e;
  pJ[5] = -ab * d + e * (a2 + c2);
  pJ[6] = pJ[3];
  pJ[7] = c * (a2 + b2 + c2);
  const double N = a * pJ[0] + b * pJ[1] + c * pJ[2];
#if DG_DEVICE_PASS
  {
    // guarded range: |N| and every non-zero numerator well inside the normal range, so that neither the quotients nor
    // the exact residuals x - q N can overflow or lose bits to underflow
    const double an = fabs(N);
    double lo = INFINITY, hi = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double v = fabs(pJ[i]);
      hi = fmax(hi, v);
      lo = (v > 0.0) ? fmin(lo, v) : lo;
    }
    if (an > 1e-100 && an < 1e100 && hi < 1e100 && lo > 1e-100) {
      const double rn = __drcp_rn(N);
#pragma unroll
      for (int i = 0; i < 8; ++i) pJ[i] = div_by_shared_rcp(pJ[i], N, rn);
      return;
    }
  }
#endif
  for (int i = 0; i < 8; ++i) pJ[i] /= N;
}

// Sampson error of a correspondence under h (reference HDs, Htools.c:161-199); the linearised
// residual pair is accumulated in the reference's coefficient order so values agree bit for bit.
DG_HD double h_resid_sampson(const double* H, double x1, double y1, double x2, double y2) {
  double r1 = 0.0, r2 = 0.0;
  r1 += H[0] * x2;
  r2 += H[1] * x2;
  r1 += H[2] * (-x1 * x2);
  r2 += H[2] * (-y1 * x2);
  r1 += H[3] * y2;
  r2 += H[4] * y2;
  r1 += H[5] * (-x1 * y2);
  r2 += H[5] * (-y1 * y2);
  r1 += H[6] * 1.0;
  r2 += H[7] * 1.0;
  r1 += H[8] * (-x1 * 1.0);
  r2 += H[8] * (-y1 * 1.0);
  const double a = H[0] - H[2] * x1;
  const double b = H[3] - H[5] * x1;
  const double c = -H[8] - H[2] * x2 - H[5] * y2;
  const double d = H[1] - H[2] * y1;
  const double e = H[4] - H[5] * y1;
  double pJ[8];
  h_pinvJ(a, b, c, d, e, pJ);
  double p = 0.0;
  for (int j = 0; j < 4; ++j) {
    const double t = pJ[j] * r1 + pJ[j + 4] * r2;
    p += t * t;
  }
  return p;
}

// Forward/backward transfer matrices for the symmetric metrics (Htools.c:202-222 and siblings):
// Hi = h transposed into row-major (maps image2 -> image1), H1 = inverse(Hi) (maps image1 -> image2).
struct HSym { double Hi[9]; double H1[9]; };
DG_HDN void h_sym_prepare(const double* H, HSym* s) {
  s->Hi[0] = H[0]; s->Hi[1] = H[3]; s->Hi[2] = H[6];
  s->Hi[3] = H[1]; s->Hi[4] = H[4]; s->Hi[5] = H[7];
  s->Hi[6] = H[2]; s->Hi[7] = H[5]; s->Hi[8] = H[8];
  for (int i = 0; i < 9; ++i) s->H1[i] = s->Hi[i];
  minv3(s->H1);   // the reference's own inverse, bit for bit (la.h)
}
// d1 = |x2 - H1 x1|^2, d2 = |x1 - Hi x2|^2 ; eps is the 1e-10 the reference adds to some denominators.
DG_HD void h_sym_d1d2(const HSym& s, double x1, double y1, double x2, double y2, double eps, double* d1, double* d2) {
  const double a = s.H1[6] * x1 + s.H1[7] * y1 + s.H1[8] + eps;
  const double b = s.Hi[6] * x2 + s.Hi[7] * y2 + s.Hi[8] + eps;
  double xa = (s.H1[0] * x1 + s.H1[1] * y1 + s.H1[2]) / a;
  double ya = (s.H1[3] * x1 + s.H1[4] * y1 + s.H1[5]) / a;
  double xd = x2 - xa, yd = y2 - ya;
  *d1 = xd * xd + yd * yd;
  xa = (s.Hi[0] * x2 + s.Hi[1] * y2 + s.Hi[2]) / b;
  ya = (s.Hi[3] * x2 + s.Hi[4] * y2 + s.Hi[5]) / b;
  xd = x1 - xa; yd = y1 - ya;
  *d2 = xd * xd + yd * yd;
}
// Full-pass metrics (HDsSymSumSq/Sum add 1e-10 to the denominators, HDsSymMax/MaxSq do not:
// Htools.c:225,233,268,276 vs :310-311,352-353).
DG_HD double h_resid_metric(int metric, const double* H, const HSym& s, double x1, double y1, double x2, double y2) {
  if (metric == H_SAMPSON) return h_resid_sampson(H, x1, y1, x2, y2);
  double d1, d2;
  if (metric == H_SYMM_SQ_SUM || metric == H_SYMM_SUM) {
    h_sym_d1d2(s, x1, y1, x2, y2, 1e-10, &d1, &d2);
    return metric == H_SYMM_SQ_SUM ? d1 + d2 : sqrt(d1) + sqrt(d2);
  }
  h_sym_d1d2(s, x1, y1, x2, y2, 0.0, &d1, &d2);
  const double m = d1 < d2 ? d2 : d1;
  return metric == H_SYMM_SQ_MAX ? m : sqrt(m);
}
// Residual of a LAF helper correspondence as the reference's "i"/"idx" metric variants compute it
// (HDsi/HDsidx and the four HDsi*Sym* / HDs*Sym*idx, Htools.c:372-815).  Two quirks are reproduced:
//   * Sampson: the linearised residual pair comes from the rows Z of the MAIN correspondence (the drivers pass
//     `Z` built from `u` together with the helper array `u_1`/`u_2`), only the Jacobian uses the helper point;
//   * symmetric metrics: helper point only, and ALL four variants add 1e-10 to both denominators (the full-pass
//     Max / MaxSq metrics do not).
DG_HD double h_resid_laf(int metric, const double* H, const HSym& s, double x1, double y1, double x2, double y2,
                         double hx1, double hy1, double hx2, double hy2) {
  if (metric == H_SAMPSON) {
    double r1 = 0.0, r2 = 0.0;
    r1 += H[0] * x2;
    r2 += H[1] * x2;
    r1 += H[2] * (-x1 * x2);
    r2 += H[2] * (-y1 * x2);
    r1 += H[3] * y2;
    r2 += H[4] * y2;
    r1 += H[5] * (-x1 * y2);
    r2 += H[5] * (-y1 * y2);
    r1 += H[6] * 1.0;
    r2 += H[7] * 1.0;
    r1 += H[8] * (-x1 * 1.0);
    r2 += H[8] * (-y1 * 1.0);
    const double a = H[0] - H[2] * hx1;
    const double b = H[3] - H[5] * hx1;
    const double c = -H[8] - H[2] * hx2 - H[5] * hy2;
    const double d = H[1] - H[2] * hy1;
    const double e = H[4] - H[5] * hy1;
    double pJ[8];
    h_pinvJ(a, b, c, d, e, pJ);
    double p = 0.0;
    for (int j = 0; j < 4; ++j) {
      const double t = pJ[j] * r1 + pJ[j + 4] * r2;
      p += t * t;
    }
    return p;
  }
  double d1, d2;
  h_sym_d1d2(s, hx1, hy1, hx2, hy2, 1e-10, &d1, &d2);
  if (metric == H_SYMM_SQ_SUM) return d1 + d2;
  if (metric == H_SYMM_SUM) return sqrt(d1) + sqrt(d2);
  const double m = d1 < d2 ? d2 : d1;
  return metric == H_SYMM_SQ_MAX ? m : sqrt(m);
}

// Symmetric-consistency gate metric: always HDsSymMaxidx, WITH the 1e-10 (Htools.c:734-774).
DG_HD double h_resid_symmax_gate(const HSym& s, double x1, double y1, double x2, double y2) {
  double d1, d2;
  h_sym_d1d2(s, x1, y1, x2, y2, 1e-10, &d1, &d2);
  return sqrt(d1 < d2 ? d2 : d1);
}

// Orientation test of a 4-point sample (reference all_Hori_valid, Htools.c:821-848); arrays are in the
// reference's samidx order a,b,c,d.
DG_HD bool oriented_ok_H(const double* sx1, const double* sy1, const double* sx2, const double* sy2) {
  double A[4][3], B[4][3], p[3], q[3];
  for (int i = 0; i < 4; ++i) {
    A[i][0] = sx1[i]; A[i][1] = sy1[i]; A[i][2] = 1.0;
    B[i][0] = sx2[i]; B[i][1] = sy2[i]; B[i][2] = 1.0;
  }
  cross3(p, A[0], A[1]);
  cross3(q, B[0], B[1]);
  if ((p[0] * A[2][0] + p[1] * A[2][1] + p[2] * A[2][2]) * (q[0] * B[2][0] + q[1] * B[2][1] + q[2] * B[2][2]) < 0) return false;
  if ((p[0] * A[3][0] + p[1] * A[3][1] + p[2] * A[3][2]) * (q[0] * B[3][0] + q[1] * B[3][1] + q[2] * B[3][2]) < 0) return false;
  cross3(p, A[2], A[3]);
  cross3(q, B[2], B[3]);
  if ((p[0] * A[0][0] + p[1] * A[0][1] + p[2] * A[0][2]) * (q[0] * B[0][0] + q[1] * B[0][1] + q[2] * B[0][2]) < 0) return false;
  if ((p[0] * A[1][0] + p[1] * A[1][1] + p[2] * A[1][2]) * (q[0] * B[1][0] + q[1] * B[1][1] + q[2] * B[1][2]) < 0) return false;
  return true;
}

// Near-singular rejection (reference HcloseToSingular, exp_ranH.c:29-44).
DG_HD bool h_close_to_singular(const double* h) {
  const double v = det3(h);
  double tol = h[8];
  if (tol == 0) {
    for (int i = 0; i < 9; ++i) tol += h[i] * h[i];
    tol = sqrt(tol);
    tol *= 0.001;
  }
  tol = tol * tol * tol;
  return fabs(v / tol) < 1e-2;
}

// De-normalisation of a column-major h (reference denormH, utools.c:72-92).
DG_HD void denorm_H(double* F, const double* A1, const double* A2) {
  double r = A2[0], x = A2[1], y = A2[2];
  F[6] += x * F[0] + y * F[3];
  F[7] += x * F[1] + y * F[4];
  F[8] += x * F[2] + y * F[5];
  F[0] *= r; F[1] *= r; F[2] *= r;
  F[3] *= r; F[4] *= r; F[5] *= r;
  r = 1 / A1[0]; x = -A1[1] * r; y = -A1[2] * r;
  #pragma unroll 1
  for (int i = 0; i < 9; i += 3) {
    F[i] = r * F[i] + x * F[i + 2];
    F[i + 1] = r * F[i + 1] + y * F[i + 2];
  }
}

// Exact 4-point homography (null space of the 8x9 DLT system; exp_ranH.c:551-566).
// Points are given in DRAW order.  Returns false unless the null space is one-dimensional.
DG_HDN bool h_from_4pt(const double* px1, const double* py1, const double* px2, const double* py2, double* h) {
  double M[81], sol[81];
  for (int i = 0; i < 4; ++i) h_lin_rows(px1[i], py1[i], px2[i], py2[i], M + 18 * i, M + 18 * i + 9);
  for (int i = 72; i < 81; ++i) M[i] = 0.0;
  const int ns = nullspace9(M, sol);
  for (int i = 0; i < 9; ++i) h[i] = sol[i];
  return ns == 1;
}

// What the reference's u2h does for len == 4 (Htools.c:108-116): lin_hg writes the 8x9 system with a
// column stride of 8, the buffer is then transposed as if it were 9x9 (stride 9) and the last row is
// zeroed, so the "null space" is taken of a scrambled matrix (9 of its entries come from uninitialised
// stack, taken as 0 here).  The outcome is a model without support, i.e. that LO repetition is a no-op;
// reproducing it keeps the LO trajectory identical to the reference's.
#ifdef DG_EIG_STATS
static long g_u2h4_calls = 0;   // (host emulation: lets the tests tell which inputs reach this undefined corner of the reference)
#endif
DG_HDN void h_from_4pt_u2h_quirk(const double* px1, const double* py1, const double* px2, const double* py2, double* h) {
#ifdef DG_EIG_STATS
  ++g_u2h4_calls;
#endif
  double Z[81], T[81], sol[81];
  #pragma unroll 1
  for (int i = 0; i < 81; ++i) { Z[i] = 0.0; sol[i] = 0.0; }
  for (int i = 0; i < 4; ++i) {
    double r0[9], r1[9];
    h_lin_rows(px1[i], py1[i], px2[i], py2[i], r0, r1);
    for (int c = 0; c < 9; ++c) { Z[8 * c + 2 * i] = r0[c]; Z[8 * c + 2 * i + 1] = r1[c]; }
  }
  for (int r = 0; r < 9; ++r)
    for (int c = 0; c < 9; ++c) T[9 * r + c] = Z[9 * c + r];
  for (int i = 72; i < 81; ++i) T[i] = 0.0;
  nullspace9(T, sol);
  for (int i = 0; i < 9; ++i) h[i] = sol[i];
}

}  // namespace dg
