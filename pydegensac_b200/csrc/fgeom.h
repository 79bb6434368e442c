// fgeom.h -- per-thread fundamental-matrix geometry: 7-point solver, oriented epipolar test,
// Sampson / symmetric-epipolar residuals, normalised 8-point pieces.
// Conventions: F[9] row-major with x2^T F x1 = 0; a correspondence is
// (x1,y1) in image 1 and (x2,y2) in image 2, homogeneous 1's implicit.
#pragma once
#include "common.h"
#include "la.h"

namespace dg {

// Row of the 7-point system for one correspondence: f-index 3k+l <-> (x2,y2,1)_k * (x1,y1,1)_l
// (reference lin_fm, Ftools.c:15-37).
DG_HD void f_lin_row(double x1, double y1, double x2, double y2, double* row) {
  row[0] = x2 * x1; row[1] = x2 * y1; row[2] = x2;
  row[3] = y2 * x1; row[4] = y2 * y1; row[5] = y2;
  row[6] = x1;      row[7] = y1;      row[8] = 1.0;
}

// Cubic in r of det(C + r*B) with C = A - B:  p[0] r^3 + p[1] r^2 + p[2] r + p[3].
// This is the reference's slcm (Ftools.c:39-81).  The 7-point cubic is ill-conditioned for a noticeable share of
// random samples (a 1e-16 change of a coefficient moves F by 1e-9), and the DEGENSAC test downstream is
// discontinuous in F, so the coefficients are accumulated in the reference's TERM ORDER: with -fmad=false they
// round identically to the x86-64 build.  Like the reference it REPLACES B by A - B afterwards, so that the
// caller mixes f = A*r + B*(1-r)  (exp_ranF.c:1366-1368).
DG_HD void seven_pt_cubic_inl(const double* A, double* B, double* p) {
  const double a11 = A[0], a12 = A[1], a13 = A[2], a21 = A[3], a22 = A[4], a23 = A[5], a31 = A[6], a32 = A[7], a33 = A[8];
  double b11 = B[0], b12 = B[1], b13 = B[2], b21 = B[3], b22 = B[4], b23 = B[5], b31 = B[6], b32 = B[7], b33 = B[8];
  p[0] = -(b13 * b22 * b31) + b12 * b23 * b31 + b13 * b21 * b32 - b11 * b23 * b32 - b12 * b21 * b33 + b11 * b22 * b33;
  p[1] = -(a33 * b12 * b21) + a32 * b13 * b21 + a33 * b11 * b22 - a31 * b13 * b22 - a32 * b11 * b23 + a31 * b12 * b23 +
         a23 * b12 * b31 - a22 * b13 * b31 - a13 * b22 * b31 + 3 * b13 * b22 * b31 + a12 * b23 * b31 -
         3 * b12 * b23 * b31 - a23 * b11 * b32 + a21 * b13 * b32 + a13 * b21 * b32 - 3 * b13 * b21 * b32 -
         a11 * b23 * b32 + 3 * b11 * b23 * b32 +
         (a22 * b11 - a21 * b12 - a12 * b21 + 3 * b12 * b21 + a11 * b22 - 3 * b11 * b22) * b33;
  p[2] = -(a21 * a33 * b12) + a21 * a32 * b13 + a13 * a32 * b21 - a12 * a33 * b21 + 2 * a33 * b12 * b21 -
         2 * a32 * b13 * b21 - a13 * a31 * b22 + a11 * a33 * b22 - 2 * a33 * b11 * b22 + 2 * a31 * b13 * b22 +
         a12 * a31 * b23 - a11 * a32 * b23 + 2 * a32 * b11 * b23 - 2 * a31 * b12 * b23 + 2 * a13 * b22 * b31 -
         3 * b13 * b22 * b31 - 2 * a12 * b23 * b31 + 3 * b12 * b23 * b31 + a13 * a21 * b32 - 2 * a21 * b13 * b32 -
         2 * a13 * b21 * b32 + 3 * b13 * b21 * b32 + 2 * a11 * b23 * b32 - 3 * b11 * b23 * b32 +
         a23 * (-(a32 * b11) + a31 * b12 + a12 * b31 - 2 * b12 * b31 - a11 * b32 + 2 * b11 * b32) +
         (-(a12 * a21) + 2 * a21 * b12 + 2 * a12 * b21 - 3 * b12 * b21 - 2 * a11 * b22 + 3 * b11 * b22) * b33 +
         a22 * (a33 * b11 - a31 * b13 - a13 * b31 + 2 * b13 * b31 + a11 * b33 - 2 * b11 * b33);
  for (int i = 0; i < 9; ++i) B[i] = A[i] - B[i];
  b11 = B[0]; b12 = B[1]; b13 = B[2]; b21 = B[3]; b22 = B[4]; b23 = B[5]; b31 = B[6]; b32 = B[7]; b33 = B[8];
  p[3] = -(b13 * b22 * b31) + b12 * b23 * b31 + b13 * b21 * b32 - b11 * b23 * b32 - b12 * b21 * b33 + b11 * b22 * b33;
}
DG_HDN void seven_pt_cubic(const double* A, double* B, double* p) { seven_pt_cubic_inl(A, B, p); }

// Real roots of po[0] x^3 + po[1] x^2 + po[2] x + po[3] (Cardano / trigonometric), the branch
// structure of the reference's rroots3 (Ftools.c:251-298): returns 1 or 3.
DG_HD int cubic_real_roots_inl(const double* po, double* r) {
  const double third_pi = 1.0471975511965967;
  const double b = po[1] / po[0];
  const double c = po[2] / po[0];
  const double b2 = b * b;
  const double bt = b / 3;
  const double p = (3 * c - b2) / 9;
  const double q = ((2 * b2 * b) / 27 - b * c / 3 + po[3] / po[0]) / 2;
  const double D = q * q + p * p * p;
  if (D > 0) {
    const double A = sqrt(D) - q;
    if (A > 0) {
      const double v = pow(A, 1.0 / 3);
      r[0] = v - p / v - bt;
    } else {
      const double v = pow(-A, 1.0 / 3);
      r[0] = p / v - v - bt;
    }
    return 1;
  }
  const double e = (q > 0) ? 1.0 : -1.0;
  const double R = e * sqrt(-p);
  const double R2 = R * 2;
  double cosphi = q / (R * R * R);
  if (cosphi > 1) cosphi = 1;
  else if (cosphi < -1) cosphi = -1;
  const double phit = acos(cosphi) / 3;
  r[0] = -R2 * cos(phit) - bt;
  r[1] = R2 * cos(third_pi - phit) - bt;
  r[2] = R2 * cos(third_pi + phit) - bt;
  return 3;
}
DG_HDN int cubic_real_roots(const double* po, double* r) { return cubic_real_roots_inl(po, r); }

// oriented_ok_F for the 7-point sample with every index a literal (registers only); same tests, same order.
DG_HD bool oriented_ok_F7(const double* F, const double (&sy1)[7], const double (&sx2)[7], const double (&sy2)[7]) {
  const double xeps = 1.9984e-15;
  double ec[3];
  cross3(ec, F, F + 6);
  bool big = false;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if ((ec[i] > xeps) || (ec[i] < -xeps)) big = true;
  if (!big) cross3(ec, F + 3, F + 6);
  double sig1 = 0.0;
  bool ok = true;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const double s1 = F[0] * sx2[i] + F[3] * sy2[i] + F[6] * 1.0;
    const double s2 = ec[1] * 1.0 - ec[2] * sy1[i];
    const double sig = s1 * s2;
    if (i == 0) sig1 = sig;
    else if (sig1 * sig < 0) ok = false;
  }
  return ok;
}

// Oriented epipolar constraint over the minimal sample (reference epipole/getorisig/all_ori_valid,
// Ftools.c:461-494).  xs/ys arrays hold the sample in the reference's samidx order.
DG_HD bool oriented_ok_F(const double* F, const double* sx1, const double* sy1, const double* sx2,
                         const double* sy2, int n) {
  const double xeps = 1.9984e-15;
  double ec[3];
  cross3(ec, F, F + 6);
  bool big = false;
  for (int i = 0; i < 3; ++i)
    if ((ec[i] > xeps) || (ec[i] < -xeps)) big = true;
  if (!big) cross3(ec, F + 3, F + 6);
  double sig1 = 0.0;
  #pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const double s1 = F[0] * sx2[i] + F[3] * sy2[i] + F[6] * 1.0;
    const double s2 = ec[1] * 1.0 - ec[2] * sy1[i];
    const double sig = s1 * s2;
    if (i == 0) sig1 = sig;
    else if (sig1 * sig < 0) return false;
  }
  (void)sx1;
  return true;
}

// Residuals.  Operation order is the reference's (Ftools.c:83-101, 147-168) so that a build without
// FMA contraction reproduces the CPU values bit for bit.
struct FRes { double rxc, ryc, r, rx, ry; };
DG_HD FRes f_res_terms(const double* F, double x1, double y1, double x2, double y2) {
  FRes t;
  t.rxc = F[0] * x2 + F[3] * y2 + F[6];
  t.ryc = F[1] * x2 + F[4] * y2 + F[7];
  const double rwc = F[2] * x2 + F[5] * y2 + F[8];
  t.r = (x1 * t.rxc + y1 * t.ryc + rwc);
  t.rx = F[0] * x1 + F[1] * y1 + F[2];
  t.ry = F[3] * x1 + F[4] * y1 + F[5];
  return t;
}
DG_HD double f_resid_sampson(const double* F, double x1, double y1, double x2, double y2) {
  const FRes t = f_res_terms(F, x1, y1, x2, y2);
  return t.r * t.r / (t.rxc * t.rxc + t.ryc * t.ryc + t.rx * t.rx + t.ry * t.ry);
}
DG_HD double f_resid_symepi(const double* F, double x1, double y1, double x2, double y2) {
  const FRes t = f_res_terms(F, x1, y1, x2, y2);
  const double a = t.rxc * t.rxc + t.ryc * t.ryc;
  const double b = t.rx * t.rx + t.ry * t.ry;
  return t.r * t.r * (a + b) / (a * b);
}
DG_HD double f_resid(int metric, const double* F, double x1, double y1, double x2, double y2) {
  return metric == F_SYMM_EPI ? f_resid_symepi(F, x1, y1, x2, y2) : f_resid_sampson(F, x1, y1, x2, y2);
}
// residual + LSQ weight (reference exFDs / exFDsSym, Ftools.c:124-146, 228-250)
DG_HD void f_resid_w(int metric, const double* F, double x1, double y1, double x2, double y2, double* e,
                     double* w) {
  const FRes t = f_res_terms(F, x1, y1, x2, y2);
  if (metric == F_SYMM_EPI) {
    const double a = t.rxc * t.rxc + t.ryc * t.ryc;
    const double b = t.rx * t.rx + t.ry * t.ry;
    const double ww = (a * b) / (a + b);
    *w = ww;
    *e = t.r * t.r / ww;
  } else {
    const double ww = t.rxc * t.rxc + t.ryc * t.ryc + t.rx * t.rx + t.ry * t.ry;
    *e = t.r * t.r / ww;
    *w = 1 / sqrt(ww);
  }
}

// De-normalisation of F estimated on Hartley-normalised points (reference denormF, utools.c:53-70).
// A = {scale, -scale*mean_x, -scale*mean_y}.
DG_HD void denorm_F(double* F, const double* A1, const double* A2) {
  double r = A2[0], x = A2[1], y = A2[2];
  F[6] += x * F[0] + y * F[3];
  F[7] += x * F[1] + y * F[4];
  F[8] += x * F[2] + y * F[5];
  F[0] *= r; F[1] *= r; F[2] *= r;
  F[3] *= r; F[4] *= r; F[5] *= r;
  r = A1[0]; x = A1[1]; y = A1[2];
  F[2] += x * F[0] + y * F[1];
  F[5] += x * F[3] + y * F[4];
  F[8] += x * F[6] + y * F[7];
  F[0] *= r; F[3] *= r; F[6] *= r;
  F[1] *= r; F[4] *= r; F[7] *= r;
}

}  // namespace dg
