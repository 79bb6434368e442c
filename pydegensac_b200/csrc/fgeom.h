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
// Same polynomial as the reference's slcm (Ftools.c:39-81), computed through cofactors; like the
// reference it REPLACES B by A - B so that the caller mixes f = A*r + B*(1-r)  (exp_ranF.c:1366-1368).
DG_HDN void seven_pt_cubic(const double* A, double* B, double* p) {
  double C[9];
  for (int i = 0; i < 9; ++i) C[i] = A[i] - B[i];
  double cb[9], cc[9];
  cb[0] = B[4] * B[8] - B[5] * B[7]; cb[1] = B[5] * B[6] - B[3] * B[8]; cb[2] = B[3] * B[7] - B[4] * B[6];
  cb[3] = B[2] * B[7] - B[1] * B[8]; cb[4] = B[0] * B[8] - B[2] * B[6]; cb[5] = B[1] * B[6] - B[0] * B[7];
  cb[6] = B[1] * B[5] - B[2] * B[4]; cb[7] = B[2] * B[3] - B[0] * B[5]; cb[8] = B[0] * B[4] - B[1] * B[3];
  cc[0] = C[4] * C[8] - C[5] * C[7]; cc[1] = C[5] * C[6] - C[3] * C[8]; cc[2] = C[3] * C[7] - C[4] * C[6];
  cc[3] = C[2] * C[7] - C[1] * C[8]; cc[4] = C[0] * C[8] - C[2] * C[6]; cc[5] = C[1] * C[6] - C[0] * C[7];
  cc[6] = C[1] * C[5] - C[2] * C[4]; cc[7] = C[2] * C[3] - C[0] * C[5]; cc[8] = C[0] * C[4] - C[1] * C[3];
  p[0] = B[0] * cb[0] + B[1] * cb[1] + B[2] * cb[2];
  p[3] = C[0] * cc[0] + C[1] * cc[1] + C[2] * cc[2];
  double s1 = 0.0, s2 = 0.0;
  for (int i = 0; i < 9; ++i) { s1 += cb[i] * C[i]; s2 += cc[i] * B[i]; }
  p[1] = s1;
  p[2] = s2;
  for (int i = 0; i < 9; ++i) B[i] = C[i];
}

// Real roots of po[0] x^3 + po[1] x^2 + po[2] x + po[3] (Cardano / trigonometric), the branch
// structure of the reference's rroots3 (Ftools.c:251-298): returns 1 or 3.
DG_HDN int cubic_real_roots(const double* po, double* r) {
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
