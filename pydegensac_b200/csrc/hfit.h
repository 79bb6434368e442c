// hfit.h -- CTA-cooperative homography LSQ and residual passes shared by the H engine and DEGENSAC.
#pragma once
#include "common.h"
#include "la.h"
#include "hgeom.h"
#include "block.h"

namespace dg {

// Homography from a list of correspondences (reference u2h, Htools.c:101-133):
//   len < 4  : nothing (h untouched);  len == 4 : the reference's scrambled-transpose branch (see
//   h_from_4pt_u2h_quirk);  len > 4 : Hartley-normalised DLT, normal matrix by block reduction,
//   smallest eigenvector (Jacobi instead of LAPACK dsyev_), de-normalisation.
DG_ENGN void blk_fit_H(const Ctx& c, const int* idx, int len, double* h) {
#ifdef DG_TRACE
  fprintf(stderr, "u2h len=%d\n", len);
#endif
  if (len < 4) return;
  if (len == 4) {
    DG_SYNC();
    if (c.tid == 0) {
      double px1[4], py1[4], px2[4], py2[4], hh[9];
      for (int i = 0; i < 4; ++i) {
        const int p = idx[i];
        px1[i] = c.x1[p]; py1[i] = c.y1[p]; px2[i] = c.x2[p]; py2[i] = c.y2[p];
      }
      h_from_4pt_u2h_quirk(px1, py1, px2, py2, hh);
      for (int i = 0; i < 9; ++i) c.sc->bc[i] = hh[i];
    }
    bc_fetch(c, h, 9);
    return;
  }
  if (len <= 32) {   // small support: whole fit inside warp 0 (see blk_fit_F)
    DG_SYNC();
    if (c.wid == 0) {
      const int W = DG_DEVICE_PASS ? 32 : 1;
      WarpScratch* ws = &c.sc->ws[0];
      double* rows = c.sc->vec;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      #pragma unroll 1
      for (int j = c.lane; j < len; j += W) {
        const int p = idx[j];
        s0 += c.x1[p]; s1 += c.y1[p]; s2 += c.x2[p]; s3 += c.y2[p];
      }
      s0 = wl_sum(s0); s1 = wl_sum(s1); s2 = wl_sum(s2); s3 = wl_sum(s3);
      double A1[3], A2[3];
      A1[1] = s0 / len; A1[2] = s1 / len; A2[1] = s2 / len; A2[2] = s3 / len;
      double d1 = 0.0, d2 = 0.0;
      #pragma unroll 1
      for (int j = c.lane; j < len; j += W) {
        const int p = idx[j];
        double a = c.x1[p] - A1[1], b = c.y1[p] - A1[2];
        d1 += sqrt(a * a + b * b);
        a = c.x2[p] - A2[1]; b = c.y2[p] - A2[2];
        d2 += sqrt(a * a + b * b);
      }
      A1[0] = wl_sum(d1); A2[0] = wl_sum(d2);
      if (A1[0] != 0) A1[0] = len * sqrt(2.0) / A1[0];
      if (A2[0] != 0) A2[0] = len * sqrt(2.0) / A2[0];
      A1[1] *= -A1[0]; A1[2] *= -A1[0];
      A2[1] *= -A2[0]; A2[2] *= -A2[0];
      #pragma unroll 1
      for (int j = c.lane; j < len; j += W) {
        const int p = idx[j];
        double a[3], b[3];
        a[0] = c.x1[p] * A1[0] + A1[1]; a[1] = c.y1[p] * A1[0] + A1[2]; a[2] = 1.0;
        b[0] = c.x2[p] * A2[0] + A2[1]; b[1] = c.y2[p] * A2[0] + A2[2]; b[2] = 1.0;
        double* r0 = rows + 18 * j;
        double* r1 = r0 + 9;
        for (int t = 0; t < 3; ++t) {
          r0[3 * t] = b[t]; r0[3 * t + 1] = 0.0; r0[3 * t + 2] = -a[0] * b[t];
          r1[3 * t] = 0.0;  r1[3 * t + 1] = b[t]; r1[3 * t + 2] = -a[1] * b[t];
        }
      }
      DG_WSYNC();
      #pragma unroll 1
      for (int t = c.lane; t < 45; t += W) {
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= t) ++i;
        const int jj = t - i * (i + 1) / 2;
        double s = 0.0;
        #pragma unroll 1
        for (int r = 0; r < 2 * len; ++r) s += rows[9 * r + i] * rows[9 * r + jj];
        ws->A[9 * i + jj] = s;
        ws->A[9 * jj + i] = s;
      }
      DG_WSYNC();
      warp_smallest_eigvec9(ws, c.lane, W);
      if (c.lane == 0) {
        double q[9];
        for (int i = 0; i < 9; ++i) q[i] = ws->cs[i];
        denorm_H(q, A1, A2);
        for (int i = 0; i < 9; ++i) c.sc->bc[i] = q[i];
      }
    }
    bc_fetch(c, h, 9);
    return;
  }
  double v[kVecRed];
  for (int i = 0; i < 4; ++i) v[i] = 0.0;
  #pragma unroll 1
  for (int j = c.tid; j < len; j += c.nt) {
    const int p = idx[j];
    v[0] += c.x1[p]; v[1] += c.y1[p]; v[2] += c.x2[p]; v[3] += c.y2[p];
  }
  blk_sum_vec(c, v, 4);
  double A1[3], A2[3];
  A1[1] = c.sc->vec_out[0] / len; A1[2] = c.sc->vec_out[1] / len;
  A2[1] = c.sc->vec_out[2] / len; A2[2] = c.sc->vec_out[3] / len;
  v[0] = 0.0; v[1] = 0.0;
  #pragma unroll 1
  for (int j = c.tid; j < len; j += c.nt) {
    const int p = idx[j];
    double a = c.x1[p] - A1[1], b = c.y1[p] - A1[2];
    v[0] += sqrt(a * a + b * b);
    a = c.x2[p] - A2[1]; b = c.y2[p] - A2[2];
    v[1] += sqrt(a * a + b * b);
  }
  blk_sum_vec(c, v, 2);
  A1[0] = c.sc->vec_out[0]; A2[0] = c.sc->vec_out[1];
  if (A1[0] != 0) A1[0] = len * sqrt(2.0) / A1[0];
  if (A2[0] != 0) A2[0] = len * sqrt(2.0) / A2[0];
  A1[1] *= -A1[0]; A1[2] *= -A1[0];
  A2[1] *= -A2[0]; A2[2] *= -A2[0];
  // Normal matrix of the 2 x len DLT rows (reference lin_hgN + cov_mat, Htools.c:60-99).  The two rows of a
  // correspondence are r0 = (b0,0,-a0 b0, b1,0,-a0 b1, b2,0,-a0 b2) and r1 = (0,b0,-a1 b0, 0,b1,-a1 b1, 0,b2,-a1 b2):
  // nine of the 45 entries are structurally zero and every other entry has one or two non-zero products.  Each
  // thread accumulates the 36 live entries in REGISTERS (literal indices only: a 45-entry accumulator indexed by a
  // running counter lived in local memory), then each entry is reduced across the warp, lane (t mod 32) adds it to
  // the warp's slot, and the slots are combined after the barrier.
  {
    double acc[45];
#pragma unroll
    for (int t = 0; t < 45; ++t) acc[t] = 0.0;
    #pragma unroll 1
    for (int j = c.tid; j < len; j += c.nt) {
      const int p = idx[j];
      const double a0 = c.x1[p] * A1[0] + A1[1], a1 = c.y1[p] * A1[0] + A1[2];
      double bb[3];
      bb[0] = c.x2[p] * A2[0] + A2[1]; bb[1] = c.y2[p] * A2[0] + A2[2]; bb[2] = 1.0;
      double r0[9], r1[9];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        r0[3 * t] = bb[t]; r0[3 * t + 1] = 0.0; r0[3 * t + 2] = -a0 * bb[t];
        r1[3 * t] = 0.0;   r1[3 * t + 1] = bb[t]; r1[3 * t + 2] = -a1 * bb[t];
      }
#pragma unroll
      for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int jj = 0; jj <= i; ++jj) {
          if (i % 3 != 1 && jj % 3 != 1) acc[i * (i + 1) / 2 + jj] += r0[i] * r0[jj];
          if (i % 3 != 0 && jj % 3 != 0) acc[i * (i + 1) / 2 + jj] += r1[i] * r1[jj];
        }
    }
    double* slot = c.sc->vec + c.wid * kVecRed;
    DG_SYNC();
    constexpr int W32 = DG_DEVICE_PASS ? 32 : 1;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
      for (int jj = 0; jj <= i; ++jj) {
        const bool live = (i % 3 != 1 && jj % 3 != 1) || (i % 3 != 0 && jj % 3 != 0);
        const double sum = live ? warp_sum(acc[i * (i + 1) / 2 + jj]) : 0.0;
        if (c.lane == (i * (i + 1) / 2 + jj) % W32) slot[i * (i + 1) / 2 + jj] = sum;
      }
    DG_SYNC();
    #pragma unroll 1
    for (int t = c.tid; t < 45; t += c.nt) {
      double sum = 0.0;
      #pragma unroll 1
      for (int wv = 0; wv < c.nw; ++wv) sum += c.sc->vec[wv * kVecRed + t];
      c.sc->vec_out[t] = sum;
    }
    DG_SYNC();
  }
  if (c.wid == 0) {
    WarpScratch* ws = &c.sc->ws[0];
    warp_min_eigvec9_packed(ws, c.sc->vec_out, c.lane, DG_DEVICE_PASS ? 32 : 1);
    if (c.lane == 0) {
      double q[9];
      for (int i = 0; i < 9; ++i) q[i] = ws->cs[i];
      denorm_H(q, A1, A2);
      for (int i = 0; i < 9; ++i) c.sc->bc[i] = q[i];
    }
  }
  bc_fetch(c, h, 9);
}

// Sampson residual row of all correspondences under h (reference HDs over lin_hg, as dHDs does).
DG_ENGN void blk_resid_H_sampson(const Ctx& c, const double* h, double* out) {
  for (int i = c.tid; i < c.N; i += c.nt) st_row(out + i, h_resid_sampson(h, ld_soa(c.x1 + i), ld_soa(c.y1 + i), ld_soa(c.x2 + i), ld_soa(c.y2 + i)));
  DG_SYNC();
}

}  // namespace dg
