// ffit.h -- per-pair workspace layout and CTA-cooperative F passes (residual rows, symmetric gate
// count, random subsets, normalised 8-point fit) shared by the F engine and DEGENSAC.
#pragma once
#include "common.h"
#include "rng.h"
#include "la.h"
#include "fgeom.h"
#include "block.h"

namespace dg {

struct Cand {
  double f[9];
  int k;
  int root;
};

struct Workspace {
  double* err[4];   // the reference's four residual rows errs[0..3] (physical storage)
  double* errBest;  // errorsBest
  double* w;        // LSQ weights
  double* dtmp[8];  // scratch rows (DEGENSAC / H paths)
  int* inliers;
  int* intbuff;
  int* intbuff_best;
  int* itmp[4];
  unsigned char* btmp[4];
  Cand* cand;
  double* nsbuf;    // per-iteration null-space bases of the wave (16 doubles each)
  int* pass;
  int cand_cap;
  uint32_t* hhash;
  int* hlen;
  int* hid;
  int hcap;
};

struct HashTab { int n; };

struct FParams {
  double th, sym_th, conf, laf_coef;
  int max_iters, metric, degen, do_sym;
  uint64_t seed;
  int chunk;
};

// ------------------------------------------------------------------ block-wide passes over the pair
DG_ENGN void blk_resid_F(const Ctx& c, int metric, const double* F, double* out) {
  DG_PROF_BEGIN(19);
  DG_PROF_COUNT(20, 1);
  for (int i = c.tid; i < c.N; i += c.nt) out[i] = f_resid(metric, F, c.x1[i], c.y1[i], c.x2[i], c.y2[i]);
  DG_SYNC();
  DG_PROF_END(19);
}
DG_ENGN void blk_resid_w_F(const Ctx& c, int metric, const double* F, double* out, double* w) {
  DG_PROF_BEGIN(19);
  DG_PROF_COUNT(20, 1);
  #pragma unroll 1
  for (int i = c.tid; i < c.N; i += c.nt) {
    double e, ww;
    f_resid_w(metric, F, c.x1[i], c.y1[i], c.x2[i], c.y2[i], &e, &ww);
    out[i] = e;
    w[i] = ww;
  }
  DG_SYNC();
  DG_PROF_END(19);
}
// symmetric-epipolar consistency count over an index list (gate at exp_ranF.c:1383-1392)
DG_ENGN unsigned blk_sym_count_F(const Ctx& c, const double* F, const int* list, int n, double sym_th) {
  int cnt = 0;
  #pragma unroll 1
  for (int j = c.tid; j < n; j += c.nt) {
    const int i = list[j];
    if (f_resid_symepi(F, c.x1[i], c.y1[i], c.x2[i], c.y2[i]) <= sym_th) ++cnt;
  }
  return (unsigned)blk_sum_i(c, cnt);
}

// Partial Fisher-Yates permutation of list[0..max_sz) drawing `siz` slots; the subset is the last
// `siz` entries (reference randsubset, rtools.c:25-39).  Sequential by nature: thread 0.
DG_ENGN void blk_randsubset(const Ctx& c, int* list, int max_sz, int siz, DrawCursor& cur) {
  DG_PROF_BEGIN(23);
  DG_SYNC();
  if (c.tid == 0) {
    DrawCursor t = cur;
    #pragma unroll 1
    for (int i = 0; i < siz; ++i) {
      const int s = (int)(next_draw(t) % (uint32_t)(max_sz - i));
      const int j = max_sz - i - 1;
      const int q = list[s];
      list[s] = list[j];
      list[j] = q;
    }
  }
  cur.j += (uint32_t)siz;
  DG_SYNC();
  DG_PROF_END(23);
}

// ---------------------------------------------------------------------------------------------
// F from a list of correspondences: reference u2f / u2fw (Ftools.c:350-458).
//   len > 8 : Hartley normalisation -> 9x9 normal matrix (block reduction) -> smallest eigenvector
//             (Jacobi) -> rank 2 -> de-normalise.
//   len <= 8: unnormalised 9 x len system -> vector orthogonal to its columns -> rank 2.  With weights
//             the reference scales the row-major 9x8 array with stride 9 (`scalmul(Z+i, w, 9, 9)`,
//             Ftools.c:431), i.e. a diagonal pattern; reproduced as is.
// Result is returned to every thread in f[9].
// ---------------------------------------------------------------------------------------------
DG_ENGN void blk_fit_F(const Ctx& c, const int* idx, int len, const double* w, double* f) {
  if (len <= 8) {
    DG_PROF_BEGIN(7);
    DG_PROF_COUNT(27, 1);
    DG_SYNC();
    if (c.wid == 0) {   // warp 0: one lane per column of the 9 x len system, Householder QR across the lanes
      WarpScratch* ws = &c.sc->ws[0];
      const int W = DG_DEVICE_PASS ? 32 : 1;
      #pragma unroll 1
      for (int i = c.lane; i < len; i += W) {
        const int p = idx[i];
        double row[9];
        f_lin_row(c.x1[p], c.y1[p], c.x2[p], c.y2[p], row);
        for (int r = 0; r < 9; ++r) ws->A[r * len + i] = row[r];
      }
      DG_WSYNC();
      if (w) {
        #pragma unroll 1
        for (int i = c.lane; i < len; i += W) {
          const double wi = w[idx[i]];
          for (int t = 0; t < 9; ++t) {
            const int lin = i + 9 * t;
            if (lin < 9 * len) ws->A[lin] *= wi;
          }
        }
        DG_WSYNC();
      }
      // len == 8 (every iterated LSQ of the LO): lane-parallel Gauss-Jordan; rank-deficient or shorter: Householder
      bool have = false;
      if (len == 8) have = warp_null_8x9(ws, c.lane, W);
      if (!have && len > 0) warp_left_null_9xk(ws, len, c.lane, W);
      if (c.lane == 0) {
        double q[9];
        for (int i = 0; i < 9; ++i) q[i] = (len > 0) ? ws->cs[i] : ((i == 8) ? 1.0 : 0.0);
        { DG_PROF_BEGIN(24); enforce_rank2(q); DG_PROF_END(24); }
        for (int i = 0; i < 9; ++i) c.sc->bc[i] = q[i];
      }
    }
    bc_fetch(c, f, 9);
    DG_PROF_END(7);
    return;
  }
  DG_PROF_BEGIN(8);
  DG_PROF_COUNT(28, 1);
  if (len <= 32) {
    // Small support (the 9..14-point inner LO samples, 10-point plane+parallax samples): the whole fit runs
    // inside warp 0 -- one lane per correspondence, no block-wide reduction, rows kept in shared memory.
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
        const double ww = w ? w[p] : 1.0;
        for (int k = 0; k < 3; ++k)
          for (int l = 0; l < 3; ++l) {
            double v = a[l] * b[k];
            if (w) v *= ww;
            rows[9 * j + 3 * k + l] = v;
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
        for (int r = 0; r < len; ++r) s += rows[9 * r + i] * rows[9 * r + jj];
        ws->A[9 * i + jj] = s;
        ws->A[9 * jj + i] = s;
      }
      DG_WSYNC();
      { DG_PROF_BEGIN(29); warp_smallest_eigvec9(ws, c.lane, W); DG_PROF_END(29); }
      if (c.lane == 0) {
        double q[9];
        for (int i = 0; i < 9; ++i) q[i] = ws->cs[i];
        { DG_PROF_BEGIN(24); enforce_rank2(q); DG_PROF_END(24); }
        denorm_F(q, A1, A2);
        for (int i = 0; i < 9; ++i) c.sc->bc[i] = q[i];
      }
    }
    bc_fetch(c, f, 9);
    DG_PROF_END(8);
    return;
  }
  // Hartley normalisation (reference normu, utools.c:7-51)
  DG_PROF_BEGIN(40);
  DG_PROF_COUNT(39, 1);
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
  // normal matrix of the normalised rows (reference lin_fmN + cov_mat, Ftools.c:300-328, utools.c:170-184)
  #pragma unroll 1
  for (int i = 0; i < 45; ++i) v[i] = 0.0;
  #pragma unroll 1
  for (int j = c.tid; j < len; j += c.nt) {
    const int p = idx[j];
    double a[3], b[3], row[9];
    a[0] = c.x1[p] * A1[0] + A1[1]; a[1] = c.y1[p] * A1[0] + A1[2]; a[2] = 1.0;
    b[0] = c.x2[p] * A2[0] + A2[1]; b[1] = c.y2[p] * A2[0] + A2[2]; b[2] = 1.0;
    for (int k = 0; k < 3; ++k)
      for (int l = 0; l < 3; ++l) row[3 * k + l] = a[l] * b[k];
    const double ww = w ? w[p] : 1.0;
    if (w) for (int k = 0; k < 9; ++k) row[k] *= ww;
    int t = 0;
    for (int i = 0; i < 9; ++i)
      #pragma unroll 1
      for (int jj = 0; jj <= i; ++jj) v[t++] += row[i] * row[jj];
  }
  blk_sum_vec(c, v, 45);
  if (c.wid == 0) {   // warp 0: parallel-order Jacobi on the 9x9 normal matrix
    WarpScratch* ws = &c.sc->ws[0];
    { DG_PROF_BEGIN(29); warp_min_eigvec9_packed(ws, c.sc->vec_out, c.lane, DG_DEVICE_PASS ? 32 : 1); DG_PROF_END(29); }
    if (c.lane == 0) {
      double q[9];
      for (int i = 0; i < 9; ++i) q[i] = ws->cs[i];
      { DG_PROF_BEGIN(24); enforce_rank2(q); DG_PROF_END(24); }
      denorm_F(q, A1, A2);
      for (int i = 0; i < 9; ++i) c.sc->bc[i] = q[i];
    }
  }
  bc_fetch(c, f, 9);
  DG_PROF_END(8);
  DG_PROF_END(40);
}

}  // namespace dg
