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
  double* dtmp[6];  // scratch rows (DEGENSAC / H paths)
  double* laf[8];   // LAF helper correspondences (see Ctx::laf)
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
  double th_laf;    // laf_coef * th (exp_ranF.c:1272)
  int do_laf;       // DO_LAF_CHECK (exp_ranF.c:1271)
  int max_iters, metric, degen, do_sym;
  uint64_t seed;
  int chunk;
  int final_lsq;    // the reference's compile-time __FINAL_LSQ__ (exp_ranF.h:28-29): one more LSQ on the inliers of the best model
};

// ------------------------------------------------------------------ block-wide passes over the pair
// Each thread works on TWO correspondences per trip: the residual is a long FP64 dependency chain (~80 operations,
// 8-16 cycles each), two independent chains interleave in the pipe and their eight loads travel together.
DG_ENGN void blk_resid_F(const Ctx& c, int metric, const double* F, double* out) {
  DG_PROF_BEGIN(19);
  DG_PROF_COUNT(20, 1);
  #pragma unroll 1
  for (int i = c.tid; i < c.N; i += 2 * c.nt) {
    const int j = i + c.nt;
    const bool two = j < c.N;
    const int jj = two ? j : i;
    const double a1 = ld_soa(c.x1 + i), b1 = ld_soa(c.y1 + i), a2 = ld_soa(c.x2 + i), b2 = ld_soa(c.y2 + i);
    const double p1 = ld_soa(c.x1 + jj), q1 = ld_soa(c.y1 + jj), p2 = ld_soa(c.x2 + jj), q2 = ld_soa(c.y2 + jj);
    const double e0 = f_resid(metric, F, a1, b1, a2, b2);
    const double e1 = f_resid(metric, F, p1, q1, p2, q2);
    st_row(out + i, e0);
    if (two) st_row(out + j, e1);
  }
  DG_SYNC();
  DG_PROF_END(19);
}
DG_ENGN void blk_resid_w_F(const Ctx& c, int metric, const double* F, double* out, double* w) {
  DG_PROF_BEGIN(19);
  DG_PROF_COUNT(20, 1);
  #pragma unroll 1
  for (int i = c.tid; i < c.N; i += 2 * c.nt) {
    const int j = i + c.nt;
    const bool two = j < c.N;
    const int jj = two ? j : i;
    const double a1 = ld_soa(c.x1 + i), b1 = ld_soa(c.y1 + i), a2 = ld_soa(c.x2 + i), b2 = ld_soa(c.y2 + i);
    const double p1 = ld_soa(c.x1 + jj), q1 = ld_soa(c.y1 + jj), p2 = ld_soa(c.x2 + jj), q2 = ld_soa(c.y2 + jj);
    double e0, w0, e1, w1;
    f_resid_w(metric, F, a1, b1, a2, b2, &e0, &w0);
    f_resid_w(metric, F, p1, q1, p2, q2, &e1, &w1);
    st_row(out + i, e0);
    st_row(w + i, w0);
    if (two) { st_row(out + j, e1); st_row(w + j, w1); }
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

#if DG_DEVICE_PASS
// Warp flavour of randsubset (siz <= MAXS): lane i draws value i, computes its slot and prefetches both entries of
// its swap; the swaps are then replayed on a register-resident log of (position, value) writes -- later entries
// override earlier ones, as in minimal_sample -- by every lane redundantly, and each lane stores the final value of
// the two positions it touched.  Same result as the sequential loop below, without eight dependent round trips
// to the list (which lives in L2: it was written by other warps a moment ago).
// Latest logged value of position P, or `dflt` when P was never written: log entry t lives on lane t (tp = position,
// tv = value, tp = -1 when empty); entries are appended in lane order, so the highest matching lane is the latest.
__device__ __forceinline__ int log_lookup(int tp, int tv, int P, int dflt) {
  const unsigned m = __ballot_sync(0xffffffffu, tp == P);
  const int v = __shfl_sync(0xffffffffu, tv, m ? 31 - __clz(m) : 0);
  return m ? v : dflt;
}
// randsubset on one warp (siz <= 16), COMPACT: lane i < siz generates draw i and prefetches both entries of its swap;
// the swaps are replayed in order on a log of (position, value) writes that is spread over the lanes (two entries
// per step), so the loop body is a handful of ballots and shuffles; each lane then stores the final value of the two
// positions it touched.  Returns in `mine` (lane q) the value drawn at step q = the entry list[max_sz - 1 - q].
// (An earlier version kept the log in registers and unrolled the O(siz^2) compare chains: 2-8 KB... 32 KB of
// straight-line code per call, which on B200 costs ~6 cycles per instruction to stream -- see DESIGN.md.)
__device__ __noinline__ int warp_subset_draw(int* list, int max_sz, int siz, uint64_t seed, uint32_t k, uint32_t j0, int lane) {
  const unsigned full = 0xffffffffu;
  int s = 0, top = 0, vs = 0, vt = 0;
  if (lane < siz) {
    s = (int)(value31(seed, k, j0 + (uint32_t)lane) % (uint32_t)(max_sz - lane));
    top = max_sz - lane - 1;
    vs = list[s];
    vt = list[top];
  }
  int tp = -1, tv = 0, mine = 0;
  #pragma unroll 1
  for (int q = 0; q < siz; ++q) {
    const int sq = __shfl_sync(full, s, q), tq = __shfl_sync(full, top, q);
    const int a = log_lookup(tp, tv, sq, __shfl_sync(full, vs, q));
    const int b = log_lookup(tp, tv, tq, __shfl_sync(full, vt, q));
    if (lane == 2 * q) { tp = sq; tv = b; }          // list[s]   <- value that sat at the top slot
    if (lane == 2 * q + 1) { tp = tq; tv = a; }      // list[top] <- the drawn value
    if (lane == q) mine = a;
  }
  int fs = vs, ft = vt;
  #pragma unroll 1
  for (int q = 0; q < siz; ++q) {
    const int a = log_lookup(tp, tv, __shfl_sync(full, s, q), __shfl_sync(full, vs, q));
    const int b = log_lookup(tp, tv, __shfl_sync(full, top, q), __shfl_sync(full, vt, q));
    if (lane == q) { fs = a; ft = b; }
  }
  if (lane < siz) { list[s] = fs; list[top] = ft; }
  return mine;
}
#endif

// LAF-consistency count over an index list (gates at exp_ranF.c:1394-1412, 1536-1555, 1664-1683): residual of the two
// helper correspondences under F with the run's own metric (FDS1idx), Ilafs = min(#p2 passing, #p1 passing).
DG_ENGN unsigned blk_laf_count_F(const Ctx& c, int metric, const double* F, const int* list, int n, double th_laf) {
  int c1 = 0, c2 = 0;
  #pragma unroll 1
  for (int j = c.tid; j < n; j += c.nt) {
    const int i = list[j];
    if (f_resid(metric, F, c.laf[0][i], c.laf[1][i], c.laf[2][i], c.laf[3][i]) <= th_laf) ++c1;
    if (f_resid(metric, F, c.laf[4][i], c.laf[5][i], c.laf[6][i], c.laf[7][i]) <= th_laf) ++c2;
  }
  const unsigned p1 = (unsigned)blk_sum_i(c, c1);
  const unsigned p2 = (unsigned)blk_sum_i(c, c2);
  return p2 < p1 ? p2 : p1;
}

// Partial Fisher-Yates permutation of list[0..max_sz) drawing `siz` slots; the subset is the last
// `siz` entries (reference randsubset, rtools.c:25-39).
DG_ENGN void blk_randsubset(const Ctx& c, int* list, int max_sz, int siz, DrawCursor& cur) {
  DG_PROF_BEGIN(23);
  DG_SYNC();
#if DG_DEVICE_PASS
  if (siz <= 16) {
    if (c.wid == 0) (void)warp_subset_draw(list, max_sz, siz, cur.seed, cur.k, cur.j, c.lane);
  } else
#endif
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

// Warp-level F fit on a SMALL support (8 < len <= 32): Hartley normalisation, 9 x 9 normal matrix, smallest
// eigenvector, rank 2, de-normalisation -- all inside the calling warp (one lane per correspondence).  `ws`: the warp's
// own tile; `rows`: 9 * len doubles of scratch (may alias ws->aux: dead before the eigen-solver starts); lane 0 writes
// the model to out[0..8].  Called by warp 0 for the block (blk_fit_F) and by several warps side by side where the
// reference's repetitions are independent (blk_inner_FH).
DG_ENGN void warp_fit_F_small(const Ctx& c, WarpScratch* ws, double* rows, const int* idx, int len, const double* w,
                              double* out) {
  const int W = DG_DEVICE_PASS ? 32 : 1;
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
    for (int i = 0; i < 9; ++i) out[i] = q[i];
  }
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
    if (c.wid == 0) warp_fit_F_small(c, &c.sc->ws[0], c.sc->vec, idx, len, w, c.sc->bc);
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
  // normal matrix of the normalised rows (reference lin_fmN + cov_mat, Ftools.c:300-328, utools.c:170-184).
  // Each thread keeps the rows of up to three correspondences in registers; every one of the 45 entries is formed
  // from them and reduced across the warp at once (no per-thread 45-entry accumulator, hence no local memory),
  // lane 0 adds the warp's partial sum to its slot in shared memory; the slots are combined after the barrier.
  {
    double* slot = c.sc->vec + c.wid * kVecRed;
    DG_SYNC();
    #pragma unroll 1
    for (int t = c.lane; t < 45; t += (DG_DEVICE_PASS ? 32 : 1)) slot[t] = 0.0;
    DG_WSYNC();
    #pragma unroll 1
    for (int base = 0; base < len; base += 3 * c.nt) {
      double r[3][9];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int j = base + q * c.nt + c.tid;
        if (j < len) {
          const int p = idx[j];
          const double a0 = c.x1[p] * A1[0] + A1[1], a1 = c.y1[p] * A1[0] + A1[2];
          const double b0 = c.x2[p] * A2[0] + A2[1], b1 = c.y2[p] * A2[0] + A2[2];
          r[q][0] = a0 * b0; r[q][1] = a1 * b0; r[q][2] = b0;
          r[q][3] = a0 * b1; r[q][4] = a1 * b1; r[q][5] = b1;
          r[q][6] = a0;      r[q][7] = a1;      r[q][8] = 1.0;
          if (w) {
            const double ww = w[p];
#pragma unroll
            for (int k = 0; k < 9; ++k) r[q][k] *= ww;
          }
        } else {
#pragma unroll
          for (int k = 0; k < 9; ++k) r[q][k] = 0.0;
        }
      }
      // every lane holds the warp total after the butterfly; lane (t mod 32) owns entry t, so the 45 slot updates
      // are spread over the lanes instead of queueing on lane 0
#pragma unroll
      for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int jj = 0; jj <= i; ++jj) {
          double sum = r[0][i] * r[0][jj] + r[1][i] * r[1][jj] + r[2][i] * r[2][jj];
          sum = warp_sum(sum);
          constexpr int W32 = DG_DEVICE_PASS ? 32 : 1;
          if (c.lane == (i * (i + 1) / 2 + jj) % W32) slot[i * (i + 1) / 2 + jj] += sum;
        }
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

// ---------------------------------------------------------------------------------------------
// Random 8-subset of list[0..max_sz) + F fit on it: exactly blk_randsubset(list, max_sz, 8) followed by
// blk_fit_F(list + max_sz - 8, 8, w), the step every iteration of the LO's re-weighted LSQ performs (the binding's
// inlLimit = 0).  Device: one trip through warp 0 -- draws, swap replay, the eight coefficient
// rows, the (quirky) weighting, Gauss-Jordan and the rank-2 projection all stay in registers; only the permuted
// list entries and the 9 results touch memory.
// ---------------------------------------------------------------------------------------------
#if DG_DEVICE_PASS
// Warp section of the fused 8-subset draw + 8-point fit (all 32 lanes of ONE warp): result to c.sc->bc[0..8], success
// flag to c.sc->bci[1].  COMPACT CODE ON PURPOSE (rolled loops, literal register indices): see DESIGN.md section 7.
__device__ __noinline__ void warp_sample8_fit(const Ctx& c, int* list, int max_sz, const double* w, uint64_t seed, uint32_t k,
                                              uint32_t j0) {
  bool fast = true;
  const unsigned full = 0xffffffffu;
  const int lane = c.lane;
  DG_PROF_BEGIN(41);
  const int mine = warp_subset_draw(list, max_sz, 8, seed, k, j0, lane);
  DG_PROF_END(41);
  DG_PROF_BEGIN(42);
  // ---- rows: correspondence of row i is list[max_sz - 8 + i] = the value drawn at step 7 - i (held by lane 7 - i)
  const int r = lane & 7;
  const int p = __shfl_sync(full, mine, 7 - r);
  double m[9];
  f_lin_row(c.x1[p], c.y1[p], c.x2[p], c.y2[p], m);
  if (w) {
    // the reference scales the row-major 9 x 8 array with stride 9 (Ftools.c:431): entry (coefficient t, row i)
    // sits at 8 t + i and is multiplied by the weight of correspondence (8 t + i) mod 9 when that is < 8
    const double wi = w[p];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int src = (8 * t + r) % 9;
      const double wv = __shfl_sync(full, wi, src & 7);
      if (src < 8) m[t] *= wv;
    }
  }
  // ---- Gauss-Jordan with partial pivoting by role; after every column the row is rotated left, so the pivot
  //      column is m[0], the right-hand side ends in m[0] after eight steps, zeros are shifted in behind it
  bool used = false;
  int mycol = 8;
  DG_PROF_END(42);
  DG_PROF_BEGIN(43);
  #pragma unroll 1
  for (int col = 0; col < 8; ++col) {
    // pivot row = an unused row whose |entry| is largest in its upper 32 bits (sign cleared; partial pivoting only
    // needs a pivot within a factor ~1 of the largest): one REDUX + one ballot instead of a three-stage
    // shuffle/compare tree.  NaN/Inf keys win and fail the test below.
    const unsigned key = used ? 0u : ((unsigned)__double2hiint(m[0]) & 0x7fffffffu);
    const unsigned kmax = __reduce_max_sync(full, key);
    const int who = (__ffs(__ballot_sync(full, key == kmax)) - 1) & 7;
    const double pv = shfl_d(m[0], who);
    if (!(fabs(pv) > 0.0) || !(fabs(pv) < 1e300) || kmax == 0u) { fast = false; break; }
    const double inv = 1.0 / pv;
    const double fm = m[0];
    const bool piv = (r == who);
#pragma unroll
    for (int j = 1; j < 9; ++j) {
      const double pj = shfl_d(m[j], who) * inv;
      m[j - 1] = piv ? pj : fma(-fm, pj, m[j]);      // eliminate and rotate in one go
    }
    m[8] = 0.0;
    if (piv) { used = true; mycol = col; }
  }
  DG_PROF_END(43);
  DG_PROF_BEGIN(44);
  if (fast) {
    double n2 = m[0] * m[0];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) n2 += __shfl_xor_sync(full, n2, o);
    const double sc = rsqrt(1.0 + n2);
    WarpScratch* ws = &c.sc->ws[0];
    __syncwarp();
    if (lane < 8) ws->cs[mycol] = -m[0] * sc;
    if (lane == 8) ws->cs[8] = sc;
    __syncwarp();
    double n[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) n[i] = ws->cs[i];
    enforce_rank2_inl(n);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 9; ++i) c.sc->bc[i] = n[i];
    }
  }
  if (lane == 0) c.sc->bci[1] = fast ? 1 : 0;
  DG_PROF_END(44);
}
#endif

DG_ENGN void blk_sample8_fit_F(const Ctx& c, int* list, int max_sz, const double* w, DrawCursor& cur, double* f) {
#if DG_DEVICE_PASS
  DG_PROF_BEGIN(7);
  DG_PROF_COUNT(27, 1);
  DG_SYNC();
  if (c.wid == 0) warp_sample8_fit(c, list, max_sz, w, cur.seed, cur.k, cur.j);
  DG_SYNC();
  const bool fast = c.sc->bci[1] != 0;
  cur.j += 8u;
  if (fast) {
#pragma unroll
    for (int i = 0; i < 9; ++i) f[i] = c.sc->bc[i];
    DG_SYNC();
    DG_PROF_END(7);
    return;
  }
  DG_PROF_END(7);
  blk_fit_F(c, list + max_sz - 8, 8, w, f);   // rank-deficient sample: Householder route on the (already permuted) list
#else
  blk_randsubset(c, list, max_sz, 8, cur);
  blk_fit_F(c, list + max_sz - 8, 8, w, f);
#endif
}

// ------------------------------------------------------------------------------ LO hash table
// The reference de-duplicates LO inlier sets with SuperFastHash + a 64-bucket chained table
// (hash.c:49-96, exp_ranF.c:675-686).  Only "(hash,len) seen under this iterID / another iterID /
// never" matters, so a flat list is equivalent.  Returns true when the refinement must abort.
#if DG_DEVICE_PASS
// Warp section of the de-duplication (all 32 lanes of ONE warp): the list is fetched 32 words at a time (coalesced,
// next block in flight while the current one is hashed), every lane runs the same serial SuperFastHash chain on words
// handed round by shuffles, and the table scan is spread over the lanes.  Verdict to c.sc->bci[0]
// (0: inserted, 1: already ours, 2: abort).
__device__ __noinline__ void warp_hash_verdict(const Ctx& c, Workspace& W, int htn, const int* list, int n, int iterID) {
  const unsigned full = 0xffffffffu;
  uint32_t h = 0u;
  if (n > 0) {
    h = sfh_init(n);
    uint32_t v = (c.lane < n) ? (uint32_t)list[c.lane] : 0u;
    #pragma unroll 1
    for (int base = 0; base < n; base += 32) {
      const int nxt = base + 32 + c.lane;
      const uint32_t vn = (nxt < n) ? (uint32_t)list[nxt] : 0u;
      if (n - base >= 32) {
#pragma unroll 4
        for (int j = 0; j < 32; ++j) h = sfh_word(h, __shfl_sync(full, v, j));
      } else {
        #pragma unroll 1
        for (int j = 0; j < n - base; ++j) h = sfh_word(h, __shfl_sync(full, v, j));
      }
      v = vn;
    }
    h = sfh_final(h);
  }
  bool same = false, other = false;
  #pragma unroll 1
  for (int i = c.lane; i < htn; i += 32) {
    if (W.hhash[i] == h && W.hlen[i] == n) {
      if (W.hid[i] == iterID) same = true; else other = true;
    }
  }
  same = __any_sync(full, same);
  other = __any_sync(full, other);
  int verdict = 0;
  if (same) verdict = 1; else if (other) verdict = 2;
  if (c.lane == 0) {
    if (verdict == 0 && htn < W.hcap) { W.hhash[htn] = h; W.hlen[htn] = n; W.hid[htn] = iterID; }
    c.sc->bci[0] = verdict;
  }
}
#endif

DG_ENGN bool hash_seen_elsewhere(const Ctx& c, Workspace& W, HashTab& ht, const int* list, int n, int iterID) {
  DG_PROF_BEGIN(4);
  DG_SYNC();
#if DG_DEVICE_PASS
  if (c.wid == 0) warp_hash_verdict(c, W, ht.n, list, n, iterID);
#else
  if (c.tid == 0) {
    const uint32_t h = superfasthash_i32(list, n);
    int same = 0, other = 0;
    #pragma unroll 1
    for (int i = 0; i < ht.n; ++i) {
      if (W.hhash[i] == h && W.hlen[i] == n) {
        if (W.hid[i] == iterID) same = 1; else other = 1;
      }
    }
    int verdict = 0;  // 0: insert, 1: already ours, 2: abort
    if (same) verdict = 1; else if (other) verdict = 2;
    if (verdict == 0 && ht.n < W.hcap) { W.hhash[ht.n] = h; W.hlen[ht.n] = n; W.hid[ht.n] = iterID; }
    c.sc->bci[0] = verdict;
  }
#endif
  DG_SYNC();
  const int verdict = c.sc->bci[0];
  DG_SYNC();
  if (verdict == 0 && ht.n < W.hcap) ++ht.n;
  DG_PROF_END(4);
  return verdict == 2;
}

#if DG_DEVICE_PASS
// De-duplication of list A and the speculative 8-point fit on the wide list B (a private copy in `spec`) side by side
// on two warps: the serial hash chain (warp 1) and the fit (warp 0) are independent until the verdict is known.
// Abort: true is returned, nothing is committed (`inl`, the draw cursor and f are untouched -- exactly the
// reference's state when it returns after the hash).  Otherwise the permuted list is copied into `inl` (byte-for-byte
// what the reference's in-place randsubset leaves there), the cursor advances by the eight draws and f holds the fit.
DG_ENGN bool blk_hash_and_fit8_F(const Ctx& c, Workspace& W, HashTab& ht, const int* listA, int nA, int iterID, int* spec,
                                 int nB, const double* w, DrawCursor& cur, int* inl, double* f) {
  DG_SYNC();
  if (c.wid == 1) warp_hash_verdict(c, W, ht.n, listA, nA, iterID);
  if (c.wid == 0) warp_sample8_fit(c, spec, nB, w, cur.seed, cur.k, cur.j);
  DG_SYNC();
  const int verdict = c.sc->bci[0];
  const bool fast = c.sc->bci[1] != 0;
  double r[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) r[i] = c.sc->bc[i];
  DG_SYNC();
  if (verdict == 0 && ht.n < W.hcap) ++ht.n;
  if (verdict == 2) return true;
  #pragma unroll 1
  for (int j = c.tid; j < nB; j += c.nt) inl[j] = spec[j];
  cur.j += 8u;
  DG_SYNC();
  if (fast) {
#pragma unroll
    for (int i = 0; i < 9; ++i) f[i] = r[i];
  } else {
    blk_fit_F(c, inl + nB - 8, 8, w, f);   // rank-deficient sample: Householder route on the permuted list
  }
  return false;
}
#endif

}  // namespace dg
