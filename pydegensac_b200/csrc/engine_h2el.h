// engine_h2el.h -- homography from correspondences of local ELLIPTICAL features: the reference's ransacH2el
// (ranH2el.c:19-191; Chum & Matas, "Homography estimation from correspondences of local elliptical features", ICPR 2012).
// Two ellipse-to-ellipse correspondences fix H (A2toRH, ranH2el.c:220-277: a 14 x 15 linear system in (h, 3 + 3
// auxiliary unknowns), null space by the reference's Gauss-Jordan `nullspace`), scored with the Sampson error on the
// ellipse centres; LO = the reference's older inner RANSAC (inHraniEl ranH2el.c:493-549 -> iterH ranH.c:18-84, shared with
// DEGENSAC's plane LO: degensac.h plane_iter_H).  SURVEY.md section 8(f).4 "unused solvers".
//
// Same WAVE + ordered REPLAY design as engine_h.h.  Differences of this driver that the replay follows: TWO scores per
// model -- J(th) against the best model and J(th * TAU), TAU = (18/7)^2, against the best sample (ranH2el.c:94-116); no
// oriented constraint, no hash de-duplication; rejected samples `continue` past the forced-LO rule and the termination
// update; the first LO is forced at the first non-rejected sample >= ITER_SAM (ranH2el.c:119-121).
// Input rows u10 = (x', y', a', b', c', x, y, a, b, c): image 1 first ("first is u', second u", ranH2el.h:4), each local
// frame the lower-triangular affinity [a 0; b c] at (x, y) (getTransf, ranH2el.c:194-214).
#pragma once
#include "common.h"
#include "rng.h"
#include "la.h"
#include "hgeom.h"
#include "block.h"
#include "ffit.h"
#include "hfit.h"
#include "degensac.h"

namespace dg {

constexpr double kTau = 18.0 * 18.0 / 7.0 / 7.0;   // ranH2el.h:31

struct H2Params {
  double th, conf;
  int max_iters;
  uint64_t seed;
  int chunk;
};

// getTransf (ranH2el.c:194-214): D de-normalises the unit circle to the ellipse of image 1, N normalises the ellipse of
// image 2 to the unit circle; both 3x3, stored column-wise.
DG_HD void el_transf(const double* u10, double* N, double* D) {
  D[0] = u10[2]; D[1] = u10[3]; D[2] = 0; D[3] = 0; D[4] = u10[4]; D[5] = 0; D[6] = u10[0]; D[7] = u10[1]; D[8] = 1;
  N[0] = 1 / u10[7];
  N[1] = -u10[8] / u10[7] / u10[9];
  N[2] = 0; N[3] = 0;
  N[4] = 1 / u10[9];
  N[5] = 0;
  N[6] = -u10[5] / u10[7];
  N[7] = (u10[8] * u10[5] - u10[7] * u10[6]) / u10[7] / u10[9];
  N[8] = 1;
}

// H from two elliptical correspondences (A2toRH with do_norm = 0, ranH2el.c:220-277).  Z is 14 x 15 stored column-wise
// with row stride S = 14 (the reference's `len*7`, len = 2).  Returns true when the null space is one-dimensional.
DG_HDN bool h_from_2el(const double* ua, const double* ub, double* h) {
  const int S = 14;
  double Z[14 * 15], ZT[15 * 15], U[15 * 15];
  #pragma unroll 1
  for (int i = 0; i < 14 * 15; ++i) Z[i] = 0.0;
  #pragma unroll 1
  for (int w = 0; w < 2; ++w) {
    const double* u = w ? ub : ua;
    double N[9], D[9];
    el_transf(u, N, D);
    double* z = Z + 7 * w;                     // Zu (ranH2el.c:358-376): rows 7w .. 7w+6 of the nine H columns
    const double u1 = u[0], u2 = u[1], u4 = u[5], u5 = u[6];
    z[0 + 0 * S] = -1; z[0 + 6 * S] = u1;
    z[1 + 1 * S] = -1; z[1 + 7 * S] = u1;
    z[2 + 2 * S] = -1; z[2 + 6 * S] = -u1 * u4; z[2 + 7 * S] = -u1 * u5;
    z[3 + 3 * S] = -1; z[3 + 6 * S] = u2;
    z[4 + 4 * S] = -1; z[4 + 7 * S] = u2;
    z[5 + 5 * S] = -1; z[5 + 6 * S] = -u2 * u4; z[5 + 7 * S] = -u2 * u5;
    z[6 + 8 * S] = -1; z[6 + 6 * S] = -u4; z[6 + 7 * S] = -u5;
    // Znd(Z + 126 + 49 w, D, N) (ranH2el.c:398-414, A = D, B = N, both read TRANSPOSED through the macros of ranH2el.h)
    double* y = Z + 2 * 7 * 9 + (2 * 7 * 3 + 7) * w;
    const double a1 = D[0], a2 = D[3], a3 = D[6], a4 = D[1], a5 = D[4], a6 = D[7];
    const double b1 = N[0], b2 = N[3], b3 = N[6], b4 = N[1], b5 = N[4], b6 = N[7];
    y[2 + 2 * S] = a3; y[5 + 2 * S] = a6; y[6 + 2 * S] = 1;
    y[0 + 0 * S] = a2 * b1 - a1 * b4; y[1 + 0 * S] = a2 * b2 - a1 * b5; y[2 + 0 * S] = a2 * b3 - a1 * b6;
    y[3 + 0 * S] = a5 * b1 - a4 * b4; y[4 + 0 * S] = a5 * b2 - a4 * b5; y[5 + 0 * S] = a5 * b3 - a4 * b6;
    y[0 + 1 * S] = a1 * b1 + a2 * b4; y[1 + 1 * S] = a1 * b2 + a2 * b5; y[2 + 1 * S] = a1 * b3 + a2 * b6;
    y[3 + 1 * S] = a4 * b1 + a5 * b4; y[4 + 1 * S] = a4 * b2 + a5 * b5; y[5 + 1 * S] = a4 * b3 + a5 * b6;
  }
  // mattr(ZT, Z, 15, 14): row-major 14 x 15, last row zero
  #pragma unroll 1
  for (int r = 0; r < 14; ++r)
    #pragma unroll 1
    for (int cc = 0; cc < 15; ++cc) ZT[r * 15 + cc] = Z[r + S * cc];
  #pragma unroll 1
  for (int i = 14 * 15; i < 15 * 15; ++i) ZT[i] = 0.0;
  #pragma unroll 1
  for (int i = 0; i < 9; ++i) U[i] = 0.0;
  const int nullsize = nullspace15(ZT, U);
  // h = first nine entries of the null vector, transposed (trnm(h, 3))
  h[0] = U[0]; h[1] = U[3]; h[2] = U[6];
  h[3] = U[1]; h[4] = U[4]; h[5] = U[7];
  h[6] = U[2]; h[7] = U[5]; h[8] = U[8];
  return nullsize == 1;
}

// the driver's own singularity test (ranH2el.c:89-92, 147, 180): |det h / h8^3| < 10e-2
DG_HD bool h2el_singular(const double* h) {
  const double v = det3(h);
  double tol = h[8];
  tol = tol * tol * tol;
  return fabs(v / tol) < 10e-2;
}

struct H2State {
  Score maxS, maxSs;
  int e[5];
  double H[9];
  int max_sam, iter_cnt;
  DrawCursor cur;
};

// inHraniEl (ranH2el.c:493-549): with loLimit = 8 the inner sample always has >= 4 correspondences, so the
// ellipse solver branch (AntoRH) is never taken and the repetition is u2h + iterH, as in inHrani (ranH.c:88-135).
DG_ENGN Score lo_inner_H2el(const Ctx& c, Workspace& W, int* e, int* inliers, int ninl, double th, double* Hout, DrawCursor& cur) {
  Score S, maxS = make_score();
  if (ninl < 8) return maxS;
  int ssiz = ninl / 2;
  if (ssiz > 12) ssiz = 12;
  double* rows[4] = {W.err[0], W.err[1], W.err[2], W.err[3]};
  int t = e[2]; e[2] = e[0]; e[0] = t;
  double h[9];
  for (int i = 0; i < 9; ++i) h[i] = Hout[i];
  #pragma unroll 1
  for (int rep = 0; rep < kRanRep; ++rep) {
    blk_randsubset(c, inliers, ninl, ssiz, cur);
    blk_fit_H(c, inliers + ninl - ssiz, ssiz, h);
    blk_resid_H_sampson(c, h, rows[e[0]]);
    e[4] = e[0];
    S = plane_iter_H(c, W, e, rows, W.intbuff, th, kTC * th, h, 0x7fffffffu, cur);
    if (score_less(maxS, S)) {
      maxS = S;
      t = e[2]; e[2] = e[0]; e[0] = t;
      for (int i = 0; i < 9; ++i) Hout[i] = h[i];
    }
  }
  t = e[2]; e[2] = e[0]; e[0] = t;
  return maxS;
}

// LO step with acceptance (ranH2el.c:124-155 in the loop, :166-188 after it).  h: working model in/out.
DG_ENGN bool run_lo_H2el(const Ctx& c, const H2Params& P, Workspace& W, H2State& st, double* h) {
  ++st.iter_cnt;
  int d = st.e[0];
  Score S = blk_inlidxs(c, W.err[st.e[4]], kTC * P.th * kTau, W.inliers);
  blk_fit_H(c, W.inliers, (int)S.I, h);
  blk_resid_H_sampson(c, h, W.err[d]);
  S = blk_inlidxs(c, W.err[d], P.th, W.inliers);
  S = lo_inner_H2el(c, W, st.e, W.inliers, (int)S.I, P.th, h, st.cur);
  if (score_less(st.maxS, S) && !h2el_singular(h)) {
    st.maxS = S;
    d = st.e[0]; st.e[0] = st.e[3]; st.e[3] = d;
    for (int i = 0; i < 9; ++i) st.H[i] = h[i];
    return true;
  }
  return false;
}

// WAVE over iterations kbeg..kend: survivors (J(th) > T1 or J(th TAU) > T2, or every valid model when passall) in W.pass,
// iteration order.  raw: the pair's u10 rows.
DG_ENGN int wave_H2el(const Ctx& c, const H2Params& P, Workspace& W, const double* raw, int kbeg, int kend, double T1,
                      double T2, bool passall) {
  DG_SYNC();
  if (c.tid == 0) { c.sc->counter[0] = 0; c.sc->counter[1] = 0; }
  DG_SYNC();
  #pragma unroll 1
  for (int k = kbeg + c.tid; k <= kend; k += c.nt) {
    int sel[2];
    minimal_sample<2>(P.seed, (uint32_t)k, c.N, sel);
    // samidx = pool + len - 2: samidx[0] is the SECOND draw, samidx[1] the first (ranH2el.c:50)
    double h[9];
    if (!h_from_2el(raw + 10 * (size_t)sel[1], raw + 10 * (size_t)sel[0], h)) continue;
    if (h2el_singular(h)) continue;
    const int slot = atomic_inc_shared(&c.sc->counter[0]);
    if (slot < W.cand_cap) {
      Cand& cd = W.cand[slot];
      for (int j = 0; j < 9; ++j) cd.f[j] = h[j];
      cd.k = k;
      cd.root = 0;
    }
  }
  DG_SYNC();
  int ncand = c.sc->counter[0];
  if (ncand > W.cand_cap) ncand = W.cand_cap;
  const double wa = P.th * 9 / 4, wb = P.th * kTau * 9 / 4;
  #pragma unroll 1
  for (int ci = c.wid; ci < ncand; ci += c.nw) {
    bool keep = passall;
    if (!passall) {
      double h[9];
      for (int j = 0; j < 9; ++j) h[j] = W.cand[ci].f[j];
      double J1 = 0.0, J2 = 0.0;
#if DG_DEVICE_PASS
      for (int i = c.lane; i < c.N; i += 32) {
#else
      for (int i = 0; i < c.N; ++i) {
#endif
        const double e = h_resid_sampson(h, c.x1[i], c.y1[i], c.x2[i], c.y2[i]);
        if (e < wa) J1 += 1 - (e / wa);
        if (e < wb) J2 += 1 - (e / wb);
      }
      J1 = warp_sum(J1);
      J2 = warp_sum(J2);
      keep = (J1 > T1 - 1e-9 * (1.0 + fabs(T1))) || (J2 > T2 - 1e-9 * (1.0 + fabs(T2)));
    }
    if (c.lane == 0 && keep) {
      const int slot = atomic_inc_shared(&c.sc->counter[1]);
      W.pass[slot] = ci;
    }
  }
  DG_SYNC();
  const int npass = c.sc->counter[1];
  if (c.tid == 0) {
    #pragma unroll 1
    for (int a = 1; a < npass; ++a) {
      const int v = W.pass[a];
      const int key = W.cand[v].k;
      int b = a - 1;
      while (b >= 0 && W.cand[W.pass[b]].k > key) { W.pass[b + 1] = W.pass[b]; --b; }
      W.pass[b + 1] = v;
    }
  }
  DG_SYNC();
  return npass;
}

// REPLAY of one surviving iteration (ranH2el.c:94-163).
DG_ENGN void replay_iteration_H2el(const Ctx& c, const H2Params& P, Workspace& W, H2State& st, int k, const Cand& cd) {
  double h[9];
  for (int j = 0; j < 9; ++j) h[j] = cd.f[j];
  st.cur.seed = P.seed; st.cur.k = (uint32_t)k; st.cur.j = 3;
  bool new_max = false, do_iterate = false;
  const int d = st.e[0];
  blk_resid_H_sampson(c, h, W.err[d]);
  Score S = blk_inlidxs(c, W.err[d], P.th, W.inliers);
  if (score_less(st.maxS, S)) {
    st.maxS = S;
    st.e[0] = st.e[3];
    st.e[3] = d;
    for (int j = 0; j < 9; ++j) st.H[j] = h[j];
    new_max = true;
  }
  S = blk_inlidxs(c, W.err[d], P.th * kTau, W.inliers);
  if (score_less(st.maxSs, S)) {
    st.maxSs = S;
    do_iterate = k > kIterSam;
    if (!new_max) { st.e[0] = st.e[2]; st.e[2] = d; }
    st.e[4] = d;
  }
  if (k >= kIterSam && st.iter_cnt == 0 && st.maxSs.I > 4) do_iterate = true;
  if (do_iterate) {
    if (run_lo_H2el(c, P, W, st, h)) new_max = true;
  }
  if (new_max) {
    const int new_sam = nsamples((int)st.maxS.I + 1, c.N, 2, P.conf);
    if (new_sam < st.max_sam) st.max_sam = new_sam;
  }
}

// One pair.  H_out: raw column-major model mapping image 2 -> image 1 (as the other H driver); stats {samples, LO runs,
// 0, inliers of the best model}.
DG_ENGN void ransac_H2el_pair(const Ctx& c, const H2Params& P, Workspace& W, const double* raw, double* H_out,
                              unsigned char* mask_out, int* stats_out) {
  H2State st;
  st.maxS = make_score(); st.maxSs = make_score();
  for (int i = 0; i < 4; ++i) st.e[i] = i;
  st.e[4] = 3;
  for (int i = 0; i < 9; ++i) st.H[i] = 0.0;
  st.max_sam = P.max_iters; st.iter_cnt = 0;
  st.cur.seed = P.seed; st.cur.k = 0; st.cur.j = 1;
  for (int r = 0; r < 4; ++r)
    #pragma unroll 1
    for (int j = c.tid; j < c.N; j += c.nt) W.err[r][j] = 0.0;
  DG_SYNC();

  int k0 = 0, no_sam = 0;
  bool finished = false;
  while (!finished && k0 < st.max_sam) {
    int kend;
    bool passall = false;
    if (st.iter_cnt == 0 && k0 < kIterSam - 1) {
      kend = k0 + P.chunk;
      if (kend > kIterSam - 1) kend = kIterSam - 1;
    } else if (st.iter_cnt == 0) {
      passall = true;
      kend = k0 + 32;
    } else {
      kend = k0 + P.chunk;
    }
    if (kend > st.max_sam) kend = st.max_sam;
    const int npass = wave_H2el(c, P, W, raw, k0 + 1, kend, st.maxS.J, st.maxSs.J, passall);
    bool rewave = false;
    #pragma unroll 1
    for (int pos = 0; pos < npass; ++pos) {
      const Cand& cd = W.cand[W.pass[pos]];
      const int k = cd.k;
      if (k > st.max_sam) break;
      const int lo_before = st.iter_cnt;
      replay_iteration_H2el(c, P, W, st, k, cd);
      if (k >= st.max_sam) { finished = true; no_sam = k; break; }
      if (passall && st.iter_cnt != lo_before && k < kend) { rewave = true; k0 = k; break; }
    }
    if (finished) break;
    if (rewave) continue;
    k0 = kend;
  }
  if (!finished) no_sam = st.max_sam;
  if ((int)st.cur.k != no_sam) { st.cur.k = (uint32_t)no_sam; st.cur.j = 3; }

  if (st.iter_cnt == 0) {   // "If there were no LO's, make at least one NOW!" (ranH2el.c:166-188)
    double h[9];
    for (int i = 0; i < 9; ++i) h[i] = st.H[i];   // (the reference's `h` is whatever the last sample left; u2h overwrites it
    run_lo_H2el(c, P, W, st, h);                  //  whenever the support has >= 4 correspondences)
  }

  const double* d = W.err[st.e[3]];
  #pragma unroll 1
  for (int j = c.tid; j < c.N; j += c.nt) mask_out[j] = (d[j] <= P.th) ? 1 : 0;
  DG_SYNC();
  if (c.tid == 0) {
    for (int i = 0; i < 9; ++i) H_out[i] = st.H[i];
    stats_out[0] = no_sam;
    stats_out[1] = st.iter_cnt;
    stats_out[2] = 0;
    stats_out[3] = (int)st.maxS.I;
  }
  DG_SYNC();
}

}  // namespace dg
