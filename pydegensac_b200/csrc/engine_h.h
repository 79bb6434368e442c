// engine_h.h -- LO-RANSAC for homographies, one CTA per image pair.
//
// Replaces the reference's exp_ransacHcustomLAF (exp_ranH.c:470-930, iter_type 4) and its LO
// (exp_inHranicustom :415-467, exp_iterHcustom :291-411) with the same WAVE + ordered REPLAY design as
// engine_f.h: one thread per 4-point sample (orientation test, 8x9 null space, near-singular rejection),
// one warp per surviving model for the MSAC score, exact in-order replay of the models that can change
// the running state.  Extra ordering rule of the H driver: the first LO is forced at the first sample
// >= ITER_SAM that survives all rejections (exp_ranH.c:639-640), so until an LO has run the wave keeps
// every valid model from iteration ITER_SAM-1 on.
#pragma once
#include "common.h"
#include "rng.h"
#include "la.h"
#include "hgeom.h"
#include "block.h"
#include "ffit.h"
#include "hfit.h"
#include "filter32.h"

namespace dg {

#ifdef DG_FILTER_CHECK
static long g_hfilter_checked = 0, g_hfilter_violations = 0;
#endif


struct HParams {
  double th, sym_th, conf, laf_coef;
  double th_laf;    // laf_coef * th (exp_ranH.c:500)
  int do_laf;       // DO_LAF_CHECK (exp_ranH.c:499)
  int max_iters, metric, do_sym;
  uint64_t seed;
  int chunk;
  int final_lsq;    // __FINAL_LSQ__ (exp_ranH.c:16, 866-870)
};

#ifdef DG_FILTER_CHECK
constexpr int kResidClassifyMinN = 64;     // host emulation: exercise the classification in the tests
#else
constexpr int kResidClassifyMinN = 2048;
#endif
#ifdef DG_FILTER_CHECK
static long g_hres_points = 0, g_hres_far = 0, g_hres_violations = 0;   // residual-row classification (host emulation)
#endif
// Residual row of all correspondences under h.  Every consumer of a row compares its entries with thresholds of at most
// TC MWM th = 8 th (inlier lists) and 9/4 of that (MSAC gain), i.e. an entry >= 18 th acts the same whatever its value.
// With the Sampson metric (HDs: ~300 FP64 instructions and 8 divisions per correspondence) the row is therefore built in
// two steps when the FP32 tile of the pair exists: (1) the FP32 LOWER bound of the error of filter32.h classifies every
// correspondence, two per instruction; those whose bound reaches 18 th get +inf; (2) the others -- inliers and near
// misses, a third of a typical pair -- are listed (order irrelevant) and evaluated exactly, densely packed over the
// threads.  A bound that is not an ordinary number (NaN / overflow) sends the correspondence to the exact path.
DG_ENGN void blk_resid_H(const Ctx& c, const HParams& P, Workspace& W, const double* h, double* out) {
  const int metric = P.metric;
  // (short rows: the per-row set-up and the extra barrier cost more than the classification saves -- measured at N = 811)
  if (metric == H_SAMPSON && c.t32 && P.th > 0 && c.N >= kResidClassifyMinN) {
    const double tmax = (kTC * kMWM * 9.0 / 4.0) * P.th;
    HFilter32 hf;
    h_filter_setup(h, *c.t32, tmax, &hf);
    int* list = W.itmp[3];
    DG_SYNC();
    if (c.tid == 0) c.sc->counter[3] = 0;
    DG_SYNC();
#if DG_DEVICE_PASS
    {
      HFilter32x2 f2;
      h_filter_pack(hf, &f2);
      const float4* tp = reinterpret_cast<const float4*>(c.t32->pts);
      const int npair = (c.N + 1) >> 1;
      const unsigned full = 0xffffffffu, lt = (1u << c.lane) - 1u;
      #pragma unroll 1
      for (int qb = c.wid * 32; qb < npair; qb += c.nt) {
        const int q = qb + c.lane;
        const bool live0 = q < npair, live1 = (2 * q + 1) < c.N;
        float r0 = 1.0f, r1 = 1.0f;
        if (live0) upk2(h_filter_raw2(f2, tp[2 * q], tp[2 * q + 1]), r0, r1);
        const bool far0 = (r0 <= 0.0f) && (r0 >= -3.0e38f), far1 = (r1 <= 0.0f) && (r1 >= -3.0e38f);
        const bool n0 = live0 && !far0, n1 = live1 && !far1;
        if (live0 && far0) st_row(out + 2 * q, INFINITY);
        if (live1 && far1) st_row(out + 2 * q + 1, INFINITY);
        const unsigned m0 = __ballot_sync(full, n0), m1 = __ballot_sync(full, n1);
        const int tot = __popc(m0) + __popc(m1);
        int base = 0;
        if (c.lane == 0 && tot) base = atomicAdd(&c.sc->counter[3], tot);
        base = __shfl_sync(full, base, 0);
        if (n0) list[base + __popc(m0 & lt)] = 2 * q;
        if (n1) list[base + __popc(m0) + __popc(m1 & lt)] = 2 * q + 1;
      }
    }
#else
    {
      int cnt = 0;
      for (int i = 0; i < c.N; ++i) {
        const float r = h_filter_raw(hf, c.t32->pts[i]);
        const bool far = (r <= 0.0f) && (r >= -3.0e38f);
        if (far) out[i] = INFINITY; else list[cnt++] = i;
#ifdef DG_FILTER_CHECK
        ++g_hres_points;
        if (far) { ++g_hres_far; if ((g_hres_far & 7) == 0) { const double e = h_resid_sampson(h, c.x1[i], c.y1[i], c.x2[i], c.y2[i]); if (!(e >= tmax)) ++g_hres_violations; } }   // (every 8th: keeps the CPU suite short)
#endif
      }
      c.sc->counter[3] = cnt;
    }
#endif
    DG_SYNC();
    const int n = c.sc->counter[3];
    #pragma unroll 1
    for (int j = c.tid; j < n; j += 2 * c.nt) {   // two independent residual chains per trip
      const int j2 = j + c.nt;
      const bool two = j2 < n;
      const int i = list[j], i2 = list[two ? j2 : j];
      const double e0 = h_resid_sampson(h, c.x1[i], c.y1[i], c.x2[i], c.y2[i]);
      const double e1 = h_resid_sampson(h, c.x1[i2], c.y1[i2], c.x2[i2], c.y2[i2]);
      st_row(out + i, e0);
      if (two) st_row(out + i2, e1);
    }
    DG_SYNC();
    return;
  }
  HSym s;
  if (metric != H_SAMPSON) h_sym_prepare(h, &s);
  #pragma unroll 1
  for (int i = c.tid; i < c.N; i += 2 * c.nt) {   // two independent residual chains per trip (cf. blk_resid_F)
    const int j = i + c.nt;
    const bool two = j < c.N;
    const int jj = two ? j : i;
    const double e0 = h_resid_metric(metric, h, s, ld_soa(c.x1 + i), ld_soa(c.y1 + i), ld_soa(c.x2 + i), ld_soa(c.y2 + i));
    const double e1 = h_resid_metric(metric, h, s, ld_soa(c.x1 + jj), ld_soa(c.y1 + jj), ld_soa(c.x2 + jj), ld_soa(c.y2 + jj));
    st_row(out + i, e0);
    if (two) st_row(out + j, e1);
  }
  DG_SYNC();
}
// symmetric-transfer consistency count over a list (gate: always HDsSymMaxidx, exp_ranH.c:588-597)
DG_ENGN unsigned blk_sym_count_H(const Ctx& c, const double* h, const int* list, int n, double sym_th) {
  HSym s;
  h_sym_prepare(h, &s);
  int cnt = 0;
  #pragma unroll 1
  for (int j = c.tid; j < n; j += c.nt) {
    const int i = list[j];
    if (h_resid_symmax_gate(s, c.x1[i], c.y1[i], c.x2[i], c.y2[i]) <= sym_th) ++cnt;
  }
  return (unsigned)blk_sum_i(c, cnt);
}

// LAF helper counts over a list: how many entries pass laf threshold for helper pair `which` (0: p1, 1: p2)
DG_ENGN int blk_laf_count_H(const Ctx& c, int metric, const double* h, const int* list, int n, double th_laf, int which) {
  HSym s;
  if (metric != H_SAMPSON) h_sym_prepare(h, &s);
  const double* const* L = c.laf + 4 * which;
  int cnt = 0;
  #pragma unroll 1
  for (int j = c.tid; j < n; j += c.nt) {
    const int i = list[j];
    if (h_resid_laf(metric, h, s, c.x1[i], c.y1[i], c.x2[i], c.y2[i], L[0][i], L[1][i], L[2][i], L[3][i]) <= th_laf) ++cnt;
  }
  return blk_sum_i(c, cnt);
}

// hash de-duplication: same table and routine as the F engine (ffit.h)
DG_ENG inline bool hash_seen_elsewhere_h(const Ctx& c, Workspace& W, HashTab& ht, const int* list, int n, int iterID) {
  return hash_seen_elsewhere(c, W, ht, list, n, iterID);
}

// Iterated LSQ with shrinking threshold (reference exp_iterHcustom, exp_ranH.c:291-411; inlLimit = 1e6
// so every fit uses the whole support).
DG_ENGN Score lo_iter_H(const Ctx& c, const HParams& P, Workspace& W, int* e, int* inl, double th, double ths,
                              double* Hio, int iterID, HashTab& ht) {
  int d = e[1];
  double h[9];
  const double dth = (ths - th) / kIlsqIters;
  Score maxS, S = make_score(), Ss;
  maxS = blk_inlidxs(c, W.err[e[4]], th, inl);
#ifdef DG_TRACE
  fprintf(stderr, "iterH start id=%d maxS.I=%u J=%.17g\n", iterID, maxS.I, maxS.J);
#endif
  if (maxS.I < 4) return S;
  S = blk_inlidxs(c, W.err[e[4]], th * kMWM, inl);
  for (int i = 0; i < 9; ++i) h[i] = Hio[i];
  blk_fit_H(c, inl, (int)S.I, h);
  #pragma unroll 1
  for (int it = 0; it < kIlsqIters; ++it) {
    blk_resid_H(c, P, W, h, W.err[d]);
    Ss = blk_inlidxs(c, W.err[d], th, inl);
#ifdef DG_TRACE
    fprintf(stderr, "iterH id=%d it=%d Ss.I=%u Ss.J=%.17g ths=%.17g\n", iterID, it, Ss.I, Ss.J, ths);
#endif
    if (hash_seen_elsewhere_h(c, W, ht, inl, (int)Ss.I, iterID)) return make_score();
    S = blk_inlidxs(c, W.err[d], ths * kMWM, inl);
    if (score_less(maxS, Ss)) {
      maxS = Ss;
      e[1] = e[0];
      e[0] = d;
      d = e[1];
      for (int i = 0; i < 9; ++i) Hio[i] = h[i];
    }
    if (S.I < 4) return maxS;
    blk_fit_H(c, inl, (int)S.I, h);
    ths -= dth;
  }
  blk_resid_H(c, P, W, h, W.err[d]);
  S = blk_inlidxs(c, W.err[d], th, inl);
  if (score_less(maxS, S)) {
    maxS = S;
    e[1] = e[0];
    e[0] = d;
    for (int i = 0; i < 9; ++i) Hio[i] = h[i];
  }
  return maxS;
}

// Inner RANSAC of the LO step (reference exp_inHranicustom, exp_ranH.c:415-467).
DG_ENGN Score lo_inner_H(const Ctx& c, const HParams& P, Workspace& W, int* e, int* inliers, int ninl, double th,
                               double* Hout, int& iterID, DrawCursor& cur, HashTab& ht) {
  Score S, maxS = make_score();
  if (ninl < 8) return maxS;
  int ssiz = ninl / 2;
  if (ssiz > 12) ssiz = 12;
  int t = e[2]; e[2] = e[0]; e[0] = t;
  double h[9];
  for (int i = 0; i < 9; ++i) h[i] = Hout[i];
  #pragma unroll 1
  for (int rep = 0; rep < kRanRep; ++rep) {
    blk_randsubset(c, inliers, ninl, ssiz, cur);
    blk_fit_H(c, inliers + ninl - ssiz, ssiz, h);
    blk_resid_H(c, P, W, h, W.err[e[0]]);
    e[4] = e[0];
    ++iterID;
    S = lo_iter_H(c, P, W, e, W.intbuff, th, kTC * th, h, iterID, ht);
    if (score_less(maxS, S)) {
      maxS = S;
      t = e[2]; e[2] = e[0]; e[0] = t;
      for (int i = 0; i < 9; ++i) Hout[i] = h[i];
    }
  }
  t = e[2]; e[2] = e[0]; e[0] = t;
  return maxS;
}

struct HState {
  Score maxS, maxSs;
  int e[5];
  double H[9];
  int max_sam, iter_cnt, iterID, no_rej;
  int p1_acc;   // the reference's `p1_inliers`: never reset, it accumulates over every LAF gate of the run (exp_ranH.c:501)
  HashTab ht;
  DrawCursor cur;
};

// LO step of iter_type 4 with acceptance (exp_ranH.c:678-747 / :793-861). h: working model in/out.
DG_ENGN bool run_lo_H(const Ctx& c, const HParams& P, Workspace& W, HState& st, double* h) {
  bool new_max = false;
  ++st.iter_cnt;
  const int d = st.e[0];
  Score S = blk_inlidxs(c, W.err[st.e[4]], kTC * P.th * kMWM, W.inliers);
  blk_fit_H(c, W.inliers, (int)S.I, h);
  blk_resid_H(c, P, W, h, W.err[d]);
  S = blk_inlidxs(c, W.err[d], P.th, W.inliers);
#ifdef DG_TRACE_DEV
  if (c.tid == 0) printf("LOH start ninl=%u J=%.12g\n", S.I, S.J);
#endif
  S = lo_inner_H(c, P, W, st.e, W.inliers, (int)S.I, P.th, h, st.iterID, st.cur, st.ht);
#ifdef DG_TRACE_DEV
  if (c.tid == 0) printf("LOH end I=%u J=%.12g\n", S.I, S.J);
#endif
  if (score_less(st.maxS, S) && !h_close_to_singular(h)) {
    bool do_update = true;
    if (P.do_sym) {
      // the reference re-lists the inliers of row `d` (the pre-LO pointer), not of the LO result
      const Score Sc = blk_inlidxs(c, W.err[d], P.th, W.itmp[0]);
      S.Is = blk_sym_count_H(c, h, W.itmp[0], (int)Sc.I, P.sym_th);
      if (S.Is < st.maxS.Is) do_update = false;
    }
    if (do_update && P.do_laf) {   // exp_ranH.c:718-736 / 834-849 (no early exit on the first helper here)
      const Score Sc = blk_inlidxs(c, W.err[d], P.th, W.itmp[0]);
      st.p1_acc += blk_laf_count_H(c, P.metric, h, W.itmp[0], (int)Sc.I, P.th_laf, 0);
      const unsigned c2 = (unsigned)blk_laf_count_H(c, P.metric, h, W.itmp[0], (int)Sc.I, P.th_laf, 1);
      S.Ilafs = c2 < (unsigned)st.p1_acc ? c2 : (unsigned)st.p1_acc;
      if (S.Ilafs < st.maxS.Ilafs) do_update = false;
    }
    if (do_update) {
      const int t = st.e[0]; st.e[0] = st.e[3]; st.e[3] = t;
      st.maxS = S;
      for (int i = 0; i < 9; ++i) st.H[i] = h[i];
      new_max = true;
    }
  }
  return new_max;
}

// WAVE over iterations kbeg..kend; survivors (J > T, or all valid models when passall) are left in
// W.pass in iteration order (one model per iteration, so sorting by k suffices).
DG_ENGN int wave_H(const Ctx& c, const HParams& P, Workspace& W, int kbeg, int kend, double T, bool passall) {
  DG_SYNC();
  if (c.tid == 0) { c.sc->counter[0] = 0; c.sc->counter[1] = 0; }
  DG_SYNC();
  #pragma unroll 1
  for (int k = kbeg + c.tid; k <= kend; k += c.nt) {
    int sel[4];
    minimal_sample<4>(P.seed, (uint32_t)k, c.N, sel);
    // px*: sample in DRAW order (rows of the DLT system); sx*: the reference's samidx order (reverse) for the
    // orientation test.  Both are filled with compile-time indices so they stay in registers.
    double px1[4], py1[4], px2[4], py2[4], sx1[4], sy1[4], sx2[4], sy2[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int p = sel[t];
      px1[t] = c.x1[p]; py1[t] = c.y1[p]; px2[t] = c.x2[p]; py2[t] = c.y2[p];
      sx1[3 - t] = px1[t]; sy1[3 - t] = py1[t]; sx2[3 - t] = px2[t]; sy2[3 - t] = py2[t];
    }
    if (!oriented_ok_H(sx1, sy1, sx2, sy2)) continue;
    double h[9];
    if (!h_from_4pt(px1, py1, px2, py2, h)) continue;
#ifdef DG_TRACE_DEV
    if (k <= 40) printf("WH2 k=%d h=%.10g %.10g %.10g sing=%d\n", k, h[0], h[4], h[8], (int)h_close_to_singular(h));
#endif
    if (h_close_to_singular(h)) continue;
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
  const double w94 = P.th * 9 / 4;
  #pragma unroll 1
  for (int ci = c.wid; ci < ncand; ci += c.nw) {
    bool keep = passall;
    if (!passall) {
      double h[9];
      for (int j = 0; j < 9; ++j) h[j] = W.cand[ci].f[j];
      if (c.t32 && P.metric == H_SAMPSON) {
        // FP32 upper bound of the MSAC score (filter32.h): a superset of the models that matter survives
        HFilter32 hf;
        h_filter_setup(h, *c.t32, w94, &hf);
        float J = 0.0f;
#if DG_DEVICE_PASS
        {
          HFilter32x2 f2;
          h_filter_pack(hf, &f2);
          const float4* tp = reinterpret_cast<const float4*>(c.t32->pts);
          const int npair = (c.N + 1) >> 1;
          const int last = (c.N & 1) ? npair - 1 : -1;
          f32x2 Ja = pk2(0.0f, 0.0f), Jb = pk2(0.0f, 0.0f);
          int i = c.lane;
          #pragma unroll 1
          for (; i + 32 < npair; i += 64) {
            const float4 A0 = tp[2 * i], B0 = tp[2 * i + 1], A1 = tp[2 * i + 64], B1 = tp[2 * i + 65];
            Ja = add2(Ja, h_filter_gain2(f2, A0, B0, i != last));
            Jb = add2(Jb, h_filter_gain2(f2, A1, B1, i + 32 != last));
          }
          if (i < npair) Ja = add2(Ja, h_filter_gain2(f2, tp[2 * i], tp[2 * i + 1], i != last));
          float j0, j1;
          upk2(add2(Ja, Jb), j0, j1);
          J = j0 + j1;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) J += __shfl_xor_sync(0xffffffffu, J, o);
        }
#else
        for (int i = 0; i < c.N; ++i) J += h_filter_gain(hf, c.t32->pts[i]);
#endif
        const double Jup = (double)J * (1.0 + 1.52587890625e-05) + 1e-3;
#ifdef DG_FILTER_CHECK
        {
          double J64 = 0.0;
          for (int i = 0; i < c.N; ++i) {
            const double e = h_resid_sampson(h, c.x1[i], c.y1[i], c.x2[i], c.y2[i]);
            if (e < w94) J64 += 1 - (e / w94);
          }
          ++g_hfilter_checked;
          if (!(Jup >= J64) && J64 == J64) ++g_hfilter_violations;
        }
#endif
        keep = Jup > T - 1e-9 * (1.0 + fabs(T));
      } else {
      HSym s;
      if (P.metric != H_SAMPSON) h_sym_prepare(h, &s);
      double J = 0.0;
#if DG_DEVICE_PASS
      for (int i = c.lane; i < c.N; i += 32) {
#else
      for (int i = 0; i < c.N; ++i) {
#endif
        const double e = h_resid_metric(P.metric, h, s, c.x1[i], c.y1[i], c.x2[i], c.y2[i]);
        if (e < w94) J += 1 - (e / w94);
      }
      J = warp_sum(J);
      keep = J > T - 1e-9 * (1.0 + fabs(T));
      }
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

// REPLAY of one surviving iteration (exp_ranH.c:580-756).
DG_ENGN void replay_iteration_H(const Ctx& c, const HParams& P, Workspace& W, HState& st, int k, const Cand& cd) {
  double h[9];
  for (int j = 0; j < 9; ++j) h[j] = cd.f[j];
  st.cur.seed = P.seed; st.cur.k = (uint32_t)k; st.cur.j = 5;
  bool new_max = false, do_iterate;
  const int d = st.e[0];
  blk_resid_H(c, P, W, h, W.err[d]);
  Score S = blk_inlidxs(c, W.err[d], P.th, W.itmp[0]);
#ifdef DG_TRACE_DEV
  if (c.tid == 0) printf("RH k=%d S.I=%u S.J=%.12g maxS.J=%.12g maxSs.J=%.12g iter=%d\n", k, S.I, S.J, st.maxS.J, st.maxSs.J, st.iter_cnt);
#endif
  if (score_less(st.maxS, S)) {
    if (P.do_sym) {
      S.Is = blk_sym_count_H(c, h, W.itmp[0], (int)S.I, P.sym_th);
      if (S.Is < st.maxS.Is) return;  // `continue`: skips LO scheduling and the termination update
    }
    if (P.do_laf) {   // exp_ranH.c:600-619; the list is the current row's inliers (same as S's list)
      st.p1_acc += blk_laf_count_H(c, P.metric, h, W.itmp[0], (int)S.I, P.th_laf, 0);
      if ((unsigned)st.p1_acc < st.maxS.Ilafs) return;
      const unsigned c2 = (unsigned)blk_laf_count_H(c, P.metric, h, W.itmp[0], (int)S.I, P.th_laf, 1);
      S.Ilafs = c2 < (unsigned)st.p1_acc ? c2 : (unsigned)st.p1_acc;
      if (S.Ilafs < st.maxS.Ilafs) return;
    }
    st.e[0] = st.e[3];
    st.e[3] = d;
    st.maxS = S;
    new_max = true;
    for (int j = 0; j < 9; ++j) st.H[j] = h[j];
  }
  if (score_less(st.maxSs, S)) {
    do_iterate = k > kIterSam;
    st.maxSs = S;
    st.e[4] = d;
  } else {
    do_iterate = false;
  }
  if ((k >= kIterSam) && (st.iter_cnt == 0) && (st.maxSs.I > 4)) do_iterate = true;
  if (do_iterate) {
    if (run_lo_H(c, P, W, st, h)) new_max = true;
  }
  if (new_max) {
    const int new_sam = nsamples((int)st.maxS.I + 1, c.N, 4, P.conf);
    if (new_sam < st.max_sam) st.max_sam = new_sam;
  }
}

// One image pair.  H_out is the RAW core output: column-major, maps image 2 -> image 1 (the Python
// layer applies inv(H.T), utils.py:108).  stats {samples, LO runs, (unused), inliers of best}.
DG_ENGN void ransac_H_pair(const Ctx& c, const HParams& P, Workspace& W, double* H_out, unsigned char* mask_out,
                                 int* stats_out) {
  HState st;
  st.maxS = make_score(); st.maxSs = make_score();
  for (int i = 0; i < 4; ++i) st.e[i] = i;
  st.e[4] = 3;
  for (int i = 0; i < 9; ++i) st.H[i] = 0.0;
  st.max_sam = P.max_iters; st.iter_cnt = 0; st.iterID = 0; st.no_rej = 0; st.p1_acc = 0;
  st.ht.n = 0;
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
    const double T = st.maxS.J < st.maxSs.J ? st.maxS.J : st.maxSs.J;
    const int npass = wave_H(c, P, W, k0 + 1, kend, T, passall);
    bool rewave = false;
    #pragma unroll 1
    for (int pos = 0; pos < npass; ++pos) {
      const Cand& cd = W.cand[W.pass[pos]];
      const int k = cd.k;
      if (k > st.max_sam) break;
      const int lo_before = st.iter_cnt;
      replay_iteration_H(c, P, W, st, k, cd);
      if (k >= st.max_sam) { finished = true; no_sam = k; break; }
      if (passall && st.iter_cnt != lo_before && k < kend) { rewave = true; k0 = k; break; }
    }
    if (finished) break;
    if (rewave) continue;
    k0 = kend;
  }
  if (!finished) no_sam = st.max_sam;
  if ((int)st.cur.k != no_sam) { st.cur.k = (uint32_t)no_sam; st.cur.j = 5; }

  // post-loop LO if none ran (exp_ranH.c:759-862)
  if (st.iter_cnt == 0) {
    double h[9];
    for (int i = 0; i < 9; ++i) h[i] = st.H[i];
    run_lo_H(c, P, W, st, h);
  }

  double* d = W.err[st.e[3]];
  if (P.final_lsq) {   // exp_ranH.c:866-870: LSQ on all inliers of the best model, residuals (and the mask) from it
    const Score Sl = blk_inlidxs(c, d, P.th, W.inliers);
    blk_fit_H(c, W.inliers, (int)Sl.I, st.H);
    blk_resid_H(c, P, W, st.H, d);
  }
  #pragma unroll 1
  for (int j = c.tid; j < c.N; j += c.nt) mask_out[j] = (d[j] <= P.th) ? 1 : 0;
  DG_SYNC();
  if (P.do_sym) {
    const Score Sc = blk_inlidxs(c, d, P.th, W.itmp[0]);
    HSym s;
    h_sym_prepare(st.H, &s);
    #pragma unroll 1
    for (int j = c.tid; j < (int)Sc.I; j += c.nt) {
      const int i = W.itmp[0][j];
      if (h_resid_symmax_gate(s, c.x1[i], c.y1[i], c.x2[i], c.y2[i]) > P.sym_th) mask_out[i] = 0;
    }
    DG_SYNC();
  }
  if (P.do_laf) {   // final LAF prune (exp_ranH.c:889-907): both helper correspondences, indexed by correspondence
    const Score Sc = blk_inlidxs(c, d, P.th, W.itmp[0]);
    HSym s;
    if (P.metric != H_SAMPSON) h_sym_prepare(st.H, &s);
    #pragma unroll 1
    for (int j = c.tid; j < (int)Sc.I; j += c.nt) {
      const int i = W.itmp[0][j];
      const double e1 = h_resid_laf(P.metric, st.H, s, c.x1[i], c.y1[i], c.x2[i], c.y2[i], c.laf[0][i], c.laf[1][i], c.laf[2][i], c.laf[3][i]);
      const double e2 = h_resid_laf(P.metric, st.H, s, c.x1[i], c.y1[i], c.x2[i], c.y2[i], c.laf[4][i], c.laf[5][i], c.laf[6][i], c.laf[7][i]);
      if (e1 > P.th_laf || e2 > P.th_laf) mask_out[i] = 0;
    }
    DG_SYNC();
  }
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
