// filter32.h -- FP32 UPPER-BOUND MSAC score used by the hypothesis wave to discard models.
//
// The wave only has to answer "can this model's MSAC score J exceed the running threshold T?".  Models that
// cannot are dropped; the few that might are re-scored exactly in FP64 by the ordered replay, so the wave may
// use any arithmetic as long as it never UNDER-estimates J.  Here J_up >= J is computed in FP32 (FMA, approximate
// division) on centroid-centred single-precision coordinates, with explicit rounding-error budgets:
//
//   exact (real arithmetic, centred form):  e = r^2 / |g|^2,  r = X2c^T F' X1c,  g = (d r/d x1, d r/d y1, d r/d x2, d r/d y2),
//                                           F' = T2^T F T1 (Sampson error is translation invariant)
//   FP32 evaluation errors:                 |r^ - r| <= Er = c_r * 2^-24 * sum_ij |F'_ij| b2_i b1_j     (b = max |coord| per axis, 1)
//                                           | |g^| - |g| | <= 2 Eg,  Eg = c_g * 2^-24 * max_k sum |F'_.k| b
//   lower bound of the residual:            e_lo = ((|r^| - Er)_+ / (|g^| (1 + 2^-20) + 2 Eg))^2 * (1 - 2^-20)  <= e
//   upper bound of the gain:                g_up = max(0, 1 - e_lo / w) >= truncQuad(e, th)              (w = 9 th / 4)
//   J_up = sum g_up (1 + 2^-16) + 1e-3 >= J.
// c_r = 16 and c_g = 8 are several times the number of roundings on each path (input conversion, 8 fused
// multiply-adds for r, 2 for each gradient component).  tests/ check J_up >= J(FP64) on every scored model.
// Cost: ~25 FP32 instructions per (model, correspondence) instead of ~75 FP64 ones, half the shared-memory bytes.
#pragma once
#include "common.h"
#include "block.h"

namespace dg {

struct alignas(16) Pt32 { float u, v, s, t; };   // (x1 - c1x, y1 - c1y, x2 - c2x, y2 - c2y)

struct Tile32 {
  const Pt32* pts;    // centred single-precision correspondences (shared memory when they fit)
  double cen[4];      // centroids c1x, c1y, c2x, c2y
  double bnd[4];      // max |u|, |v|, |s|, |t|
};

// Fill the FP32 tile of the current pair (block-wide) and its error-model constants.
DG_ENGN void blk_prepare_tile32(const Ctx& c, Pt32* dst, Tile32* T) {
  double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int i = c.tid; i < c.N; i += c.nt) { v[0] += c.x1[i]; v[1] += c.y1[i]; v[2] += c.x2[i]; v[3] += c.y2[i]; }
  blk_sum_vec(c, v, 4);
  for (int k = 0; k < 4; ++k) T->cen[k] = c.sc->vec_out[k] / c.N;
  double m[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int i = c.tid; i < c.N; i += c.nt) {
    Pt32 p;
    p.u = (float)(c.x1[i] - T->cen[0]); p.v = (float)(c.y1[i] - T->cen[1]);
    p.s = (float)(c.x2[i] - T->cen[2]); p.t = (float)(c.y2[i] - T->cen[3]);
    dst[i] = p;
    m[0] = fmax(m[0], fabs((double)p.u)); m[1] = fmax(m[1], fabs((double)p.v));
    m[2] = fmax(m[2], fabs((double)p.s)); m[3] = fmax(m[3], fabs((double)p.t));
  }
  // block max through the sum-reduction scratch: max == -min(-x); done with one lane per warp then a scan
  DG_SYNC();
#if DG_DEVICE_PASS
#pragma unroll
  for (int k = 0; k < 4; ++k)
    for (int o = 16; o > 0; o >>= 1) m[k] = fmax(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
#endif
  if (c.lane == 0) for (int k = 0; k < 4; ++k) c.sc->vec[c.wid * kVecRed + k] = m[k];
  DG_SYNC();
  for (int k = 0; k < 4; ++k) {
    double mm = 0.0;
    for (int w = 0; w < c.nw; ++w) mm = fmax(mm, c.sc->vec[w * kVecRed + k]);
    T->bnd[k] = mm;
  }
  DG_SYNC();
  T->pts = dst;
}

struct FFilter32 {
  float F[9];
  float Er, Eg2, winv, sdk;
  int sym;    // 0: Sampson (FDs), 1: symmetric epipolar (FDsSym):  e = r^2 (1/|g1|^2 + 1/|g2|^2)
};

// Per-model constants (every lane computes them redundantly: they are warp-uniform).
DG_HD void f_filter_setup(int metric, const double* F, const Tile32& T, double w, FFilter32* o) {
  o->sym = (metric == F_SYMM_EPI) ? 1 : 0;
  double G[9], Fp[9];
  for (int i = 0; i < 3; ++i) {
    G[3 * i] = F[3 * i];
    G[3 * i + 1] = F[3 * i + 1];
    G[3 * i + 2] = F[3 * i] * T.cen[0] + F[3 * i + 1] * T.cen[1] + F[3 * i + 2];
  }
  for (int j = 0; j < 3; ++j) {
    Fp[j] = G[j];
    Fp[3 + j] = G[3 + j];
    Fp[6 + j] = T.cen[2] * G[j] + T.cen[3] * G[3 + j] + G[6 + j];
  }
  double mx = 0.0;
  for (int i = 0; i < 9; ++i) mx = fmax(mx, fabs(Fp[i]));
  const double sc = (mx > 0.0 && mx < 1e300) ? 1.0 / mx : 1.0;
  double a[9];
  for (int i = 0; i < 9; ++i) { o->F[i] = (float)(Fp[i] * sc); a[i] = fabs((double)o->F[i]); }
  const double b1[3] = {T.bnd[0], T.bnd[1], 1.0}, b2[3] = {T.bnd[2], T.bnd[3], 1.0};
  double Kr = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Kr += a[3 * i + j] * b2[i] * b1[j];
  // gradient components: d/dx1 = column 0 of F' against X2, d/dy1 = column 1; d/dx2 = row 0 against X1, d/dy2 = row 1
  const double g0 = a[0] * b2[0] + a[3] * b2[1] + a[6], g1 = a[1] * b2[0] + a[4] * b2[1] + a[7];
  const double g2 = a[0] * b1[0] + a[1] * b1[1] + a[2], g3 = a[3] * b1[0] + a[4] * b1[1] + a[5];
  const double Kg = fmax(fmax(g0, g1), fmax(g2, g3));
  const double u24 = 5.9604644775390625e-08;   // 2^-24
  o->Er = (float)(16.0 * u24 * Kr * 1.0001) ;
  o->Eg2 = (float)(2.0 * 8.0 * u24 * Kg * 1.0001);
  o->winv = (float)((1.0 / w) * (1.0 - 1e-6));
  o->sdk = 1.0f + 9.5367431640625e-07f;        // 1 + 2^-20
}

DG_HD float f_filter_gain(const FFilter32& f, const Pt32& p) {
#if DG_DEVICE_PASS
  const float rxc = __fmaf_rn(f.F[0], p.s, __fmaf_rn(f.F[3], p.t, f.F[6]));
  const float ryc = __fmaf_rn(f.F[1], p.s, __fmaf_rn(f.F[4], p.t, f.F[7]));
  const float rwc = __fmaf_rn(f.F[2], p.s, __fmaf_rn(f.F[5], p.t, f.F[8]));
  const float r = __fmaf_rn(p.u, rxc, __fmaf_rn(p.v, ryc, rwc));
  const float rx = __fmaf_rn(f.F[0], p.u, __fmaf_rn(f.F[1], p.v, f.F[2]));
  const float ry = __fmaf_rn(f.F[3], p.u, __fmaf_rn(f.F[4], p.v, f.F[5]));
  const float rl = fmaxf(fabsf(r) - f.Er, 0.0f);
  float e_lo;
  if (f.sym) {
    const float Da = __fmaf_rn(__fsqrt_rn(__fmaf_rn(rxc, rxc, ryc * ryc)), f.sdk, f.Eg2);
    const float Db = __fmaf_rn(__fsqrt_rn(__fmaf_rn(rx, rx, ry * ry)), f.sdk, f.Eg2);
    const float qa = __fdividef(rl, Da) * 0.999999f, qb = __fdividef(rl, Db) * 0.999999f;
    e_lo = __fmaf_rn(qa, qa, qb * qb) * 0.999999f;
  } else {
    const float den = __fmaf_rn(rxc, rxc, __fmaf_rn(ryc, ryc, __fmaf_rn(rx, rx, ry * ry)));
    const float D = __fmaf_rn(__fsqrt_rn(den), f.sdk, f.Eg2);
    const float q = __fdividef(rl, D) * 0.999999f;
    e_lo = q * q;
  }
#else
  const float rxc = fmaf(f.F[0], p.s, fmaf(f.F[3], p.t, f.F[6]));
  const float ryc = fmaf(f.F[1], p.s, fmaf(f.F[4], p.t, f.F[7]));
  const float rwc = fmaf(f.F[2], p.s, fmaf(f.F[5], p.t, f.F[8]));
  const float r = fmaf(p.u, rxc, fmaf(p.v, ryc, rwc));
  const float rx = fmaf(f.F[0], p.u, fmaf(f.F[1], p.v, f.F[2]));
  const float ry = fmaf(f.F[3], p.u, fmaf(f.F[4], p.v, f.F[5]));
  const float rl = fmaxf(fabsf(r) - f.Er, 0.0f);
  float e_lo;
  if (f.sym) {
    const float Da = fmaf(sqrtf(fmaf(rxc, rxc, ryc * ryc)), f.sdk, f.Eg2);
    const float Db = fmaf(sqrtf(fmaf(rx, rx, ry * ry)), f.sdk, f.Eg2);
    const float qa = (rl / Da) * 0.999999f, qb = (rl / Db) * 0.999999f;
    e_lo = fmaf(qa, qa, qb * qb) * 0.999999f;
  } else {
    const float den = fmaf(rxc, rxc, fmaf(ryc, ryc, fmaf(rx, rx, ry * ry)));
    const float D = fmaf(sqrtf(den), f.sdk, f.Eg2);
    const float q = (rl / D) * 0.999999f;
    e_lo = q * q;
  }
#endif
  const float g = 1.0f - e_lo * f.winv;
  return g > 0.0f ? g : 0.0f;   // NaN (0/0) -> 0: such a model has a NaN FP64 score and can never be accepted
}

}  // namespace dg
