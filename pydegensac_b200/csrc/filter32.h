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
//   lower bound of the residual:            D = |g^| (1 + 2^-20) + 2 Eg >= |g|;  e >= ((|r^| - Er)_+ / D)^2, and, SQUARE-ROOT FREE,
//                                           D^2 <= |g^|^2 c1 + c2  with  c1 = (1 + 2^-19)(1 + 2^-12), c2 = (2 Eg)^2 (1 + 2^12)
//                                           (2ab <= 2^-12 a^2 + 2^12 b^2), so  e_lo = (|r^| - Er)_+^2 / (|g^|^2 c1 + c2) (1 - 2^-17) <= e
//   upper bound of the gain:                g_up = max(0, 1 - e_lo / w) >= truncQuad(e, th)              (w = 9 th / 4)
//   J_up = sum g_up (1 + 2^-16) + 1e-3 >= J.
// On the device the loop runs on PACKED pairs of correspondences with Blackwell's two-wide FP32 instructions
// (fma.rn.f32x2 / mul / add -> FFMA2, FMUL2, FADD2 in SASS): the tile is stored pair-interleaved
// {u0 u1 v0 v1 | s0 s1 t0 t1} so that two LDS.128 deliver the four operand pairs, ~16 instructions per (model,
// correspondence) instead of ~34.
// c_r = 16 and c_g = 8 are several times the number of roundings on each path (input conversion, 8 fused
// multiply-adds for r, 2 for each gradient component).  tests/ check J_up >= J(FP64) on every scored model.
// Cost: ~16 FP32 instructions per (model, correspondence) instead of ~75 FP64 ones, half the shared-memory bytes.
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
#if DG_DEVICE_PASS
    {   // pair-interleaved: correspondences 2q and 2q+1 share two float4 {u0 u1 v0 v1} {s0 s1 t0 t1}
      float* base = reinterpret_cast<float*>(dst) + 8 * (i >> 1) + (i & 1);
      base[0] = p.u; base[2] = p.v; base[4] = p.s; base[6] = p.t;
      if (i == c.N - 1 && !(i & 1)) { base[1] = 0.f; base[3] = 0.f; base[5] = 0.f; base[7] = 0.f; }   // odd N: padding slot (masked in the loop)
    }
#else
    dst[i] = p;
#endif
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
  float Er, c1, c2, winv;
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
  const double Eg2 = 2.0 * 8.0 * u24 * Kg * 1.0001;
  o->c1 = (float)((1.0 + 1.9073486328125e-06) * (1.0 + 2.44140625e-04) * (1.0 + 1e-6));   // (1 + 2^-19)(1 + 2^-12), rounded up
  o->c2 = (float)(Eg2 * Eg2 * 4097.0 * 1.0001);
  o->winv = (float)((1.0 / w) * (1.0 - 7.62939453125e-06) * (1.0 - 1e-6));               // carries the (1 - 2^-17) of e_lo
}

DG_HD float f_filter_gain(const FFilter32& f, const Pt32& p) {
#if DG_DEVICE_PASS
#define DG_FMAF __fmaf_rn
#define DG_RCPF(x) __fdividef(1.0f, (x))
#else
#define DG_FMAF fmaf
#define DG_RCPF(x) (1.0f / (x))
#endif
  const float rxc = DG_FMAF(f.F[0], p.s, DG_FMAF(f.F[3], p.t, f.F[6]));
  const float ryc = DG_FMAF(f.F[1], p.s, DG_FMAF(f.F[4], p.t, f.F[7]));
  const float rwc = DG_FMAF(f.F[2], p.s, DG_FMAF(f.F[5], p.t, f.F[8]));
  const float r = DG_FMAF(p.u, rxc, DG_FMAF(p.v, ryc, rwc));
  const float rx = DG_FMAF(f.F[0], p.u, DG_FMAF(f.F[1], p.v, f.F[2]));
  const float ry = DG_FMAF(f.F[3], p.u, DG_FMAF(f.F[4], p.v, f.F[5]));
  const float rl = fmaxf(fabsf(r) - f.Er, 0.0f);
  const float num = rl * rl;
  float e_lo;
  if (f.sym) {
    const float a = DG_FMAF(rxc, rxc, ryc * ryc), b = DG_FMAF(rx, rx, ry * ry);
    e_lo = num * (DG_RCPF(DG_FMAF(a, f.c1, f.c2)) + DG_RCPF(DG_FMAF(b, f.c1, f.c2)));
  } else {
    const float den = DG_FMAF(rxc, rxc, DG_FMAF(ryc, ryc, DG_FMAF(rx, rx, ry * ry)));
    e_lo = num * DG_RCPF(DG_FMAF(den, f.c1, f.c2));
  }
#undef DG_FMAF
#undef DG_RCPF
  const float g = 1.0f - e_lo * f.winv;
  return g > 0.0f ? g : 0.0f;   // NaN (0/0) -> 0: such a model has a NaN FP64 score and can never be accepted
}

#if DG_DEVICE_PASS
// ---- packed pairs: Blackwell FFMA2 / FMUL2 / FADD2 (PTX fma.rn.f32x2 ...), operands are 64-bit register pairs
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

struct FFilter32x2 {     // per-model constants broadcast into both halves
  f32x2 F[9];
  f32x2 c1, c2, nwinv, one;
  float Er;
  int sym;
};
__device__ __forceinline__ void f_filter_pack(const FFilter32& f, FFilter32x2* o) {
#pragma unroll
  for (int i = 0; i < 9; ++i) o->F[i] = pk2(f.F[i], f.F[i]);
  o->c1 = pk2(f.c1, f.c1); o->c2 = pk2(f.c2, f.c2); o->nwinv = pk2(-f.winv, -f.winv); o->one = pk2(1.0f, 1.0f);
  o->Er = f.Er; o->sym = f.sym;
}
// gains of the two correspondences of one tile pair {u0 u1 v0 v1} {s0 s1 t0 t1}; `hi_live` = the second one exists
__device__ __forceinline__ f32x2 f_filter_gain2(const FFilter32x2& f, const float4 A, const float4 B, bool hi_live) {
  const f32x2 u = pk2(A.x, A.y), v = pk2(A.z, A.w), s = pk2(B.x, B.y), t = pk2(B.z, B.w);
  const f32x2 rxc = fma2(f.F[0], s, fma2(f.F[3], t, f.F[6]));
  const f32x2 ryc = fma2(f.F[1], s, fma2(f.F[4], t, f.F[7]));
  const f32x2 rwc = fma2(f.F[2], s, fma2(f.F[5], t, f.F[8]));
  const f32x2 r = fma2(u, rxc, fma2(v, ryc, rwc));
  const f32x2 rx = fma2(f.F[0], u, fma2(f.F[1], v, f.F[2]));
  const f32x2 ry = fma2(f.F[3], u, fma2(f.F[4], v, f.F[5]));
  float r0, r1;
  upk2(r, r0, r1);
  const f32x2 rl = pk2(fmaxf(fabsf(r0) - f.Er, 0.0f), fmaxf(fabsf(r1) - f.Er, 0.0f));
  const f32x2 num = mul2(rl, rl);
  f32x2 e;
  if (f.sym) {
    const f32x2 a = fma2(rxc, rxc, mul2(ryc, ryc)), b = fma2(rx, rx, mul2(ry, ry));
    float a0, a1, b0, b1;
    upk2(fma2(a, f.c1, f.c2), a0, a1);
    upk2(fma2(b, f.c1, f.c2), b0, b1);
    e = mul2(num, pk2(__fdividef(1.0f, a0) + __fdividef(1.0f, b0), __fdividef(1.0f, a1) + __fdividef(1.0f, b1)));
  } else {
    const f32x2 den = fma2(rxc, rxc, fma2(ryc, ryc, fma2(rx, rx, mul2(ry, ry))));
    float d0, d1;
    upk2(fma2(den, f.c1, f.c2), d0, d1);
    e = mul2(num, pk2(__fdividef(1.0f, d0), __fdividef(1.0f, d1)));
  }
  float g0, g1;
  upk2(fma2(e, f.nwinv, f.one), g0, g1);
  g0 = fmaxf(g0, 0.0f);                      // NaN -> 0 (fmaxf returns the other operand)
  g1 = hi_live ? fmaxf(g1, 0.0f) : 0.0f;
  return pk2(g0, g1);
}
#endif

}  // namespace dg
