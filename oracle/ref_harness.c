/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C harness around the UNMODIFIED reference C sources, which are compiled
 * where they lie under /root/reference by oracle/build_ref.sh and linked with
 * -Wl,--wrap=time,--wrap=srand,--wrap=rand,--wrap=random so that:
 *
 *   mode 0 (passthrough): the reference runs on glibc's own rand()/random(),
 *                         seeded from the settable "time" -> used for CPU timing;
 *   mode 1 (replay)     : every random draw the reference makes is served from the
 *                         counter-based Philox4x32-10 stream that the B200 engine
 *                         uses, keyed by (seed, iteration k, draw j)  -> identical
 *                         sampling order on both sides ("fixed sampling order" parity).
 *
 * Stream contract (shared with pydegensac_b200/csrc/rng.h, restated here on purpose):
 *   value31(seed,k,j) = philox4x32_10(ctr=(j>>2, k, 0, 0), key=(seed_lo, seed_hi))[j&3] >> 1
 *   Iteration k >= 1 is the k-th pass of the reference's main loop (it re-seeds with
 *   srand() at the top of every pass: exp_ranF.c:1337, exp_ranH.c:550).
 *     draws j = 0..m-1 (m = 7 for F, 4 for H): minimal-sample draws. The engine defines the
 *         sample statelessly: a partial Fisher-Yates over a FRESH identity pool,
 *         s_i = value31 % (N-i), swap slots s_i <-> N-1-i (same arithmetic as
 *         rtools.c:12-23 `sample`).  The reference keeps a PERSISTENT pool, so the
 *         wrapper mirrors that pool and hands the reference the slot that currently
 *         holds the index the stateless rule selected.
 *     draw j = m: the reference's `seed = rand()` (value unused: srand() is wrapped).
 *     draws j > m: LO / DEGENSAC draws (randsubset, rFtH, dual_sample), returned raw.
 *   k = 0 is the single pre-loop rand() (exp_ranF.c:1331, exp_ranH.c:539).
 *
 * The entry points below mirror what the reference's pybind layer does before it
 * calls the C core (bindings.cpp:19-251 for H, :253-467 for F): metric selection,
 * threshold conventions, u[6N] packing, LAF helper points.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ---- reference entry points / metric functions (declared, not copied) ---- */
typedef struct { unsigned I; double J; unsigned Is; unsigned Ilafs; } RefScore; /* rtools.h:18-29 */
typedef void (*FDsPtr)(const double *, const double *, double *, int);
typedef void (*exFDsPtr)(const double *, const double *, double *, double *, int);
typedef void (*FDsidxPtr)(const double *, const double *, double *, int, int *, int);
typedef void (*HDsPtr)(const double *, const double *, const double *, double *, int);
typedef void (*HDsiPtr)(const double *, const double *, const double *, double *, int, int *, int);
typedef void (*HDsidxPtr)(const double *, const double *, const double *, double *, int, int *, int);

extern int exp_ransacFcustomLAF(double *u, double *u_1, double *u_2, int len, double th, double laf_coef,
                                double conf, int max_sam, double *F, unsigned char *inl, int *data_out,
                                int do_lo, unsigned inlLimit, double **resids, double *H_best, int *Ih,
                                exFDsPtr, FDsPtr, FDsidxPtr, double SymCheck_th, int enable_degen_check);
extern RefScore exp_ransacHcustomLAF(double *u, double *u_1, double *u_2, int len, double th, double laf_coef,
                                     double conf, int max_sam, double *H, unsigned char *inl, int iter_type,
                                     int *data_out, int oriented_constraint, unsigned inlLimit, double **resids,
                                     HDsPtr, HDsiPtr, HDsidxPtr, double SymCheck_th);
extern void FDs(const double *, const double *, double *, int);
extern void FDsSym(const double *, const double *, double *, int);
extern void FDsidx(const double *, const double *, double *, int, int *, int);
extern void FDsSymidx(const double *, const double *, double *, int, int *, int);
extern void exFDs(const double *, const double *, double *, double *, int);
extern void exFDsSym(const double *, const double *, double *, double *, int);
extern void HDs(const double *, const double *, const double *, double *, int);
extern void HDsi(const double *, const double *, const double *, double *, int, int *, int);
extern void HDsidx(const double *, const double *, const double *, double *, int, int *, int);
extern void HDsSymMaxSq(const double *, const double *, const double *, double *, int);
extern void HDsiSymMaxSq(const double *, const double *, const double *, double *, int, int *, int);
extern void HDsSymMaxSqidx(const double *, const double *, const double *, double *, int, int *, int);
extern void HDsSymMax(const double *, const double *, const double *, double *, int);
extern void HDsiSymMax(const double *, const double *, const double *, double *, int, int *, int);
extern void HDsSymMaxidx(const double *, const double *, const double *, double *, int, int *, int);
extern void HDsSymSumSq(const double *, const double *, const double *, double *, int);
extern void HDsiSymSumSq(const double *, const double *, const double *, double *, int, int *, int);
extern void HDsSymSumSqidx(const double *, const double *, const double *, double *, int, int *, int);
extern void HDsSymSum(const double *, const double *, const double *, double *, int);
extern void HDsiSymSum(const double *, const double *, const double *, double *, int, int *, int);
extern void HDsSymSumidx(const double *, const double *, const double *, double *, int, int *, int);

/* ------------------------------ RNG interposition ------------------------------ */
extern void __real_srand(unsigned);
extern int __real_rand(void);
extern long __real_random(void);

static int g_mode = 0;          /* 0 passthrough, 1 Philox replay */
static uint64_t g_seed = 0;
static long g_time = 12345;     /* what the reference's time(NULL) sees */
static int g_N = 0, g_m = 0;    /* correspondences, minimal sample size */
static int *g_pool = 0, *g_pos = 0;
static long g_k = 0;            /* iteration (0 = pre-loop) */
static unsigned g_j = 0;        /* draw counter inside iteration */
static long g_srand_calls = 0, g_draws_total = 0;
static int g_sel[8];

static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  int r;
  for (r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static uint32_t value31(uint64_t seed, uint32_t k, uint32_t j) {
  uint32_t o[4];
  philox4x32_10(j >> 2, k, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  return o[j & 3] >> 1;
}

/* stateless sample of iteration k: partial Fisher-Yates on a fresh identity pool */
static void stateless_sample(uint64_t seed, uint32_t k, int N, int m, int *sel) {
  int touched_pos[16], touched_val[16], nt = 0, i, t;
  for (i = 0; i < m; ++i) {
    int s = (int)(value31(seed, k, (uint32_t)i) % (uint32_t)(N - i));
    int j = N - i - 1, vs = s, vj = j;
    for (t = 0; t < nt; ++t) { if (touched_pos[t] == s) vs = touched_val[t]; if (touched_pos[t] == j) vj = touched_val[t]; }
    /* pool[s] = vj ; pool[j] = vs */
    for (t = 0; t < nt && touched_pos[t] != s; ++t) {}
    if (t == nt) { touched_pos[nt] = s; ++nt; }
    touched_val[t] = vj;
    for (t = 0; t < nt && touched_pos[t] != j; ++t) {}
    if (t == nt) { touched_pos[nt] = j; ++nt; }
    touched_val[t] = vs;
    sel[i] = vs;
  }
}

static long replay_draw(void) {
  unsigned j = g_j++;
  ++g_draws_total;
  if (g_k >= 1 && (int)j < g_m) {
    int i = (int)j, want, s, top, q;
    if (i == 0) stateless_sample(g_seed, (uint32_t)g_k, g_N, g_m, g_sel);
    want = g_sel[i];
    s = g_pos[want];         /* slot of the reference's persistent pool that holds `want` */
    top = g_N - i - 1;
    /* mirror rtools.c:12-23: swap pool[s] <-> pool[top] */
    q = g_pool[s]; g_pool[s] = g_pool[top]; g_pool[top] = q;
    g_pos[g_pool[s]] = s; g_pos[g_pool[top]] = top;
    return (long)s;          /* the reference computes s % (N-i) == s */
  }
  return (long)value31(g_seed, (uint32_t)g_k, j);
}

time_t __wrap_time(time_t *t) { if (t) *t = (time_t)g_time; return (time_t)g_time; }
void __wrap_srand(unsigned s) {
  if (!g_mode) { __real_srand(s); return; }
  if (g_srand_calls++ == 0) g_k = 0; else ++g_k;
  g_j = 0;
}
int __wrap_rand(void) { return g_mode ? (int)replay_draw() : __real_rand(); }
long __wrap_random(void) { return g_mode ? replay_draw() : __real_random(); }

static void rng_begin(int mode, uint64_t seed, int N, int m) {
  int i;
  g_mode = mode; g_seed = seed; g_N = N; g_m = m; g_k = 0; g_j = 0; g_srand_calls = 0; g_draws_total = 0;
  g_time = (long)(seed & 0x7fffffff);
  free(g_pool); free(g_pos);
  g_pool = (int *)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
  g_pos = (int *)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
  for (i = 0; i < N; ++i) { g_pool[i] = i; g_pos[i] = i; }
}

void ref_set_rng(int mode, uint64_t seed, int N, int m) { rng_begin(mode, seed, N, m); }
uint32_t ref_value31(uint64_t seed, uint32_t k, uint32_t j) { return value31(seed, k, j); }
void ref_stateless_sample(uint64_t seed, uint32_t k, int N, int m, int *sel) { stateless_sample(seed, k, N, m, sel); }

/* ------------------------------ u[6N] packing ------------------------------ */
/* bindings.cpp:126-197 / 337-408: (x1,y1,1,x2,y2,1); LAF helper points p1 = x + (a12,a22), p2 = x + (a11,a21) */
static void pack_u(const double *x1y1, const double *x2y2, int n, int dim, int laf, double *u, double *u1, double *u2) {
  int i;
  for (i = 0; i < n; ++i) {
    const double *a = x1y1 + (size_t)dim * i, *b = x2y2 + (size_t)dim * i;
    double *p = u + 6 * (size_t)i;
    p[0] = a[0]; p[1] = a[1]; p[2] = 1.0; p[3] = b[0]; p[4] = b[1]; p[5] = 1.0;
    if (laf) {
      double *q = u1 + 6 * (size_t)i, *r = u2 + 6 * (size_t)i;
      q[0] = a[0] + a[3]; q[1] = a[1] + a[5]; q[2] = 1.0; q[3] = b[0] + b[3]; q[4] = b[1] + b[5]; q[5] = 1.0;
      r[0] = a[0] + a[2]; r[1] = a[1] + a[4]; r[2] = 1.0; r[3] = b[0] + b[2]; r[4] = b[1] + b[4]; r[5] = 1.0;
    }
  }
}

/* stats_out[0]=samples drawn, [1]=LO runs, [2]=rejections(H)/plane inliers Ih (F), [3]=returned inlier count */
int ref_find_fundamental(const double *x1y1, const double *x2y2, int n, int dim, double px_th, double conf,
                         int max_iters, int error_type, int sym_check, double laf_coef, int degen_check,
                         int rng_mode, uint64_t seed, double *F_out, unsigned char *mask_out, int *stats_out) {
  int laf = laf_coef > 0, i, Ih = 0, I;
  double *u, *u1, *u2, *resids = 0, HinF[9], F[9];
  int *data_out;
  FDsPtr f = error_type == 1 ? &FDsSym : &FDs;
  exFDsPtr ef = error_type == 1 ? &exFDsSym : &exFDs;
  FDsidxPtr fi = error_type == 1 ? &FDsSymidx : &FDsidx;
  double th = px_th * px_th, sym_th = px_th * px_th * (3.0 * (sym_check ? 1 : 0)); /* bindings.cpp:299-318 */
  if (n < 8 || (dim != 2 && dim != 6)) return -1;
  u = (double *)malloc(sizeof(double) * 6 * (size_t)n);
  u1 = (double *)malloc(sizeof(double) * 6 * (size_t)(laf ? n : 1));
  u2 = (double *)malloc(sizeof(double) * 6 * (size_t)(laf ? n : 1));
  data_out = (int *)calloc((size_t)n * 18, sizeof(int));
  pack_u(x1y1, x2y2, n, dim, laf, u, u1, u2);
  for (i = 0; i < 9; ++i) F[i] = 0.0;
  memset(mask_out, 0, (size_t)n);
  rng_begin(rng_mode, seed, n, 7);
  I = exp_ransacFcustomLAF(u, u1, u2, n, th, laf_coef, conf, max_iters, F, mask_out, data_out, 1, 0, &resids,
                           HinF, &Ih, ef, f, fi, sym_th, degen_check);
  for (i = 0; i < 9; ++i) F_out[i] = F[i];
  if (stats_out) { stats_out[0] = data_out[0]; stats_out[1] = data_out[1]; stats_out[2] = Ih; stats_out[3] = I; }
  free(resids); free(data_out); free(u); free(u1); free(u2);
  g_mode = 0;
  return 0;
}

int ref_find_homography(const double *x1y1, const double *x2y2, int n, int dim, double px_th, double conf,
                        int max_iters, int error_type, int sym_check, double laf_coef,
                        int rng_mode, uint64_t seed, double *H_out, unsigned char *mask_out, int *stats_out) {
  int laf = laf_coef > 0, i;
  double *u, *u1, *u2, *resids = 0, H[9];
  int *data_out;
  HDsPtr h; HDsiPtr hi; HDsidxPtr hx;
  double th, sym_th, coef = 3.0 * (sym_check ? 1 : 0);
  RefScore S;
  if (n < 4 || (dim != 2 && dim != 6)) return -1;
  switch (error_type) { /* bindings.cpp:66-107 */
    case 0: h = &HDs; hi = &HDsi; hx = &HDsidx; th = px_th * px_th; sym_th = px_th * coef; break;
    case 1: h = &HDsSymMaxSq; hi = &HDsiSymMaxSq; hx = &HDsSymMaxSqidx; th = px_th * px_th; sym_th = 0; break;
    case 2: h = &HDsSymMax; hi = &HDsiSymMax; hx = &HDsSymMaxidx; th = px_th; sym_th = 0; break;
    case 3: h = &HDsSymSumSq; hi = &HDsiSymSumSq; hx = &HDsSymSumSqidx; th = px_th * px_th; sym_th = px_th * coef; break;
    case 4: h = &HDsSymSum; hi = &HDsiSymSum; hx = &HDsSymSumidx; th = px_th; sym_th = px_th * coef; break;
    default: return -2;
  }
  u = (double *)malloc(sizeof(double) * 6 * (size_t)n);
  u1 = (double *)malloc(sizeof(double) * 6 * (size_t)(laf ? n : 1));
  u2 = (double *)malloc(sizeof(double) * 6 * (size_t)(laf ? n : 1));
  data_out = (int *)calloc((size_t)n * 18, sizeof(int));
  pack_u(x1y1, x2y2, n, dim, laf, u, u1, u2);
  for (i = 0; i < 9; ++i) H[i] = 0.0;
  memset(mask_out, 0, (size_t)n);
  rng_begin(rng_mode, seed, n, 4);
  S = exp_ransacHcustomLAF(u, u1, u2, n, th, laf_coef, conf, max_iters, H, mask_out, 4, data_out, 1, 0, &resids,
                           h, hi, hx, sym_th);
  for (i = 0; i < 9; ++i) H_out[i] = H[i]; /* raw: column-major, maps image 2 -> image 1 */
  if (stats_out) { stats_out[0] = data_out[0]; stats_out[1] = data_out[1]; stats_out[2] = data_out[2]; stats_out[3] = (int)S.I; }
  free(resids); free(data_out); free(u); free(u1); free(u2);
  g_mode = 0;
  return 0;
}

/* ---- homography from elliptical features: ransacH2el (ranH2el.c:19), no binding in the reference; rows u10 =
 * (x', y', a', b', c', x, y, a, b, c).  The driver has no pre-loop srand(): one is issued here so that its first pass is
 * iteration k = 1 of the stream contract like the other two drivers.  th = px_th^2 (Sampson error, as error_type 0). */
extern RefScore ransacH2el(double *u10, int len, double th, double conf, int max_sam, double *H, unsigned char *inl,
                           int *data_out, int do_lo, int inlLimit);
extern int A2toRH(double *N1, double *D1, double *N2, double *D2, double *u, int *samidx, double *h);
extern void getTransf(double *u10, double *N, double *D);
int ref_find_homography_2el(const double *u10, int n, double px_th, double conf, int max_iters, int rng_mode,
                            uint64_t seed, double *H_out, unsigned char *mask_out, int *stats_out) {
  int i, data_out[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double H[9], *u;
  RefScore S;
  if (n < 4) return -1;
  u = (double *)malloc(sizeof(double) * 10 * (size_t)n);
  memcpy(u, u10, sizeof(double) * 10 * (size_t)n);
  for (i = 0; i < 9; ++i) H[i] = 0.0;
  memset(mask_out, 0, (size_t)n);
  rng_begin(rng_mode, seed, n, 2);
  if (rng_mode) __wrap_srand(0); else __real_srand((unsigned)g_time);
  S = ransacH2el(u, n, px_th * px_th, conf, max_iters, H, mask_out, data_out, 1, 0);
  for (i = 0; i < 9; ++i) H_out[i] = H[i];
  if (stats_out) { stats_out[0] = data_out[0]; stats_out[1] = data_out[1]; stats_out[2] = data_out[2]; stats_out[3] = (int)S.I; }
  free(u);
  g_mode = 0;
  return 0;
}
int ref_h_from_2el(const double *ua, const double *ub, double *h) {
  double u[20], N1[9], D1[9], N2[9], D2[9];
  int samidx[2] = {0, 1};
  memcpy(u, ua, 80); memcpy(u + 10, ub, 80);
  getTransf(u, N1, D1); getTransf(u + 10, N2, D2);
  return A2toRH(N1, D1, N2, D2, u, samidx, h) ? 0 : 1;
}

/* CCMATH minv (matutls/minv.c) exposed for the bit-for-bit test of the engine's restatement. */
extern int minv(double *a, int n);
int ref_minv3(double *a) { return minv(a, 3); }
