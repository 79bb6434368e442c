#include <stdio.h>
#include <stdlib.h>
// hostemu.cpp -- TEST HELPER ONLY.  Compiles the engine headers with plain g++ as a one-thread CTA
// (tid=0, nt=1, barriers are no-ops) so the ordered-replay control flow can be checked against the
// reference on a machine without a GPU.  Never loaded by the pydegensac_b200 package.
#include <stdlib.h>
#include <string.h>
#define DG_FILTER_CHECK 1
#define DG_EIG_STATS 1
#include "../../pydegensac_b200/csrc/engine_f.h"
#include "../../pydegensac_b200/csrc/engine_h.h"
#include "../../pydegensac_b200/csrc/engine_h2el.h"
#include "../../pydegensac_b200/csrc/workspace.h"

using namespace dg;

static int g_use_filter32 = 1;
extern "C" void emu_set_filter32(int on) { g_use_filter32 = on; }
extern "C" void emu_hfilter_stats(long* checked, long* violations) { *checked = g_hfilter_checked; *violations = g_hfilter_violations; }
extern "C" void emu_hres_stats(long* points, long* far, long* violations) { *points = g_hres_points; *far = g_hres_far; *violations = g_hres_violations; }
extern "C" void emu_pp_stats(long* checked, long* violations, long* settled) { *checked = g_pp_checked; *violations = g_pp_violations; *settled = g_pp_settled; }
extern "C" void emu_filter_stats(long* checked, long* violations, double* maxslack) {
  *checked = g_filter_checked; *violations = g_filter_violations; *maxslack = g_filter_maxslack;
}

struct Emu {
  Tile32 t32;
  Pt32* tile;
  Ctx c;
  Workspace W;
  BlockScratch sc;
  unsigned char* slab;
  double* soa;
};

static int g_final_lsq = 0;
extern "C" void emu_set_final_lsq(int on) { g_final_lsq = on; }
static int g_poison = 0;
extern "C" void emu_set_poison(int v) { g_poison = v; }
static void emu_setup(Emu& E, const double* x1y1, const double* x2y2, int n, int dim, int chunk) {
  memset(&E.sc, g_poison, sizeof(E.sc));
  const size_t bytes = workspace_bytes(n, chunk, dim == 6);
  E.slab = (unsigned char*)calloc(bytes + 256, 1);
  workspace_carve(E.slab, n, chunk, dim == 6, &E.W, &E.soa);
  const size_t row = align_up(sizeof(double) * (size_t)n, 128) / sizeof(double);
  double* x1 = E.soa; double* y1 = E.soa + row; double* x2 = E.soa + 2 * row; double* y2 = E.soa + 3 * row;
  for (int i = 0; i < n; ++i) {
    x1[i] = x1y1[(size_t)dim * i]; y1[i] = x1y1[(size_t)dim * i + 1];
    x2[i] = x2y2[(size_t)dim * i]; y2[i] = x2y2[(size_t)dim * i + 1];
  }
  E.c.tid = 0; E.c.nt = 1; E.c.lane = 0; E.c.wid = 0; E.c.nw = 1; E.c.N = n;
  E.c.x1 = x1; E.c.y1 = y1; E.c.x2 = x2; E.c.y2 = y2;
  E.c.sc = &E.sc;
  E.c.t32 = nullptr;
  for (int k = 0; k < 8; ++k) E.c.laf[k] = nullptr;
  if (dim == 6) {
    for (int i = 0; i < n; ++i) {
      const double* q1 = x1y1 + (size_t)6 * i; const double* q2 = x2y2 + (size_t)6 * i;
      E.W.laf[0][i] = q1[0] + q1[3]; E.W.laf[1][i] = q1[1] + q1[5]; E.W.laf[2][i] = q2[0] + q2[3]; E.W.laf[3][i] = q2[1] + q2[5];
      E.W.laf[4][i] = q1[0] + q1[2]; E.W.laf[5][i] = q1[1] + q1[4]; E.W.laf[6][i] = q2[0] + q2[2]; E.W.laf[7][i] = q2[1] + q2[4];
    }
    for (int k = 0; k < 8; ++k) E.c.laf[k] = E.W.laf[k];
  }
  E.tile = (Pt32*)malloc(sizeof(Pt32) * (size_t)n);
}

extern "C" int emu_find_fundamental(const double* x1y1, const double* x2y2, int n, int dim, double px_th, double conf,
                                    int max_iters, int error_type, int sym_check, double laf_coef, int degen,
                                    uint64_t seed, int chunk, double* F, unsigned char* mask, int* stats) {
  if (n < 8 || (dim != 2 && dim != 6)) return -1;
  if (laf_coef > 0 && dim != 6) return -1;
  Emu E;
  emu_setup(E, x1y1, x2y2, n, dim, chunk);
  FParams P;
  f_thresholds(px_th, sym_check, &P.th, &P.sym_th);
  P.conf = conf; P.laf_coef = laf_coef; P.max_iters = max_iters; P.metric = error_type; P.degen = degen;
  P.do_laf = laf_coef > 0 ? 1 : 0; P.th_laf = laf_coef * P.th;
  P.do_sym = P.sym_th > 0; P.seed = seed; P.chunk = chunk; P.final_lsq = g_final_lsq;
  if (g_use_filter32) { blk_prepare_tile32(E.c, E.tile, &E.t32); E.c.t32 = &E.t32; }
  ransac_F_pair(E.c, P, E.W, F, mask, stats);
  free(E.slab); free(E.tile);
  return 0;
}

extern "C" int emu_find_homography(const double* x1y1, const double* x2y2, int n, int dim, double px_th, double conf,
                                   int max_iters, int error_type, int sym_check, double laf_coef, uint64_t seed,
                                   int chunk, double* H, unsigned char* mask, int* stats) {
  if (n < 4 || (dim != 2 && dim != 6)) return -1;
  if (laf_coef > 0 && dim != 6) return -1;
  Emu E;
  emu_setup(E, x1y1, x2y2, n, dim, chunk);
  HParams P;
  if (h_thresholds(error_type, px_th, sym_check, &P.th, &P.sym_th)) return -2;
  P.conf = conf; P.laf_coef = laf_coef; P.max_iters = max_iters; P.metric = error_type;
  P.do_laf = laf_coef > 0 ? 1 : 0; P.th_laf = laf_coef * P.th;
  P.do_sym = P.sym_th > 0; P.seed = seed; P.chunk = chunk; P.final_lsq = g_final_lsq;
  if (g_use_filter32 && error_type == H_SAMPSON) { blk_prepare_tile32(E.c, E.tile, &E.t32); E.c.t32 = &E.t32; }
  ransac_H_pair(E.c, P, E.W, H, mask, stats);
  free(E.slab); free(E.tile);
  return 0;
}

extern "C" int emu_find_homography_2el(const double* u10, int n, double px_th, double conf, int max_iters, uint64_t seed,
                                       int chunk, double* H, unsigned char* mask, int* stats) {
  if (n < 4) return -1;
  Emu E;
  emu_setup(E, u10, u10 + 5, n, 10, chunk);
  H2Params P;
  P.th = px_th * px_th; P.conf = conf; P.max_iters = max_iters; P.seed = seed; P.chunk = chunk;
  ransac_H2el_pair(E.c, P, E.W, u10, H, mask, stats);
  free(E.slab); free(E.tile);
  return 0;
}
extern "C" long emu_u2h4_calls(int reset) { const long v = g_u2h4_calls; if (reset) g_u2h4_calls = 0; return v; }
extern "C" int emu_h_from_2el(const double* ua, const double* ub, double* h) { return h_from_2el(ua, ub, h) ? 1 : 0; }

// ---- leaf exports for known-answer tests against the reference's own leaves ----
extern "C" int emu_minv3(double* a) { return minv3(a); }
extern "C" int emu_nullspace9(double* M, double* ns) { return nullspace9(M, ns); }
extern "C" void emu_seven_pt_cubic(const double* A, double* B, double* p) { seven_pt_cubic(A, B, p); }
extern "C" int emu_cubic_real_roots(const double* po, double* r) { return cubic_real_roots(po, r); }
extern "C" double emu_f_resid(int metric, const double* F, double x1, double y1, double x2, double y2) { return f_resid(metric, F, x1, y1, x2, y2); }
extern "C" double emu_h_resid(int metric, const double* H, double x1, double y1, double x2, double y2) {
  HSym s; h_sym_prepare(H, &s); return h_resid_metric(metric, H, s, x1, y1, x2, y2);
}
extern "C" int emu_nullspace_qr7x9(const double* A, double* N) { return nullspace_qr7x9(A, N); }
extern "C" uint32_t emu_hash(const int* idx, int n) { return superfasthash_i32(idx, n); }
extern "C" int emu_nsamples(int a, int b, int c, double d) { return nsamples(a, b, c, d); }
extern "C" void emu_min_eigvec9(double* C, double* v) { min_eigvec9(C, v); }
extern "C" void emu_enforce_rank2(double* F) { enforce_rank2(F); }
extern "C" void emu_enforce_rank2_slow(double* F) { enforce_rank2_slow(F); }
extern "C" int emu_smallest_right_sv3_fast(const double* F, double* v) { return smallest_right_sv3_fast(F, v) ? 1 : 0; }
extern "C" void emu_left_null(double* Z, int len, double* q) { left_null_9xk(Z, len, q); }
extern "C" void emu_minimal_sample(uint64_t seed, uint32_t k, int N, int m, int* sel) {
  if (m == 7) minimal_sample<7>(seed, k, N, sel); else minimal_sample<4>(seed, k, N, sel);
}
extern "C" uint32_t emu_value31(uint64_t seed, uint32_t k, uint32_t j) { return value31(seed, k, j); }
extern "C" int emu_checksample(const double* F, const double* u7, double th, double* H) {
  static WarpScratch ws;
  for (int t = 0; t < 5; ++t) if (warp_checksample_triplet(&ws, F, u7, t, th, H, 0, 1)) return 1;
  return 0;
}
extern "C" void emu_gkr_v3(const double* A, double* v) { gkr_third_right_vector3(A, v); }

extern "C" void emu_eig_stats(long* calls, long* fallbacks, long* iters) { *calls = g_eig_calls; *fallbacks = g_eig_fallbacks; *iters = g_eig_iters; }
