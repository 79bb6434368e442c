// legacy_shim.cpp -- the reference's own C entry points (include/degensac_legacy.h) on top of the C ABI of
// libdegensac_b200.so: argument conventions of exp_ransacFcustomLAF / exp_ransacHcustomLAF are translated back into
// (px_th, error_type, sym_check, laf_coef, ...) -- the inverse of what the reference's binding layer does in
// bindings.cpp:64-107, 126-197, 297-318, 337-408 -- and the engine is called for one pair.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <vector>

#include "../../include/degensac_b200.h"
#include "../../include/degensac_legacy.h"

namespace {
uint64_t legacy_seed() {
  if (const char* e = getenv("DGB200_LEGACY_SEED")) return strtoull(e, nullptr, 10);
  return (uint64_t)time(nullptr);     // exp_ranF.c:1277, exp_ranH.c:510: srand(time(NULL))
}
bool close(double a, double b) { return fabs(a - b) <= 1e-9 * (fabs(a) + fabs(b)) + 1e-300; }
// u[6N] (x1 y1 1 x2 y2 1) (+ helper arrays u_1 = p1, u_2 = p2) -> [N,dim] rows of the ABI
void unpack(const double* u, const double* u1, const double* u2, int len, bool laf, std::vector<double>& a, std::vector<double>& b) {
  const int dim = laf ? 6 : 2;
  a.assign((size_t)len * dim, 0.0); b.assign((size_t)len * dim, 0.0);
  for (int i = 0; i < len; ++i) {
    const double x1 = u[6 * i], y1 = u[6 * i + 1], x2 = u[6 * i + 3], y2 = u[6 * i + 4];
    a[(size_t)i * dim] = x1; a[(size_t)i * dim + 1] = y1; b[(size_t)i * dim] = x2; b[(size_t)i * dim + 1] = y2;
    if (laf) {   // p1 = x + (a12, a22), p2 = x + (a11, a21) (bindings.cpp:355-385): rows are (x, y, a11, a12, a21, a22)
      a[(size_t)i * 6 + 3] = u1[6 * i] - x1;     a[(size_t)i * 6 + 5] = u1[6 * i + 1] - y1;
      a[(size_t)i * 6 + 2] = u2[6 * i] - x1;     a[(size_t)i * 6 + 4] = u2[6 * i + 1] - y1;
      b[(size_t)i * 6 + 3] = u1[6 * i + 3] - x2; b[(size_t)i * 6 + 5] = u1[6 * i + 4] - y2;
      b[(size_t)i * 6 + 2] = u2[6 * i + 3] - x2; b[(size_t)i * 6 + 4] = u2[6 * i + 4] - y2;
    }
  }
}
void give_resids(double** resids) { if (resids) *resids = (double*)calloc(1, sizeof(double)); }
}  // namespace

extern "C" {

#define TOKEN3(name) void name(const double*, const double*, double*, int) {}
#define TOKEN4(name) void name(const double*, const double*, double*, double*, int) {}
#define TOKENI(name) void name(const double*, const double*, double*, int, int*, int) {}
#define HTOKEN(name) void name(const double*, const double*, const double*, double*, int) {}
#define HTOKENI(name) void name(const double*, const double*, const double*, double*, int, int*, int) {}
TOKEN3(FDs) TOKEN4(exFDs) TOKENI(FDsidx) TOKEN3(FDsSym) TOKEN4(exFDsSym) TOKENI(FDsSymidx)
HTOKEN(HDs) HTOKENI(HDsi) HTOKENI(HDsidx)
HTOKEN(HDsSymMaxSq) HTOKENI(HDsiSymMaxSq) HTOKENI(HDsSymMaxSqidx)
HTOKEN(HDsSymMax) HTOKENI(HDsiSymMax) HTOKENI(HDsSymMaxidx)
HTOKEN(HDsSymSumSq) HTOKENI(HDsiSymSumSq) HTOKENI(HDsSymSumSqidx)
HTOKEN(HDsSymSum) HTOKENI(HDsiSymSum) HTOKENI(HDsSymSumidx)

int exp_ransacFcustomLAF(double* u, double* u_1, double* u_2, int len, double th, double laf_coef, double conf, int max_sam,
                         double* F, unsigned char* inl, int* data_out, int do_lo, unsigned inlLimit, double** resids,
                         double* H_best, int* Ih, exFDsPtr EXFDS1, FDsPtr FDS1, FDsidxPtr FDS1idx, double SymCheck_th,
                         int enable_degen_check) {
  (void)H_best;
  give_resids(resids);
  if (F) for (int i = 0; i < 9; ++i) F[i] = 0.0;
  if (Ih) *Ih = 0;
  int metric = -1;
  if (FDS1 == &FDs && EXFDS1 == &exFDs && FDS1idx == &FDsidx) metric = DGB200_F_SAMPSON;
  if (FDS1 == &FDsSym && EXFDS1 == &exFDsSym && FDS1idx == &FDsSymidx) metric = DGB200_F_SYMM_EPIPOLAR;
  const bool sym = SymCheck_th > 0;
  if (metric < 0 || !do_lo || inlLimit != 0 || !(th > 0) || (sym && !close(SymCheck_th, 3.0 * th))) {
    fprintf(stderr, "degensac_b200 legacy shim: exp_ransacFcustomLAF called with an argument combination the reference's "
                    "binding layer never produces (metric pointers / do_lo / inlLimit / SymCheck_th)\n");
    return 0;
  }
  const bool laf = laf_coef > 0 && u_1 && u_2;
  std::vector<double> a, b;
  unpack(u, u_1, u_2, len, laf, a, b);
  int32_t stats[4] = {0, 0, 0, 0};
  const int rc = dgb200_find_fundamental(a.data(), b.data(), len, laf ? 6 : 2, sqrt(th), conf, max_sam, metric, sym ? 1 : 0,
                                         laf ? laf_coef : 0.0, enable_degen_check, legacy_seed(), F, inl, stats);
  if (rc != 0) { fprintf(stderr, "degensac_b200 legacy shim: %s\n", dgb200_last_error()); return 0; }
  if (data_out) { data_out[0] = stats[0]; data_out[1] = stats[1]; }
  if (Ih) *Ih = stats[2];
  return stats[3];
}

Score exp_ransacHcustomLAF(double* u, double* u_1, double* u_2, int len, double th, double laf_coef, double conf,
                           int max_sam, double* H, unsigned char* inl, int iter_type, int* data_out,
                           int oriented_constraint, unsigned inlLimit, double** resids, HDsPtr HDS1, HDsiPtr HDSi1,
                           HDsidxPtr HDSidx1, double SymCheck_th) {
  Score S = {0, 0.0, 0, 0};
  give_resids(resids);
  if (H) for (int i = 0; i < 9; ++i) H[i] = 0.0;
  int metric = -1;
  if (HDS1 == &HDs && HDSi1 == &HDsi && HDSidx1 == &HDsidx) metric = DGB200_H_SAMPSON;
  if (HDS1 == &HDsSymMaxSq && HDSi1 == &HDsiSymMaxSq && HDSidx1 == &HDsSymMaxSqidx) metric = DGB200_H_SYMM_SQ_MAX;
  if (HDS1 == &HDsSymMax && HDSi1 == &HDsiSymMax && HDSidx1 == &HDsSymMaxidx) metric = DGB200_H_SYMM_MAX;
  if (HDS1 == &HDsSymSumSq && HDSi1 == &HDsiSymSumSq && HDSidx1 == &HDsSymSumSqidx) metric = DGB200_H_SYMM_SQ_SUM;
  if (HDS1 == &HDsSymSum && HDSi1 == &HDsiSymSum && HDSidx1 == &HDsSymSumidx) metric = DGB200_H_SYMM_SUM;
  // error_threshold: px^2 for sampson / squared metrics, px for the others (bindings.cpp:64-107)
  const bool squared = metric == DGB200_H_SAMPSON || metric == DGB200_H_SYMM_SQ_MAX || metric == DGB200_H_SYMM_SQ_SUM;
  const double px = squared ? sqrt(th) : th;
  const bool sym = SymCheck_th > 0;
  const bool sym_allowed = metric == DGB200_H_SAMPSON || metric == DGB200_H_SYMM_SQ_SUM || metric == DGB200_H_SYMM_SUM;
  if (metric < 0 || iter_type != 4 || !oriented_constraint || inlLimit != 0 || !(th > 0) ||
      (sym && (!sym_allowed || !close(SymCheck_th, 3.0 * px)))) {
    fprintf(stderr, "degensac_b200 legacy shim: exp_ransacHcustomLAF called with an argument combination the reference's "
                    "binding layer never produces (metric pointers / iter_type / oriented_constraint / inlLimit / SymCheck_th)\n");
    return S;
  }
  const bool laf = laf_coef > 0 && u_1 && u_2;
  std::vector<double> a, b;
  unpack(u, u_1, u_2, len, laf, a, b);
  int32_t stats[4] = {0, 0, 0, 0};
  // the binding's two "max" metrics never enable the gate (SymCheck_th = 0); for the others sym <=> sym_check_enable
  const int rc = dgb200_find_homography(a.data(), b.data(), len, laf ? 6 : 2, px, conf, max_sam, metric, sym ? 1 : 0,
                                        laf ? laf_coef : 0.0, legacy_seed(), H, inl, stats);
  if (rc != 0) { fprintf(stderr, "degensac_b200 legacy shim: %s\n", dgb200_last_error()); return S; }
  if (data_out) { data_out[0] = stats[0]; data_out[1] = stats[1]; data_out[2] = 0; }
  S.I = (unsigned)stats[3];
  return S;
}

}  // extern "C"
