/*
 * degensac_legacy.h -- the reference's OWN C entry points, served by the B200 engine (libdegensac_b200_legacy.so).
 *
 * SURVEY.md section 8(b), proposal 4: symbols with the exact legacy signatures of
 *     exp_ransacFcustomLAF   (/root/reference/src/pydegensac/degensac/exp_ranF.h:69-74)
 *     exp_ransacHcustomLAF   (/root/reference/src/pydegensac/degensac/exp_ranH.h:27-33)
 * so that a C caller written against the reference (its binding layer, bindings.cpp:228-240 / 420-435, or any other
 * program) links against this library instead of libpydegensac_support.a without a source change.  The metric
 * function pointers a legacy caller passes (&FDs, &exFDs, &FDsidx, &FDsSym ...; &HDs, &HDsi, &HDsidx, &HDsSymMaxSq ...)
 * are exported here as identity tokens with the reference's names and prototypes (Fcustomdef.h:3-5, Htools.h:1-3): the
 * shim recognises them by address and maps them to the engine's error_type enum; they are never called.
 *
 * Supported argument combinations = what the reference's binding layer passes: do_lo = 1, inlLimit = 0 (F);
 * iter_type = 4, oriented_constraint = 1, inlLimit = 0 (H); thresholds in the binding's conventions
 * (bindings.cpp:64-107, 297-318: th = px^2 or px, SymCheck_th = 0 or 3 th (F) / 3 px (H)).  Anything else returns 0
 * inliers and leaves the model zeroed (message on stderr).  `*resids` receives a malloc'd buffer the caller frees, as
 * with the reference (bindings.cpp:242, 458); its contents (a diagnostic dump nobody reads) are not reproduced.
 * The reference seeds libc rand() from time(NULL) (exp_ranF.c:1277); so does the shim, unless the environment variable
 * DGB200_LEGACY_SEED pins the seed.  Score.J of the H entry point is not reproduced (0): the binding ignores it.
 */
#ifndef DEGENSAC_LEGACY_H
#define DEGENSAC_LEGACY_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { unsigned I; double J; unsigned Is; unsigned Ilafs; } Score;   /* rtools.h:18-29 */
typedef void (*FDsPtr)(const double*, const double*, double*, int);                                  /* Fcustomdef.h:3 */
typedef void (*exFDsPtr)(const double*, const double*, double*, double*, int);                       /* Fcustomdef.h:4 */
typedef void (*FDsidxPtr)(const double*, const double*, double*, int, int*, int);                    /* Fcustomdef.h:5 */
typedef void (*HDsPtr)(const double*, const double*, const double*, double*, int);                   /* Htools.h:1 */
typedef void (*HDsiPtr)(const double*, const double*, const double*, double*, int, int*, int);       /* Htools.h:2 */
typedef void (*HDsidxPtr)(const double*, const double*, const double*, double*, int, int*, int);     /* Htools.h:3 */

int exp_ransacFcustomLAF(double* u, double* u_1, double* u_2, int len, double th, double laf_coef, double conf, int max_sam,
                         double* F, unsigned char* inl, int* data_out, int do_lo, unsigned inlLimit, double** resids,
                         double* H_best, int* Ih, exFDsPtr EXFDS1, FDsPtr FDS1, FDsidxPtr FDS1idx, double SymCheck_th,
                         int enable_degen_check);
Score exp_ransacHcustomLAF(double* u, double* u_1, double* u_2, int len, double th, double laf_coef, double conf,
                           int max_sam, double* H, unsigned char* inl, int iter_type, int* data_out,
                           int oriented_constraint, unsigned inlLimit, double** resids, HDsPtr HDS1, HDsiPtr HDSi1,
                           HDsidxPtr HDSidx1, double SymCheck_th);

/* metric identity tokens (Ftools.h / Htools.h names) */
void FDs(const double*, const double*, double*, int);
void exFDs(const double*, const double*, double*, double*, int);
void FDsidx(const double*, const double*, double*, int, int*, int);
void FDsSym(const double*, const double*, double*, int);
void exFDsSym(const double*, const double*, double*, double*, int);
void FDsSymidx(const double*, const double*, double*, int, int*, int);
void HDs(const double*, const double*, const double*, double*, int);
void HDsi(const double*, const double*, const double*, double*, int, int*, int);
void HDsidx(const double*, const double*, const double*, double*, int, int*, int);
void HDsSymMaxSq(const double*, const double*, const double*, double*, int);
void HDsiSymMaxSq(const double*, const double*, const double*, double*, int, int*, int);
void HDsSymMaxSqidx(const double*, const double*, const double*, double*, int, int*, int);
void HDsSymMax(const double*, const double*, const double*, double*, int);
void HDsiSymMax(const double*, const double*, const double*, double*, int, int*, int);
void HDsSymMaxidx(const double*, const double*, const double*, double*, int, int*, int);
void HDsSymSumSq(const double*, const double*, const double*, double*, int);
void HDsiSymSumSq(const double*, const double*, const double*, double*, int, int*, int);
void HDsSymSumSqidx(const double*, const double*, const double*, double*, int, int*, int);
void HDsSymSum(const double*, const double*, const double*, double*, int);
void HDsiSymSum(const double*, const double*, const double*, double*, int, int*, int);
void HDsSymSumidx(const double*, const double*, const double*, double*, int, int*, int);

#ifdef __cplusplus
}
#endif
#endif /* DEGENSAC_LEGACY_H */
