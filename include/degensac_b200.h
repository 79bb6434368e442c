/*
 * degensac_b200.h -- C ABI of the B200-native LO-RANSAC / DEGENSAC engine (libdegensac_b200.so).
 *
 * Drop-in boundary for the reference's hot path.  Each entry point states the reference interface it
 * replaces; conventions (thresholds, metric numbering, model layout) are those of the reference's
 * binding layer so that a maintainer can rebind `pydegensac.findHomography_/findFundamentalMatrix_`
 * one-to-one.  Plain pointers and sizes only: no torch / pybind types.
 *
 * Common conventions
 *   x1y1, x2y2 : row-major [n_pairs][n][dim] float64, dim = 2 (x,y) or 6 (x,y,a11,a12,a21,a22); only
 *                columns 0-1 are correspondences (bindings.cpp:180-197, 391-408).
 *   px_th, conf, max_iters, error_type, sym_check, laf_coef : exactly the arguments of
 *                findFundamentalMatrix_/findHomography_ (bindings.cpp:484-503); thresholds are squared /
 *                scaled inside, as the binding does (bindings.cpp:64-107, 297-318).
 *   seeds      : one uint64 per pair (NULL -> pair index).  The reference seeds libc rand() from
 *                time(NULL) (exp_ranF.c:1277, exp_ranH.c:510); here sampling is a counter-based Philox
 *                stream keyed by (seed, iteration, draw) -> reproducible and order-independent.
 *   model_out  : [n_pairs][9] float64.  F: row-major, x2^T F x1 = 0 (as the reference returns it).
 *                H: RAW core output = column-major, maps image 2 -> image 1, i.e. what
 *                exp_ransacHcustomLAF writes; the Python layer applies inv(H.T) (utils.py:108).
 *                All-zero model = "no model found" (then the mask is all zero, utils.py:104-107,143-145).
 *   mask_out   : [n_pairs][n] uint8 (1 = inlier).
 *   stats_out  : [n_pairs][4] int32 or NULL: {samples drawn, LO runs, plane inliers (F) / 0 (H),
 *                inlier count of the returned model}  (the reference's data_out[0..1], *Ih, return value).
 *   return     : 0 ok; <0 error (DGB200_E_*), message via dgb200_last_error().
 */
#ifndef DEGENSAC_B200_H
#define DEGENSAC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGB200_OK 0
#define DGB200_E_ARG (-1)          /* bad shape: n < 8 (F) / n < 4 (H), dim not 2 or 6 (bindings.cpp:32-47, 267-282) */
#define DGB200_E_METRIC (-2)       /* unknown error_type (bindings.cpp:10-17) */
#define DGB200_E_UNSUPPORTED (-3)  /* reserved (the LAF-consistency gate, SURVEY.md §8(f).1, is implemented for [n,6] inputs) */
#define DGB200_E_CUDA (-10)        /* no device / CUDA runtime failure: the engine has no CPU fallback */

/* error_type numbering of the reference (bindings.cpp:10-17) */
#define DGB200_F_SAMPSON 0
#define DGB200_F_SYMM_EPIPOLAR 1
#define DGB200_H_SAMPSON 0
#define DGB200_H_SYMM_SQ_MAX 1
#define DGB200_H_SYMM_MAX 2
#define DGB200_H_SYMM_SQ_SUM 3
#define DGB200_H_SYMM_SUM 4

/* Replaces exp_ransacFcustomLAF (exp_ranF.h:69-74, exp_ranF.c:1244) as called by findFundamentalMatrix_
 * (bindings.cpp:253-467), batched over independent image pairs.  HOST buffers; synchronous (copies in,
 * runs the wave/replay kernel, copies out). */
int dgb200_find_fundamental_batch(const double* x1y1, const double* x2y2, int n_pairs, int n, int dim,
                                  double px_th, double conf, int max_iters, int error_type, int sym_check,
                                  double laf_coef, int degen_check, const uint64_t* seeds,
                                  double* F_out, uint8_t* mask_out, int32_t* stats_out);

/* Replaces exp_ransacHcustomLAF (exp_ranH.h:27-33, exp_ranH.c:470; iter_type 4, oriented constraint on,
 * inlLimit 0) as called by findHomography_ (bindings.cpp:19-251).  HOST buffers; synchronous. */
int dgb200_find_homography_batch(const double* x1y1, const double* x2y2, int n_pairs, int n, int dim,
                                 double px_th, double conf, int max_iters, int error_type, int sym_check,
                                 double laf_coef, const uint64_t* seeds,
                                 double* H_out, uint8_t* mask_out, int32_t* stats_out);

/* Same two paths with DEVICE pointers (inputs already resident in HBM, outputs left in HBM) on a CUDA
 * stream (cudaStream_t passed as void*; NULL = default stream).  Asynchronous: returns after enqueueing.
 * Re-entrant: every launch in flight owns its scratch slabs and work counter (a pool keyed by stream), so calls on
 * different streams -- F and H mixed -- may overlap; calls on one stream are ordered by the stream.  Input pointers
 * need only the natural 8-byte alignment of float64 (16-byte aligned [n,2] inputs are read with 128-bit loads). */
int dgb200_find_fundamental_batch_dev(const double* d_x1y1, const double* d_x2y2, int n_pairs, int n, int dim,
                                      double px_th, double conf, int max_iters, int error_type, int sym_check,
                                      double laf_coef, int degen_check, const uint64_t* d_seeds,
                                      double* d_F_out, uint8_t* d_mask_out, int32_t* d_stats_out, void* stream);
int dgb200_find_homography_batch_dev(const double* d_x1y1, const double* d_x2y2, int n_pairs, int n, int dim,
                                     double px_th, double conf, int max_iters, int error_type, int sym_check,
                                     double laf_coef, const uint64_t* d_seeds,
                                     double* d_H_out, uint8_t* d_mask_out, int32_t* d_stats_out, void* stream);

/* Flagged variants.  DGB200_FLAG_FINAL_LSQ = the reference's compile-time option __FINAL_LSQ__ (exp_ranF.h:28-29;
 * exp_ranF.c:1701-1705, exp_ranH.c:866-870): after the loop one more least-squares fit on all inliers of the best
 * model, the inlier mask is derived from the polished model's residuals. */
#define DGB200_FLAG_FINAL_LSQ 1u
int dgb200_find_fundamental_batch_ex(const double* x1y1, const double* x2y2, int n_pairs, int n, int dim,
                                     double px_th, double conf, int max_iters, int error_type, int sym_check,
                                     double laf_coef, int degen_check, const uint64_t* seeds,
                                     double* F_out, uint8_t* mask_out, int32_t* stats_out, unsigned flags);
int dgb200_find_homography_batch_ex(const double* x1y1, const double* x2y2, int n_pairs, int n, int dim,
                                    double px_th, double conf, int max_iters, int error_type, int sym_check,
                                    double laf_coef, const uint64_t* seeds,
                                    double* H_out, uint8_t* mask_out, int32_t* stats_out, unsigned flags);
int dgb200_find_fundamental_batch_dev_ex(const double* d_x1y1, const double* d_x2y2, int n_pairs, int n, int dim,
                                         double px_th, double conf, int max_iters, int error_type, int sym_check,
                                         double laf_coef, int degen_check, const uint64_t* d_seeds,
                                         double* d_F_out, uint8_t* d_mask_out, int32_t* d_stats_out, void* stream,
                                         unsigned flags);
int dgb200_find_homography_batch_dev_ex(const double* d_x1y1, const double* d_x2y2, int n_pairs, int n, int dim,
                                        double px_th, double conf, int max_iters, int error_type, int sym_check,
                                        double laf_coef, const uint64_t* d_seeds,
                                        double* d_H_out, uint8_t* d_mask_out, int32_t* d_stats_out, void* stream,
                                        unsigned flags);

/* Ragged batches (proposal 3): real tentative sets never share n.  The correspondences of all
 * pairs are concatenated ([offsets[n_pairs]][dim] float64); pair p owns rows offsets[p] .. offsets[p+1]-1
 * (offsets[0] = 0, int32, every pair n >= 8 for F / n >= 4 for H); mask_out is concatenated the same way
 * ([offsets[n_pairs]] uint8); models and stats stay [n_pairs][9] / [n_pairs][4].  Same kernel, same results as one
 * call per pair.  HOST buffers; synchronous. */
int dgb200_find_fundamental_ragged(const double* x1y1, const double* x2y2, const int32_t* offsets, int n_pairs, int dim,
                                   double px_th, double conf, int max_iters, int error_type, int sym_check,
                                   double laf_coef, int degen_check, const uint64_t* seeds,
                                   double* F_out, uint8_t* mask_out, int32_t* stats_out);
int dgb200_find_homography_ragged(const double* x1y1, const double* x2y2, const int32_t* offsets, int n_pairs, int dim,
                                  double px_th, double conf, int max_iters, int error_type, int sym_check,
                                  double laf_coef, const uint64_t* seeds,
                                  double* H_out, uint8_t* mask_out, int32_t* stats_out);
/* DEVICE-pointer flavour of the ragged batches (d_offsets in device memory; n_max = the largest pair, it sizes the
 * per-CTA scratch).  Asynchronous on `stream`. */
int dgb200_find_fundamental_ragged_dev(const double* d_x1y1, const double* d_x2y2, const int32_t* d_offsets, int n_pairs,
                                       int n_max, int dim, double px_th, double conf, int max_iters, int error_type,
                                       int sym_check, double laf_coef, int degen_check, const uint64_t* d_seeds,
                                       double* d_F_out, uint8_t* d_mask_out, int32_t* d_stats_out, void* stream);
int dgb200_find_homography_ragged_dev(const double* d_x1y1, const double* d_x2y2, const int32_t* d_offsets, int n_pairs,
                                      int n_max, int dim, double px_th, double conf, int max_iters, int error_type,
                                      int sym_check, double laf_coef, const uint64_t* d_seeds,
                                      double* d_H_out, uint8_t* d_mask_out, int32_t* d_stats_out, void* stream);

/* Homography from correspondences of local ELLIPTICAL features: the reference's ransacH2el (ranH2el.c:19-208; no binding
 * in the reference -- the solver its C core carries for two-correspondence samples).  u10: [n_pairs][n][10] rows
 * (x', y', a', b', c', x, y, a, b, c) as ranH2el.h:4 (image 1 first; each local frame the lower-triangular affinity
 * [a 0; b c] at (x, y)).  th = px_th^2 on the Sampson error of the centres, do_lo = 1, inlLimit = 0.  H_out: RAW core
 * output like dgb200_find_homography (column-major; inv(H^T) maps (x', y') -> (x, y)).  stats: {samples, LO runs, 0,
 * inliers of the returned model}. */
int dgb200_find_homography_2el_batch(const double* u10, int n_pairs, int n, double px_th, double conf, int max_iters,
                                     const uint64_t* seeds, double* H_out, uint8_t* mask_out, int32_t* stats_out);
int dgb200_find_homography_2el_batch_dev(const double* d_u10, int n_pairs, int n, double px_th, double conf,
                                         int max_iters, const uint64_t* d_seeds, double* d_H_out, uint8_t* d_mask_out,
                                         int32_t* d_stats_out, void* stream);

/* One pair (what one findFundamentalMatrix_/findHomography_ call does): batch of 1 with one seed. */
int dgb200_find_fundamental(const double* x1y1, const double* x2y2, int n, int dim, double px_th, double conf,
                            int max_iters, int error_type, int sym_check, double laf_coef, int degen_check,
                            uint64_t seed, double* F_out, uint8_t* mask_out, int32_t* stats_out);
int dgb200_find_homography(const double* x1y1, const double* x2y2, int n, int dim, double px_th, double conf,
                           int max_iters, int error_type, int sym_check, double laf_coef,
                           uint64_t seed, double* H_out, uint8_t* mask_out, int32_t* stats_out);

/* ---- the steps either side of the path, device-resident, asynchronous on `stream` ----
 *
 * Descriptor matching: replaces the host side of the reference's pipeline, cv2.BFMatcher().knnMatch(descs1, descs2, k=2)
 * + SNN ratio test `m.distance < ratio * n.distance` (examples/simple-example.py:46-53), optionally with a mutual
 * nearest-neighbour check.  d_desc1 [n1][D], d_desc2 [n2][D] float32 (D a multiple of 4, <= 256); accepted matches in
 * ascending query order: d_match_q/d_match_t [capacity] int32, *d_count; when d_x1y1 != NULL the first `out_dim`
 * columns of the keypoint rows d_kp1 [n1][kp_dim], d_kp2 [n2][kp_dim] (float64) are gathered into d_x1y1/d_x2y2
 * [capacity][out_dim] -- exactly the arrays dgb200_find_*_batch_dev take.  d_workspace: dgb200_match_workspace_bytes(). */
size_t dgb200_match_workspace_bytes(int n1, int n2);
int dgb200_match_descriptors_dev(const float* d_desc1, int n1, const float* d_desc2, int n2, int D, float ratio, int mutual,
                                 const double* d_kp1, const double* d_kp2, int kp_dim, int* d_match_q, int* d_match_t,
                                 double* d_x1y1, double* d_x2y2, int out_dim, int capacity, int* d_count, void* d_workspace,
                                 void* stream);
/* Pose from fundamental matrices: E = K2^T F K1, the four (R, t) candidates of its SVD, cheirality vote over the
 * correspondences with mask != 0 (NULL: all).  d_K1/d_K2: one 3x3 (k_per_pair = 0) or one per pair, row-major.
 * Outputs R [n_pairs][9] row-major, t [n_pairs][3] (unit length, x2 ~ R x1 + t), good [n_pairs] = supporters or NULL.
 * (The reference stops at F; this is the survey's "step after".) */
int dgb200_pose_from_fundamental_batch_dev(const double* d_F, const double* d_K1, const double* d_K2, int k_per_pair,
                                           const double* d_x1y1, const double* d_x2y2, const uint8_t* d_mask, int n_pairs,
                                           int n, int dim, double* d_R_out, double* d_t_out, int32_t* d_good_out,
                                           void* stream);
/* The reference's alternative 7-point null-space solver nullspace_qr7x9 (Ftools.c:594-668, compile-time USE_QR,
 * exp_ranF.c:1346-1349) over `count` systems: d_A [count][7*9] row-major -> d_N [count][2*9], d_rc [count] or NULL. */
int dgb200_nullspace_qr7x9_batch_dev(const double* d_A, double* d_N, int32_t* d_rc, int count, void* stream);
/* Test hook: compares `count` random quotients of the shared-reciprocal division used by the homography Sampson residual
 * (csrc/hgeom.h: h_pinvJ) bit for bit with IEEE division on the device; *h_bad = number of mismatches (must be 0). */
int dgb200_debug_div_check(unsigned long long seed, long long count, unsigned long long* h_bad);
const char* dgb200_frontend_last_error(void);

/* Housekeeping */
int dgb200_version(void);              /* ABI version */
int dgb200_device_count(void);         /* CUDA devices visible (0 or <0: the engine cannot run) */
int dgb200_set_device(int device);     /* device used by subsequent calls of this thread's process */
const char* dgb200_last_error(void);   /* last error message of the calling thread (thread-local storage) */
long long dgb200_kernel_launches(void);/* RANSAC kernels launched so far by this process */
double dgb200_last_kernel_ms(void);    /* device time of the most recent HOST-buffer call's kernel (CUDA events) */
void dgb200_release(void);             /* free cached device buffers */

#ifdef __cplusplus
}
#endif
#endif /* DEGENSAC_B200_H */
