/*
 * degensac_port.c -- CPU RESTATEMENT of the reference's hot path (TEST INFRASTRUCTURE ONLY).
 *
 * Plain sequential C following the reference's own control flow -- one hypothesis at a time, physical residual
 * rows behind rotating pointers, LO and DEGENSAC triggered in place -- with the two substitutions every
 * non-reference build needs:
 *   * sampling: the counter-based Philox stream of the B200 engine (value31(seed,k,j), stateless minimal sample)
 *     instead of libc rand()/random() re-seeded per iteration (exp_ranF.c:1277,1331-1342; exp_ranH.c:510,539-552);
 *   * LAPACK dsyev_/dgesvd_ (lapwrap.c:21,67; unpinned system dependency): cyclic Jacobi / one-sided Jacobi here.
 * Everything else cites the reference lines it restates.  Parity is PINNED: tests/test_port_oracle.py checks this
 * file against oracle/_ref (the unmodified reference compiled from /root/reference) on the golden vectors and on
 * randomised inputs (identical masks).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it;
 * the product (pydegensac_b200) never does.
 *
 * Layouts: F[9] row-major, x2^T F x1 = 0.  h[9] column-major, maps image 2 -> image 1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ constants (rtools.h:4-15,31-41) */
#define ITER_SAM 50
#define RAN_REP 10
#define ILSQ 4
#define TCF 4
#define MWM_INT 2           /* (9/4) in integer arithmetic */
#define MAX_SAMPLES 1000000
#define P_EPS 2.2204e-16

typedef struct { unsigned I; double J; unsigned Is; unsigned Ilafs; } Sc;
typedef struct { int n; const double *x1, *y1, *x2, *y2; } Pts;     /* SoA view of the correspondences */
typedef struct { uint64_t seed; uint32_t k, j; } Stream;

/* ------------------------------------------------------------------ sampling stream */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t o[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
static uint32_t value31(uint64_t seed, uint32_t k, uint32_t j) {
  uint32_t o[4];
  philox(j >> 2, k, 0, 0, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  return o[j & 3] >> 1;
}
static uint32_t draw(Stream *s) { return value31(s->seed, s->k, s->j++); }

/* minimal sample of iteration k: partial Fisher-Yates on a fresh identity pool, draw order (cf. rtools.c:12-23) */
static void minimal_sample(uint64_t seed, uint32_t k, int N, int m, int *sel) {
  int pos[16], val[16], nl = 0;
  for (int i = 0; i < m; ++i) {
    int s = (int)(value31(seed, k, (uint32_t)i) % (uint32_t)(N - i)), top = N - i - 1, vs = s, vt = top;
    for (int t = 0; t < nl; ++t) { if (pos[t] == s) vs = val[t]; if (pos[t] == top) vt = val[t]; }
    pos[nl] = s; val[nl++] = vt; pos[nl] = top; val[nl++] = vs;
    sel[i] = vs;
  }
}
/* randsubset, rtools.c:25-39: subset = last `siz` entries after the swaps */
static int *randsubset(int *pool, int max_sz, int siz, Stream *st) {
  for (int i = 0; i < siz; ++i) {
    int s = (int)(draw(st) % (uint32_t)(max_sz - i)), j = max_sz - i - 1, q = pool[s];
    pool[s] = pool[j]; pool[j] = q;
  }
  return pool + max_sz - siz;
}

/* ------------------------------------------------------------------ small linear algebra */
/* utools.c:97-167 */
static int nullspace9(double *M, double *ns) {
  int freec[9], pivc[9], nf = 0, np = 0, row = 0;
  for (int col = 0; col < 9; ++col) {
    int best = row; double mag = fabs(M[9 * row + col]);
    for (int r = row + 1; r < 9; ++r) { double t = fabs(M[9 * r + col]); if (mag < t) { mag = t; best = r; } }
    if (mag < 1e-12) { freec[nf++] = col; for (int r = row; r < 9; ++r) M[9 * r + col] = 0; continue; }
    pivc[np++] = col;
    for (int c = col; c < 9; ++c) { double t = M[9 * row + c]; M[9 * row + c] = M[9 * best + c]; M[9 * best + c] = t; }
    double p = M[9 * row + col];
    for (int c = col; c < 9; ++c) M[9 * row + c] /= p;
    for (int r = 0; r < 9; ++r) if (r != row) { double a = M[9 * r + col]; for (int c = col; c < 9; ++c) M[9 * r + c] -= a * M[9 * row + c]; }
    ++row;
  }
  for (int k = 0; k < nf; ++k) {
    int j = freec[k];
    for (int l = 0; l < np; ++l) ns[k * 9 + pivc[l]] = -M[l * 9 + j];
    for (int l = 0; l < nf; ++l) ns[k * 9 + freec[l]] = (j == freec[l]) ? 1.0 : 0.0;
  }
  return nf;
}
/* smallest-eigenvalue eigenvector of a symmetric 9x9 (replaces lap_eig + "column of the minimum", Ftools.c:368-389) */
static void min_eigvec9(double *A, double *v) {
  double V[81];
  for (int i = 0; i < 81; ++i) V[i] = (i / 9 == i % 9);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, dia = 0;
    for (int p = 0; p < 9; ++p) { dia += A[p * 10] * A[p * 10]; for (int q = p + 1; q < 9; ++q) off += A[p * 9 + q] * A[p * 9 + q]; }
    if (!(off > 1e-31 * dia)) break;
    for (int p = 0; p < 8; ++p) for (int q = p + 1; q < 9; ++q) {
      double apq = A[p * 9 + q];
      if (apq == 0) continue;
      double th = (A[q * 10] - A[p * 10]) / (2 * apq), t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1)), c = 1 / sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < 9; ++k) { double a = A[k * 9 + p], b = A[k * 9 + q]; A[k * 9 + p] = c * a - s * b; A[k * 9 + q] = s * a + c * b; }
      for (int k = 0; k < 9; ++k) { double a = A[p * 9 + k], b = A[q * 9 + k]; A[p * 9 + k] = c * a - s * b; A[q * 9 + k] = s * a + c * b; }
      for (int k = 0; k < 9; ++k) { double a = V[k * 9 + p], b = V[k * 9 + q]; V[k * 9 + p] = c * a - s * b; V[k * 9 + q] = s * a + c * b; }
    }
  }
  int j = 0;
  for (int i = 1; i < 9; ++i) if (A[i * 10] < A[j * 10]) j = i;
  for (int i = 0; i < 9; ++i) v[i] = V[i * 9 + j];
}
/* rank-2 projection (singulF, Ftools.c:330-347): F - sigma_min u v^T via one-sided Jacobi */
static void rank2(double *F) {
  double G[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(G, F, sizeof G);
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rot = 0;
    for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
      double al = 0, be = 0, ga = 0;
      for (int i = 0; i < 3; ++i) { al += G[3 * i + p] * G[3 * i + p]; be += G[3 * i + q] * G[3 * i + q]; ga += G[3 * i + p] * G[3 * i + q]; }
      if (ga == 0 || ga * ga <= 1e-30 * al * be) continue;
      rot = 1;
      double z = (be - al) / (2 * ga), t = (z >= 0 ? 1.0 : -1.0) / (fabs(z) + sqrt(1 + z * z)), c = 1 / sqrt(1 + t * t), s = c * t;
      for (int i = 0; i < 3; ++i) {
        double a = G[3 * i + p], b = G[3 * i + q]; G[3 * i + p] = c * a - s * b; G[3 * i + q] = s * a + c * b;
        a = V[3 * i + p]; b = V[3 * i + q]; V[3 * i + p] = c * a - s * b; V[3 * i + q] = s * a + c * b;
      }
    }
    if (!rot) break;
  }
  int m = 0; double best = -1;
  for (int c = 0; c < 3; ++c) { double s = G[c] * G[c] + G[3 + c] * G[3 + c] + G[6 + c] * G[6 + c]; if (best < 0 || s < best) { best = s; m = c; } }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) F[3 * i + j] -= G[3 * i + m] * V[3 * j + m];
}
/* vector orthogonal to the columns of the 9 x len matrix (last column of U of svduv, Ftools.c:373,383) */
static void left_null(double *Z, int len, double *q) {
  double vs[8][9], beta[8];
  for (int c = 0; c < len; ++c) {
    double nrm = 0; for (int r = c; r < 9; ++r) nrm += Z[r * len + c] * Z[r * len + c];
    nrm = sqrt(nrm); memset(vs[c], 0, sizeof vs[c]); beta[c] = 0;
    if (nrm == 0) continue;
    double x0 = Z[c * len + c], alpha = x0 >= 0 ? -nrm : nrm, vn = 0;
    for (int r = c; r < 9; ++r) vs[c][r] = Z[r * len + c];
    vs[c][c] = x0 - alpha;
    for (int r = c; r < 9; ++r) vn += vs[c][r] * vs[c][r];
    beta[c] = vn > 0 ? 2 / vn : 0;
    for (int cc = c; cc < len; ++cc) { double d = 0; for (int r = c; r < 9; ++r) d += vs[c][r] * Z[r * len + cc]; d *= beta[c]; for (int r = c; r < 9; ++r) Z[r * len + cc] -= d * vs[c][r]; }
  }
  memset(q, 0, 9 * sizeof(double)); q[8] = 1;
  for (int c = len - 1; c >= 0; --c) { double d = 0; for (int r = c; r < 9; ++r) d += vs[c][r] * q[r]; d *= beta[c]; for (int r = c; r < 9; ++r) q[r] -= d * vs[c][r]; }
}
/* 3x3 inverse with the arithmetic of CCMATH minv, operation for operation (matutls/minv.c:10-71): column-wise Crout LU
 * with row pivoting, both factors inverted in place, product, interchanges undone on the columns.  Bit-for-bit equal to
 * the reference's (tests/test_small_la.py); the symmetric-transfer metrics need that on scenes where exact four-point fits
 * compete with scores that differ by rounding noise only. */
static int inv3(double *a) {
  enum { n = 3 };
  int le[n]; double q0[n], tq = 0, zr = 1.e-15;
  for (int j = 0; j < n; ++j) {
    if (j > 0) {
      for (int i = 0; i < n; ++i) q0[i] = a[i * n + j];
      for (int i = 1; i < n; ++i) { int lc = i < j ? i : j; double t = 0; for (int k = 0; k < lc; ++k) t += a[i * n + k] * q0[k]; q0[i] -= t; }
      for (int i = 0; i < n; ++i) a[i * n + j] = q0[i];
    }
    double s = fabs(a[j * n + j]); int lc = j;
    for (int k = j + 1; k < n; ++k) { double t = fabs(a[k * n + j]); if (t > s) { s = t; lc = k; } }
    tq = tq > s ? tq : s;
    if (s < zr * tq) return -1;
    le[j] = lc;
    if (lc != j) for (int k = 0; k < n; ++k) { double t = a[j * n + k]; a[j * n + k] = a[lc * n + k]; a[lc * n + k] = t; }
    double t = 1. / a[j * n + j];
    for (int k = j + 1; k < n; ++k) a[k * n + j] *= t;
    a[j * n + j] = t;
  }
  for (int j = 1; j < n; ++j) for (int k = 0; k < j; ++k) a[k * n + j] *= a[j * n + j];
  for (int j = 1; j < n; ++j) {
    for (int i = 0; i < j; ++i) q0[i] = a[i * n + j];
    for (int k = 0; k < j; ++k) { double t = 0; for (int i = k; i < j; ++i) t -= a[k * n + i] * q0[i]; q0[k] = t; }
    for (int i = 0; i < j; ++i) a[i * n + j] = q0[i];
  }
  for (int j = n - 2; j >= 0; --j) {
    int m = n - j - 1;
    for (int i = 0; i < m; ++i) q0[i] = a[(j + 1 + i) * n + j];
    for (int k = n - 1; k > j; --k) { double t = -a[k * n + j]; for (int i = j + 1; i < k; ++i) t -= a[k * n + i] * q0[i - j - 1]; q0[--m] = t; }
    m = n - j - 1;
    for (int i = 0; i < m; ++i) a[(j + 1 + i) * n + j] = q0[i];
  }
  for (int k = 0; k < n - 1; ++k) {
    for (int i = 0; i < n; ++i) q0[i] = a[i * n + k];
    for (int j = 0; j < n; ++j) {
      double t; int i;
      if (j > k) { t = 0; i = j; } else { t = q0[j]; i = k + 1; }
      for (; i < n; ++i) t += a[j * n + i] * q0[i];
      q0[j] = t;
    }
    for (int i = 0; i < n; ++i) a[i * n + k] = q0[i];
  }
  for (int j = n - 2; j >= 0; --j) { int lc = le[j]; for (int k = 0; k < n; ++k) { double t = a[k * n + j]; a[k * n + j] = a[k * n + lc]; a[k * n + lc] = t; } }
  return 0;
}
static void cross(double *o, const double *a, const double *b) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
static double det3(const double *A) { double r = A[0] * A[4] * A[8] + A[2] * A[3] * A[7] + A[1] * A[5] * A[6]; r -= A[2] * A[4] * A[6] + A[0] * A[5] * A[7] + A[1] * A[3] * A[8]; return r; }

/* Column 2 of V as CCMATH svduv leaves it for a 3x3 input (svduv.c, ldvmat.c, qrbdv.c): Householder
 * bidiagonalisation + implicit-shift QR, singular values UNSORTED.  Hdetect (DegUtils.c:109-110) uses that column
 * as the epipole whether or not it belongs to the vanishing singular value, so the sweep order is followed. */
static void ccmath_v3(const double *Ain, double *vout) {
  double a[9], d[3], e[2] = {0, 0}, V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memcpy(a, Ain, sizeof a);
  { double w0 = a[0], w1 = a[3], w2 = a[6], s = w0 * w0 + w1 * w1 + w2 * w2, h = 0;
    if (s > 0) { h = sqrt(s); if (a[0] < 0) h = -h; s += a[0] * h; s = 1 / s; w0 += h;
      for (int k = 1; k < 3; ++k) { double r = (w0 * a[k] + w1 * a[3 + k] + w2 * a[6 + k]) * s; a[k] -= r * w0; a[3 + k] -= r * w1; a[6 + k] -= r * w2; } }
    d[0] = -h; }
  double hb = 0, u2 = 0;
  { double s = a[1] * a[1] + a[2] * a[2], h = 0;
    if (s > 0) { h = sqrt(s); if (a[1] < 0) h = -h; hb = 1 + fabs(a[1] / h); s += a[1] * h; s = 1 / s;
      double p0 = a[1] + h, t = 1 / p0;
      for (int row = 1; row < 3; ++row) { double r = (p0 * a[3 * row + 1] + a[2] * a[3 * row + 2]) * s; a[3 * row + 1] -= r * p0; a[3 * row + 2] -= r * a[2]; }
      u2 = a[2] * t; }
    e[0] = -h; }
  { double w0 = a[4], w1 = a[7], s = w0 * w0 + w1 * w1, h = 0;
    if (s > 0) { h = sqrt(s); if (a[4] < 0) h = -h; s += a[4] * h; s = 1 / s; w0 += h; double r = (w0 * a[5] + w1 * a[8]) * s; a[5] -= r * w0; a[8] -= r * w1; }
    d[1] = -h; }
  e[1] = a[5]; d[2] = a[8];
  if (hb != 0) { double sd = hb * u2; V[4] = 1 - hb; V[7] = -hb * u2; V[8] = 1 - sd * u2; V[5] = -sd; }
  int m = 3; double t = fabs(d[0]);
  for (int j = 1; j < 3; ++j) { double s = fabs(d[j]) + fabs(e[j - 1]); if (s > t) t = s; }
  t *= 1e-15;
  for (int it = 0; m > 1 && it < 300; ++it) {
    int k;
    for (k = m - 1; k > 0; --k) {
      if (fabs(e[k - 1]) < t) break;
      if (fabs(d[k - 1]) < t) { double s = 1, c = 0; for (int i = k; i < m; ++i) { double aa = s * e[i - 1], bb = d[i]; e[i - 1] *= c; double u = sqrt(aa * aa + bb * bb); d[i] = u; s = -aa / u; c = bb / u; } break; }
    }
    double y = d[k], x = d[m - 1], u = e[m - 2], aa = (y + x) * (y - x) - u * u, s = y * e[k], bb = s + s;
    u = sqrt(aa * aa + bb * bb);
    if (u != 0) {
      double c = sqrt((u + aa) / (u + u));
      if (c != 0) s /= (c * u); else s = 1;
      for (int i = k; i < m - 1; ++i) {
        bb = e[i];
        if (i > k) { aa = s * e[i]; bb *= c; e[i - 1] = u = sqrt(x * x + aa * aa); c = x / u; s = aa / u; }
        aa = c * y + s * bb; bb = c * bb - s * y;
        for (int r = 0; r < 3; ++r) { double w = c * V[3 * r + i] + s * V[3 * r + i + 1]; V[3 * r + i + 1] = c * V[3 * r + i + 1] - s * V[3 * r + i]; V[3 * r + i] = w; }
        s *= d[i + 1]; d[i] = u = sqrt(aa * aa + s * s); y = c * d[i + 1]; c = aa / u; s /= u; x = c * bb + s * y; y = c * y - s * bb;
      }
    }
    e[m - 2] = x; d[m - 1] = y;
    if (fabs(x) < t) --m;
    if (m == k + 1) --m;
  }
  vout[0] = V[2]; vout[1] = V[5]; vout[2] = V[8];
}

/* hash.c:4-47 on an int32 list; hash.c:49-96 reduces to a flat (hash,len,iterID) list */
static uint32_t sfh(const int *idx, int n) {
  if (n <= 0) return 0;
  uint32_t h = (uint32_t)(4 * n);
  for (int i = 0; i < n; ++i) { uint32_t w = (uint32_t)idx[i]; h += w & 0xffff; uint32_t tmp = ((w >> 16) << 11) ^ h; h = (h << 16) ^ tmp; h += h >> 11; }
  h ^= h << 3; h += h >> 5; h ^= h << 4; h += h >> 17; h ^= h << 25; h += h >> 6;
  return h;
}
typedef struct { uint32_t *h; int *len, *id; int n, cap; } HT;
static int ht_check(HT *t, const int *list, int n, int id) {   /* 1: abort (seen under another id) */
  uint32_t h = sfh(list, n); int same = 0, other = 0;
  for (int i = 0; i < t->n; ++i) if (t->h[i] == h && t->len[i] == n) { if (t->id[i] == id) same = 1; else other = 1; }
  if (same) return 0;
  if (other) return 1;
  if (t->n == t->cap) { t->cap *= 2; t->h = realloc(t->h, t->cap * sizeof *t->h); t->len = realloc(t->len, t->cap * sizeof *t->len); t->id = realloc(t->id, t->cap * sizeof *t->id); }
  t->h[t->n] = h; t->len[t->n] = n; t->id[t->n] = id; ++t->n;
  return 0;
}

/* rtools.c:202-225, :228-236, :160-171 */
static int nsamples(int ninl, int ptNum, int samsiz, double conf) {
  double a = 1, b = 1;
  for (int i = 0; i < samsiz; ++i) { a *= ninl - i; b *= ptNum - i; }
  a = a / b;
  if (a < P_EPS) return MAX_SAMPLES;
  a = 1 - a;
  if (a < P_EPS) return 1;
  b = log(1 - conf) / log(a);
  return b > MAX_SAMPLES ? MAX_SAMPLES : (int)ceil(b);
}
static double tquad(double e, double thr) { if (thr == 0) return 0; if (e >= thr * 9 / 4) return 0; return 1 - (e / (thr * 9 / 4)); }
static Sc inlidxs(const double *err, int len, double th, int *inl) {
  Sc s = {0, 0, 0, 0};
  for (int i = 0; i < len; ++i) { s.J += tquad(err[i], th); if (err[i] <= th) inl[s.I++] = i; }
  return s;
}

/* ------------------------------------------------------------------ F geometry (Ftools.c) */
typedef struct { double rxc, ryc, r, rx, ry; } FT;
static FT fterms(const double *F, const Pts *P, int i) {   /* Ftools.c:91-96 */
  FT t; double x1 = P->x1[i], y1 = P->y1[i], x2 = P->x2[i], y2 = P->y2[i];
  t.rxc = F[0] * x2 + F[3] * y2 + F[6]; t.ryc = F[1] * x2 + F[4] * y2 + F[7];
  double rwc = F[2] * x2 + F[5] * y2 + F[8];
  t.r = (x1 * t.rxc + y1 * t.ryc + rwc); t.rx = F[0] * x1 + F[1] * y1 + F[2]; t.ry = F[3] * x1 + F[4] * y1 + F[5];
  return t;
}
static double fres(int metric, const double *F, const Pts *P, int i) {   /* FDs :83, FDsSym :147 */
  FT t = fterms(F, P, i);
  if (metric == 1) { double a = t.rxc * t.rxc + t.ryc * t.ryc, b = t.rx * t.rx + t.ry * t.ry; return t.r * t.r * (a + b) / (a * b); }
  return t.r * t.r / (t.rxc * t.rxc + t.ryc * t.ryc + t.rx * t.rx + t.ry * t.ry);
}
static void fres_all(int metric, const double *F, const Pts *P, double *d) { for (int i = 0; i < P->n; ++i) d[i] = fres(metric, F, P, i); }
static void fres_w(int metric, const double *F, const Pts *P, double *d, double *w) {   /* exFDs :124, exFDsSym :228 */
  for (int i = 0; i < P->n; ++i) {
    FT t = fterms(F, P, i);
    if (metric == 1) { double a = t.rxc * t.rxc + t.ryc * t.ryc, b = t.rx * t.rx + t.ry * t.ry; w[i] = (a * b) / (a + b); d[i] = t.r * t.r / w[i]; }
    else { double ww = t.rxc * t.rxc + t.ryc * t.ryc + t.rx * t.rx + t.ry * t.ry; d[i] = t.r * t.r / ww; w[i] = 1 / sqrt(ww); }
  }
}
static void frow(const Pts *P, int p, double *row) {   /* lin_fm, Ftools.c:15-37 */
  double x1 = P->x1[p], y1 = P->y1[p], x2 = P->x2[p], y2 = P->y2[p];
  row[0] = x2 * x1; row[1] = x2 * y1; row[2] = x2; row[3] = y2 * x1; row[4] = y2 * y1; row[5] = y2; row[6] = x1; row[7] = y1; row[8] = 1;
}
/* slcm, Ftools.c:39-81 (term order kept: the cubic is ill-conditioned, see fgeom.h of the engine) */
static void slcm(const double *A, double *B, double *p) {
  double a11 = A[0], a12 = A[1], a13 = A[2], a21 = A[3], a22 = A[4], a23 = A[5], a31 = A[6], a32 = A[7], a33 = A[8];
  double b11 = B[0], b12 = B[1], b13 = B[2], b21 = B[3], b22 = B[4], b23 = B[5], b31 = B[6], b32 = B[7], b33 = B[8];
  p[0] = -(b13 * b22 * b31) + b12 * b23 * b31 + b13 * b21 * b32 - b11 * b23 * b32 - b12 * b21 * b33 + b11 * b22 * b33;
  p[1] = -(a33 * b12 * b21) + a32 * b13 * b21 + a33 * b11 * b22 - a31 * b13 * b22 - a32 * b11 * b23 + a31 * b12 * b23 +
         a23 * b12 * b31 - a22 * b13 * b31 - a13 * b22 * b31 + 3 * b13 * b22 * b31 + a12 * b23 * b31 - 3 * b12 * b23 * b31 -
         a23 * b11 * b32 + a21 * b13 * b32 + a13 * b21 * b32 - 3 * b13 * b21 * b32 - a11 * b23 * b32 + 3 * b11 * b23 * b32 +
         (a22 * b11 - a21 * b12 - a12 * b21 + 3 * b12 * b21 + a11 * b22 - 3 * b11 * b22) * b33;
  p[2] = -(a21 * a33 * b12) + a21 * a32 * b13 + a13 * a32 * b21 - a12 * a33 * b21 + 2 * a33 * b12 * b21 - 2 * a32 * b13 * b21 -
         a13 * a31 * b22 + a11 * a33 * b22 - 2 * a33 * b11 * b22 + 2 * a31 * b13 * b22 + a12 * a31 * b23 - a11 * a32 * b23 +
         2 * a32 * b11 * b23 - 2 * a31 * b12 * b23 + 2 * a13 * b22 * b31 - 3 * b13 * b22 * b31 - 2 * a12 * b23 * b31 +
         3 * b12 * b23 * b31 + a13 * a21 * b32 - 2 * a21 * b13 * b32 - 2 * a13 * b21 * b32 + 3 * b13 * b21 * b32 +
         2 * a11 * b23 * b32 - 3 * b11 * b23 * b32 + a23 * (-(a32 * b11) + a31 * b12 + a12 * b31 - 2 * b12 * b31 - a11 * b32 + 2 * b11 * b32) +
         (-(a12 * a21) + 2 * a21 * b12 + 2 * a12 * b21 - 3 * b12 * b21 - 2 * a11 * b22 + 3 * b11 * b22) * b33 +
         a22 * (a33 * b11 - a31 * b13 - a13 * b31 + 2 * b13 * b31 + a11 * b33 - 2 * b11 * b33);
  for (int i = 0; i < 9; ++i) B[i] = A[i] - B[i];
  b11 = B[0]; b12 = B[1]; b13 = B[2]; b21 = B[3]; b22 = B[4]; b23 = B[5]; b31 = B[6]; b32 = B[7]; b33 = B[8];
  p[3] = -(b13 * b22 * b31) + b12 * b23 * b31 + b13 * b21 * b32 - b11 * b23 * b32 - b12 * b21 * b33 + b11 * b22 * b33;
}
static int rroots3(const double *po, double *r) {   /* Ftools.c:251-298 */
  double b = po[1] / po[0], c = po[2] / po[0], b2 = b * b, bt = b / 3, p = (3 * c - b2) / 9;
  double q = ((2 * b2 * b) / 27 - b * c / 3 + po[3] / po[0]) / 2, D = q * q + p * p * p;
  if (D > 0) {
    double A = sqrt(D) - q;
    if (A > 0) { double v = pow(A, 1.0 / 3); r[0] = v - p / v - bt; } else { double v = pow(-A, 1.0 / 3); r[0] = p / v - v - bt; }
    return 1;
  }
  double e = q > 0 ? 1 : -1, R = e * sqrt(-p), R2 = R * 2, cp = q / (R * R * R);
  if (cp > 1) cp = 1; else if (cp < -1) cp = -1;
  double ph = acos(cp) / 3, pit = 1.0471975511965967;
  r[0] = -R2 * cos(ph) - bt; r[1] = R2 * cos(pit - ph) - bt; r[2] = R2 * cos(pit + ph) - bt;
  return 3;
}
static int ori_ok_F(const double *F, const Pts *P, const int *idx, int n) {   /* Ftools.c:461-494 */
  double ec[3]; int big = 0;
  cross(ec, F, F + 6);
  for (int i = 0; i < 3; ++i) if (ec[i] > 1.9984e-15 || ec[i] < -1.9984e-15) big = 1;
  if (!big) cross(ec, F + 3, F + 6);
  double sig1 = 0;
  for (int i = 0; i < n; ++i) {
    int p = idx[i];
    double s1 = F[0] * P->x2[p] + F[3] * P->y2[p] + F[6] * 1.0, s2 = ec[1] * 1.0 - ec[2] * P->y1[p], sig = s1 * s2;
    if (i == 0) sig1 = sig; else if (sig1 * sig < 0) return 0;
  }
  return 1;
}
/* normu, utools.c:7-51 */
static void normu(const Pts *P, const int *inl, int len, double *A1, double *A2) {
  A1[0] = A1[1] = A1[2] = A2[0] = A2[1] = A2[2] = 0;
  for (int j = 0; j < len; ++j) { int p = inl[j]; A1[1] += P->x1[p]; A1[2] += P->y1[p]; A2[1] += P->x2[p]; A2[2] += P->y2[p]; }
  if (len > 0) for (int i = 1; i < 3; ++i) { A1[i] /= len; A2[i] /= len; }
  for (int j = 0; j < len; ++j) {
    int p = inl[j]; double a = P->x1[p] - A1[1], b = P->y1[p] - A1[2];
    A1[0] += sqrt(a * a + b * b); a = P->x2[p] - A2[1]; b = P->y2[p] - A2[2]; A2[0] += sqrt(a * a + b * b);
  }
  if (A1[0] != 0) A1[0] = len * sqrt(2) / A1[0];
  if (A2[0] != 0) A2[0] = len * sqrt(2) / A2[0];
  A1[1] *= -A1[0]; A1[2] *= -A1[0]; A2[1] *= -A2[0]; A2[2] *= -A2[0];
}
static void denormF(double *F, const double *A1, const double *A2) {   /* utools.c:53-70 */
  double r = A2[0], x = A2[1], y = A2[2];
  F[6] += x * F[0] + y * F[3]; F[7] += x * F[1] + y * F[4]; F[8] += x * F[2] + y * F[5];
  for (int i = 0; i < 6; ++i) F[i] *= r;
  r = A1[0]; x = A1[1]; y = A1[2];
  F[2] += x * F[0] + y * F[1]; F[5] += x * F[3] + y * F[4]; F[8] += x * F[6] + y * F[7];
  F[0] *= r; F[3] *= r; F[6] *= r; F[1] *= r; F[4] *= r; F[7] *= r;
}
/* u2f / u2fw, Ftools.c:350-458 (w == NULL: unweighted) */
static void u2f(const Pts *P, const int *inl, int len, const double *w, double *F) {
  if (len <= 8) {
    double Z[72], row[9];
    for (int i = 0; i < len; ++i) { frow(P, inl[i], row); for (int r = 0; r < 9; ++r) Z[r * len + i] = row[r]; }
    if (w) for (int i = 0; i < len; ++i) for (int t = 0; t < 9; ++t) { int lin = i + 9 * t; if (lin < 9 * len) Z[lin] *= w[inl[i]]; }  /* scalmul(Z+i,w,9,9), Ftools.c:431 */
    if (len > 0) left_null(Z, len, F); else { memset(F, 0, 72); F[8] = 1; }
    rank2(F);
    return;
  }
  double A1[3], A2[3], C[81];
  normu(P, inl, len, A1, A2);
  memset(C, 0, sizeof C);
  for (int j = 0; j < len; ++j) {   /* lin_fmN :300-328 + cov_mat utools.c:170-184 */
    int p = inl[j]; double a[3], b[3], row[9];
    a[0] = P->x1[p] * A1[0] + A1[1]; a[1] = P->y1[p] * A1[0] + A1[2]; a[2] = 1;
    b[0] = P->x2[p] * A2[0] + A2[1]; b[1] = P->y2[p] * A2[0] + A2[2]; b[2] = 1;
    for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) row[3 * k + l] = a[l] * b[k];
    if (w) for (int k = 0; k < 9; ++k) row[k] *= w[p];
    for (int i = 0; i < 9; ++i) for (int jj = 0; jj <= i; ++jj) C[9 * i + jj] += row[i] * row[jj];
  }
  for (int i = 0; i < 9; ++i) for (int jj = 0; jj < i; ++jj) C[9 * jj + i] = C[9 * i + jj];
  min_eigvec9(C, F);
  rank2(F);
  denormF(F, A1, A2);
}

/* ------------------------------------------------------------------ H geometry (Htools.c) */
static void hrows(const Pts *P, int p, double *r0, double *r1) {   /* lin_hg :20-58 */
  double x1 = P->x1[p], y1 = P->y1[p], x2 = P->x2[p], y2 = P->y2[p];
  r0[0] = x2; r0[1] = 0; r0[2] = -x1 * x2; r0[3] = y2; r0[4] = 0; r0[5] = -x1 * y2; r0[6] = 1; r0[7] = 0; r0[8] = -x1 * 1.0;
  r1[0] = 0; r1[1] = x2; r1[2] = -y1 * x2; r1[3] = 0; r1[4] = y2; r1[5] = -y1 * y2; r1[6] = 0; r1[7] = 1; r1[8] = -y1 * 1.0;
}
static double hres_sampson(const double *H, const Pts *P, int i) {   /* HDs :161-199 + pinvJ :135-159 */
  double x1 = P->x1[i], y1 = P->y1[i], x2 = P->x2[i], y2 = P->y2[i], r0[9], r1[9], ra = 0, rb = 0, pJ[8];
  hrows(P, i, r0, r1);
  for (int j = 0; j < 9; ++j) { ra += H[j] * r0[j]; rb += H[j] * r1[j]; }
  double a = H[0] - H[2] * x1, b = H[3] - H[5] * x1, c = -H[8] - H[2] * x2 - H[5] * y2, d = H[1] - H[2] * y1, e = H[4] - H[5] * y1;
  double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d, e2 = e * e, c2pd2 = c2 + d2, ab = a * b, de = d * e, Q = c * (c2pd2 + e2);
  pJ[0] = -b * de + a * (c2 + e2); pJ[1] = b * c2pd2 - a * de; pJ[2] = Q; pJ[3] = -c * (a * d + b * e);
  pJ[4] = d * (b2 + c2) - ab * e; pJ[5] = -ab * d + e * (a2 + c2); pJ[6] = pJ[3]; pJ[7] = c * (a2 + b2 + c2);
  double N = a * pJ[0] + b * pJ[1] + c * pJ[2], p = 0;
  for (int j = 0; j < 8; ++j) pJ[j] /= N;
  for (int j = 0; j < 4; ++j) { double t = pJ[j] * ra + pJ[j + 4] * rb; p += t * t; }
  return p;
}
typedef struct { double Hi[9], H1[9]; } HS;
static void hsym(const double *H, HS *s) {   /* the Hinv/H1 prologue of every HDsSym* (Htools.c:209-221) */
  s->Hi[0] = H[0]; s->Hi[1] = H[3]; s->Hi[2] = H[6]; s->Hi[3] = H[1]; s->Hi[4] = H[4]; s->Hi[5] = H[7]; s->Hi[6] = H[2]; s->Hi[7] = H[5]; s->Hi[8] = H[8];
  memcpy(s->H1, s->Hi, sizeof s->Hi); inv3(s->H1);
}
static void hd1d2(const HS *s, const Pts *P, int i, double eps, double *d1, double *d2) {
  double x1 = P->x1[i], y1 = P->y1[i], x2 = P->x2[i], y2 = P->y2[i];
  double a = s->H1[6] * x1 + s->H1[7] * y1 + s->H1[8] + eps, b = s->Hi[6] * x2 + s->Hi[7] * y2 + s->Hi[8] + eps;
  double xa = (s->H1[0] * x1 + s->H1[1] * y1 + s->H1[2]) / a, ya = (s->H1[3] * x1 + s->H1[4] * y1 + s->H1[5]) / a, xd = x2 - xa, yd = y2 - ya;
  *d1 = xd * xd + yd * yd;
  xa = (s->Hi[0] * x2 + s->Hi[1] * y2 + s->Hi[2]) / b; ya = (s->Hi[3] * x2 + s->Hi[4] * y2 + s->Hi[5]) / b; xd = x1 - xa; yd = y1 - ya;
  *d2 = xd * xd + yd * yd;
}
static double hres(int metric, const double *H, const HS *s, const Pts *P, int i) {   /* :161, :202-370 */
  double d1, d2;
  if (metric == 0) return hres_sampson(H, P, i);
  if (metric == 3 || metric == 4) { hd1d2(s, P, i, 1e-10, &d1, &d2); return metric == 3 ? d1 + d2 : sqrt(d1) + sqrt(d2); }
  hd1d2(s, P, i, 0.0, &d1, &d2);
  double m = d1 < d2 ? d2 : d1;
  return metric == 1 ? m : sqrt(m);
}
static void hres_all(int metric, const double *H, const Pts *P, double *d) { HS s; if (metric) hsym(H, &s); for (int i = 0; i < P->n; ++i) d[i] = hres(metric, H, &s, P, i); }
static double hgate(const HS *s, const Pts *P, int i) { double d1, d2; hd1d2(s, P, i, 1e-10, &d1, &d2); return sqrt(d1 < d2 ? d2 : d1); }   /* HDsSymMaxidx :734 */
/* Residual of a LAF helper correspondence as the "i"/"idx" metric variants compute it (Htools.c:372-815): for the
 * Sampson metric the linearised pair comes from the rows Z of the MAIN correspondence and only the Jacobian from the
 * helper point; the four symmetric variants use the helper point alone and all add 1e-10 to both denominators. */
static double hres_laf(int metric, const double *H, const HS *s, const Pts *P, const Pts *L, int i) {
  if (metric == 0) {
    double r0[9], r1[9], ra = 0, rb = 0, pJ[8];
    hrows(P, i, r0, r1);
    for (int j = 0; j < 9; ++j) { ra += H[j] * r0[j]; rb += H[j] * r1[j]; }
    double x1 = L->x1[i], y1 = L->y1[i], x2 = L->x2[i], y2 = L->y2[i];
    double a = H[0] - H[2] * x1, b = H[3] - H[5] * x1, c = -H[8] - H[2] * x2 - H[5] * y2, d = H[1] - H[2] * y1, e = H[4] - H[5] * y1;
    double a2 = a * a, b2 = b * b, c2 = c * c, d2 = d * d, e2 = e * e, c2pd2 = c2 + d2, ab = a * b, de = d * e, Q = c * (c2pd2 + e2);
    pJ[0] = -b * de + a * (c2 + e2); pJ[1] = b * c2pd2 - a * de; pJ[2] = Q; pJ[3] = -c * (a * d + b * e);
    pJ[4] = d * (b2 + c2) - ab * e; pJ[5] = -ab * d + e * (a2 + c2); pJ[6] = pJ[3]; pJ[7] = c * (a2 + b2 + c2);
    double N = a * pJ[0] + b * pJ[1] + c * pJ[2], p = 0;
    for (int j = 0; j < 8; ++j) pJ[j] /= N;
    for (int j = 0; j < 4; ++j) { double t = pJ[j] * ra + pJ[j + 4] * rb; p += t * t; }
    return p;
  }
  double d1, d2;
  hd1d2(s, L, i, 1e-10, &d1, &d2);
  if (metric == 3) return d1 + d2;
  if (metric == 4) return sqrt(d1) + sqrt(d2);
  double m = d1 < d2 ? d2 : d1;
  return metric == 1 ? m : sqrt(m);
}
static int lafcountH(int metric, const double *H, const Pts *P, const Pts *L, const int *list, unsigned n, double th_laf) {
  HS s; if (metric) hsym(H, &s);
  int c = 0;
  for (unsigned j = 0; j < n; ++j) if (hres_laf(metric, H, &s, P, L, list[j]) <= th_laf) ++c;
  return c;
}
static int ori_ok_H(const Pts *P, const int *idx) {   /* all_Hori_valid :821-848 */
  double A[4][3], B[4][3], p[3], q[3];
  for (int i = 0; i < 4; ++i) { A[i][0] = P->x1[idx[i]]; A[i][1] = P->y1[idx[i]]; A[i][2] = 1; B[i][0] = P->x2[idx[i]]; B[i][1] = P->y2[idx[i]]; B[i][2] = 1; }
#define DOT3(u, v) ((u)[0] * (v)[0] + (u)[1] * (v)[1] + (u)[2] * (v)[2])
  cross(p, A[0], A[1]); cross(q, B[0], B[1]);
  if (DOT3(p, A[2]) * DOT3(q, B[2]) < 0) return 0;
  if (DOT3(p, A[3]) * DOT3(q, B[3]) < 0) return 0;
  cross(p, A[2], A[3]); cross(q, B[2], B[3]);
  if (DOT3(p, A[0]) * DOT3(q, B[0]) < 0) return 0;
  if (DOT3(p, A[1]) * DOT3(q, B[1]) < 0) return 0;
  return 1;
}
static int hsingular(const double *h) {   /* exp_ranH.c:29-44 */
  double v = det3(h), tol = h[8];
  if (tol == 0) { for (int i = 0; i < 9; ++i) tol += h[i] * h[i]; tol = sqrt(tol); tol *= 0.001; }
  tol = tol * tol * tol;
  return fabs(v / tol) < 1e-2;
}
static void denormH(double *F, const double *A1, const double *A2) {   /* utools.c:72-92 */
  double r = A2[0], x = A2[1], y = A2[2];
  F[6] += x * F[0] + y * F[3]; F[7] += x * F[1] + y * F[4]; F[8] += x * F[2] + y * F[5];
  for (int i = 0; i < 6; ++i) F[i] *= r;
  r = 1 / A1[0]; x = -A1[1] * r; y = -A1[2] * r;
  for (int i = 0; i < 9; i += 3) { F[i] = r * F[i] + x * F[i + 2]; F[i + 1] = r * F[i + 1] + y * F[i + 2]; }
}
/* u2h, Htools.c:101-133.  len == 4 restates what the reference actually executes: an 8-stride buffer transposed
 * as 9x9 (the 9 uninitialised entries taken as 0), last row zeroed, first null vector. */
static void u2h(const Pts *P, const int *inl, int len, double *H) {
  if (len < 4) return;
  if (len == 4) {
    double Z[81], T[81], sol[81], r0[9], r1[9];
    memset(Z, 0, sizeof Z); memset(sol, 0, sizeof sol);
    for (int i = 0; i < 4; ++i) { hrows(P, inl[i], r0, r1); for (int c = 0; c < 9; ++c) { Z[8 * c + 2 * i] = r0[c]; Z[8 * c + 2 * i + 1] = r1[c]; } }
    for (int r = 0; r < 9; ++r) for (int c = 0; c < 9; ++c) T[9 * r + c] = Z[9 * c + r];
    for (int i = 72; i < 81; ++i) T[i] = 0;
    nullspace9(T, sol);
    memcpy(H, sol, 72);
    return;
  }
  double A1[3], A2[3], C[81];
  normu(P, inl, len, A1, A2);
  memset(C, 0, sizeof C);
  for (int j = 0; j < len; ++j) {   /* lin_hgN :60-99 */
    int p = inl[j]; double a[3], b[3], r0[9], r1[9];
    a[0] = P->x1[p] * A1[0] + A1[1]; a[1] = P->y1[p] * A1[0] + A1[2]; a[2] = 1;
    b[0] = P->x2[p] * A2[0] + A2[1]; b[1] = P->y2[p] * A2[0] + A2[2]; b[2] = 1;
    for (int t = 0; t < 3; ++t) { r0[3 * t] = b[t]; r0[3 * t + 1] = 0; r0[3 * t + 2] = -a[0] * b[t]; r1[3 * t] = 0; r1[3 * t + 1] = b[t]; r1[3 * t + 2] = -a[1] * b[t]; }
    for (int i = 0; i < 9; ++i) for (int jj = 0; jj <= i; ++jj) { C[9 * i + jj] += r0[i] * r0[jj]; C[9 * i + jj] += r1[i] * r1[jj]; }
  }
  for (int i = 0; i < 9; ++i) for (int jj = 0; jj < i; ++jj) C[9 * jj + i] = C[9 * i + jj];
  min_eigvec9(C, H);
  denormH(H, A1, A2);
}

/* ------------------------------------------------------------------ DEGENSAC (DegUtils.c) */
static void hdetect(const double *F, const Pts *S7, const int *tri, double *H) {   /* :93-161 */
  double ec[3], A[9], b[3], M[9], v[3];
  ccmath_v3(F, ec);
  double Ex[9] = {0, -ec[2], ec[1], ec[2], 0, -ec[0], -ec[1], ec[0], 0};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += Ex[3 * i + k] * F[3 * j + k]; A[3 * i + j] = s; }
  for (int t = 0; t < 3; ++t) {
    int p = tri[t]; double a1[3] = {S7->x1[p], S7->y1[p], 1}, a2[3] = {S7->x2[p], S7->y2[p], 1}, Ab[3], p1[3], p2[3];
    for (int i = 0; i < 3; ++i) { double s = 0; for (int k = 0; k < 3; ++k) s += A[3 * i + k] * a2[k]; Ab[i] = s; }
    cross(p1, a1, Ab);
    for (int i = 0; i < 3; ++i) { double s = 0; for (int k = 0; k < 3; ++k) s += (-Ex[3 * i + k]) * a1[k]; p2[i] = s; }
    b[t] = (p1[0] * p2[0] + p1[1] * p2[1] + p1[2] * p2[2]) / (p2[0] * p2[0] + p2[1] * p2[1] + p2[2] * p2[2]);
    M[3 * t] = a2[0]; M[3 * t + 1] = a2[1]; M[3 * t + 2] = a2[2];
  }
  int sing = inv3(M);
  for (int i = 0; i < 3; ++i) v[i] = M[3 * i] * b[0] + M[3 * i + 1] * b[1] + M[3 * i + 2] * b[2];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) H[i + 3 * j] = A[3 * i + j] - ec[i] * v[j];
  if (isnan(H[0]) || isinf(H[0]) || sing) { memset(H, 0, 72); H[0] = H[4] = H[8] = 1; }
}
static int checksample(const double *F, const Pts *S7, double th, double *H) {   /* :42-82 */
  static const int TRI[5][3] = {{0, 1, 2}, {3, 4, 5}, {0, 1, 6}, {3, 4, 6}, {2, 5, 6}};
  for (int t = 0; t < 5; ++t) {
    double Ds[7]; int idx[7], cnt = 0;
    hdetect(F, S7, TRI[t], H);
    for (int j = 0; j < 7; ++j) { Ds[j] = hres_sampson(H, S7, j); idx[j] = j; }
    for (int i = 0; i < 7; ++i) for (int j = i + 1; j < 7; ++j) if (Ds[j] < Ds[i]) { double td = Ds[j]; Ds[j] = Ds[i]; Ds[i] = td; int ti = idx[j]; idx[j] = idx[i]; idx[i] = ti; }   /* sortDs :164 */
    u2h(S7, idx, 5, H);
    for (int j = 0; j < 7; ++j) if (hres_sampson(H, S7, j) < th) ++cnt;
    if (cnt > 4) return 1;
  }
  return 0;
}
/* iterH / inHrani (ranH.c:18-135) as reached from innerH (DegUtils.c:693-731) */
static Sc iterH_old(const Pts *P, int *inliers, double th, double ths, double *H, double **errs, unsigned lim, Stream *st) {
  double *d = errs[1], h[9], dth = (ths - th) / ILSQ;
  Sc S = {0, 0, 0, 0}, Ss, maxS = inlidxs(errs[4], P->n, th, inliers);
  if (maxS.I < 4) return S;
  memcpy(h, H, 72);
  if (maxS.I <= lim) u2h(P, inliers, (int)maxS.I, h); else u2h(P, randsubset(inliers, (int)maxS.I, (int)lim, st), (int)lim, h);
  for (int it = 0; it < ILSQ; ++it) {
    hres_all(0, h, P, d);
    S = inlidxs(d, P->n, th, inliers); Ss = inlidxs(d, P->n, ths, inliers);
    if (maxS.J < S.J) { maxS = S; errs[1] = errs[0]; errs[0] = d; d = errs[1]; memcpy(H, h, 72); }
    if (Ss.I < 4) return maxS;
    if (Ss.I <= lim) u2h(P, inliers, (int)Ss.I, h); else u2h(P, randsubset(inliers, (int)Ss.I, (int)lim, st), (int)lim, h);
    ths -= dth;
  }
  hres_all(0, h, P, d);
  S = inlidxs(d, P->n, th, inliers);
  if (maxS.J < S.J) { maxS = S; errs[1] = errs[0]; errs[0] = d; memcpy(H, h, 72); }
  return maxS;
}
static unsigned innerH(double *H, const Pts *P, double th, unsigned lim, unsigned char *mask, Stream *st) {
  int n = P->n, *inliers = malloc(n * sizeof(int)), *intbuff = malloc(n * sizeof(int));
  double *err = malloc(4 * n * sizeof(double)), *errs[5] = {err, err + n, err + 2 * n, err + 3 * n, err + 3 * n};
  hres_all(0, H, P, errs[0]);
  Sc S = inlidxs(errs[0], n, th, inliers);
  int ninl = (int)S.I;
  if (ninl >= 8) {
    Sc maxS = {0, 0, 0, 0}; int ssiz = ninl / 2 > 12 ? 12 : ninl / 2; double h[9], *t;
    memcpy(h, H, 72);
    t = errs[2]; errs[2] = errs[0]; errs[0] = t;
    for (int rep = 0; rep < RAN_REP; ++rep) {
      u2h(P, randsubset(inliers, ninl, ssiz, st), ssiz, h);
      hres_all(0, h, P, errs[0]); errs[4] = errs[0];
      S = iterH_old(P, intbuff, th, TCF * th, h, errs, lim, st);
      if (maxS.J < S.J) { maxS = S; t = errs[2]; errs[2] = errs[0]; errs[0] = t; memcpy(H, h, 72); }
    }
    t = errs[2]; errs[2] = errs[0]; errs[0] = t;
  }
  unsigned I = 0;
  for (int j = 0; j < n; ++j) { mask[j] = errs[0][j] <= th; I += mask[j]; }
  free(err); free(inliers); free(intbuff);
  return I;
}
static unsigned u2Fit(const Pts *P, double *F, unsigned char *inl, double th, double ths, unsigned iters) {   /* :635-690 */
  int n = P->n, *list = malloc(n * sizeof(int)); double *Ds = malloc(n * sizeof(double)), dth = (ths - th) / (iters - 1); unsigned no_i = 0;
  for (unsigned it = 0; it < iters; ++it) {
    fres_all(0, F, P, Ds); no_i = 0;
    for (int i = 0; i < n; ++i) { inl[i] = Ds[i] < ths; no_i += inl[i]; }
    if (no_i < 8) { free(list); free(Ds); return no_i; }
    no_i = 0; for (int i = 0; i < n; ++i) if (inl[i]) list[no_i++] = i;
    u2f(P, list, (int)no_i, NULL, F);
    ths -= dth;
  }
  fres_all(0, F, P, Ds); no_i = 0;
  for (int i = 0; i < n; ++i) { inl[i] = Ds[i] < th; no_i += inl[i]; }
  free(list); free(Ds);
  return no_i;
}
static void innerFH(const Pts *P, const int *uH, int nH, const int *uO, int nO, double th, double *F, unsigned char *inl, Stream *st) {   /* :488-632 */
  int n = P->n; unsigned char *v = malloc(n); double *Ds = malloc(n * sizeof(double)), aF[9]; unsigned max_i = 0, max_s = 0;
  for (int i = 0; i < 9; ++i) F[i] = 1;
  memset(inl, 0, n);
  for (int rep = 0; rep < 15; ++rep) {
    int usam[10], *pa = malloc(nH * sizeof(int)), *pb = malloc(nO * sizeof(int));
    for (int i = 0; i < nH; ++i) pa[i] = i;
    for (int i = 0; i < nO; ++i) pb[i] = i;
    for (int pos = 0; pos < 6; ++pos) { int idx = (int)(draw(st) % (uint32_t)nH), t = pa[pos]; pa[pos] = pa[idx]; pa[idx] = t; }
    for (int pos = 0; pos < 4; ++pos) { int idx = (int)(draw(st) % (uint32_t)nO), t = pb[pos]; pb[pos] = pb[idx]; pb[idx] = t; }
    for (int i = 0; i < 6; ++i) usam[i] = uH[pa[i]];
    for (int i = 0; i < 4; ++i) usam[6 + i] = uO[pb[i]];
    free(pa); free(pb);
    u2f(P, usam, 10, NULL, aF);
    fres_all(0, aF, P, Ds);
    unsigned no_i = 0;
    for (int i = 0; i < n; ++i) { v[i] = Ds[i] < th; no_i += v[i]; }
    if (max_i < no_i) { memcpy(inl, v, n); memcpy(F, aF, 72); max_i = no_i; }
    if (no_i > max_s) { max_s = no_i; no_i = u2Fit(P, aF, v, th, th * 3, 4); if (max_i < no_i) { memcpy(inl, v, n); memcpy(F, aF, 72); max_i = no_i; } }
  }
  free(v); free(Ds);
}
static unsigned rFtH(const Pts *P, const unsigned char *hinl, double th, const double *H, double *F, Stream *st) {   /* :254-444 */
  int n = P->n, nN = 0, nH = 0, *uN = malloc(n * sizeof(int)), *uHl = malloc(n * sizeof(int)), *uV = malloc(n * sizeof(int));
  unsigned char *nh = malloc(n), *inl = malloc(n); double *Ds = malloc(n * sizeof(double));
  for (int i = 0; i < n; ++i) { nh[i] = hres_sampson(H, P, i) > 100 * th; if (nh[i]) uN[nN++] = i; if (hinl[i]) uHl[nH++] = i; }
  unsigned max_i = 3, m_i = 4, max_sam = 10000;
  if (nN < 4 || nH < 6) max_i = 0;
  else {
    unsigned *ptr = malloc(nN * sizeof(unsigned));
    for (int i = 0; i < nN; ++i) ptr[i] = i;
    for (unsigned no_sam = 1; no_sam < 2 * max_sam; ++no_sam) {
      for (int pos = 0; pos < 2; ++pos) { unsigned idx = pos + 1 + draw(st) % (uint32_t)(nN - pos - 1), t = ptr[pos]; ptr[pos] = ptr[idx]; ptr[idx] = t; }
      int a = uN[ptr[0]], b = uN[ptr[1]];
      double ua[3] = {P->x1[a], P->y1[a], 1}, ub[3] = {P->x1[b], P->y1[b], 1}, ha[3], hb[3], c1[3], c2[3], ec[3], aF[9];
      for (int i = 0; i < 3; ++i) { ha[i] = H[i] * P->x2[a] + H[3 + i] * P->y2[a] + H[6 + i] * 1.0; hb[i] = H[i] * P->x2[b] + H[3 + i] * P->y2[b] + H[6 + i] * 1.0; }
      cross(c1, ua, ha); cross(c2, ub, hb); cross(ec, c1, c2);
      double nr = sqrt(ec[0] * ec[0] + ec[1] * ec[1] + ec[2] * ec[2]);
      ec[0] = ec[0] / nr; ec[1] = ec[1] / nr; ec[2] = ec[2] / nr;
      double S[9] = {0, -ec[2], ec[1], ec[2], 0, -ec[0], -ec[1], ec[0], 0};
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += S[3 * i + k] * H[j * 3 + k]; aF[3 * j + i] = s; }
      unsigned no_i = 0;
      for (int i = 0; i < nN; ++i) { Ds[i] = fres(0, aF, P, uN[i]); if (Ds[i] < th * 2) ++no_i; }
      if (no_i > m_i) {
        no_i = 0;
        for (int i = 0; i < nN; ++i) if (Ds[i] < th * 2) uV[no_i++] = uN[i];
        m_i = no_i;
        double Fn[9];
        innerFH(P, uHl, nH, uV, (int)no_i, th, Fn, inl, st);
        unsigned ninl = 0, both = 0;
        for (int i = 0; i < n; ++i) if (inl[i]) { ++ninl; if (nh[i]) ++both; }
        if (ninl > max_i) { max_i = ninl; memcpy(F, Fn, 72); unsigned ns = (unsigned)nsamples((int)both, nN, 2, 0.999); if (ns < max_sam) max_sam = ns; }
      }
    }
    free(ptr);
  }
  free(uN); free(uHl); free(uV); free(nh); free(inl); free(Ds);
  return max_i;
}

/* ------------------------------------------------------------------ F: LO (exp_ranF.c:621-806) and driver (:1244-1767) */
typedef struct { int metric; double th, sym_th, conf; int do_sym, degen; } FP;
static Sc iterF(const Pts *P, const FP *fp, int *inl, double th, double ths, double *F, double **errs, double *w, int id, HT *ht, Stream *st) {
  double *d = errs[1], f[9], dth = (ths - th) / ILSQ; int n = P->n;
  Sc S = {0, 0, 0, 0}, Ss, maxS = inlidxs(errs[4], n, th, inl);
  if (maxS.I < 8) return S;
  S = inlidxs(errs[4], n, th * MWM_INT, inl);
  if (8 >= S.I) u2f(P, inl, (int)S.I, NULL, f); else u2f(P, randsubset(inl, (int)S.I, 8, st), 8, NULL, f);   /* inlLimit=0 -> 8 (App. A#5) */
  for (int it = 0; it < ILSQ; ++it) {
    fres_w(fp->metric, f, P, d, w);
    S = inlidxs(d, n, th, inl);
    if (ht_check(ht, inl, (int)S.I, id)) { Sc z = {0, 0, 0, 0}; return z; }
    if (maxS.J < S.J) { maxS = S; errs[1] = errs[0]; errs[0] = d; d = errs[1]; memcpy(F, f, 72); }
    Ss = inlidxs(d, n, ths * MWM_INT, inl);
    if (Ss.I < 8) return maxS;
    if (8 >= Ss.I) u2f(P, inl, (int)Ss.I, w, f); else u2f(P, randsubset(inl, (int)Ss.I, 8, st), 8, w, f);
    ths -= dth;
  }
  fres_all(fp->metric, f, P, d);
  S = inlidxs(d, n, th, inl);
  if (maxS.J < S.J) { maxS = S; errs[1] = errs[0]; errs[0] = d; memcpy(F, f, 72); }
  return maxS;
}
static Sc inFrani(const Pts *P, const FP *fp, int *inliers, int ninl, double **errs, double *w, double *F, int *iterID, HT *ht, Stream *st) {
  Sc S, maxS = {0, 0, 0, 0}; int n = P->n;
  if (ninl < 16) return maxS;
  int ssiz = ninl / 2 > 14 ? 14 : ninl / 2, *ib = malloc(n * sizeof(int)), *best = malloc(n * sizeof(int)); double f[9], *t;
  t = errs[2]; errs[2] = errs[0]; errs[0] = t;
  for (int rep = 0; rep < RAN_REP; ++rep) {
    u2f(P, randsubset(inliers, ninl, ssiz, st), ssiz, NULL, f);
    fres_all(fp->metric, f, P, errs[0]); errs[4] = errs[0];
    S = iterF(P, fp, ib, fp->th, TCF * fp->th, f, errs, w, ++*iterID, ht, st);
    if (maxS.J < S.J) { maxS = S; t = errs[2]; errs[2] = errs[0]; errs[0] = t; memcpy(F, f, 72); for (unsigned j = 0; j < maxS.I; ++j) best[j] = ib[j]; }
  }
  t = errs[2]; errs[2] = errs[0]; errs[0] = t;
  for (unsigned j = 0; j < maxS.I; ++j) inliers[j] = best[j];
  free(ib); free(best);
  return maxS;
}
static unsigned symcountF(const double *F, const Pts *P, const int *list, unsigned n, double sym_th) { unsigned c = 0; for (unsigned j = 0; j < n; ++j) if (fres(1, F, P, list[j]) <= sym_th) ++c; return c; }

static void soa(const double *x1y1, const double *x2y2, int n, int dim, double **buf, Pts *P) {
  double *b = malloc(4 * (size_t)n * sizeof(double));
  for (int i = 0; i < n; ++i) { b[i] = x1y1[(size_t)dim * i]; b[n + i] = x1y1[(size_t)dim * i + 1]; b[2 * n + i] = x2y2[(size_t)dim * i]; b[3 * n + i] = x2y2[(size_t)dim * i + 1]; }
  P->n = n; P->x1 = b; P->y1 = b + n; P->x2 = b + 2 * n; P->y2 = b + 3 * n; *buf = b;
}

/* LAF helper correspondences (bindings.cpp:337-389): p1 = x + (a12, a22), p2 = x + (a11, a21) in each image */
static void soa_laf(const double *x1y1, const double *x2y2, int n, double **buf, Pts *L1, Pts *L2) {
  double *b = malloc(8 * (size_t)n * sizeof(double));
  for (int i = 0; i < n; ++i) {
    const double *q1 = x1y1 + 6 * (size_t)i, *q2 = x2y2 + 6 * (size_t)i;
    b[i] = q1[0] + q1[3]; b[n + i] = q1[1] + q1[5]; b[2 * n + i] = q2[0] + q2[3]; b[3 * n + i] = q2[1] + q2[5];
    b[4 * n + i] = q1[0] + q1[2]; b[5 * n + i] = q1[1] + q1[4]; b[6 * n + i] = q2[0] + q2[2]; b[7 * n + i] = q2[1] + q2[4];
  }
  L1->n = n; L1->x1 = b; L1->y1 = b + n; L1->x2 = b + 2 * n; L1->y2 = b + 3 * n;
  L2->n = n; L2->x1 = b + 4 * n; L2->y1 = b + 5 * n; L2->x2 = b + 6 * n; L2->y2 = b + 7 * n; *buf = b;
}
/* F LAF gate count (exp_ranF.c:1394-1412): min(#p2 passing, #p1 passing) with the run's own metric (FDS1idx) */
static unsigned lafcountF(int metric, const double *F, const Pts *L1, const Pts *L2, const int *list, unsigned n, double th_laf) {
  unsigned c1 = 0, c2 = 0;
  for (unsigned j = 0; j < n; ++j) { if (fres(metric, F, L1, list[j]) <= th_laf) ++c1; if (fres(metric, F, L2, list[j]) <= th_laf) ++c2; }
  return c2 < c1 ? c2 : c1;
}

static int g_final_lsq = 0;   /* the reference's compile-time option __FINAL_LSQ__ (exp_ranF.h:28-29, exp_ranH.c:16) as a run-time switch */
void port_set_final_lsq(int on) { g_final_lsq = on; }

int port_find_fundamental(const double *x1y1, const double *x2y2, int n, int dim, double px_th, double conf, int max_iters,
                          int error_type, int sym_check, double laf_coef, int degen, uint64_t seed, double *F_out,
                          unsigned char *mask, int *stats) {
  if (n < 8 || (dim != 2 && dim != 6)) return -1;
  if (laf_coef > 0 && dim != 6) return -1;
  double *pb; Pts P; soa(x1y1, x2y2, n, dim, &pb, &P);
  const int do_laf = laf_coef > 0;                         /* exp_ranF.c:1271-1272 */
  double *lb = NULL; Pts L1, L2; if (do_laf) soa_laf(x1y1, x2y2, n, &lb, &L1, &L2);
  FP fp = {error_type, px_th * px_th, px_th * px_th * (3.0 * (sym_check ? 1 : 0)), conf, 0, degen};   /* bindings.cpp:299-318 */
  fp.do_sym = fp.sym_th > 0;
  double th = fp.th;
  const double th_laf = laf_coef * th;
  double *err = calloc(4 * (size_t)n, sizeof(double)), *errs[5] = {err, err + n, err + 2 * n, err + 3 * n, err + 3 * n};
  double *errorsBest = calloc(n, sizeof(double)), *w = malloc(n * sizeof(double)), *HDsb = malloc(n * sizeof(double));
  int *inliers = malloc(n * sizeof(int)); unsigned char *hmask = calloc(n, 1);
  HT ht = {malloc(64 * 4), malloc(64 * 4), malloc(64 * 4), 0, 64};
  Sc maxS = {8, 0, 0, 0}, maxSs = {8, 0, 0, 0}, S;
  double F[9] = {0}, FBest[9] = {0}, f[9], H[9];
  int samidxBest[7] = {0}, max_sam = max_iters, no_sam = 0, iter_cnt = 0, degen_cnt = 0, iterID = 0, Ihmax = 0;
  unsigned non_degen = 0;
  Stream st = {seed, 0, 1};
  while (no_sam < max_sam) {
    ++no_sam;
    int sel[7], samidx[7], new_max = 0, do_iterate = 0;
    minimal_sample(seed, (uint32_t)no_sam, n, 7, sel);
    for (int t = 0; t < 7; ++t) samidx[t] = sel[6 - t];   /* sample = last 7 pool slots, exp_ranF.c:1302 */
    st.k = (uint32_t)no_sam; st.j = 8;
    double A[81], sol[81], poly[4], roots[3];
    for (int i = 0; i < 7; ++i) frow(&P, sel[i], A + 9 * i);
    for (int i = 63; i < 81; ++i) A[i] = 0;
    if (nullspace9(A, sol) != 2) continue;
    slcm(sol, sol + 9, poly);
    int nsol = rroots3(poly, roots);
    for (int i = 0; i < nsol; ++i) {
      for (int j = 0; j < 9; ++j) f[j] = sol[j] * roots[i] + sol[9 + j] * (1 - roots[i]);
      if (!ori_ok_F(f, &P, samidx, 7)) continue;
      double *d = errs[i];
      fres_all(fp.metric, f, &P, d);
      S = inlidxs(d, n, th, inliers);
      if (maxS.J < S.J) {   /* :1381-1421 */
        if (fp.do_sym) { S.Is = symcountF(f, &P, inliers, S.I, fp.sym_th); if (S.Is < maxS.Is) continue; }
        if (do_laf) { S.Ilafs = lafcountF(fp.metric, f, &L1, &L2, inliers, S.I, th_laf); if (S.Ilafs < maxS.Ilafs) continue; }   /* :1394-1412 */
        errs[i] = errs[3]; errs[3] = d; maxS = S; memcpy(F, f, 72); new_max = 1;
      }
      if (maxSs.J < S.J) {   /* :1425-1492 */
        maxSs = S;
        int deg = 0;
        if (fp.degen) {
          double sb[28]; Pts S7 = {7, sb, sb + 7, sb + 14, sb + 21};
          for (int t = 0; t < 7; ++t) { sb[t] = P.x1[samidx[t]]; sb[7 + t] = P.y1[samidx[t]]; sb[14 + t] = P.x2[samidx[t]]; sb[21 + t] = P.y2[samidx[t]]; }
          deg = checksample(f, &S7, 3 * th, H);
        }
        if (deg) {
          unsigned I = 0;
          for (int j = 0; j < n; ++j) { HDsb[j] = hres_sampson(H, &P, j); if (HDsb[j] < th * 3) ++I; }
          if (I < 8) break;
          I = innerH(H, &P, 16 * th, 10, hmask, &st);
          if ((int)I > Ihmax) Ihmax = (int)I;
          if (I > 6) {
            I = rFtH(&P, hmask, th, H, f, &st);
            if (I > maxS.I) { fres_all(fp.metric, f, &P, errs[3]); maxS.I = I; memcpy(F, f, 72); new_max = 1; d = errs[3]; }
            else { fres_all(fp.metric, f, &P, errs[i]); d = errs[i]; }
            double jj = 0;
            for (int j = 0; j < n; ++j) jj += tquad(d[j], th);
            if (new_max) maxS.J = jj;
            ++degen_cnt;
          }
        } else {
          do_iterate = no_sam > ITER_SAM; errs[4] = d; ++non_degen;
          memcpy(samidxBest, samidx, sizeof samidx); memcpy(errorsBest, d, n * sizeof(double)); memcpy(FBest, f, 72);
        }
      }
    }
    if (no_sam == ITER_SAM && non_degen) do_iterate = 1;   /* :1497-1499 */
    if (do_iterate) {   /* :1501-1577 */
      ++iter_cnt;
      double *d = errs[0];
      S = inlidxs(errs[4], n, TCF * th * MWM_INT, inliers);
      u2f(&P, inliers, (int)S.I, NULL, f);
      fres_all(fp.metric, f, &P, d);
      S = inlidxs(d, n, th, inliers);
      S = inFrani(&P, &fp, inliers, (int)S.I, errs, w, f, &iterID, &ht, &st);
      if (maxS.J < S.J) {
        int upd = 1;
        if (fp.do_sym) { S.Is = symcountF(f, &P, inliers, S.I, fp.sym_th); if (S.Is < maxS.Is) upd = 0; }
        if (do_laf && upd) { S.Ilafs = lafcountF(fp.metric, f, &L1, &L2, inliers, S.I, th_laf); if (S.Ilafs < maxS.Ilafs) upd = 0; }   /* :1536-1555 */
        if (upd) { d = errs[0]; errs[0] = errs[3]; errs[3] = d; maxS = S; memcpy(F, f, 72); new_max = 1; }
      }
      if (new_max) { int ns = nsamples((int)maxS.I + 1, n, 7, conf); if (ns < max_sam) max_sam = ns; }   /* nested in do_iterate: App. A#3 */
    }
  }
  if (st.k != (uint32_t)no_sam) { st.k = (uint32_t)no_sam; st.j = 8; }
  if (!iter_cnt && !degen_cnt && non_degen) {   /* post-loop LO, :1580-1697 */
    int deg = 0;
    if (fp.degen) {
      double sb[28]; Pts S7 = {7, sb, sb + 7, sb + 14, sb + 21};
      for (int t = 0; t < 7; ++t) { sb[t] = P.x1[samidxBest[t]]; sb[7 + t] = P.y1[samidxBest[t]]; sb[14 + t] = P.x2[samidxBest[t]]; sb[21 + t] = P.y2[samidxBest[t]]; }
      deg = checksample(FBest, &S7, 3 * th, H);
    }
    if (deg) {
      unsigned I = 0;
      for (int j = 0; j < n; ++j) if (hres_sampson(H, &P, j) < th * 3) ++I;
      if (I >= 8) I = innerH(H, &P, 16 * th, 10, hmask, &st);
      if ((int)I > Ihmax) Ihmax = (int)I;
      if (I > 6) {
        int nm = 0; double *d;
        memcpy(f, FBest, 72);
        I = rFtH(&P, hmask, th, H, f, &st);
        if (I > maxS.I) { fres_all(fp.metric, f, &P, errs[3]); maxS.I = I; memcpy(F, f, 72); nm = 1; d = errs[3]; }
        else { fres_all(fp.metric, f, &P, errs[0]); d = errs[0]; }   /* reference: errs[i] with a stale loop index */
        double jj = 0; for (int j = 0; j < n; ++j) jj += tquad(d[j], th);
        if (nm) maxS.J = jj;
        ++degen_cnt;
      }
    } else {
      ++iter_cnt;
      double *d = errs[0];
      S = inlidxs(errorsBest, n, TCF * th * MWM_INT, inliers);
      u2f(&P, inliers, (int)S.I, NULL, f);
      fres_all(fp.metric, f, &P, d);
      S = inlidxs(d, n, th, inliers);
      S = inFrani(&P, &fp, inliers, (int)S.I, errs, w, f, &iterID, &ht, &st);
      if (maxS.J < S.J) {
        int upd = 1;
        if (fp.do_sym) { S.Is = symcountF(f, &P, inliers, S.I, fp.sym_th); if (S.Is < maxS.Is) upd = 0; }
        if (do_laf && upd) { S.Ilafs = lafcountF(fp.metric, f, &L1, &L2, inliers, S.I, th_laf); if (S.Ilafs < maxS.Ilafs) upd = 0; }   /* :1664-1683 */
        if (upd) { d = errs[0]; errs[0] = errs[3]; errs[3] = d; maxS = S; memcpy(F, f, 72); }
      }
    }
  }
  { double *d = errs[3];   /* :1699-1723, including the list-position indexing of the symmetric prune (App. A#4) */
    if (g_final_lsq) {   /* #ifdef __FINAL_LSQ__ (:1701-1705).  The reference's text `I = inlidxs(...)` assigns a Score to an
                            unsigned and does not compile (exp_ranF.c:1702); restated with the evident `.I`. */
      S = inlidxs(d, n, th, inliers); u2f(&P, inliers, (int)S.I, NULL, F); fres_all(fp.metric, F, &P, d); }
    for (int j = 0; j < n; ++j) mask[j] = d[j] <= th;
    if (fp.do_sym) { S = inlidxs(d, n, th, inliers); for (unsigned j = 0; j < S.I; ++j) if (fres(1, F, &P, inliers[j]) > fp.sym_th) mask[j] = 0; } }
  double asum = 0; for (int i = 0; i < 9; ++i) { F_out[i] = F[i]; asum += fabs(F[i]); }
  if (asum == 0) memset(mask, 0, n);
  if (stats) { stats[0] = no_sam; stats[1] = iter_cnt; stats[2] = Ihmax; stats[3] = (int)maxS.I; }
  /* (the final LAF prune, :1724-1740, is guarded by a function-scope `do_update` that is never assigned: undefined
      behaviour; in the compiled reference it never runs, and it is not restated) */
  free(pb); free(lb); free(err); free(errorsBest); free(w); free(HDsb); free(inliers); free(hmask); free(ht.h); free(ht.len); free(ht.id);
  return 0;
}

/* ------------------------------------------------------------------ H: LO (exp_ranH.c:291-467) and driver (:470-930) */
typedef struct { int metric; double th, sym_th, conf; int do_sym; int do_laf; double th_laf; const Pts *L1, *L2; int *p1_inliers; } HP;
static Sc iterHc(const Pts *P, const HP *hp, int *inl, double th, double ths, double *H, double **errs, int id, HT *ht) {
  double *d = errs[1], h[9], dth = (ths - th) / ILSQ; int n = P->n;
  Sc S = {0, 0, 0, 0}, Ss, maxS = inlidxs(errs[4], n, th, inl);
  if (maxS.I < 4) return S;
  S = inlidxs(errs[4], n, th * MWM_INT, inl);
  memcpy(h, H, 72);
  u2h(P, inl, (int)S.I, h);   /* inlLimit = 1e6: whole support (exp_ranH.c:505-507) */
  for (int it = 0; it < ILSQ; ++it) {
    hres_all(hp->metric, h, P, d);
    Ss = inlidxs(d, n, th, inl);
    if (ht_check(ht, inl, (int)Ss.I, id)) { Sc z = {0, 0, 0, 0}; return z; }
    S = inlidxs(d, n, ths * MWM_INT, inl);
    if (maxS.J < Ss.J) { maxS = Ss; errs[1] = errs[0]; errs[0] = d; d = errs[1]; memcpy(H, h, 72); }
    if (S.I < 4) return maxS;
    u2h(P, inl, (int)S.I, h);
    ths -= dth;
  }
  hres_all(hp->metric, h, P, d);
  S = inlidxs(d, n, th, inl);
  if (maxS.J < S.J) { maxS = S; errs[1] = errs[0]; errs[0] = d; memcpy(H, h, 72); }
  return maxS;
}
static Sc inHranic(const Pts *P, const HP *hp, int *inliers, int ninl, double **errs, double *H, int *iterID, HT *ht, Stream *st) {
  Sc S, maxS = {0, 0, 0, 0}; int n = P->n;
  if (ninl < 8) return maxS;
  int ssiz = ninl / 2 > 12 ? 12 : ninl / 2, *ib = malloc(n * sizeof(int)); double h[9], *t;
  memcpy(h, H, 72);
  t = errs[2]; errs[2] = errs[0]; errs[0] = t;
  for (int rep = 0; rep < RAN_REP; ++rep) {
    u2h(P, randsubset(inliers, ninl, ssiz, st), ssiz, h);
    hres_all(hp->metric, h, P, errs[0]); errs[4] = errs[0];
    S = iterHc(P, hp, ib, hp->th, TCF * hp->th, h, errs, ++*iterID, ht);
    if (maxS.J < S.J) { maxS = S; t = errs[2]; errs[2] = errs[0]; errs[0] = t; memcpy(H, h, 72); }
  }
  t = errs[2]; errs[2] = errs[0]; errs[0] = t;
  free(ib);
  return maxS;
}
static int lo_step_H(const Pts *P, const HP *hp, double **errs, int *inliers, int *inliersS, double *h, double *Hbest, Sc *maxS, int *iterID, HT *ht, Stream *st) {
  int n = P->n, new_max = 0; double *d = errs[0];   /* iter_type 4: exp_ranH.c:678-747 */
  Sc S = inlidxs(errs[4], n, TCF * hp->th * MWM_INT, inliers);
  u2h(P, inliers, (int)S.I, h);
  hres_all(hp->metric, h, P, d);
  S = inlidxs(d, n, hp->th, inliers);
  S = inHranic(P, hp, inliers, (int)S.I, errs, h, iterID, ht, st);
  if (maxS->J < S.J && !hsingular(h)) {
    int upd = 1;
    if (hp->do_sym) {   /* re-lists row `d` (the pre-LO pointer), :708-716 */
      Sc Sc2 = inlidxs(d, n, hp->th, inliersS); HS s; hsym(h, &s); S.Is = 0;
      for (unsigned j = 0; j < Sc2.I; ++j) if (hgate(&s, P, inliersS[j]) <= hp->sym_th) ++S.Is;
      if (S.Is < maxS->Is) upd = 0;
    }
    if (upd && hp->do_laf) {   /* :718-736 / :834-849; `p1_inliers` is never reset in the reference: it accumulates */
      Sc Sc2 = inlidxs(d, n, hp->th, inliersS);
      *hp->p1_inliers += lafcountH(hp->metric, h, P, hp->L1, inliersS, Sc2.I, hp->th_laf);
      unsigned c2 = (unsigned)lafcountH(hp->metric, h, P, hp->L2, inliersS, Sc2.I, hp->th_laf);
      S.Ilafs = c2 < (unsigned)*hp->p1_inliers ? c2 : (unsigned)*hp->p1_inliers;
      if (S.Ilafs < maxS->Ilafs) upd = 0;
    }
    if (upd) { double *t = errs[0]; errs[0] = errs[3]; errs[3] = t; *maxS = S; memcpy(Hbest, h, 72); new_max = 1; }
  }
  return new_max;
}
int port_find_homography(const double *x1y1, const double *x2y2, int n, int dim, double px_th, double conf, int max_iters,
                         int error_type, int sym_check, double laf_coef, uint64_t seed, double *H_out, unsigned char *mask,
                         int *stats) {
  if (n < 4 || (dim != 2 && dim != 6)) return -1;
  if (error_type < 0 || error_type > 4) return -2;
  if (laf_coef > 0 && dim != 6) return -1;
  double *pb; Pts P; soa(x1y1, x2y2, n, dim, &pb, &P);
  double *lb = NULL; Pts L1, L2; int p1_inliers = 0;
  if (laf_coef > 0) soa_laf(x1y1, x2y2, n, &lb, &L1, &L2);
  double coef = 3.0 * (sym_check ? 1 : 0);   /* bindings.cpp:64-107 */
  HP hp = {error_type, 0, 0, conf, 0, laf_coef > 0, 0, &L1, &L2, &p1_inliers};
  switch (error_type) { case 0: hp.th = px_th * px_th; hp.sym_th = px_th * coef; break; case 1: hp.th = px_th * px_th; break; case 2: hp.th = px_th; break;
                        case 3: hp.th = px_th * px_th; hp.sym_th = px_th * coef; break; default: hp.th = px_th; hp.sym_th = px_th * coef; }
  hp.do_sym = hp.sym_th > 0;
  hp.th_laf = laf_coef * hp.th;   /* exp_ranH.c:500 */
  double th = hp.th, *err = calloc(4 * (size_t)n, sizeof(double)), *errs[5] = {err, err + n, err + 2 * n, err + 3 * n, err + 3 * n};
  int *inliers = malloc(n * sizeof(int)), *inliersS = malloc(n * sizeof(int));
  HT ht = {malloc(64 * 4), malloc(64 * 4), malloc(64 * 4), 0, 64};
  Sc maxS = {0, 0, 0, 0}, maxSs = {0, 0, 0, 0}, S;
  double H[9] = {0}, h[9] = {0};
  int max_sam = max_iters, no_sam = 0, iter_cnt = 0, iterID = 0, no_rej = 0;
  Stream st = {seed, 0, 1};
  while (no_sam < max_sam) {
    ++no_sam;
    int sel[4], samidx[4], new_max = 0, do_iterate;
    minimal_sample(seed, (uint32_t)no_sam, n, 4, sel);
    for (int t = 0; t < 4; ++t) samidx[t] = sel[3 - t];
    st.k = (uint32_t)no_sam; st.j = 5;
    if (!ori_ok_H(&P, samidx)) { ++no_rej; continue; }
    double M[81], sol[81];
    for (int i = 0; i < 4; ++i) hrows(&P, sel[i], M + 18 * i, M + 18 * i + 9);
    for (int i = 72; i < 81; ++i) M[i] = 0;
    if (nullspace9(M, sol) != 1) { ++no_rej; continue; }
    memcpy(h, sol, 72);
    if (hsingular(h)) { ++no_rej; continue; }
    double *d = errs[0];
    hres_all(hp.metric, h, &P, d);
    S = inlidxs(d, n, th, inliersS);
    if (maxS.J < S.J) {   /* :585-627 */
      if (hp.do_sym) { HS s; hsym(h, &s); S.Is = 0; for (unsigned j = 0; j < S.I; ++j) if (hgate(&s, &P, inliersS[j]) <= hp.sym_th) ++S.Is; if (S.Is < maxS.Is) continue; }
      if (hp.do_laf) {   /* :600-619 */
        p1_inliers += lafcountH(hp.metric, h, &P, &L1, inliersS, S.I, hp.th_laf);
        if ((unsigned)p1_inliers < maxS.Ilafs) continue;
        unsigned c2 = (unsigned)lafcountH(hp.metric, h, &P, &L2, inliersS, S.I, hp.th_laf);
        S.Ilafs = c2 < (unsigned)p1_inliers ? c2 : (unsigned)p1_inliers;
        if (S.Ilafs < maxS.Ilafs) continue;
      }
      errs[0] = errs[3]; errs[3] = d; maxS = S; new_max = 1; memcpy(H, h, 72);
    }
    if (maxSs.J < S.J) { do_iterate = no_sam > ITER_SAM; maxSs = S; errs[4] = d; } else do_iterate = 0;
    if (no_sam >= ITER_SAM && iter_cnt == 0 && maxSs.I > 4) do_iterate = 1;   /* :639-640 */
    if (do_iterate) { ++iter_cnt; if (lo_step_H(&P, &hp, errs, inliers, inliersS, h, H, &maxS, &iterID, &ht, &st)) new_max = 1; }
    if (new_max) { int ns = nsamples((int)maxS.I + 1, n, 4, conf); if (ns < max_sam) max_sam = ns; }
  }
  if (st.k != (uint32_t)no_sam) { st.k = (uint32_t)no_sam; st.j = 5; }
  if (iter_cnt == 0) { ++iter_cnt; memcpy(h, H, 72); lo_step_H(&P, &hp, errs, inliers, inliersS, h, H, &maxS, &iterID, &ht, &st); }   /* :759-862 */
  { double *d = errs[3];
    if (g_final_lsq) { Sc Sl = inlidxs(d, n, th, inliers); u2h(&P, inliers, (int)Sl.I, H); hres_all(hp.metric, H, &P, d); }   /* #ifdef __FINAL_LSQ__, exp_ranH.c:866-870 */
    for (int j = 0; j < n; ++j) mask[j] = d[j] <= th;
    if (hp.do_sym) { Sc Sc2 = inlidxs(d, n, th, inliersS); HS s; hsym(H, &s); for (unsigned j = 0; j < Sc2.I; ++j) if (hgate(&s, &P, inliersS[j]) > hp.sym_th) mask[inliersS[j]] = 0; }
    if (hp.do_laf) {   /* final LAF prune, :889-907 (HDSidx1 on both helper correspondences) */
      Sc Sc2 = inlidxs(d, n, th, inliersS); HS s; if (hp.metric) hsym(H, &s);
      for (unsigned j = 0; j < Sc2.I; ++j) { int i = inliersS[j];
        if (hres_laf(hp.metric, H, &s, &P, &L1, i) > hp.th_laf || hres_laf(hp.metric, H, &s, &P, &L2, i) > hp.th_laf) mask[i] = 0; } } }
  double asum = 0; for (int i = 0; i < 9; ++i) { H_out[i] = H[i]; asum += fabs(H[i]); }
  if (asum == 0) memset(mask, 0, n);
  if (stats) { stats[0] = no_sam; stats[1] = iter_cnt; stats[2] = no_rej; stats[3] = (int)maxS.I; }
  free(pb); free(lb); free(err); free(inliers); free(inliersS); free(ht.h); free(ht.len); free(ht.id);
  return 0;
}
