/*
 * oracle/dense.c -- the BLAS/LAPACK subset hp3D's element path calls (TEST INFRASTRUCTURE ONLY).
 *
 * BLAS/LAPACK are third-party to the reference (not vendored under /root/reference; resolved at link
 * time from PETSc's fblaslapack or MKL, m_options_files/m_options_linux:78; no version pin).  They are
 * standard dense kernels whose results are defined up to rounding.  Call sites restated here:
 *   DSYRK  POISSON/GALERKIN/elem_opt.F90:131        ZSYRK  MAXWELL/GALERKIN/elem_opt.F90:146,148
 *   DGEMM/DSYRK/ZHERK/ZPOTRF/ZTRTRS  MAXWELL/ULTRAWEAK_DPG/elem/elem_opt.F90:341-465,647,841,852,862
 *   DSFRK/DGEMM/DPFTRF/DTFSM/DSYRK   POISSON/PRIMAL_DPG/elem_opt.F90:260-269,383,417,424,430
 *   ?TRTTF/?PFTRF/?PFTRS/?GEMM/?GETRF/?GETRS  src/modules/stc.F90:356-413,460-506
 * (RFP routines ?PFTRF/?PFTRS/?TFSM/?SFRK are storage variants of POTRF/POTRS/TRSM/SYRK.)
 *
 * Two back-ends: textbook loops (always available, used to cross-check) and, when
 * orc_dense_use_blas() can dlopen an OpenBLAS (the one bundled with scipy in this image), its routines,
 * so that the CPU baseline is timed against an optimised BLAS like the reference would be linked to.
 */
#include "hp3d_oracle.h"
#include "dense.h"
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void (*gemm_f)(const char *, const char *, const int *, const int *, const int *, const void *,
                       const void *, const int *, const void *, const int *, const void *, void *, const int *,
                       size_t, size_t);
typedef void (*syrk_f)(const char *, const char *, const int *, const int *, const void *, const void *,
                       const int *, const void *, void *, const int *, size_t, size_t);
typedef void (*trsm_f)(const char *, const char *, const char *, const char *, const int *, const int *,
                       const void *, const void *, const int *, void *, const int *, size_t, size_t, size_t, size_t);
typedef void (*potrf_f)(const char *, const int *, void *, const int *, int *, size_t);
typedef void (*getrf_f)(const int *, const int *, void *, const int *, int *, int *);
typedef void (*getrs_f)(const char *, const int *, const int *, const void *, const int *, const int *, void *,
                        const int *, int *, size_t);
typedef void (*setthr_f)(int);

static struct {
  void *h;
  gemm_f dgemm, zgemm;
  syrk_f dsyrk, zsyrk, zherk;
  trsm_f dtrsm, ztrsm;
  potrf_f dpotrf, zpotrf;
  getrf_f dgetrf, zgetrf;
  getrs_f dgetrs, zgetrs;
  setthr_f setthr;
} B;

static void *sym2(void *h, const char *a, const char *b) { void *p = dlsym(h, a); return p ? p : dlsym(h, b); }

int orc_dense_use_blas(const char *libpath) {
  if (!libpath || !*libpath) { memset(&B, 0, sizeof B); return 0; }
  void *h = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 0;
  B.h = h;
#define LD(field, name) B.field = (__typeof__(B.field))sym2(h, "scipy_" name, name)
  LD(dgemm, "dgemm_"); LD(zgemm, "zgemm_"); LD(dsyrk, "dsyrk_"); LD(zsyrk, "zsyrk_"); LD(zherk, "zherk_");
  LD(dtrsm, "dtrsm_"); LD(ztrsm, "ztrsm_"); LD(dpotrf, "dpotrf_"); LD(zpotrf, "zpotrf_");
  LD(dgetrf, "dgetrf_"); LD(zgetrf, "zgetrf_"); LD(dgetrs, "dgetrs_"); LD(zgetrs, "zgetrs_");
  B.setthr = (setthr_f)sym2(h, "scipy_openblas_set_num_threads", "openblas_set_num_threads");
  if (!(B.dgemm && B.zgemm && B.dsyrk && B.zsyrk && B.zherk && B.dtrsm && B.ztrsm && B.dpotrf && B.zpotrf &&
        B.dgetrf && B.zgetrf && B.dgetrs && B.zgetrs)) { memset(&B, 0, sizeof B); return 0; }
  return 1;
}
void orc_dense_set_threads(int n) { if (B.setthr) B.setthr(n); }

/* ---------------------------------------------------------------- textbook loops (column-major) */
#define IDX(A, ld, i, j) (A)[(size_t)(i) + (size_t)(ld) * (size_t)(j)]

#define DEF_GEMM(NAME, T, CONJ)                                                                            \
  static void NAME(char ta, char tb, int m, int n, int k, T alpha, const T *A, int lda, const T *Bm,       \
                   int ldb, T beta, T *C, int ldc) {                                                       \
    for (int j = 0; j < n; j++) {                                                                          \
      for (int i = 0; i < m; i++) IDX(C, ldc, i, j) = (beta == 0) ? 0 : beta * IDX(C, ldc, i, j);          \
      for (int l = 0; l < k; l++) {                                                                        \
        T b = (tb == 'N') ? IDX(Bm, ldb, l, j) : (tb == 'T' ? IDX(Bm, ldb, j, l) : CONJ(IDX(Bm, ldb, j, l))); \
        T ab = alpha * b;                                                                                  \
        if (ta == 'N') for (int i = 0; i < m; i++) IDX(C, ldc, i, j) += ab * IDX(A, lda, i, l);            \
        else if (ta == 'T') for (int i = 0; i < m; i++) IDX(C, ldc, i, j) += ab * IDX(A, lda, l, i);       \
        else for (int i = 0; i < m; i++) IDX(C, ldc, i, j) += ab * CONJ(IDX(A, lda, l, i));                \
      }                                                                                                    \
    }                                                                                                      \
  }
#define NOCONJ(x) (x)
DEF_GEMM(ref_dgemm, double, NOCONJ)
DEF_GEMM(ref_zgemm, zdouble, conj)

/* upper Cholesky A = U^H U (in place, upper triangle), LAPACK ?POTRF('U') semantics incl. info */
#define DEF_POTRF(NAME, T, CONJ, REAL)                                                  \
  static int NAME(int n, T *A, int lda) {                                               \
    for (int j = 0; j < n; j++) {                                                       \
      double d = REAL(IDX(A, lda, j, j));                                               \
      for (int k = 0; k < j; k++) d -= REAL(CONJ(IDX(A, lda, k, j)) * IDX(A, lda, k, j)); \
      if (!(d > 0.0)) return j + 1;                                                     \
      d = sqrt(d);                                                                      \
      IDX(A, lda, j, j) = d;                                                            \
      for (int i = j + 1; i < n; i++) {                                                 \
        T s = IDX(A, lda, j, i);                                                        \
        for (int k = 0; k < j; k++) s -= CONJ(IDX(A, lda, k, j)) * IDX(A, lda, k, i);   \
        IDX(A, lda, j, i) = s / d;                                                      \
      }                                                                                 \
    }                                                                                   \
    return 0;                                                                           \
  }
#define REALD(x) (x)
DEF_POTRF(ref_dpotrf, double, NOCONJ, REALD)
DEF_POTRF(ref_zpotrf, zdouble, conj, creal)

/* solve U^H X = B (trans='C') or U X = B (trans='N'), U upper, left side */
#define DEF_TRSM_U(NAME, T, CONJ)                                                          \
  static void NAME(char trans, int n, int nrhs, const T *U, int ldu, T *X, int ldx) {      \
    for (int c = 0; c < nrhs; c++) {                                                       \
      T *x = X + (size_t)ldx * c;                                                          \
      if (trans == 'N') {                                                                  \
        for (int i = n - 1; i >= 0; i--) {                                                 \
          x[i] = x[i] / IDX(U, ldu, i, i);                                                 \
          T xi = x[i];                                                                     \
          for (int k = 0; k < i; k++) x[k] -= IDX(U, ldu, k, i) * xi;                      \
        }                                                                                  \
      } else {                                                                             \
        for (int i = 0; i < n; i++) {                                                      \
          T s = x[i];                                                                      \
          for (int k = 0; k < i; k++) s -= CONJ(IDX(U, ldu, k, i)) * x[k];                 \
          x[i] = s / CONJ(IDX(U, ldu, i, i));                                              \
        }                                                                                  \
      }                                                                                    \
    }                                                                                      \
  }
DEF_TRSM_U(ref_dtrsm_u, double, NOCONJ)
DEF_TRSM_U(ref_ztrsm_u, zdouble, conj)

/* LU with partial pivoting (?GETRF) and solve (?GETRS 'N'); ipiv 1-based like LAPACK */
#define DEF_GETRF(NAME, T, ABS)                                                   \
  static int NAME(int n, T *A, int lda, int *ipiv) {                              \
    int info = 0;                                                                 \
    for (int j = 0; j < n; j++) {                                                 \
      int p = j; double best = ABS(IDX(A, lda, j, j));                            \
      for (int i = j + 1; i < n; i++) { double v = ABS(IDX(A, lda, i, j)); if (v > best) { best = v; p = i; } } \
      ipiv[j] = p + 1;                                                            \
      if (best == 0.0) { if (!info) info = j + 1; continue; }                     \
      if (p != j) for (int c = 0; c < n; c++) { T t = IDX(A, lda, j, c); IDX(A, lda, j, c) = IDX(A, lda, p, c); IDX(A, lda, p, c) = t; } \
      T piv = IDX(A, lda, j, j);                                                  \
      for (int i = j + 1; i < n; i++) IDX(A, lda, i, j) = IDX(A, lda, i, j) / piv; \
      for (int c = j + 1; c < n; c++) {                                           \
        T u = IDX(A, lda, j, c);                                                  \
        for (int i = j + 1; i < n; i++) IDX(A, lda, i, c) -= IDX(A, lda, i, j) * u; \
      }                                                                           \
    }                                                                             \
    return info;                                                                  \
  }
static double cabs1(zdouble z) { return fabs(creal(z)) + fabs(cimag(z)); } /* IZAMAX uses |re|+|im| */
DEF_GETRF(ref_dgetrf, double, fabs)
DEF_GETRF(ref_zgetrf, zdouble, cabs1)
#define DEF_GETRS(NAME, T)                                                         \
  static void NAME(int n, int nrhs, const T *A, int lda, const int *ipiv, T *Bm, int ldb) { \
    for (int c = 0; c < nrhs; c++) {                                               \
      T *x = Bm + (size_t)ldb * c;                                                 \
      for (int j = 0; j < n; j++) { int p = ipiv[j] - 1; if (p != j) { T t = x[j]; x[j] = x[p]; x[p] = t; } } \
      for (int j = 0; j < n; j++) { T xj = x[j]; for (int i = j + 1; i < n; i++) x[i] -= IDX(A, lda, i, j) * xj; } \
      for (int j = n - 1; j >= 0; j--) { x[j] = x[j] / IDX(A, lda, j, j); T xj = x[j]; for (int i = 0; i < j; i++) x[i] -= IDX(A, lda, i, j) * xj; } \
    }                                                                              \
  }
DEF_GETRS(ref_dgetrs, double)
DEF_GETRS(ref_zgetrs, zdouble)

/* ---------------------------------------------------------------- public wrappers */
void orc_dgemm(char ta, char tb, int m, int n, int k, double alpha, const double *A, int lda, const double *Bm,
               int ldb, double beta, double *C, int ldc) {
  if (m <= 0 || n <= 0) return;
  if (B.dgemm && k > 0) B.dgemm(&ta, &tb, &m, &n, &k, &alpha, A, &lda, Bm, &ldb, &beta, C, &ldc, 1, 1);
  else ref_dgemm(ta, tb, m, n, k, alpha, A, lda, Bm, ldb, beta, C, ldc);
}
void orc_zgemm(char ta, char tb, int m, int n, int k, zdouble alpha, const zdouble *A, int lda, const zdouble *Bm,
               int ldb, zdouble beta, zdouble *C, int ldc) {
  if (m <= 0 || n <= 0) return;
  if (B.zgemm && k > 0) B.zgemm(&ta, &tb, &m, &n, &k, &alpha, A, &lda, Bm, &ldb, &beta, C, &ldc, 1, 1);
  else ref_zgemm(ta, tb, m, n, k, alpha, A, lda, Bm, ldb, beta, C, ldc);
}
/* C(upper) = alpha*op(A)*op(A)^T + beta*C ; trans='N': A is n x k, 'T': A is k x n.  Only the upper
 * triangle of C is referenced/updated (like ?SYRK 'U'). */
void orc_dsyrk_u(char trans, int n, int k, double alpha, const double *A, int lda, double beta, double *C, int ldc) {
  if (n <= 0) return;
  if (B.dsyrk) { char u = 'U'; B.dsyrk(&u, &trans, &n, &k, &alpha, A, &lda, &beta, C, &ldc, 1, 1); return; }
  for (int j = 0; j < n; j++)
    for (int i = 0; i <= j; i++) {
      double s = 0;
      if (trans == 'N') for (int l = 0; l < k; l++) s += IDX(A, lda, i, l) * IDX(A, lda, j, l);
      else for (int l = 0; l < k; l++) s += IDX(A, lda, l, i) * IDX(A, lda, l, j);
      IDX(C, ldc, i, j) = alpha * s + (beta == 0 ? 0 : beta * IDX(C, ldc, i, j));
    }
}
void orc_zsyrk_u(char trans, int n, int k, zdouble alpha, const zdouble *A, int lda, zdouble beta, zdouble *C, int ldc) {
  if (n <= 0) return;
  if (B.zsyrk) { char u = 'U'; B.zsyrk(&u, &trans, &n, &k, &alpha, A, &lda, &beta, C, &ldc, 1, 1); return; }
  for (int j = 0; j < n; j++)
    for (int i = 0; i <= j; i++) {
      zdouble s = 0;
      if (trans == 'N') for (int l = 0; l < k; l++) s += IDX(A, lda, i, l) * IDX(A, lda, j, l);
      else for (int l = 0; l < k; l++) s += IDX(A, lda, l, i) * IDX(A, lda, l, j);
      IDX(C, ldc, i, j) = alpha * s + (beta == 0 ? 0 : beta * IDX(C, ldc, i, j));
    }
}
/* ZHERK 'U': trans='N': C = alpha*A*A^H + beta*C ; trans='C': C = alpha*A^H*A + beta*C (alpha,beta real) */
void orc_zherk_u(char trans, int n, int k, double alpha, const zdouble *A, int lda, double beta, zdouble *C, int ldc) {
  if (n <= 0) return;
  if (B.zherk) { char u = 'U'; B.zherk(&u, &trans, &n, &k, &alpha, A, &lda, &beta, C, &ldc, 1, 1); return; }
  for (int j = 0; j < n; j++)
    for (int i = 0; i <= j; i++) {
      zdouble s = 0;
      if (trans == 'N') for (int l = 0; l < k; l++) s += IDX(A, lda, i, l) * conj(IDX(A, lda, j, l));
      else for (int l = 0; l < k; l++) s += conj(IDX(A, lda, l, i)) * IDX(A, lda, l, j);
      zdouble c = alpha * s + (beta == 0 ? 0 : beta * IDX(C, ldc, i, j));
      IDX(C, ldc, i, j) = (i == j) ? creal(c) : c;
    }
}
int orc_dpotrf_u(int n, double *A, int lda) {
  if (B.dpotrf) { char u = 'U'; int info; B.dpotrf(&u, &n, A, &lda, &info, 1); return info; }
  return ref_dpotrf(n, A, lda);
}
int orc_zpotrf_u(int n, zdouble *A, int lda) {
  if (B.zpotrf) { char u = 'U'; int info; B.zpotrf(&u, &n, A, &lda, &info, 1); return info; }
  return ref_zpotrf(n, A, lda);
}
void orc_dtrsm_u(char trans, int n, int nrhs, const double *U, int ldu, double *X, int ldx) {
  if (n <= 0 || nrhs <= 0) return;
  if (B.dtrsm) { char s = 'L', u = 'U', d = 'N', t = (trans == 'N') ? 'N' : 'T'; double one = 1.0;
    B.dtrsm(&s, &u, &t, &d, &n, &nrhs, &one, U, &ldu, X, &ldx, 1, 1, 1, 1); return; }
  ref_dtrsm_u(trans, n, nrhs, U, ldu, X, ldx);
}
void orc_ztrsm_u(char trans, int n, int nrhs, const zdouble *U, int ldu, zdouble *X, int ldx) {
  if (n <= 0 || nrhs <= 0) return;
  if (B.ztrsm) { char s = 'L', u = 'U', d = 'N', t = (trans == 'N') ? 'N' : 'C'; zdouble one = 1.0;
    B.ztrsm(&s, &u, &t, &d, &n, &nrhs, &one, U, &ldu, X, &ldx, 1, 1, 1, 1); return; }
  ref_ztrsm_u(trans, n, nrhs, U, ldu, X, ldx);
}
int orc_dgetrf(int n, double *A, int lda, int *ipiv) {
  if (B.dgetrf) { int info; B.dgetrf(&n, &n, A, &lda, ipiv, &info); return info; }
  return ref_dgetrf(n, A, lda, ipiv);
}
int orc_zgetrf(int n, zdouble *A, int lda, int *ipiv) {
  if (B.zgetrf) { int info; B.zgetrf(&n, &n, A, &lda, ipiv, &info); return info; }
  return ref_zgetrf(n, A, lda, ipiv);
}
void orc_dgetrs(int n, int nrhs, const double *A, int lda, const int *ipiv, double *Bm, int ldb) {
  if (n <= 0 || nrhs <= 0) return;
  if (B.dgetrs) { char t = 'N'; int info; B.dgetrs(&t, &n, &nrhs, A, &lda, ipiv, Bm, &ldb, &info, 1); return; }
  ref_dgetrs(n, nrhs, A, lda, ipiv, Bm, ldb);
}
void orc_zgetrs(int n, int nrhs, const zdouble *A, int lda, const int *ipiv, zdouble *Bm, int ldb) {
  if (n <= 0 || nrhs <= 0) return;
  if (B.zgetrs) { char t = 'N'; int info; B.zgetrs(&t, &n, &nrhs, A, &lda, ipiv, Bm, &ldb, &info, 1); return; }
  ref_zgetrs(n, nrhs, A, lda, ipiv, Bm, ldb);
}
