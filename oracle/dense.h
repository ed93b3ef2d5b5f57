/* oracle/dense.h -- internal prototypes of the dense kernels in dense.c (TEST INFRASTRUCTURE ONLY). */
#ifndef HP3D_ORACLE_DENSE_H
#define HP3D_ORACLE_DENSE_H
#include "hp3d_oracle.h"
void orc_dgemm(char ta, char tb, int m, int n, int k, double alpha, const double *A, int lda, const double *B,
               int ldb, double beta, double *C, int ldc);
void orc_zgemm(char ta, char tb, int m, int n, int k, zdouble alpha, const zdouble *A, int lda, const zdouble *B,
               int ldb, zdouble beta, zdouble *C, int ldc);
void orc_dsyrk_u(char trans, int n, int k, double alpha, const double *A, int lda, double beta, double *C, int ldc);
void orc_zsyrk_u(char trans, int n, int k, zdouble alpha, const zdouble *A, int lda, zdouble beta, zdouble *C, int ldc);
void orc_zherk_u(char trans, int n, int k, double alpha, const zdouble *A, int lda, double beta, zdouble *C, int ldc);
int orc_dpotrf_u(int n, double *A, int lda);
int orc_zpotrf_u(int n, zdouble *A, int lda);
void orc_dtrsm_u(char trans, int n, int nrhs, const double *U, int ldu, double *X, int ldx);
void orc_ztrsm_u(char trans, int n, int nrhs, const zdouble *U, int ldu, zdouble *X, int ldx);
int orc_dgetrf(int n, double *A, int lda, int *ipiv);
int orc_zgetrf(int n, zdouble *A, int lda, int *ipiv);
void orc_dgetrs(int n, int nrhs, const double *A, int lda, const int *ipiv, double *B, int ldb);
void orc_zgetrs(int n, int nrhs, const zdouble *A, int lda, const int *ipiv, zdouble *B, int ldb);
#endif
