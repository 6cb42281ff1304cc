/* oracle/shim/lapacke.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal LAPACKE declaration shim (see cblas.h in this directory): the three
 * drivers libcd's chomp.c calls, mapped onto scipy's bundled LP64 OpenBLAS. */
#ifndef ORACLE_SHIM_LAPACKE_H
#define ORACLE_SHIM_LAPACKE_H
#ifdef __cplusplus
extern "C" {
#endif
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
#define LAPACKE_dgetrf scipy_LAPACKE_dgetrf
#define LAPACKE_dgetri scipy_LAPACKE_dgetri
#define LAPACKE_dgesv  scipy_LAPACKE_dgesv
int LAPACKE_dgetrf(int layout, int m, int n, double *a, int lda, int *ipiv);
int LAPACKE_dgetri(int layout, int n, double *a, int lda, const int *ipiv);
int LAPACKE_dgesv(int layout, int n, int nrhs, double *a, int lda, int *ipiv, double *b, int ldb);
#ifdef __cplusplus
}
#endif
#endif
