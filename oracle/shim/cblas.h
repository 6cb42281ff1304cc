/* oracle/shim/cblas.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal CBLAS declaration shim so the reference's libcd sources
 * (/root/reference/src/libcd/{chomp,kin,spatial}.c, which #include <cblas.h>)
 * compile unmodified against the LP64 OpenBLAS bundled with scipy
 * (symbols are exported with a scipy_ prefix).  Only the six routines libcd
 * calls are declared. */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
#define cblas_dgemm scipy_cblas_dgemm
#define cblas_dgemv scipy_cblas_dgemv
#define cblas_daxpy scipy_cblas_daxpy
#define cblas_ddot  scipy_cblas_ddot
#define cblas_dnrm2 scipy_cblas_dnrm2
#define cblas_dscal scipy_cblas_dscal
void cblas_dgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                 int m, int n, int k, double alpha, const double *a, int lda,
                 const double *b, int ldb, double beta, double *c, int ldc);
void cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, int m, int n,
                 double alpha, const double *a, int lda, const double *x, int incx,
                 double beta, double *y, int incy);
void cblas_daxpy(int n, double alpha, const double *x, int incx, double *y, int incy);
double cblas_ddot(int n, const double *x, int incx, const double *y, int incy);
double cblas_dnrm2(int n, const double *x, int incx);
void cblas_dscal(int n, double alpha, double *x, int incx);
#ifdef __cplusplus
}
#endif
#endif
