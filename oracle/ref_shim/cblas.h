/* TEST INFRASTRUCTURE (oracle/): minimal CBLAS front for building the reference's
 * CPU path (src/math_functions_cpu.cpp:44,63,99,105 via src/mkl_alternate.hpp:35).
 * No BLAS headers exist in this image; the Fortran BLAS symbols sgemm_/dgemm_ and
 * cblas_{s,d}axpy are exported by torch's libtorch_cpu.so (bundled MKL, LP64), so
 * the two gemm entry points are forwarded there. Written for this repo; nothing
 * here comes from the reference tree. */
#ifndef MEB200_ORACLE_CBLAS_SHIM_H
#define MEB200_ORACLE_CBLAS_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif

typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE {
  CblasNoTrans = 111,
  CblasTrans = 112,
  CblasConjTrans = 113
} CBLAS_TRANSPOSE;

void sgemm_(const char *ta, const char *tb, const int *m, const int *n, const int *k,
            const float *alpha, const float *a, const int *lda, const float *b,
            const int *ldb, const float *beta, float *c, const int *ldc);
void dgemm_(const char *ta, const char *tb, const int *m, const int *n, const int *k,
            const double *alpha, const double *a, const int *lda, const double *b,
            const int *ldb, const double *beta, double *c, const int *ldc);
void cblas_saxpy(const int n, const float alpha, const float *x, const int incx,
                 float *y, const int incy);
void cblas_daxpy(const int n, const double alpha, const double *x, const int incx,
                 double *y, const int incy);

/* Row-major C = op(A) op(B) is column-major C^T = op(B)^T op(A)^T: swap operands
 * and the M/N extents, keep the transpose flags attached to their operands. */
#define MEB200_DEFINE_GEMM(NAME, T, F77)                                               \
  static inline void NAME(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb,   \
                          int M, int N, int K, T alpha, const T *A, int lda,           \
                          const T *B, int ldb, T beta, T *C, int ldc) {                \
    char ca = (ta == CblasNoTrans) ? 'N' : 'T';                                        \
    char cb = (tb == CblasNoTrans) ? 'N' : 'T';                                        \
    if (order == CblasColMajor) {                                                      \
      F77(&ca, &cb, &M, &N, &K, &alpha, A, &lda, B, &ldb, &beta, C, &ldc);             \
    } else {                                                                           \
      F77(&cb, &ca, &N, &M, &K, &alpha, B, &ldb, A, &lda, &beta, C, &ldc);             \
    }                                                                                  \
  }
MEB200_DEFINE_GEMM(cblas_sgemm, float, sgemm_)
MEB200_DEFINE_GEMM(cblas_dgemm, double, dgemm_)
#undef MEB200_DEFINE_GEMM

#ifdef __cplusplus
}
#endif
#endif
