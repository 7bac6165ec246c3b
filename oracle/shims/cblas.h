/* cblas prototypes for the four gemm flavours the PanguLU reference CPU
 * platform calls (reference src/platforms/01_SHAREDMEM/00_CPU/000_CPU/
 * pangulu_platform_0100000.c:317-327).  Test infrastructure only.
 * Implementation: cblas_gemm.c (dlopens a real OpenBLAS when one is found,
 * otherwise a plain blocked C gemm). */
#ifndef PANGULU_ORACLE_CBLAS_SHIM_H
#define PANGULU_ORACLE_CBLAS_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
/* src/pangulu.c:117-119 pins BLAS to one thread when this macro exists */
#define OPENBLAS_VERSION "shim"
void openblas_set_num_threads(int n);
void cblas_dgemm(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                 double alpha, const double *a, int lda, const double *b, int ldb, double beta, double *c, int ldc);
void cblas_sgemm(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                 float alpha, const float *a, int lda, const float *b, int ldb, float beta, float *c, int ldc);
void cblas_zgemm(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                 const void *alpha, const void *a, int lda, const void *b, int ldb, const void *beta, void *c, int ldc);
void cblas_cgemm(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                 const void *alpha, const void *a, int lda, const void *b, int ldb, const void *beta, void *c, int ldc);
/* reports which backend served the last gemm: "openblas:<path>" or "builtin" */
const char *pangulu_oracle_blas_backend(void);
#ifdef __cplusplus
}
#endif
#endif
