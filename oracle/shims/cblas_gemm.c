/* cblas_?gemm for the PanguLU reference CPU platform (test infrastructure only).
 *
 * The reference links OpenBLAS (README.md:32); this container has no system
 * BLAS.  At first use we look for a real OpenBLAS shared object
 * (PANGULU_ORACLE_OPENBLAS, then the copy bundled in a Python wheel of this
 * image) and forward to it; if none loads, a plain cache-blocked C gemm is used.
 * Only ColMajor / NoTrans / NoTrans is ever requested
 * (pangulu_platform_0100000.c:317-327).
 */
#define _GNU_SOURCE
#include "cblas.h"
#include <dlfcn.h>
#include <glob.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>

typedef void (*dgemm_fn)(int, int, int, int, int, int, double, const double *, int, const double *, int, double, double *, int);
typedef void (*sgemm_fn)(int, int, int, int, int, int, float, const float *, int, const float *, int, float, float *, int);
typedef void (*zgemm_fn)(int, int, int, int, int, int, const void *, const void *, int, const void *, int, const void *, void *, int);
typedef void (*setthr_fn)(int);

static int g_tried = 0;
static void *g_handle = NULL;
static dgemm_fn g_dgemm = NULL;
static sgemm_fn g_sgemm = NULL;
static zgemm_fn g_zgemm = NULL, g_cgemm = NULL;
static setthr_fn g_setthr = NULL;
static char g_backend[512] = "builtin";

/* wheels ship libgfortran/libquadmath beside the BLAS without an rpath: load them first */
static void preload_siblings(const char *path)
{
    char dir[400];
    snprintf(dir, sizeof dir, "%s", path);
    char *slash = strrchr(dir, '/');
    if (!slash) return;
    *slash = 0;
    static const char *deps[] = {"libquadmath*.so*", "libgfortran*.so*", NULL};
    for (int i = 0; deps[i]; i++) {
        char pat[512];
        snprintf(pat, sizeof pat, "%s/%s", dir, deps[i]);
        glob_t g;
        if (glob(pat, 0, NULL, &g) == 0)
            for (size_t k = 0; k < g.gl_pathc; k++) dlopen(g.gl_pathv[k], RTLD_NOW | RTLD_GLOBAL);
        globfree(&g);
    }
}

static void *sym2(void *h, const char *name)
{
    void *p = dlsym(h, name);
    if (p) return p;
    char alt[128];
    snprintf(alt, sizeof alt, "scipy_%s", name); /* scipy-openblas wheels prefix every symbol */
    return dlsym(h, alt);
}

static int try_open(const char *path)
{
    if (strstr(path, "openblas64_")) return 0; /* ILP64 build: 64-bit ints, not our prototypes */
    preload_siblings(path);
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return 0;
    dgemm_fn d = (dgemm_fn)sym2(h, "cblas_dgemm");
    if (!d) { dlclose(h); return 0; }
    g_handle = h;
    g_dgemm = d;
    g_sgemm = (sgemm_fn)sym2(h, "cblas_sgemm");
    g_zgemm = (zgemm_fn)sym2(h, "cblas_zgemm");
    g_cgemm = (zgemm_fn)sym2(h, "cblas_cgemm");
    g_setthr = (setthr_fn)sym2(h, "openblas_set_num_threads");
    snprintf(g_backend, sizeof g_backend, "openblas:%s", path);
    return 1;
}

static void blas_init(void)
{
    if (g_tried) return;
    g_tried = 1;
    const char *env = getenv("PANGULU_ORACLE_OPENBLAS");
    if (env && strcmp(env, "builtin") == 0) return;
    if (env && try_open(env)) goto done;
    {
        static const char *pats[] = {
            "/opt/prime-rl/.venv/lib/python3*/site-packages/scipy.libs/libscipy_openblas*.so*",
            "/opt/prime-rl/.venv/lib/python3*/site-packages/*.libs/libopenblas*.so*",
            "/usr/lib/x86_64-linux-gnu/libopenblas.so*",
            "/usr/lib/x86_64-linux-gnu/openblas-*/libopenblas.so*",
            NULL};
        for (int i = 0; pats[i]; i++) {
            glob_t g;
            if (glob(pats[i], 0, NULL, &g) == 0) {
                for (size_t k = 0; k < g.gl_pathc; k++)
                    if (try_open(g.gl_pathv[k])) { globfree(&g); goto done; }
            }
            globfree(&g);
        }
    }
    return;
done:
    if (g_setthr) g_setthr(1);
}

const char *pangulu_oracle_blas_backend(void) { blas_init(); return g_backend; }

void openblas_set_num_threads(int n) { blas_init(); if (g_setthr) g_setthr(n); }

#define GEMM_BODY(T)                                                                              \
    for (int j = 0; j < n; j++) {                                                                 \
        T *cj = c + (size_t)j * ldc;                                                              \
        if (beta == (T)0) for (int i = 0; i < m; i++) cj[i] = (T)0;                               \
        else if (beta != (T)1) for (int i = 0; i < m; i++) cj[i] *= beta;                         \
    }                                                                                             \
    for (int j0 = 0; j0 < n; j0 += 64)                                                            \
        for (int p0 = 0; p0 < k; p0 += 128) {                                                     \
            int j1 = j0 + 64 < n ? j0 + 64 : n, p1 = p0 + 128 < k ? p0 + 128 : k;                 \
            for (int j = j0; j < j1; j++) {                                                       \
                T *cj = c + (size_t)j * ldc;                                                      \
                for (int p = p0; p < p1; p++) {                                                   \
                    T bpj = alpha * b[(size_t)j * ldb + p];                                       \
                    const T *ap = a + (size_t)p * lda;                                            \
                    for (int i = 0; i < m; i++) cj[i] += ap[i] * bpj;                             \
                }                                                                                 \
            }                                                                                     \
        }

static void builtin_dgemm(int m, int n, int k, double alpha, const double *a, int lda, const double *b, int ldb, double beta, double *c, int ldc)
{ GEMM_BODY(double) }
static void builtin_sgemm(int m, int n, int k, float alpha, const float *a, int lda, const float *b, int ldb, float beta, float *c, int ldc)
{ GEMM_BODY(float) }
static void builtin_zgemm(int m, int n, int k, double _Complex alpha, const double _Complex *a, int lda, const double _Complex *b, int ldb, double _Complex beta, double _Complex *c, int ldc)
{ GEMM_BODY(double _Complex) }
static void builtin_cgemm(int m, int n, int k, float _Complex alpha, const float _Complex *a, int lda, const float _Complex *b, int ldb, float _Complex beta, float _Complex *c, int ldc)
{ GEMM_BODY(float _Complex) }

static void check_layout(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb)
{
    if (o != CblasColMajor || ta != CblasNoTrans || tb != CblasNoTrans) {
        fprintf(stderr, "[pangulu cblas shim] only ColMajor/NoTrans/NoTrans is implemented\n");
        abort();
    }
}

void cblas_dgemm(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                 double alpha, const double *a, int lda, const double *b, int ldb, double beta, double *c, int ldc)
{
    blas_init();
    if (g_dgemm) { g_dgemm(o, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc); return; }
    check_layout(o, ta, tb);
    builtin_dgemm(m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void cblas_sgemm(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                 float alpha, const float *a, int lda, const float *b, int ldb, float beta, float *c, int ldc)
{
    blas_init();
    if (g_sgemm) { g_sgemm(o, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc); return; }
    check_layout(o, ta, tb);
    builtin_sgemm(m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void cblas_zgemm(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                 const void *alpha, const void *a, int lda, const void *b, int ldb, const void *beta, void *c, int ldc)
{
    blas_init();
    if (g_zgemm) { g_zgemm(o, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc); return; }
    check_layout(o, ta, tb);
    builtin_zgemm(m, n, k, *(const double _Complex *)alpha, (const double _Complex *)a, lda, (const double _Complex *)b, ldb,
                  *(const double _Complex *)beta, (double _Complex *)c, ldc);
}
void cblas_cgemm(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb, int m, int n, int k,
                 const void *alpha, const void *a, int lda, const void *b, int ldb, const void *beta, void *c, int ldc)
{
    blas_init();
    if (g_cgemm) { g_cgemm(o, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc); return; }
    check_layout(o, ta, tb);
    builtin_cgemm(m, n, k, *(const float _Complex *)alpha, (const float _Complex *)a, lda, (const float _Complex *)b, ldb,
                  *(const float _Complex *)beta, (float _Complex *)c, ldc);
}
