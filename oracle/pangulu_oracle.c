/* pangulu_oracle.c -- CPU restatement of PanguLU's block kernels (see pangulu_oracle.h).
 * TEST INFRASTRUCTURE ONLY -- never linked into, called by, or shipped with the product path.
 *
 * The reference walks pairs of sorted index lists with two-pointer merges; this restatement gathers one
 * operand into a dense work vector of length nb and sweeps the other, touching only entries that exist in
 * the destination pattern.  Per destination entry the sequence of floating-point operations (one division
 * by the clamped pivot, then "x -= l*u" in ascending elimination order) is the same as the reference's,
 * so GETRF/TSTRF/GESSM agree with it bit for bit; SSSSM accumulates the product column in a different
 * order than the reference's gather->gemm->scatter and agrees to rounding.
 */
#include "pangulu_oracle.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_TOL 1e-16 /* PANGULU_TOL, src/pangulu_common.h:133 */

typedef oracle_value_t T;

int oracle_sizeof_value(void) { return (int)sizeof(T); }

/* if (fabs(pivot) < TOL) pivot = TOL -- fabs() of a complex value converts it to double first, i.e. only
 * the real part is inspected (pangulu_platform_0100000.c:80-84,153-157) */
static T clamp_pivot(T p)
{
    if (fabs((double)p) < ORACLE_TOL) return (T)ORACLE_TOL;
    return p;
}

static uint32_t lo(const uint32_t *ptr, int i) { return i == 0 ? 0u : ptr[i]; }

/* ---- GETRF: pangulu_platform_0100000.c:57-135 ------------------------------------------------------ */
void oracle_getrf(int nb, const uint32_t *lp, const uint16_t *li, T *lv, const uint32_t *up, const uint16_t *ui, T *uv)
{
    T *urow = (T *)calloc((size_t)nb, sizeof(T));   /* U(level, :) scattered by column */
    T *lcol = (T *)calloc((size_t)nb, sizeof(T));   /* L(:, level) scattered by row    */
    char *in_urow = (char *)calloc((size_t)nb, 1);
    char *in_lcol = (char *)calloc((size_t)nb, 1);
    for (int level = 0; level < nb; level++)
    {
        if (up[level] == up[level + 1]) continue; /* :74-77 padded row */
        const T diag = clamp_pivot(uv[up[level]]);
        for (uint32_t e = lp[level]; e < lp[level + 1]; e++)
        {
            lv[e] /= diag;                            /* :85-88 */
            lcol[li[e]] = lv[e];
            in_lcol[li[e]] = 1;
        }
        for (uint32_t e = up[level] + 1; e < up[level + 1]; e++)
        {
            urow[ui[e]] = uv[e];
            in_urow[ui[e]] = 1;
        }
        /* U(r, c) -= L(r, level) U(level, c) for r in L(:,level), c in U(r,:) n U(level,:)   :90-110 */
        for (uint32_t e = lp[level]; e < lp[level + 1]; e++)
        {
            const int r = li[e];
            const T l = lv[e];
            for (uint32_t q = up[r]; q < up[r + 1]; q++)
                if (in_urow[ui[q]]) uv[q] -= l * urow[ui[q]];
        }
        /* L(r, c) -= L(r, level) U(level, c) for c in U(level,:), r in L(:,c) n L(:,level)   :112-132 */
        for (uint32_t e = up[level] + 1; e < up[level + 1]; e++)
        {
            const int c = ui[e];
            const T u = uv[e];
            for (uint32_t q = lp[c]; q < lp[c + 1]; q++)
                if (in_lcol[li[q]]) lv[q] -= lcol[li[q]] * u;
        }
        for (uint32_t e = lp[level]; e < lp[level + 1]; e++) in_lcol[li[e]] = 0;
        for (uint32_t e = up[level] + 1; e < up[level + 1]; e++) in_urow[ui[e]] = 0;
    }
    free(urow); free(lcol); free(in_urow); free(in_lcol);
}

/* ---- TSTRF: pangulu_platform_0100000.c:137-175 ------------------------------------------------------ */
void oracle_tstrf(int nb, const uint32_t *brp, const uint16_t *bci, const uint32_t *bmap, T *bv,
                  const uint32_t *up, const uint16_t *ui, const T *uv)
{
    int64_t *pos = (int64_t *)malloc(sizeof(int64_t) * (size_t)nb); /* column -> index into bv, or -1 */
    for (int r = 0; r < nb; r++)
    {
        for (int c = 0; c < nb; c++) pos[c] = -1;
        for (uint32_t e = brp[r]; e < brp[r + 1]; e++) pos[bci[e]] = bmap[e];
        for (uint32_t e = brp[r]; e < brp[r + 1]; e++)
        {
            const int c = bci[e];
            const T diag = clamp_pivot(uv[up[c]]);
            bv[bmap[e]] /= diag;                         /* :158 */
            const T m = bv[bmap[e]];
            for (uint32_t q = up[c] + 1; q < up[c + 1]; q++) /* later entries of the row that U(c,:) reaches   :160-172 */
                if (pos[ui[q]] >= 0 && ui[q] > c) bv[pos[ui[q]]] -= m * uv[q];
        }
    }
    free(pos);
}

/* ---- GESSM: pangulu_platform_0100000.c:178-209 ------------------------------------------------------ */
void oracle_gessm(int nb, const uint32_t *bp, const uint16_t *bi, T *bv, const uint32_t *lp, const uint16_t *li, const T *lv)
{
    int64_t *pos = (int64_t *)malloc(sizeof(int64_t) * (size_t)nb); /* row -> index into bv, or -1 */
    for (int c = 0; c < nb; c++)
    {
        for (int r = 0; r < nb; r++) pos[r] = -1;
        for (uint32_t e = bp[c]; e < bp[c + 1]; e++) pos[bi[e]] = e;
        for (uint32_t e = bp[c]; e < bp[c + 1]; e++)
        {
            const int r = bi[e];
            const T m = bv[e];
            for (uint32_t q = lp[r]; q < lp[r + 1]; q++)
                if (pos[li[q]] >= 0 && li[q] > r) bv[pos[li[q]]] -= m * lv[q];
        }
    }
    free(pos);
}

/* ---- SSSSM: pangulu_platform_0100000.c:211-397 -------------------------------------------------------
 * C <- C - A B restricted to C's pattern.  (The reference compacts A and B to their non-empty rows/columns,
 * multiplies with cblas_?gemm and scatters; the result is the same sum of products.) */
static void product_column(int nb, int j, const uint32_t *ap, const uint16_t *ai, const T *av,
                           const uint32_t *bp, const uint16_t *bi, const T *bv, T *acc)
{
    memset(acc, 0, sizeof(T) * (size_t)nb);
    for (uint32_t e = lo(bp, j); e < bp[j + 1]; e++)
    {
        const int k = bi[e];
        const T b = bv[e];
        for (uint32_t q = lo(ap, k); q < ap[k + 1]; q++) acc[ai[q]] += av[q] * b;
    }
}

void oracle_ssssm(int nb, const uint32_t *cp, const uint16_t *ci, T *cv, const uint32_t *ap, const uint16_t *ai, const T *av,
                  const uint32_t *bp, const uint16_t *bi, const T *bv)
{
    T *acc = (T *)malloc(sizeof(T) * (size_t)nb);
    for (int j = 0; j < nb; j++)
    {
        if (lo(bp, j) == bp[j + 1]) continue;
        product_column(nb, j, ap, ai, av, bp, bi, bv, acc);
        for (uint32_t e = lo(cp, j); e < cp[j + 1]; e++) cv[e] -= acc[ci[e]]; /* :376-396 */
    }
    free(acc);
}

void oracle_ssssm_diag(int nb, const uint32_t *lp, const uint16_t *li, T *lv, const uint32_t *up, const uint16_t *ui, T *uv,
                       const uint32_t *ap, const uint16_t *ai, const T *av, const uint32_t *bp, const uint16_t *bi, const T *bv)
{
    /* full product, column major nb x nb; the reference subtracts it from the strict-lower CSC half
     * column by column (:344-359) and from the upper CSR half row by row (:360-374) */
    T *prod = (T *)calloc((size_t)nb * nb, sizeof(T));
    for (int j = 0; j < nb; j++)
        if (lo(bp, j) != bp[j + 1]) product_column(nb, j, ap, ai, av, bp, bi, bv, prod + (size_t)j * nb);
    for (int j = 0; j < nb; j++)
        for (uint32_t e = lp[j]; e < lp[j + 1]; e++) lv[e] -= prod[(size_t)j * nb + li[e]];
    for (int r = 0; r < nb; r++)
        for (uint32_t e = up[r]; e < up[r + 1]; e++) uv[e] -= prod[(size_t)ui[e] * nb + r];
    free(prod);
}

/* ---- flop model: src/pangulu_kernel_interface.c:4-176 ------------------------------------------------------
 * The reference counts, with two-pointer merges, how many index pairs coincide.  Here the same count
 * is taken with a membership table. */
static long long common_count(const uint16_t *a, uint32_t a0, uint32_t a1, const uint16_t *b, uint32_t b0, uint32_t b1, char *mark)
{
    long long n = 0;
    for (uint32_t e = a0; e < a1; e++) mark[a[e]] = 1;
    for (uint32_t e = b0; e < b1; e++) n += mark[b[e]];
    for (uint32_t e = a0; e < a1; e++) mark[a[e]] = 0;
    return n;
}

long long oracle_getrf_flop(int nb, const uint32_t *lp, const uint16_t *li, const uint32_t *up, const uint16_t *ui)
{
    long long flop = 0;
    char *mark = (char *)calloc((size_t)nb, 1);
    for (int level = 0; level < nb; level++)
    {
        if (up[level] == up[level + 1]) continue;
        flop += lp[level + 1] - lp[level];                                           /* :28 divisions */
        for (uint32_t e = lp[level]; e < lp[level + 1]; e++)                       /* :30-54 */
            flop += 2 * common_count(ui, up[level], up[level + 1], ui, up[li[e]], up[li[e] + 1], mark);
        for (uint32_t e = up[level] + 1; e < up[level + 1]; e++)                   /* :56-80 */
            flop += 2 * common_count(li, lp[level], lp[level + 1], li, lp[ui[e]], lp[ui[e] + 1], mark);
    }
    free(mark);
    return flop;
}

long long oracle_tstrf_flop(int nb, const uint32_t *brp, const uint16_t *bci, const uint32_t *up, const uint16_t *ui)
{
    long long flop = 0;
    char *mark = (char *)calloc((size_t)nb, 1);
    for (int r = 0; r < nb; r++)
        for (uint32_t e = brp[r]; e < brp[r + 1]; e++)
        {
            flop++;                                                                   /* :101 */
            flop += 2 * common_count(bci, e + 1, brp[r + 1], ui, up[bci[e]], up[bci[e] + 1], mark); /* :102-118 */
        }
    free(mark);
    return flop;
}

long long oracle_gessm_flop(int nb, const uint32_t *bp, const uint16_t *bi, const uint32_t *lp, const uint16_t *li)
{
    long long flop = 0;
    char *mark = (char *)calloc((size_t)nb, 1);
    for (int c = 0; c < nb; c++)
        for (uint32_t e = bp[c]; e < bp[c + 1]; e++)
            flop += 2 * common_count(bi, e + 1, bp[c + 1], li, lp[bi[e]], lp[bi[e] + 1], mark);       /* :140-156 */
    free(mark);
    return flop;
}

long long oracle_ssssm_flop(int nb, const uint32_t *ap, const uint32_t *bp, const uint16_t *bi)
{
    long long flop = 0;
    for (int c = 0; c < nb; c++)
        for (uint32_t e = bp[c]; e < bp[c + 1]; e++) flop += 2LL * (ap[bi[e] + 1] - ap[bi[e]]);      /* :168-175 */
    return flop;
}

/* ---- solve-phase kernels: pangulu_platform_0100000.c:435-506 -------------------------------------------- */
void oracle_spmv(int nb, const uint32_t *cp, const uint16_t *ri, const T *v, const T *x, T *y)
{
    for (int c = 0; c < nb; c++)
        for (uint32_t e = lo(cp, c); e < cp[c + 1]; e++) y[ri[e]] -= v[e] * x[c];
}

void oracle_sptrsv(int nb, const uint32_t *p, const uint16_t *idx, const T *v, T *x, int uplo)
{
    if (uplo == 0)
    {
        for (int c = 0; c < nb; c++)
            for (uint32_t e = p[c]; e < p[c + 1]; e++) x[idx[e]] -= v[e] * x[c];
    }
    else
    {
        for (int r = nb - 1; r >= 0; r--)
        {
            if (p[r] == p[r + 1]) continue;
            for (uint32_t e = p[r] + 1; e < p[r + 1]; e++) x[r] -= v[e] * x[idx[e]];
            if (fabs((double)v[p[r]]) > ORACLE_TOL) x[r] = x[r] / v[p[r]];
            else x[r] = x[r] / ORACLE_TOL;
        }
    }
}
