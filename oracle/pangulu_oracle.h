/* pangulu_oracle.h -- CPU restatement of PanguLU's four block kernels and its flop model.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against the unmodified
 * reference CPU platform (oracle/_ref, built in place from /root/reference by oracle/Makefile) on the
 * reference's shipped fixture examples/Trefethen_20b.mtx and on seeded stencil matrices, and against the
 * golden vectors under tests/golden/ that were generated from that same reference build
 * (tests/golden/make_golden.py).
 *
 * Blocks are passed as plain arrays (no structs): colptr u32[nb+1], rowidx u16[nnz], value T[nnz];
 * for the upper half of a diagonal block the same trio holds rowptr / colidx / value (CSR, diagonal
 * entry first in each row).  T is chosen at compile time by -DCALCULATE_TYPE_{R64,R32,CR64,CR32}.
 */
#ifndef PANGULU_ORACLE_H
#define PANGULU_ORACLE_H
#include <stdint.h>

#if defined(CALCULATE_TYPE_CR64)
typedef double _Complex oracle_value_t;
#elif defined(CALCULATE_TYPE_CR32)
typedef float _Complex oracle_value_t;
#elif defined(CALCULATE_TYPE_R32)
typedef float oracle_value_t;
#else
typedef double oracle_value_t;
#endif

int oracle_sizeof_value(void);

/* GETRF, pangulu_platform_0100000.c:57-135 */
void oracle_getrf(int nb, const uint32_t *l_colptr, const uint16_t *l_rowidx, oracle_value_t *l_val,
                  const uint32_t *u_rowptr, const uint16_t *u_colidx, oracle_value_t *u_val);
/* TSTRF, pangulu_platform_0100000.c:137-175; B given by its CSR mirror + value index into b_val */
void oracle_tstrf(int nb, const uint32_t *b_rowptr, const uint16_t *b_colidx, const uint32_t *b_csr2csc, oracle_value_t *b_val,
                  const uint32_t *u_rowptr, const uint16_t *u_colidx, const oracle_value_t *u_val);
/* GESSM, pangulu_platform_0100000.c:178-209 */
void oracle_gessm(int nb, const uint32_t *b_colptr, const uint16_t *b_rowidx, oracle_value_t *b_val,
                  const uint32_t *l_colptr, const uint16_t *l_rowidx, const oracle_value_t *l_val);
/* SSSSM with an off-diagonal destination, pangulu_platform_0100000.c:211-331,376-396 */
void oracle_ssssm(int nb, const uint32_t *c_colptr, const uint16_t *c_rowidx, oracle_value_t *c_val,
                  const uint32_t *a_colptr, const uint16_t *a_rowidx, const oracle_value_t *a_val,
                  const uint32_t *b_colptr, const uint16_t *b_rowidx, const oracle_value_t *b_val);
/* SSSSM with a diagonal destination (both halves updated), pangulu_platform_0100000.c:332-375 */
void oracle_ssssm_diag(int nb, const uint32_t *l_colptr, const uint16_t *l_rowidx, oracle_value_t *l_val,
                       const uint32_t *u_rowptr, const uint16_t *u_colidx, oracle_value_t *u_val,
                       const uint32_t *a_colptr, const uint16_t *a_rowidx, const oracle_value_t *a_val,
                       const uint32_t *b_colptr, const uint16_t *b_rowidx, const oracle_value_t *b_val);

/* flop model, src/pangulu_kernel_interface.c:4-176 (merge-based, exactly as the reference counts) */
long long oracle_getrf_flop(int nb, const uint32_t *l_colptr, const uint16_t *l_rowidx, const uint32_t *u_rowptr, const uint16_t *u_colidx);
long long oracle_tstrf_flop(int nb, const uint32_t *b_rowptr, const uint16_t *b_colidx, const uint32_t *u_rowptr, const uint16_t *u_colidx);
long long oracle_gessm_flop(int nb, const uint32_t *b_colptr, const uint16_t *b_rowidx, const uint32_t *l_colptr, const uint16_t *l_rowidx);
long long oracle_ssssm_flop(int nb, const uint32_t *a_colptr, const uint32_t *b_colptr, const uint16_t *b_rowidx);

/* solve-phase block kernels, pangulu_platform_0100000.c:435-506 */
void oracle_spmv(int nb, const uint32_t *colptr, const uint16_t *rowidx, const oracle_value_t *val, const oracle_value_t *x, oracle_value_t *y);
void oracle_sptrsv(int nb, const uint32_t *ptr, const uint16_t *idx, const oracle_value_t *val, oracle_value_t *x, int uplo);

#endif
