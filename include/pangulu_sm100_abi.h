/* pangulu_sm100_abi.h -- the data layout shared between PanguLU's host code and
 * the sm_100a platform module, restated so the module also builds OUTSIDE the
 * PanguLU tree (tests, bench, the host driver in pangulu_b200/csrc/host).
 *
 * Every declaration here mirrors one in the reference's src/pangulu_common.h
 * (file:line given beside it) as compiled with -DGPU_OPEN, i.e. with the
 * device twin pointers present.  tests/test_abi.py checks each field offset
 * against the reference build (oracle/_ref, GPU-flavoured compile of the
 * harness) and this header static_asserts the x86-64 numbers of SURVEY.md
 * Appendix A.
 *
 * When the module is compiled INSIDE the PanguLU tree define
 * PANGULU_SM100_USE_REFERENCE_HEADER: the reference's own pangulu_common.h is
 * included instead and this file only adds the static checks.
 *
 * Value type: exactly one of CALCULATE_TYPE_{R64,R32,CR64,CR32}, as in the
 * reference (src/pangulu_common.h:11-33); default R64.
 */
#ifndef PANGULU_SM100_ABI_H
#define PANGULU_SM100_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef PANGULU_SM100_USE_REFERENCE_HEADER
#ifndef GPU_OPEN
#error "the sm_100a platform needs the reference header compiled with -DGPU_OPEN (device twin pointers)"
#endif
#define PANGULU_PLATFORM_ENV
#include "pangulu_common.h"
#if defined(__cplusplus)
/* helper pair types of the kernels (real value types only in this mode: PanguLU's header gives a C++ translation unit
 * `double _Complex`, which the kernels do not take) */
struct pangulu_sm100_c64 { double re, im; };
struct pangulu_sm100_c32 { float re, im; };
#if defined(CALCULATE_TYPE_CR64) || defined(CALCULATE_TYPE_CR32)
#error "PANGULU_SM100_USE_REFERENCE_HEADER supports the real value types; build the complex types with the restated header"
#endif
#endif
#else

#if !defined(CALCULATE_TYPE_R64) && !defined(CALCULATE_TYPE_R32) && !defined(CALCULATE_TYPE_CR64) && !defined(CALCULATE_TYPE_CR32)
#define CALCULATE_TYPE_R64
#endif

/* src/pangulu_common.h:35-65 */
typedef long long int pangulu_int64_t;
typedef unsigned long long int pangulu_uint64_t;
typedef int pangulu_int32_t;
typedef unsigned int pangulu_uint32_t;
typedef short int pangulu_int16_t;
typedef unsigned short int pangulu_uint16_t;
typedef pangulu_uint64_t pangulu_exblock_ptr; /* pointer into the whole matrix      */
typedef pangulu_uint32_t pangulu_exblock_idx; /* row/column of the whole matrix     */
typedef pangulu_uint32_t pangulu_inblock_ptr; /* pointer inside one nb x nb block   */
typedef pangulu_uint16_t pangulu_inblock_idx; /* row/column inside one block        */

/* src/pangulu_common.h:11-33.  In C++/CUDA translation units the complex types are
 * plain {re, im} pairs with the same size and alignment as C99 _Complex. */
#if defined(__cplusplus)
struct pangulu_sm100_c64 { double re, im; };
struct pangulu_sm100_c32 { float re, im; };
#if defined(CALCULATE_TYPE_CR64)
typedef pangulu_sm100_c64 calculate_type;
typedef double calculate_real_type;
#define PANGULU_COMPLEX
#elif defined(CALCULATE_TYPE_CR32)
typedef pangulu_sm100_c32 calculate_type;
typedef float calculate_real_type;
#define PANGULU_COMPLEX
#elif defined(CALCULATE_TYPE_R32)
typedef float calculate_type;
typedef float calculate_real_type;
#else
typedef double calculate_type;
typedef double calculate_real_type;
#endif
#else
#if defined(CALCULATE_TYPE_CR64)
typedef double _Complex calculate_type;
typedef double calculate_real_type;
#define PANGULU_COMPLEX
#elif defined(CALCULATE_TYPE_CR32)
typedef float _Complex calculate_type;
typedef float calculate_real_type;
#define PANGULU_COMPLEX
#elif defined(CALCULATE_TYPE_R32)
typedef float calculate_type;
typedef float calculate_real_type;
#else
typedef double calculate_type;
typedef double calculate_real_type;
#endif
#endif

/* src/pangulu_common.h:124-133 */
#define PANGULU_TASK_GETRF 1
#define PANGULU_TASK_TSTRF 2
#define PANGULU_TASK_GESSM 3
#define PANGULU_TASK_SSSSM 4
#define PANGULU_LOWER 0
#define PANGULU_UPPER 1
#define PANGULU_DATA_INVALID 0
#define PANGULU_DATA_PREPARING 1
#define PANGULU_DATA_READY 2
#define PANGULU_TOL 1e-16
#define PANGULU_SPTRSV_TOL 1e-16
/* owner of block (row, col) on a p x q rank grid, src/pangulu_common.h:135 */
#define PANGULU_CALC_RANK(row, col, p, q) (((row) % (p)) * (q) + ((col) % (q)))

/* src/pangulu_common.h:200-205 (opaque to the platform) */
typedef struct pangulu_aggregate_queue_t
{
    unsigned long long capacity;
    unsigned long long length;
    void *task_descriptors;
} pangulu_aggregate_queue_t;

/* One nb x nb sparse block, src/pangulu_common.h:207-231.
 *  - off-diagonal block and LOWER half of a diagonal block: CSC in
 *    columnpointer/rowindex/value (d_columnpointer/d_rowindex/d_value on device);
 *  - UPPER half of a diagonal block: CSR (diagonal entry first in each row) kept in the SAME
 *    host fields, but in d_rowpointer/d_columnindex on the device
 *    (src/pangulu_storage.c:346-352);
 *  - strictly-lower off-diagonal blocks (brow > bcol) also carry a CSR mirror of the pattern and
 *    idx_of_csc_value_for_csr[] mapping CSR position -> index into value[]. */
typedef struct pangulu_storage_slot_t
{
    pangulu_exblock_idx brow_pos;
    pangulu_exblock_idx bcol_pos;
    pangulu_inblock_ptr *columnpointer;
    pangulu_inblock_idx *rowindex;
    calculate_type *value;
    pangulu_inblock_ptr *rowpointer;
    pangulu_inblock_idx *columnindex;
    pangulu_inblock_ptr *idx_of_csc_value_for_csr;
    volatile char data_status;
    struct pangulu_storage_slot_t *related_block;
    pangulu_int32_t is_upper;
    pangulu_int32_t bin_id;
    pangulu_int32_t slot_idx;
    pangulu_aggregate_queue_t *task_queue;
    /* #ifdef PANGULU_NONSHAREDMEM in the reference */
    pangulu_inblock_ptr *d_columnpointer;
    pangulu_inblock_idx *d_rowindex;
    calculate_type *d_value;
    pangulu_inblock_ptr *d_rowpointer;
    pangulu_inblock_idx *d_columnindex;
    pangulu_inblock_ptr *d_idx_of_csc_value_for_csr;
} pangulu_storage_slot_t;

/* One block task, src/pangulu_common.h:233-243.  TSTRF/GESSM: op1 = either half of the diagonal
 * block (callee normalises through related_block), op2 = NULL.  SSSSM: op1 = L-panel block (i,k),
 * op2 = U-panel block (k,j). */
typedef struct pangulu_task_t
{
    pangulu_exblock_idx row;
    pangulu_exblock_idx col;
    pangulu_int16_t kernel_id;
    pangulu_exblock_idx task_level;
    pangulu_int64_t compare_flag;
    pangulu_storage_slot_t *opdst;
    pangulu_storage_slot_t *op1;
    pangulu_storage_slot_t *op2;
} pangulu_task_t;

#endif /* PANGULU_SM100_USE_REFERENCE_HEADER */

/* x86-64 layout the module was written against */
#if defined(__cplusplus)
#define PANGULU_SM100_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define PANGULU_SM100_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif
PANGULU_SM100_STATIC_ASSERT(sizeof(pangulu_storage_slot_t) == 144, "slot size");
PANGULU_SM100_STATIC_ASSERT(offsetof(pangulu_storage_slot_t, value) == 24, "slot.value");
PANGULU_SM100_STATIC_ASSERT(offsetof(pangulu_storage_slot_t, related_block) == 64, "slot.related_block");
PANGULU_SM100_STATIC_ASSERT(offsetof(pangulu_storage_slot_t, is_upper) == 72, "slot.is_upper");
PANGULU_SM100_STATIC_ASSERT(offsetof(pangulu_storage_slot_t, d_columnpointer) == 96, "slot.d_columnpointer");
PANGULU_SM100_STATIC_ASSERT(offsetof(pangulu_storage_slot_t, d_value) == 112, "slot.d_value");
PANGULU_SM100_STATIC_ASSERT(offsetof(pangulu_storage_slot_t, d_idx_of_csc_value_for_csr) == 136, "slot.d_idx");
PANGULU_SM100_STATIC_ASSERT(sizeof(pangulu_task_t) == 48, "task size");
PANGULU_SM100_STATIC_ASSERT(offsetof(pangulu_task_t, kernel_id) == 8, "task.kernel_id");
PANGULU_SM100_STATIC_ASSERT(offsetof(pangulu_task_t, opdst) == 24, "task.opdst");

/* Packed block layout in host memory, device memory and on the wire
 * (src/pangulu_communication.c:1294-1322, src/pangulu_storage.c:297-326):
 *   [-32] u64 nnz | u32 brow | u32 bcol | u32 is_upper | 12 B pad     (header, filled at send time)
 *   [  0] T   value[nnz]
 *         u32 columnpointer[nb+1]
 *         u16 rowindex[nnz]           then pad to 8 B
 *   only if brow > bcol:
 *         u32 idx_of_csc_value_for_csr[nnz]
 *         u32 rowpointer[nb+1]
 *         u16 columnindex[nnz]        then pad to 8 B                                   */
#define PANGULU_SM100_BLOCK_HEADER_BYTES 32

static inline size_t pangulu_sm100_align8(size_t x) { return (x + 7u) & ~(size_t)7u; }

/* bytes of one packed block including its 32 B header */
static inline size_t pangulu_sm100_packed_bytes(size_t sizeof_value, size_t nb, size_t nnz, int with_csr_mirror)
{
    size_t s = PANGULU_SM100_BLOCK_HEADER_BYTES + sizeof_value * nnz + 4u * (nb + 1) + 2u * nnz;
    s = pangulu_sm100_align8(s);
    if (with_csr_mirror)
    {
        s += 4u * nnz + 4u * (nb + 1) + 2u * nnz;
        s = pangulu_sm100_align8(s);
    }
    return s;
}

#endif /* PANGULU_SM100_ABI_H */
