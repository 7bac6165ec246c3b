/* Harness around the UNMODIFIED PanguLU reference (compiled in place from
 * /root/reference by oracle/Makefile into oracle/_ref/).  Test infrastructure
 * only: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs load it; the product library never does.
 *
 * It gives Python (ctypes) three things:
 *   1. the reference's own public API driven end to end
 *      (pangulu_init -> pangulu_gstrf -> pangulu_gstrs, include/pangulu.h:11-15)
 *      with a wall clock around pangulu_gstrf;
 *   2. a walk over handle->block_smatrix->storage->bins[0]
 *      (src/pangulu_storage.c:239-293) so every owned block's pattern and
 *      values can be read back -- the golden factors;
 *   3. direct calls into the reference CPU platform kernels
 *      pangulu_platform_0100000_{getrf,tstrf,gessm,ssssm} on caller-built slots
 *      (the per-kernel oracle), after allocating the scratch tables the CPU
 *      SSSSM expects (src/pangulu_preprocessing.c:596-611).
 */
#include "pangulu_common.h"

#ifdef PANGULU_PERF
extern pangulu_stat_t global_stat;
#endif

extern pangulu_int32_t **ssssm_hash_lu;
extern pangulu_int32_t **ssssm_hash_l_row;
extern pangulu_int32_t **ssssm_hash_l_row_inv;
extern pangulu_int32_t **ssssm_hash_u_col;
extern pangulu_int32_t **ssssm_hash_u_col_inv;
extern calculate_type **ssssm_l_value;
extern calculate_type **ssssm_u_value;
extern calculate_type **temp_a_value;

int refh_sizeof_value(void) { return (int)sizeof(calculate_type); }
int refh_sizeof_slot(void) { return (int)sizeof(pangulu_storage_slot_t); }
int refh_sizeof_task(void) { return (int)sizeof(pangulu_task_t); }
int refh_is_gpu_build(void)
{
#ifdef GPU_OPEN
    return 1;
#else
    return 0;
#endif
}
int refh_has_metis(void)
{
#ifdef METIS
    return 1;
#else
    return 0;
#endif
}
int refh_has_mc64(void)
{
#ifdef PANGULU_MC64
    return 1;
#else
    return 0;
#endif
}
int refh_has_perf(void)
{
#ifdef PANGULU_PERF
    return 1;
#else
    return 0;
#endif
}

/* Field offsets of the reference structs, for the ABI check in tests/
 * (include/pangulu_sm100_abi.h must agree with the GPU build). */
void refh_slot_offsets(long long *out)
{
    int k = 0;
    out[k++] = offsetof(pangulu_storage_slot_t, brow_pos);
    out[k++] = offsetof(pangulu_storage_slot_t, bcol_pos);
    out[k++] = offsetof(pangulu_storage_slot_t, columnpointer);
    out[k++] = offsetof(pangulu_storage_slot_t, rowindex);
    out[k++] = offsetof(pangulu_storage_slot_t, value);
    out[k++] = offsetof(pangulu_storage_slot_t, rowpointer);
    out[k++] = offsetof(pangulu_storage_slot_t, columnindex);
    out[k++] = offsetof(pangulu_storage_slot_t, idx_of_csc_value_for_csr);
    out[k++] = offsetof(pangulu_storage_slot_t, data_status);
    out[k++] = offsetof(pangulu_storage_slot_t, related_block);
    out[k++] = offsetof(pangulu_storage_slot_t, is_upper);
    out[k++] = offsetof(pangulu_storage_slot_t, bin_id);
    out[k++] = offsetof(pangulu_storage_slot_t, slot_idx);
    out[k++] = offsetof(pangulu_storage_slot_t, task_queue);
#ifdef PANGULU_NONSHAREDMEM
    out[k++] = offsetof(pangulu_storage_slot_t, d_columnpointer);
    out[k++] = offsetof(pangulu_storage_slot_t, d_rowindex);
    out[k++] = offsetof(pangulu_storage_slot_t, d_value);
    out[k++] = offsetof(pangulu_storage_slot_t, d_rowpointer);
    out[k++] = offsetof(pangulu_storage_slot_t, d_columnindex);
    out[k++] = offsetof(pangulu_storage_slot_t, d_idx_of_csc_value_for_csr);
#else
    for (int i = 0; i < 6; i++) out[k++] = -1;
#endif
    out[k++] = sizeof(pangulu_storage_slot_t);
    out[k++] = offsetof(pangulu_task_t, row);
    out[k++] = offsetof(pangulu_task_t, col);
    out[k++] = offsetof(pangulu_task_t, kernel_id);
    out[k++] = offsetof(pangulu_task_t, task_level);
    out[k++] = offsetof(pangulu_task_t, compare_flag);
    out[k++] = offsetof(pangulu_task_t, opdst);
    out[k++] = offsetof(pangulu_task_t, op1);
    out[k++] = offsetof(pangulu_task_t, op2);
    out[k++] = sizeof(pangulu_task_t);
}

/* ---- 1. end-to-end API ------------------------------------------------- */

typedef struct refh_ctx
{
    void *handle;
    double t_gstrf;
    double t_init;
    double t_gstrs;
} refh_ctx;

static double now_s(void)
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

/* The reference frees nothing it is given, and keeps no reference to the CSC
 * arrays after init except through its own copies. */
refh_ctx *refh_init(pangulu_exblock_idx n, pangulu_exblock_ptr nnz, pangulu_exblock_ptr *colptr,
                    pangulu_exblock_idx *rowidx, calculate_type *value, int nb, int nthread,
                    float recv_buffer_level)
{
    static int mpi_up = 0;
    if (!mpi_up)
    {
        int provided = 0;
        MPI_Init_thread(NULL, NULL, MPI_THREAD_MULTIPLE, &provided);
        mpi_up = 1;
    }
    refh_ctx *ctx = (refh_ctx *)calloc(1, sizeof(refh_ctx));
    pangulu_init_options opt;
    memset(&opt, 0, sizeof opt);
    opt.nb = nb;
    opt.nthread = nthread;
    opt.gpu_kernel_warp_per_block = 4;
    opt.gpu_data_move_warp_per_block = 4;
    opt.sizeof_value = sizeof(calculate_type);
#ifdef PANGULU_COMPLEX
    opt.is_complex_matrix = 1;
#else
    opt.is_complex_matrix = 0;
#endif
    opt.mpi_recv_buffer_level = recv_buffer_level;
    double t0 = now_s();
    pangulu_init(n, nnz, colptr, rowidx, value, &opt, &ctx->handle);
    ctx->t_init = now_s() - t0;
    return ctx;
}

void refh_gstrf(refh_ctx *ctx)
{
    pangulu_gstrf_options opt;
    double t0 = now_s();
    pangulu_gstrf(&opt, &ctx->handle);
    ctx->t_gstrf = now_s() - t0;
}

void refh_gstrs(refh_ctx *ctx, calculate_type *rhs)
{
    pangulu_gstrs_options opt;
    double t0 = now_s();
    pangulu_gstrs(rhs, &opt, &ctx->handle);
    ctx->t_gstrs = now_s() - t0;
}

double refh_time(refh_ctx *ctx, int which) { return which == 0 ? ctx->t_init : which == 1 ? ctx->t_gstrf : ctx->t_gstrs; }

long long refh_flop(void)
{
#ifdef PANGULU_PERF
    return global_stat.flop;
#else
    return -1;
#endif
}

void refh_reset_flop(void)
{
#ifdef PANGULU_PERF
    memset(&global_stat, 0, sizeof global_stat);
#endif
}

/* ---- 2. factor walk ------------------------------------------------------ */

static pangulu_handle_t *H(refh_ctx *ctx) { return (pangulu_handle_t *)ctx->handle; }

long long refh_block_length(refh_ctx *ctx) { return H(ctx)->block_common->block_length; }
long long refh_nb(refh_ctx *ctx) { return H(ctx)->block_common->nb; }
long long refh_symbolic_nnz(refh_ctx *ctx) { return (long long)H(ctx)->block_smatrix->symbolic_nnz; }
long long refh_num_slots(refh_ctx *ctx) { return H(ctx)->block_smatrix->storage->bins[0].slot_count; }
long long refh_num_nondiag(refh_ctx *ctx) { return H(ctx)->block_smatrix->storage->bins[0].nondiag_slot_cnt; }
void *refh_slot(refh_ctx *ctx, long long i) { return &H(ctx)->block_smatrix->storage->bins[0].slots[i]; }

/* info = {brow, bcol, is_upper, nnz, has_csr_mirror} */
void refh_slot_info(refh_ctx *ctx, long long i, long long *info)
{
    pangulu_storage_slot_t *s = &H(ctx)->block_smatrix->storage->bins[0].slots[i];
    pangulu_inblock_idx nb = H(ctx)->block_common->nb;
    info[0] = s->brow_pos;
    info[1] = s->bcol_pos;
    /* bins[0].slots is malloc'ed, not zeroed (src/pangulu_storage.c:243): is_upper/rowpointer are only
     * meaningful where pangulu_storage_init wrote them (diag pairs; CSR mirror when brow > bcol) */
    info[2] = (s->brow_pos == s->bcol_pos) ? s->is_upper : (s->brow_pos < s->bcol_pos);
    info[3] = s->columnpointer[nb];
    info[4] = (s->brow_pos > s->bcol_pos);
}

/* which: 0 columnpointer, 1 rowindex, 2 value, 3 rowpointer, 4 columnindex, 5 idx_of_csc_value_for_csr */
void *refh_slot_array(refh_ctx *ctx, long long i, int which)
{
    pangulu_storage_slot_t *s = &H(ctx)->block_smatrix->storage->bins[0].slots[i];
    switch (which)
    {
    case 0: return s->columnpointer;
    case 1: return s->rowindex;
    case 2: return s->value;
    case 3: return s->rowpointer;
    case 4: return s->columnindex;
    case 5: return s->idx_of_csc_value_for_csr;
    }
    return NULL;
}

/* which: 0 row_perm, 1 col_perm, 2 metis_perm (u32[n]); 3 row_scale, 4 col_scale (T[n]) -- rank 0 only,
 * src/pangulu_reordering.c:1259-1263 */
void *refh_perm(refh_ctx *ctx, int which)
{
    pangulu_block_smatrix *b = H(ctx)->block_smatrix;
    switch (which)
    {
    case 0: return b->row_perm;
    case 1: return b->col_perm;
    case 2: return b->metis_perm;
    case 3: return b->row_scale;
    case 4: return b->col_scale;
    }
    return NULL;
}

/* element-level symbolic structure (lower incl. diagonal, by column): src/pangulu_symbolic.c:242-244.
 * NOTE: pangulu_preprocessing frees it on some paths; only valid right after init if still present. */
void *refh_symbolic(refh_ctx *ctx, int which)
{
    pangulu_block_smatrix *b = H(ctx)->block_smatrix;
    return which == 0 ? (void *)b->symbolic_rowpointer : (void *)b->symbolic_columnindex;
}

/* ---- 3. per-kernel oracle --------------------------------------------------- */

void refh_kernel_scratch_init(void)
{
    if (ssssm_hash_lu) return;
    int nthr = 1;
    ssssm_hash_lu = (pangulu_int32_t **)calloc(nthr, sizeof(void *));
    ssssm_hash_l_row = (pangulu_int32_t **)calloc(nthr, sizeof(void *));
    ssssm_hash_l_row_inv = (pangulu_int32_t **)calloc(nthr, sizeof(void *));
    ssssm_hash_u_col = (pangulu_int32_t **)calloc(nthr, sizeof(void *));
    ssssm_hash_u_col_inv = (pangulu_int32_t **)calloc(nthr, sizeof(void *));
    ssssm_l_value = (calculate_type **)calloc(nthr, sizeof(void *));
    ssssm_u_value = (calculate_type **)calloc(nthr, sizeof(void *));
    temp_a_value = (calculate_type **)calloc(nthr, sizeof(void *));
}

/* The CPU SSSSM caches nb-sized scratch on first use; tests that change nb must drop it. */
void refh_kernel_scratch_reset(void)
{
    if (!ssssm_hash_lu) return;
    free(ssssm_hash_lu[0]); ssssm_hash_lu[0] = NULL;
    free(ssssm_hash_l_row[0]); ssssm_hash_l_row[0] = NULL;
    free(ssssm_hash_l_row_inv[0]); ssssm_hash_l_row_inv[0] = NULL;
    free(ssssm_hash_u_col[0]); ssssm_hash_u_col[0] = NULL;
    free(ssssm_hash_u_col_inv[0]); ssssm_hash_u_col_inv[0] = NULL;
    free(ssssm_l_value[0]); ssssm_l_value[0] = NULL;
    free(ssssm_u_value[0]); ssssm_u_value[0] = NULL;
    free(temp_a_value[0]); temp_a_value[0] = NULL;
}

/* Build a host slot around caller-owned arrays (CSC trio, optional CSR mirror). */
pangulu_storage_slot_t *refh_make_slot(pangulu_exblock_idx brow, pangulu_exblock_idx bcol, int is_upper,
                                       pangulu_inblock_ptr *colptr, pangulu_inblock_idx *rowidx, calculate_type *value,
                                       pangulu_inblock_ptr *rowptr, pangulu_inblock_idx *colidx, pangulu_inblock_ptr *csr2csc)
{
    pangulu_storage_slot_t *s = (pangulu_storage_slot_t *)calloc(1, sizeof(pangulu_storage_slot_t));
    s->brow_pos = brow;
    s->bcol_pos = bcol;
    s->is_upper = is_upper;
    s->columnpointer = colptr;
    s->rowindex = rowidx;
    s->value = value;
    s->rowpointer = rowptr;
    s->columnindex = colidx;
    s->idx_of_csc_value_for_csr = csr2csc;
    s->data_status = PANGULU_DATA_READY;
    return s;
}
void refh_link_diag(pangulu_storage_slot_t *lower, pangulu_storage_slot_t *upper)
{
    lower->related_block = upper;
    upper->related_block = lower;
    lower->is_upper = 0;
    upper->is_upper = 1;
}
void refh_free_slot(pangulu_storage_slot_t *s) { free(s); }

void refh_getrf(int nb, pangulu_storage_slot_t *d) { pangulu_platform_0100000_getrf(nb, d, 0); }
void refh_tstrf(int nb, pangulu_storage_slot_t *b, pangulu_storage_slot_t *diag) { pangulu_platform_0100000_tstrf(nb, b, diag, 0); }
void refh_gessm(int nb, pangulu_storage_slot_t *b, pangulu_storage_slot_t *diag) { pangulu_platform_0100000_gessm(nb, b, diag, 0); }
void refh_ssssm(int nb, pangulu_storage_slot_t *c, pangulu_storage_slot_t *a, pangulu_storage_slot_t *b)
{
    refh_kernel_scratch_init();
    pangulu_platform_0100000_ssssm(nb, c, a, b, 0);
}

/* solve-phase block kernels of the CPU platform (pangulu_platform_0100000.c:433-506) */
void refh_spmv(int nb, pangulu_storage_slot_t *a, calculate_type *x, calculate_type *y) { pangulu_platform_0100000_spmv(nb, a, x, y); }
void refh_vecadd(long long n, calculate_type *b, calculate_type *x) { pangulu_platform_0100000_vecadd(n, b, x); }
void refh_sptrsv(int nb, pangulu_storage_slot_t *s, calculate_type *x, long long uplo) { pangulu_platform_0100000_sptrsv(nb, s, x, uplo); }

/* CSC -> CSR mirror with value index, the reference's own routine (src/pangulu_utils.c:479-511) */
void refh_transpose_with_valueidx(int nb, pangulu_inblock_ptr *colptr, pangulu_inblock_idx *rowidx,
                                  pangulu_inblock_ptr *rowptr, pangulu_inblock_idx *colidx, pangulu_inblock_ptr *csr2csc)
{
    pangulu_transpose_struct_with_valueidx_inblock(nb, colptr, rowidx, rowptr, colidx, csr2csc);
}
