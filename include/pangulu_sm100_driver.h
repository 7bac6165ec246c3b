/* pangulu_sm100_driver.h -- the host driver that ships with the sm_100a platform module.
 *
 * It exports PanguLU's public API for the numeric-factorisation path with the reference's own
 * prototypes (reference include/pangulu.h:11-15, include/pangulu_interface_common.h:3-20), so a program
 * written against PanguLU (e.g. its examples/example.c) runs against this library unchanged:
 *
 *     pangulu_init  -> reorder (METIS_NodeND when available) + symbolic + block storage + static schedule
 *     pangulu_gstrf -> numeric LU on the GPU through the platform C ABI (pangulu_platform_0201001.h)
 *     pangulu_gstrs -> block forward/backward substitution on the device-resident L and U (one rank)
 *
 * Out of scope here (kept as the reference's own code when the module is used inside PanguLU): MC64,
 * MPI distribution, the dynamic scheduler.  See DESIGN.md.
 */
#ifndef PANGULU_SM100_DRIVER_H
#define PANGULU_SM100_DRIVER_H

#include "pangulu_platform_0201001.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/pangulu_common.h:67-70 */
typedef pangulu_exblock_ptr sparse_pointer_t;
typedef pangulu_exblock_idx sparse_index_t;
typedef calculate_type sparse_value_t;

/* reference include/pangulu_interface_common.h:3-20 */
typedef struct pangulu_init_options
{
    int nthread;
    int nb;
    int gpu_kernel_warp_per_block;
    int gpu_data_move_warp_per_block;
    int sizeof_value;
    int is_complex_matrix;
    float mpi_recv_buffer_level;
} pangulu_init_options;
typedef struct pangulu_gstrf_options { char unused_; } pangulu_gstrf_options;
typedef struct pangulu_gstrs_options { char unused_; } pangulu_gstrs_options;

/* reference include/pangulu.h:11-15 */
void pangulu_init(sparse_index_t pangulu_n, sparse_pointer_t pangulu_nnz, sparse_pointer_t *csc_colptr, sparse_index_t *csc_rowidx,
                  sparse_value_t *csc_value, pangulu_init_options *init_options, void **pangulu_handle);
void pangulu_gstrf(pangulu_gstrf_options *gstrf_options, void **pangulu_handle);
void pangulu_gstrs(sparse_value_t *rhs, pangulu_gstrs_options *gstrs_options, void **pangulu_handle);
void pangulu_gssv(sparse_value_t *rhs, pangulu_gstrf_options *gstrf_options, pangulu_gstrs_options *gstrs_options, void **pangulu_handle);
void pangulu_finalize(void **pangulu_handle);

/* Driver options, set BEFORE pangulu_init; a handle keeps the values it was created with (the reference takes
 * rank/size from MPI_COMM_WORLD):
 *   "rank", "nproc"   position in the p x q rank grid (one rank per GPU)
 *   "ordering"        0 natural (input already permuted), 1 METIS_NodeND, -1 (default) METIS when built in
 *   "split_diag"      1 (default): independent groups of diagonal blocks get their own GETRF task and dependency
 *                     counters (single rank), 0: one GETRF per diagonal block
 *   "mode"            0 reference-facing gstrf (host buffers, per-wave hybrid_batched), 1 resident plan,
 *                     2 host buffers + resident plan (upload values, run the plan, download factors)
 *   "schedule"        wave of an SSSSM update: 0 lazy (the wave that finalises its destination), 1 eager (first wave
 *                     after both operands are final), 2 (default) windowed: max(eager, destination's wave - "lookahead")
 *   "lookahead"       window of policy 2 in waves (default 2)
 *   "device_solve"    1 (default): pangulu_gstrs runs the block substitution on the device-resident factors (only the
 *                     right-hand side travels); 0: on the host copies of the factors, as the reference does
 *   "verbose"         1 prints the reference's info lines
 *   "host_only"       1 builds ordering/symbolic/blocks/schedule in host memory only (CPU-side tests of the
 *                     host logic); pangulu_gstrf on such a handle is a fatal error -- there is no CPU path
 * returns 0, -1 for an unknown key. */
int pangulu_sm100_driver_config(const char *key, long long value);

/* numbers about an initialised handle: "n","nb","block_length","symbolic_nnz","nnz_blocks","num_slots",
 * "num_own_offdiag","num_own_diag","nwave","ntask","n_getrf","n_tstrf","n_gessm","n_ssssm","flop",
 * "flop_<class>","bytes_<class>","value_bytes","own_value_bytes","buffer_bytes","p","q","time_<phase>" */
double pangulu_sm100_driver_info(void *handle, const char *key);
/* i-th block slot this rank holds (order of the reference's bins[0], then mirrored remote blocks) */
pangulu_storage_slot_t *pangulu_sm100_driver_slot(void *handle, long long i);
/* 0 row_perm, 1 col_perm, 2 metis_perm (u32[n]); 3 row_scale, 4 col_scale (T[n]) */
const void *pangulu_sm100_driver_perm(void *handle, int which);
/* the static schedule: tasks sorted by wave, wave_ptr has nwave+1 entries */
const pangulu_task_t *pangulu_sm100_driver_tasks(void *handle, const pangulu_uint64_t **wave_ptr);
/* switch an initialised handle between the gstrf modes (0 reference-facing, 1 resident plan, 2 host buffers + plan) */
void pangulu_sm100_driver_set_mode(void *handle, int mode);
/* put the values of A back into the host block buffers (and the device copy) so gstrf can run again */
void pangulu_sm100_driver_reset_values(void *handle, int also_device);
/* Preprocessing on the device.
 * update_values: the matrix again in the CSC arrays given to pangulu_init -- same pattern, new values; nnz(A) values travel to
 *   the device, the block value region is cleared there and the values are scattered by a kernel (first call: the positions are
 *   computed once).  The host copy of the reordered matrix follows; the host block buffers do not.  Returns seconds.
 * device_mirrors: rebuild the CSR mirrors (rowpointer / columnindex / idx_of_csc_value_for_csr) of the owned strictly-lower
 *   blocks on the device; returns the number of bytes in which the device index region differs from the host-built one. */
double pangulu_sm100_driver_update_values(void *handle, const sparse_pointer_t *colptr, const sparse_index_t *rowidx, const sparse_value_t *val);
long long pangulu_sm100_driver_device_mirrors(void *handle, int clear_first);
/* resident mode: enqueue the whole factorisation on the module's stream (pangulu_sm100_stream()) without
 * synchronising; <id>_synchronize() or a stream event completes it */
void pangulu_sm100_driver_gstrf_enqueue(void *handle);
/* Multi-GPU (one rank per GPU, 2D block-cyclic ownership as src/pangulu_common.h:135).  Every rank builds its own
 * handle with "rank"/"nproc" set, then the ranks exchange ipc_handles + mirror_table (any host channel, e.g.
 * torch.distributed.all_gather_object), call connect_peer once per peer and commit.  pangulu_gstrf then needs a
 * host barrier (set_barrier).  Finished blocks travel GPU-to-GPU inside the factorisation kernel. */
void pangulu_sm100_driver_set_barrier(void (*fn)(void));
void pangulu_sm100_driver_ipc_handles(void *handle, unsigned char *out192);
long long pangulu_sm100_driver_mirror_table(void *handle, long long *out, long long cap);
long long pangulu_sm100_driver_connect_peer(void *handle, const unsigned char *handles192, const long long *table, long long n);
void pangulu_sm100_driver_commit(void *handle);
/* close the peer mappings of an earlier round of connect_peer calls (reconnect after rebuild_plan; finalize does it too) */
void pangulu_sm100_driver_disconnect(void *handle);
void pangulu_sm100_driver_gstrf_reset(void *handle);
void pangulu_sm100_driver_gstrf_launch(void *handle);
/* independent groups of diagonal block k: bounds[0..ng] of the index ranges, returns ng (1 = not split) */
int pangulu_sm100_driver_diag_groups(void *handle, long long k, unsigned short *bounds5);
/* drop and rebuild the device-resident plan with the platform options in force now (e.g. "trace") */
void pangulu_sm100_driver_rebuild_plan(void *handle);
/* task trace of the last resident gstrf (see pangulu_sm100_plan_trace) */
long long pangulu_sm100_driver_trace(void *handle, unsigned long long *out, long long cap);
/* per traced task: DMMA instructions it issues (see pangulu_sm100_plan_task_work) */
long long pangulu_sm100_driver_task_work(void *handle, float *out, long long cap);
/* device -> host copy of the owned factors (resident mode) */
void pangulu_sm100_driver_download(void *handle);
/* the reference's gstrf self-check ||L(U1) - A1|| / ||A1|| (src/pangulu_numeric.c:1082-1341) on host factors */
double pangulu_sm100_driver_factor_check(void *handle);
/* The same check with the factors spread over several ranks, in the three passes the reference makes
 * (src/pangulu_numeric.c:1082-1341: every rank multiplies the blocks it owns, the partial vectors are summed over the
 * ranks).  All vectors have block_length*nb doubles (real part for the complex types):
 *   check_ux  : ux  += U_owned * 1                     -> sum over ranks
 *   check_lux : lux += L_owned * ux (+ ux on rank 0: the unit diagonal)  -> sum over ranks
 *   check_ax  : ax = A_reordered * 1 (every rank holds A), returns n
 * residual = ||lux - ax||_2 / ||ax||_2 over the first n entries. */
void pangulu_sm100_driver_check_ux(void *handle, double *ux);
void pangulu_sm100_driver_check_lux(void *handle, const double *ux, double *lux);
long long pangulu_sm100_driver_check_ax(void *handle, double *ax);

/* Factor files: the owned blocks in PanguLU's own packed block format (32-byte header, value / columnpointer / rowindex and,
 * for strictly-lower blocks, the CSR mirror: src/pangulu_storage.c:295-357, src/pangulu_communication.c:1929-1942) behind a
 * header with the permutations.  save returns the bytes written (-1: cannot open); load puts the values into a handle
 * initialised on the same matrix (patterns and permutation are compared; a mismatch is a fatal error) and returns the number
 * of blocks read: pangulu_init + load_factors + pangulu_gstrs solves without factoring again. */
long long pangulu_sm100_driver_save_factors(void *handle, const char *path);
long long pangulu_sm100_driver_load_factors(void *handle, const char *path);

/* The fast path behind PanguLU's own pangulu_gstrf (one rank): called INSTEAD of pangulu_numeric (src/pangulu.c:236-238)
 * with the block slots PanguLU's storage layer built (storage->bins[0]: all off-diagonal blocks in block-CSC order, then a
 * (lower, upper) pair per diagonal block, src/pangulu_storage.c:239-293).  Builds the static schedule and the device plan
 * on the first call (cached per slot array), runs the persistent kernel on PanguLU's device arrays and refreshes PanguLU's
 * host value arrays.  See INTEGRATION.md for the 12-line overlay that makes the call. */
void pangulu_sm100_reference_gstrf(int nb, unsigned int n, unsigned int block_length, pangulu_storage_slot_t *slots, long long slot_count,
                                   long long nondiag_slot_cnt);
/* seconds: 0 schedule (first call), 1 plan creation (first call), 2 factorisation on the device (last call), 3 download */
double pangulu_sm100_reference_gstrf_time(pangulu_storage_slot_t *slots, int which);
void pangulu_sm100_reference_release(pangulu_storage_slot_t *slots);

#ifdef __cplusplus
}
#endif
#endif
