/* pangulu_platform_0201001.h -- C ABI of the sm_100a (B200) platform module for
 * PanguLU's numeric factorisation.  Platform id 0201001, name GPU_SM100:
 * directory src/platforms/02_NONSHAREDMEM/01_GPU/001_SM100/ in a PanguLU tree
 * (INTEGRATION.md).
 *
 * Part 1 is the drop-in boundary: the 21 functions every PanguLU platform
 * exports (reference build_helper.py:8-32; prototypes of the CUDA platform it
 * replaces: src/platforms/02_NONSHAREDMEM/01_GPU/000_CUDA/pangulu_platform_0201000.h:7-83).
 * Same names (modulo the id), argument meaning and error behaviour
 * (print "[PanguLU ERROR] ..." and exit(EXIT_FAILURE); no return codes).
 *
 * Part 2 are additive entry points (prefix pangulu_sm100_) that PanguLU's
 * scheduler does not know about: bulk factor download, run-time options, a
 * device-resident static task plan, peer-to-peer block transport and the
 * instrumentation bench.py reads.  A PanguLU build that never calls them
 * behaves exactly like one using platform 0201000.
 *
 * The id in the symbol names is a build parameter (PANGULU_SM100_ID, default
 * 0201001).  Building with -DPANGULU_SM100_ID=0201000 yields a library that
 * takes the place of the stock GPU_CUDA object in an UNMODIFIED, un-regenerated
 * PanguLU tree (whose checked-in dispatch only names 0100000 and 0201000); the
 * parity tests use that flavour to run the reference scheduler on top of this
 * module.
 */
#ifndef PANGULU_PLATFORM_SM100_H
#define PANGULU_PLATFORM_SM100_H

#include "pangulu_sm100_abi.h"

#ifndef PANGULU_SM100_ID
#define PANGULU_SM100_ID 0201001
#endif
#define PANGULU_SM100_CAT3(a, b, c) a##b##c
#define PANGULU_SM100_CAT3X(a, b, c) PANGULU_SM100_CAT3(a, b, c)
/* PSM(getrf) -> pangulu_platform_0201001_getrf */
#define PSM(fn) PANGULU_SM100_CAT3X(pangulu_platform_, PANGULU_SM100_ID, _##fn)

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ *
 * Part 1 -- the PanguLU platform interface                            *
 * ------------------------------------------------------------------ */

/* replaces pangulu_platform_0201000_malloc (…0201000.cu:52-56): cudaMalloc */
void PSM(malloc)(void **platform_address, size_t size);
/* …0201000.cu:58-62: page-locked host memory */
void PSM(malloc_pinned)(void **platform_address, size_t size);
/* …0201000.cu:64-68: block until every queued kernel/copy of this rank's device has finished.
 * The kernel interface calls it after every getrf / hybrid_batched (src/pangulu_kernel_interface.c:191,267). */
void PSM(synchronize)(void);
/* …0201000.cu:70-74 */
void PSM(memset)(void *s, int c, size_t n);
/* …0201000.cu:76-80 */
void PSM(create_stream)(void **stream);
/* …0201000.cu:82-103: kind 0 = host->device, 1 = device->host, 2 = device->device; anything else is fatal */
void PSM(memcpy)(void *dst, const void *src, size_t count, unsigned int kind);
/* …0201000.cu:105-126: same on a stream (NULL = legacy default stream, as pangulu_cm_recv_block passes,
 * src/pangulu_communication.c:1850,1880) */
void PSM(memcpy_async)(void *dst, const void *src, size_t count, unsigned int kind, void *stream);
/* …0201000.cu:128-132 */
void PSM(free)(void *devptr);
/* …0201000.cu:134-138 */
void PSM(get_device_num)(int *device_num);
/* …0201000.cu:140-147: cudaSetDevice; also publishes sharedMemPerBlock in the host global
 * pangulu_gpu_shared_mem_size when the PanguLU core is linked in */
void PSM(set_default_device)(int device_num);
/* …0201000.cu:149-155 */
void PSM(get_device_name)(char *name, int device_num);
/* …0201000.cu:157-163: bytes in use on the current device */
void PSM(get_device_memory_usage)(size_t *used_byte);

/* GETRF: in-place LU without pivoting of one diagonal block; opdst is either half, the other one is
 * opdst->related_block.  CPU semantics followed (tiny pivots clamped to PANGULU_TOL):
 * src/platforms/01_SHAREDMEM/00_CPU/000_CPU/pangulu_platform_0100000.c:57-135; replaces …0201000.cu:547-641. */
void PSM(getrf)(pangulu_inblock_idx nb, pangulu_storage_slot_t *opdst, int tid);
/* TSTRF: B <- B * U^-1 for a strictly-lower block, …0100000.c:137-175 / …0201000.cu:643-681 */
void PSM(tstrf)(pangulu_inblock_idx nb, pangulu_storage_slot_t *opdst, pangulu_storage_slot_t *opdiag, int tid);
/* GESSM: B <- L^-1 * B for an upper block, …0100000.c:178-209 / …0201000.cu:683-715 */
void PSM(gessm)(pangulu_inblock_idx nb, pangulu_storage_slot_t *opdst, pangulu_storage_slot_t *opdiag, int tid);
/* SSSSM: C <- C - A*B, both halves when C is a diagonal block, …0100000.c:211-397 / …0201000.cu:717-873 */
void PSM(ssssm)(pangulu_inblock_idx nb, pangulu_storage_slot_t *opdst, pangulu_storage_slot_t *op1, pangulu_storage_slot_t *op2, int tid);
/* batches: one grid over all tasks (the reference loops serially, …0201000.cu:875-909).
 * Same-opdst SSSSM tasks inside one call are legal and are combined per destination. */
void PSM(ssssm_batched)(pangulu_inblock_idx nb, pangulu_uint64_t ntask, pangulu_task_t *tasks);
void PSM(hybrid_batched)(pangulu_inblock_idx nb, pangulu_uint64_t ntask, pangulu_task_t *tasks);

/* solve-phase block kernels on DEVICE arrays (the reference leaves these empty, …0201000.cu:958-979;
 * CPU semantics: …0100000.c:435-506).  x, y, bval, xval are device pointers. */
void PSM(spmv)(pangulu_inblock_idx nb, pangulu_storage_slot_t *a, calculate_type *x, calculate_type *y);
void PSM(vecadd)(pangulu_int64_t length, calculate_type *bval, calculate_type *xval);
void PSM(sptrsv)(pangulu_inblock_idx nb, pangulu_storage_slot_t *s, calculate_type *xval, pangulu_int64_t uplo);

/* ------------------------------------------------------------------ *
 * Part 2 -- additive entry points                                      *
 * ------------------------------------------------------------------ */

/* Run-time options.  Keys:
 *   "eager_download"  1 (default): after GETRF/TSTRF/GESSM the block's values are also copied to the
 *                     host slot->value, as pangulu_cuda_download_block does (…0201000.cu:196-201,639-640,680,714)
 *                     -- but asynchronously; complete after PSM(synchronize).  0: the caller promises to
 *                     call pangulu_sm100_download_blocks() itself (one bulk copy at the end of gstrf).
 *   "dense_threshold" per-mille operand density sqrt(nnzA*nnzB)/nb^2 from which an SSSSM update runs on the dense
 *                     FP64 tensor-core (DMMA) / FP32 FMA tile path (default 120)
 *   "ssssm_mode"      0 auto, 1 force sparse path, 2 force tile path
 *   "getrf_mode"      0 auto, 1 no shared-memory staging, 2 no dense promotion, 3 always dense-promoted (nb <= 256)
 *   "executor"        plans only: 1 (default) persistent megakernel -- one launch per factorisation, device-side
 *                     dependency flags; 0 one launch per kernel class per wave
 *   "compact", "compact_threshold", "compact_block_threshold", "cpanel", "cpanel_block_threshold"
 *                     plans only: which updates / panel solves take the compact dense paths (sm100_compact.cuh); defaults: every
 *                     update, panel blocks of density >= 1 per mille
 *   "spin_timeout_ms" persistent executor: a dependency wait longer than this (default 30000) ends the run with
 *                     "[PanguLU ERROR] ..." and exit, the reference's error convention (…0201000.cu:31-50), instead of
 *                     hanging the GPU (dead peer rank, CTAs not co-resident)
 *   "split_pairs", "tile_variant", "profile", "trace"   tuning / instrumentation, see DESIGN.md
 * returns 0, or -1 for an unknown key. */
int pangulu_sm100_set_option(const char *key, long long value);
long long pangulu_sm100_get_option(const char *key);

/* release memory obtained from <id>_malloc_pinned (the reference interface has no such call and leaks it) */
void pangulu_sm100_free_pinned(void *host_ptr);
/* Independent groups of a diagonal block (index ranges bounds[g] .. bounds[g+1], g < ngroups <= 4, over which the block
 * is block diagonal).  Once registered for the block's LOWER-half slot, a GETRF task factors only the group named in
 * bits 32.. of its compare_flag, updates of the block count towards the groups they touch and TSTRF/GESSM wait only
 * for the groups they read.  ngroups <= 1 removes the entry; lower_half == NULL clears the registry. */
void pangulu_sm100_set_diag_groups(const pangulu_storage_slot_t *lower_half, int ngroups, const unsigned short *bounds);

/* Copy the values of n blocks device -> host (slot->d_value -> slot->value) in one pass. */
void pangulu_sm100_download_blocks(pangulu_inblock_idx nb, pangulu_uint64_t n, pangulu_storage_slot_t **slots);
/* ... and host -> device (re-factorisation with new values on an unchanged pattern). */
void pangulu_sm100_upload_blocks(pangulu_inblock_idx nb, pangulu_uint64_t n, pangulu_storage_slot_t **slots);

/* Preprocessing on the device.
 * scatter_values: clear clear_bytes bytes at d_region (0: leave it), then region[d_dst[i]] = d_src[i] (* d_scale[i] when
 *   d_scale is not NULL) for i < n; entries with d_dst[i] == ~0 are skipped.  This is the value scatter of
 *   src/pangulu_conversion.c:241-350 with the positions precomputed once per pattern; all pointers are device pointers,
 *   values are calculate_type.
 * mirror_patterns: rebuild d_rowpointer / d_columnindex / d_idx_of_csc_value_for_csr of n strictly-lower blocks from their
 *   d_columnpointer / d_rowindex (pangulu_transpose_struct_with_valueidx_inblock, src/pangulu_utils.c:479-511); bit-identical
 *   to the host routine.  Returns after the device work is complete. */
void pangulu_sm100_scatter_values(void *d_region, size_t clear_bytes, const unsigned long long *d_dst, const void *d_src, const void *d_scale,
                                  unsigned long long n);
void pangulu_sm100_mirror_patterns(pangulu_inblock_idx nb, pangulu_uint64_t n, pangulu_storage_slot_t **slots);

/* Static plan: the whole task graph of one rank, translated once into device-resident descriptors.
 * waves[w] .. waves[w+1] index tasks[]; waves execute in order, tasks inside a wave are independent
 * except SSSSM tasks sharing opdst (combined).  plan_execute only enqueues; PSM(synchronize) completes it. */
typedef struct pangulu_sm100_plan pangulu_sm100_plan;
pangulu_sm100_plan *pangulu_sm100_plan_create(pangulu_inblock_idx nb, pangulu_uint64_t nwave,
                                              const pangulu_uint64_t *waves, const pangulu_task_t *tasks);
void pangulu_sm100_plan_execute(pangulu_sm100_plan *plan);
/* Execute the plan and copy the values of the n given blocks to their host arrays (slot->d_value -> slot->value, pinned) as
 * the blocks become final, overlapped with the rest of the factorisation; returns when kernel and copies are complete.
 * (The reference downloads every block after the call that finishes it, pangulu_platform_0201000.cu:196-201.) */
void pangulu_sm100_plan_execute_download(pangulu_sm100_plan *plan, pangulu_uint64_t n, pangulu_storage_slot_t **slots);
/* execute waves [w0, w1) only (multi-GPU driver interleaves block exchange between waves) */
void pangulu_sm100_plan_execute_range(pangulu_sm100_plan *plan, pangulu_uint64_t w0, pangulu_uint64_t w1);
void pangulu_sm100_plan_destroy(pangulu_sm100_plan *plan);
/* persistent executor with option "trace" = 1 at plan creation: per task {kind, fetch_ns, start_ns, end_ns}
 * (4 u64 per task) of the last execution; returns the number of tasks (call with out = NULL to size the buffer) */
long long pangulu_sm100_plan_trace(pangulu_sm100_plan *plan, unsigned long long *out, long long cap);
/* per task, in the order of plan_trace: FP64 tensor (DMMA.8x8x4) instructions the task issues (compact tile tasks, else 0) */
long long pangulu_sm100_plan_task_work(pangulu_sm100_plan *plan, float *out, long long cap);

/* pangulu_gstrs on device-resident factors: the block forward / backward substitution of
 * src/pangulu_sptrsv.c:24-174 with <id>_spmv / <id>_sptrsv semantics, one launch per block row and phase.  slots = every
 * block the rank holds (off-diagonal blocks and both halves of the diagonal blocks); x_host has block_length*nb values
 * (already permuted / scaled) and is overwritten with the solution.  Factors never return to the host. */
typedef struct pangulu_sm100_solve_plan pangulu_sm100_solve_plan;
pangulu_sm100_solve_plan *pangulu_sm100_solve_plan_create(pangulu_inblock_idx nb, pangulu_uint64_t block_length, pangulu_uint64_t nslot,
                                                          pangulu_storage_slot_t **slots);
void pangulu_sm100_solve_plan_execute(pangulu_sm100_solve_plan *plan, calculate_type *x_host);
void pangulu_sm100_solve_plan_destroy(pangulu_sm100_solve_plan *plan);

/* Multi-GPU wiring of a persistent plan (one rank per GPU).  A block finalised by its owner is written by the
 * producing kernel itself into the mirror copy every consumer rank keeps (peer-mapped pointers over NVLink /
 * NVSwitch, obtained with CUDA IPC), followed by a system-scope release of the consumer's "done" flag -- this
 * replaces pangulu_cm_isend_block / pangulu_cm_recv_block (src/pangulu_communication.c:1902-1944, 1786-1900):
 * no host staging, no MPI, no probe loop.
 *   sync_base / flag_offset   where the plan keeps the done flag of the block whose device value array is d_value
 *                             (the allocation is exported with pangulu_sm100_ipc_export)
 *   expect_remote             the block is produced by a peer: its flag starts cleared
 *   add_push                  after finalising the local block d_value_local, copy its values to peer_value and
 *                             raise *peer_flag (both pointers are peer-mapped)
 *   commit                    upload the push lists;  reset / launch = the two halves of plan_execute (all ranks
 *                             must have reset before any rank launches: put a host barrier in between) */
void *pangulu_sm100_plan_sync_base(pangulu_sm100_plan *plan, size_t *bytes);
long long pangulu_sm100_plan_flag_offset(pangulu_sm100_plan *plan, const void *d_value);
void pangulu_sm100_plan_expect_remote(pangulu_sm100_plan *plan, const void *d_value);
void pangulu_sm100_plan_add_push(pangulu_sm100_plan *plan, const void *d_value_local, void *peer_value, void *peer_flag);
/* compact dense operand copies (sm100_compact.cuh) across ranks: where the plan keeps them (one allocation, exported
 * with pangulu_sm100_ipc_export), the byte offset of the copy of the block whose device value array is d_value (-1:
 * none), and a push that also writes the owner's copy into the consumer's (peer_dense, peer-mapped; may be NULL) */
/* work statistics of a plan: "ct_tasks", "ct_pairtiles", "ct_chunks", "ct_dmma" (DMMA instructions of the compact tile
 * path), "ct_bytes" (operand bytes it streams), "dense_bytes", "dense_blocks" */
double pangulu_sm100_plan_stat(pangulu_sm100_plan *plan, const char *key);
void *pangulu_sm100_plan_dense_base(pangulu_sm100_plan *plan, size_t *bytes);
long long pangulu_sm100_plan_dense_offset(pangulu_sm100_plan *plan, const void *d_value);
void pangulu_sm100_plan_add_push2(pangulu_sm100_plan *plan, const void *d_value_local, void *peer_value, void *peer_flag, void *peer_dense);
/* the same per independent group of a diagonal half (pangulu_sm100_set_diag_groups): flag of group `group`, and a push
 * of the entries [off, off+cnt) of the half's value array when that group's GETRF task is done */
long long pangulu_sm100_plan_flag_offset2(pangulu_sm100_plan *plan, const void *d_value, int group);
void pangulu_sm100_plan_expect_remote2(pangulu_sm100_plan *plan, const void *d_value, int group);
void pangulu_sm100_plan_add_push3(pangulu_sm100_plan *plan, const void *d_value_local, int group, unsigned int off, unsigned int cnt,
                                  void *peer_value, void *peer_flag);
void pangulu_sm100_plan_commit(pangulu_sm100_plan *plan);
void pangulu_sm100_plan_reset(pangulu_sm100_plan *plan);
void pangulu_sm100_plan_launch(pangulu_sm100_plan *plan);
/* CUDA IPC: 64-byte handle of a device allocation of this module / map a peer's allocation into this process */
int pangulu_sm100_ipc_export(void *devptr, void *handle64);
int pangulu_sm100_ipc_open(const void *handle64, void **devptr);
int pangulu_sm100_ipc_close(void *devptr);

/* Instrumentation (bench.py): kernels launched / device time per kernel class since the last reset.
 * cls: 0 getrf, 1 tstrf, 2 gessm, 3 ssssm-sparse, 4 ssssm-tile, 5 copy/aux.  Timing is off unless
 * option "profile" is 1 (it brackets every launch with CUDA events on the module's stream). */
void pangulu_sm100_stats_reset(void);
long long pangulu_sm100_stats_launches(int cls);
double pangulu_sm100_stats_ms(int cls);
/* the CUDA stream all kernels of this module are enqueued on (cudaStream_t) */
void *pangulu_sm100_stream(void);

/* library identity: "pangulu-sm100 <type> id=<id> sm_100a" */
const char *pangulu_sm100_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PANGULU_PLATFORM_SM100_H */
