// host_internal.h -- internals of the host driver that sits above the sm_100a platform C ABI.
//
// The driver mirrors the reference's public API (include/pangulu.h:11-15: pangulu_init / pangulu_gstrf /
// pangulu_gstrs / pangulu_gssv / pangulu_finalize) for the numeric-factorisation path so that the module
// can be exercised and timed without a PanguLU checkout: it reorders (METIS when present), runs the
// symmetric-pattern symbolic factorisation, cuts the result into nb x nb blocks in PanguLU's packed
// block layout, builds the static task schedule and drives the platform entry points.
// It calls the platform ONLY through include/pangulu_platform_0201001.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "pangulu_platform_0201001.h"

namespace pdrv {

typedef calculate_type T;
typedef uint64_t eptr;  // pangulu_exblock_ptr
typedef uint32_t eidx;  // pangulu_exblock_idx
typedef uint32_t iptr;  // pangulu_inblock_ptr
typedef uint16_t iidx;  // pangulu_inblock_idx

// ---- host value arithmetic (solve phase, value scatter, residuals) ----
#ifdef PANGULU_COMPLEX
inline T t_zero() { T z; z.re = 0; z.im = 0; return z; }
inline T t_from_real(double r) { T z; z.re = (calculate_real_type)r; z.im = 0; return z; }
inline T t_mul(T a, T b) { T r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }
inline T t_sub(T a, T b) { T r; r.re = a.re - b.re; r.im = a.im - b.im; return r; }
inline T t_add(T a, T b) { T r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }
inline T t_div(T a, T b)
{
    double d = (double)b.re * b.re + (double)b.im * b.im;
    T r;
    r.re = (calculate_real_type)(((double)a.re * b.re + (double)a.im * b.im) / d);
    r.im = (calculate_real_type)(((double)a.im * b.re - (double)a.re * b.im) / d);
    return r;
}
inline double t_real(T a) { return a.re; }
inline double t_abs2(T a) { return (double)a.re * a.re + (double)a.im * a.im; }
#else
inline T t_zero() { return (T)0; }
inline T t_from_real(double r) { return (T)r; }
inline T t_mul(T a, T b) { return a * b; }
inline T t_sub(T a, T b) { return a - b; }
inline T t_add(T a, T b) { return a + b; }
inline T t_div(T a, T b) { return a / b; }
inline double t_real(T a) { return (double)a; }
inline double t_abs2(T a) { return (double)a * (double)a; }
#endif

struct Csc
{
    eidx n = 0;
    std::vector<eptr> ptr;
    std::vector<eidx> idx;
    std::vector<T> val;
};

// ---- reordering (reference: src/pangulu_reordering.c) ----
enum OrderingKind { ORDER_NATURAL = 0, ORDER_METIS = 1 };
struct Reordering
{
    // names follow pangulu_block_smatrix (src/pangulu_common.h:284-288)
    std::vector<eidx> row_perm, col_perm;  // MC64 (identity until the MC64 row is built)
    std::vector<eidx> metis_perm;          // new index of old index (METIS iperm)
    std::vector<T> row_scale, col_scale;
};
bool metis_available();
// A (sorted CSC with full or partial diagonal) -> reordered, sorted CSC; fills R.
void reorder_matrix(const Csc& A, int ordering, Reordering& R, Csc& out);

// ---- symbolic (reference: src/pangulu_symbolic.c) ----
// pattern of L (lower incl. diagonal, by column, ascending rows) of the symmetrised matrix
struct Symbolic
{
    std::vector<eptr> ptr;
    std::vector<eidx> idx;
    eptr symbolic_nnz = 0;  // 2*nnz(L) - n, src/pangulu_symbolic.c:242
};
void symbolic_factor(const Csc& A, Symbolic& S);

// ---- block storage in PanguLU's packed layout (reference: src/pangulu_storage.c, Appendix A) ----
struct BlockStore
{
    int nb = 0;
    eidx n = 0;
    eidx block_length = 0;
    int rank = 0, nproc = 1, p = 1, q = 1;
    // block-level CSC of all OFF-DIAGONAL blocks of the whole matrix (both triangles)
    std::vector<eptr> bcol_ptr;
    std::vector<eidx> bcol_row;
    std::vector<uint32_t> bnnz;           // nnz per off-diagonal block (global block id)
    std::vector<uint32_t> diag_nnz_lower, diag_nnz_upper;
    // slots of the blocks this rank holds: [0, n_offdiag) in block-CSC order of the OWNED off-diagonal
    // blocks, then (lower, upper) pairs of the owned diagonal blocks -- the order of bins[0]
    // (src/pangulu_storage.c:239-293) -- then slots of remote blocks mirrored on this rank (multi-GPU)
    std::vector<pangulu_storage_slot_t> slots;
    std::vector<int64_t> slot_of_block;   // global off-diagonal block id -> slot index or -1
    std::vector<int64_t> slot_of_diag;    // block row -> slot index of the LOWER half or -1
    size_t n_own_offdiag = 0, n_own_diag = 0;
    char* h_buf = nullptr;                // pinned; byte-identical to the device buffer
    char* d_buf = nullptr;
    size_t buf_bytes = 0;                 // value region + index region
    size_t value_bytes = 0;               // [0, value_bytes): value arrays of all kept blocks
    size_t own_value_bytes = 0;           // prefix of the value region holding OWNED blocks
    uint64_t nnz_total = 0;
    bool host_only = false;               // structure-only mode (CPU tests): no device memory

    int owner(eidx brow, eidx bcol) const { return (int)((brow % p) * q + (bcol % q)); }
    int64_t find_block(eidx brow, eidx bcol) const;  // global id of an off-diagonal block, -1 if absent
};
void build_blocks(const Csc& A, const Symbolic& S, int nb, int rank, int nproc, BlockStore& B);
void free_blocks(BlockStore& B);

// ---- static schedule ----
#define DIAG_MAX_GROUPS 4
struct Schedule
{
    std::vector<pangulu_task_t> tasks;  // sorted by wave
    std::vector<uint64_t> wave_ptr;     // nwave + 1
    std::vector<int32_t> depth_offdiag; // per global off-diagonal block
    std::vector<int32_t> depth_diag;    // per block row and group: [k * DIAG_MAX_GROUPS + g]
    // independent parts of the diagonal blocks (schedule.cpp: diag_groups)
    int split_diag = 1;
    std::vector<uint16_t> group_bounds; // [k * (DIAG_MAX_GROUPS+1) + g] .. [+ g + 1): index range of group g of diagonal block k
    std::vector<uint8_t> group_count;   // per block row
    std::vector<uint8_t> colgroups;     // per global off-diagonal block (I,J): groups of diagonal block J its columns touch
    std::vector<uint8_t> rowgroups;     //                                       groups of diagonal block I its rows touch
    uint64_t n_getrf = 0, n_tstrf = 0, n_gessm = 0, n_ssssm = 0;
    int policy = 2;                     // 0 lazy (updates in the destination's wave), 1 eager (as soon as operands are final), 2 windowed
    int lookahead = 2;                  // policy 2: waves before the destination's finalisation at which slack updates run
};
// independent groups of the diagonal blocks, from the element-level symbolic pattern (call before build_schedule)
void diag_groups(const BlockStore& B, const Symbolic& Sy, Schedule& S);
void build_schedule(BlockStore& B, Schedule& S);

// ---- flop / byte models (reference: src/pangulu_kernel_interface.c:4-176) ----
struct WorkModel
{
    double flop_getrf = 0, flop_tstrf = 0, flop_gessm = 0, flop_ssssm = 0;
    double bytes_getrf = 0, bytes_tstrf = 0, bytes_gessm = 0, bytes_ssssm = 0;
    double flop_total() const { return flop_getrf + flop_tstrf + flop_gessm + flop_ssssm; }
};
void work_model(const BlockStore& B, const Schedule& S, WorkModel& W);

// ---- solve (reference: src/pangulu_sptrsv.c, host arithmetic as in the reference) ----
void solve_host(const BlockStore& B, std::vector<T>& x);

double wall_seconds();

}  // namespace pdrv
