// driver.cpp -- host driver above the sm_100a platform C ABI.
//
// Exports the reference's public API for this path with the reference's signatures and error behaviour
// (include/pangulu.h:11-15, src/pangulu.c:11-345):
//     pangulu_init / pangulu_gstrf / pangulu_gstrs / pangulu_gssv / pangulu_finalize
// plus pangulu_sm100_driver_* accessors used by the tests and bench.py.
//
// gstrf has two modes (driver option "mode"):
//   0  reference-facing: factor values start in HOST block buffers, are uploaded, every wave of the static
//      schedule is handed to <id>_hybrid_batched() as a HOST pangulu_task_t array (the call PanguLU's
//      scheduler makes, src/pangulu_kernel_interface.c:256-290), and the finished factors are downloaded
//      back to the host buffers -- all inside the call (what bench.py reports as "e2e");
//   1  resident: the schedule was translated once into a device-resident plan; gstrf only enqueues the
//      plan on values already in HBM and synchronises (bench.py's "value").
//   2  host buffers + resident plan: upload the value region, run the plan, download the factors (bench.py's "e2e").
#include <algorithm>
#include <complex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>

#include "host_internal.h"
#include "pangulu_sm100_driver.h"

using namespace pdrv;

namespace {

struct DriverConfig
{
    long long rank = 0, nproc = 1;
    long long ordering = -1;  // -1: METIS when built in, else natural
    long long mode = 0;
    long long verbose = 0;
    long long host_only = 0;
    long long schedule = 2;
    long long lookahead = 2;
    long long split_diag = 1;  // schedule independent groups of diagonal blocks separately (single rank)
    long long device_solve = 1;  // pangulu_gstrs: block substitution on the device-resident factors (0: on the host copies)
};
DriverConfig g_cfg;
void (*g_barrier)(void) = nullptr;  // host barrier across ranks (multi-GPU), installed by the launcher

struct Handle
{
    int nb = 0;
    eidx n = 0;
    Csc A_reordered;  // kept for value reset and residual checks
    Reordering R;
    Symbolic S;
    BlockStore B;
    Schedule sched;
    WorkModel work;
    pangulu_sm100_plan* plan = nullptr;
    pangulu_sm100_solve_plan* solve_plan = nullptr;
    bool device_solve = true;
    int mode = 0;
    int rank = 0, nproc = 1;  // fixed at pangulu_init (later driver_config calls configure FUTURE handles only)
    bool factored = false;
    // multi-GPU wiring: host barrier across the ranks, peer allocations mapped with CUDA IPC (closed in finalize /
    // disconnect), and whether the current plan has been wired to the peers
    void (*barrier)(void) = nullptr;
    std::vector<void*> ipc_opened;
    bool wired = false;
    std::map<std::string, double> times;
    // preprocessing on the device (pangulu_sm100_driver_update_values): per entry of the caller's CSC arrays its position in the
    // value region (device, u64; ~0 = held by another rank) and in A_reordered (host), built on the first call; staging for the values
    unsigned long long* d_scatter_dst = nullptr;
    T* d_scatter_src = nullptr;
    std::vector<uint64_t> pos_in_reordered;
    uint64_t scatter_nnz = 0, scatter_extra = 0;
};

void scatter_values(Handle* h);

}  // namespace

extern "C" {

int pangulu_sm100_driver_config(const char* key, long long value)
{
    std::string k(key ? key : "");
    if (k == "rank") g_cfg.rank = value;
    else if (k == "nproc") g_cfg.nproc = value;
    else if (k == "ordering") g_cfg.ordering = value;
    else if (k == "mode") g_cfg.mode = value;
    else if (k == "verbose") g_cfg.verbose = value;
    else if (k == "host_only") g_cfg.host_only = value;
    else if (k == "schedule") g_cfg.schedule = value;
    else if (k == "lookahead") g_cfg.lookahead = value;
    else if (k == "split_diag") g_cfg.split_diag = value;
    else if (k == "device_solve") g_cfg.device_solve = value;
    else return -1;
    return 0;
}

void pangulu_init(sparse_index_t pangulu_n, sparse_pointer_t pangulu_nnz, sparse_pointer_t* csc_colptr, sparse_index_t* csc_rowidx,
                  sparse_value_t* csc_value, pangulu_init_options* init_options, void** pangulu_handle)
{
    // option checks of src/pangulu.c:28-56
    if (init_options == NULL)
    {
        printf("[PanguLU Error] initialisation options pointer is NULL. (pangulu_init)\n");
        exit(1);
    }
    if ((!init_options->is_complex_matrix) != (sizeof(calculate_type) == sizeof(calculate_real_type)))
    {
        printf("[PanguLU Error] Input matrix type (complex/real) does not match PanguLU configuration. Exiting.\n");
        exit(1);
    }
    if (init_options->sizeof_value != (int)sizeof(calculate_type))
    {
        printf("[PanguLU Error] Input element size is %lld B, but expected %lld B in PanguLU. Exiting.\n", (long long)init_options->sizeof_value,
               (long long)sizeof(calculate_type));
        exit(1);
    }
    if (init_options->nb == 0)
    {
        printf("[PanguLU Error] Block size 'nb' is zero.\n");
        exit(1);
    }
    if (pangulu_n == 0)
    {
        printf("[PanguLU Error] Input matrix is invalid: order is zero.\n");
        exit(1);
    }
    Handle* h = new Handle();
    h->nb = init_options->nb <= 0 ? 256 : init_options->nb;  // src/pangulu.c:52-56
    if (h->nb > 65535)
    {
        printf("[PanguLU Error] nb exceeds the 16-bit in-block index range.\n");
        exit(1);
    }
    h->n = pangulu_n;
    h->mode = (int)g_cfg.mode;
    h->device_solve = g_cfg.device_solve != 0;
    h->rank = (int)g_cfg.rank;
    h->nproc = (int)g_cfg.nproc;
    // several ranks: a peer may read any owned panel block through its compact dense copy
    if (!g_cfg.host_only) pangulu_sm100_set_option("dense_all_panels", h->nproc > 1 ? 1 : 0);

    // one rank <-> one GPU: rank % device_count (src/pangulu.c:126-130)
    h->B.host_only = g_cfg.host_only != 0;
    if (!h->B.host_only)
    {
        int ndev = 0;
        PSM(get_device_num)(&ndev);
        if (ndev <= 0)
        {
            fprintf(stderr, "[PanguLU ERROR] no CUDA device: the sm_100a platform has no CPU path\n");
            exit(EXIT_FAILURE);
        }
        PSM(set_default_device)((int)(g_cfg.rank % ndev));
    }

    Csc A;
    A.n = pangulu_n;
    A.ptr.assign(csc_colptr, csc_colptr + pangulu_n + 1);
    A.idx.assign(csc_rowidx, csc_rowidx + pangulu_nnz);
    A.val.assign((T*)csc_value, (T*)csc_value + pangulu_nnz);

    double t0 = wall_seconds();
    int ordering = (int)g_cfg.ordering;
    if (ordering < 0) ordering = metis_available() ? ORDER_METIS : ORDER_NATURAL;
    reorder_matrix(A, ordering, h->R, h->A_reordered);
    h->times["reorder"] = wall_seconds() - t0;
    t0 = wall_seconds();
    symbolic_factor(h->A_reordered, h->S);
    h->times["symbolic"] = wall_seconds() - t0;
    if (g_cfg.verbose) printf("[PanguLU Info] Symbolic nonzero count is %llu.\n", (unsigned long long)h->S.symbolic_nnz);
    t0 = wall_seconds();
    build_blocks(h->A_reordered, h->S, h->nb, (int)g_cfg.rank, (int)g_cfg.nproc, h->B);
    h->times["blocks"] = wall_seconds() - t0;
    t0 = wall_seconds();
    h->sched.policy = (int)g_cfg.schedule;
    h->sched.lookahead = (int)g_cfg.lookahead;
    h->sched.split_diag = (int)g_cfg.split_diag;
    diag_groups(h->B, h->S, h->sched);
    // the element-level pattern is not needed any more
    std::vector<eidx>().swap(h->S.idx);
    build_schedule(h->B, h->sched);
    if (!h->B.host_only)
        for (eidx k = 0; k < h->B.block_length; ++k)
            if (h->sched.group_count[k] > 1 && h->B.slot_of_diag[k] >= 0)
                pangulu_sm100_set_diag_groups(&h->B.slots[h->B.slot_of_diag[k]], h->sched.group_count[k],
                                              &h->sched.group_bounds[(size_t)k * (DIAG_MAX_GROUPS + 1)]);
    h->times["schedule"] = wall_seconds() - t0;
    t0 = wall_seconds();
    work_model(h->B, h->sched, h->work);
    h->times["work_model"] = wall_seconds() - t0;
    if (h->mode == 1 && !h->B.host_only)
    {
        t0 = wall_seconds();
        h->plan = pangulu_sm100_plan_create((pangulu_inblock_idx)h->nb, h->sched.wave_ptr.size() - 1, (const pangulu_uint64_t*)h->sched.wave_ptr.data(),
                                            h->sched.tasks.data());
        h->times["plan"] = wall_seconds() - t0;
    }
    *pangulu_handle = h;
}

void pangulu_gstrf(pangulu_gstrf_options* gstrf_options, void** pangulu_handle)
{
    if (gstrf_options == NULL)
    {
        printf("[PanguLU Error] Options pointer is NULL. (pangulu_gstrf)\n");
        exit(1);
    }
    Handle* h = (Handle*)*pangulu_handle;
    BlockStore& B = h->B;
    if (B.host_only)
    {
        fprintf(stderr, "[PanguLU ERROR] handle was created in host_only (structure) mode; numeric factorisation needs the GPU -- there is no CPU path\n");
        exit(EXIT_FAILURE);
    }
    if (h->nproc > 1) h->mode = 1;  // (the mode of a handle can be switched with pangulu_sm100_driver_set_mode)
    if (h->mode == 1 && !h->plan)
    {
        const double tp = wall_seconds();
        h->plan = pangulu_sm100_plan_create((pangulu_inblock_idx)h->nb, h->sched.wave_ptr.size() - 1, (const pangulu_uint64_t*)h->sched.wave_ptr.data(),
                                            h->sched.tasks.data());
        h->times["plan"] = wall_seconds() - tp;
    }
    const double t0 = wall_seconds();
    if (h->nproc > 1)
    {
        // one rank per GPU: the plan must have been wired to the peers (pangulu_sm100_driver_connect_peer/commit);
        // every rank clears its flags, all ranks meet, then the persistent kernels run and push blocks to each other
        if (!h->barrier || !h->plan || !h->wired)
        {
            fprintf(stderr, "[PanguLU ERROR] multi-rank gstrf needs a plan wired to the peers (pangulu_sm100_driver_connect_peer + commit; again after "
                            "rebuild_plan) and a host barrier (pangulu_sm100_driver_set_barrier)\n");
            exit(EXIT_FAILURE);
        }
        pangulu_sm100_plan_reset(h->plan);
        PSM(synchronize)();
        h->barrier();
        pangulu_sm100_plan_launch(h->plan);
        PSM(synchronize)();
        h->barrier();
    }
    else if (h->mode == 1)
    {
        pangulu_sm100_plan_execute(h->plan);
        PSM(synchronize)();
    }
    else if (h->mode == 2)
    {
        // host buffers in, host buffers out, resident plan in between: what a PanguLU user sees from pangulu_gstrf
        if (!h->plan)
            h->plan = pangulu_sm100_plan_create((pangulu_inblock_idx)h->nb, h->sched.wave_ptr.size() - 1, (const pangulu_uint64_t*)h->sched.wave_ptr.data(),
                                                h->sched.tasks.data());
        PSM(memcpy)(B.d_buf, B.h_buf, B.value_bytes, 0);
        // the finished factors travel back while the factorisation is still running (pieces of consecutive owned blocks, as
        // their done flags come up)
        std::vector<pangulu_storage_slot_t*> own;
        own.reserve(B.n_own_offdiag + 2 * B.n_own_diag);
        for (size_t i = 0; i < B.n_own_offdiag + 2 * B.n_own_diag; ++i) own.push_back(&B.slots[i]);
        pangulu_sm100_plan_execute_download(h->plan, own.size(), own.data());
        PSM(synchronize)();
    }
    else
    {
        const long long eager = pangulu_sm100_get_option("eager_download");
        pangulu_sm100_set_option("eager_download", 0);
        // host -> device: the value region of the block buffer
        PSM(memcpy)(B.d_buf, B.h_buf, B.value_bytes, 0);
        const size_t nwave = h->sched.wave_ptr.size() - 1;
        for (size_t w = 0; w < nwave; ++w)
        {
            const uint64_t a = h->sched.wave_ptr[w], b = h->sched.wave_ptr[w + 1];
            if (b > a) PSM(hybrid_batched)((pangulu_inblock_idx)h->nb, b - a, h->sched.tasks.data() + a);
        }
        PSM(synchronize)();
        // device -> host: finished factors (gstrs reads host values, src/pangulu_sptrsv.c)
        PSM(memcpy)(B.h_buf, B.d_buf, B.own_value_bytes, 1);
        pangulu_sm100_set_option("eager_download", eager);
    }
    h->times["gstrf"] = wall_seconds() - t0;
    h->factored = true;
}

void pangulu_gstrs(sparse_value_t* rhs, pangulu_gstrs_options* gstrs_options, void** pangulu_handle)
{
    if (gstrs_options == NULL)
    {
        printf("[PanguLU Error] Options pointer is NULL. (pangulu_gstrs)\n");
        exit(1);
    }
    Handle* h = (Handle*)*pangulu_handle;
    const eidx n = h->n;
    if (h->nproc > 1)
    {
        // the triangular solves here walk the blocks of ONE rank; with the factors spread over several GPUs that would
        // silently drop the remote blocks.  The reference's distributed gstrs (src/pangulu_sptrsv.c) is not part of
        // this module's scope: refuse instead of returning a wrong x.
        fprintf(stderr, "[PanguLU ERROR] pangulu_gstrs / pangulu_gssv on %d ranks: the sm_100a driver solves on one rank only; gather the factors "
                        "(run gstrf with nproc = 1) or use PanguLU's own distributed gstrs\n", h->nproc);
        exit(EXIT_FAILURE);
    }
    const double t0 = wall_seconds();
    const bool on_device = h->device_solve && !h->B.host_only;
    if (!on_device && h->mode == 1)
    {
        // resident mode keeps factors in HBM; the host solve needs them once
        PSM(memcpy)(h->B.h_buf, h->B.d_buf, h->B.own_value_bytes, 1);
    }
    T* b = (T*)rhs;
    // b' = P (Dr b): src/pangulu_reordering.c:683-697
    std::vector<T> x((size_t)h->B.block_length * h->nb, t_zero());
    for (eidx i = 0; i < n; ++i) x[h->R.metis_perm[h->R.row_perm[i]]] = t_mul(b[i], h->R.row_scale[i]);
    if (on_device)
    {
        // the factors stay in HBM (every gstrf mode leaves the finished values in the device block buffer): block
        // forward / backward substitution on the device, only the right-hand side travels
        if (!h->solve_plan)
        {
            std::vector<pangulu_storage_slot_t*> sl;
            for (pangulu_storage_slot_t& s : h->B.slots)
                if (s.bin_id == 0) sl.push_back(&s);
            h->solve_plan = pangulu_sm100_solve_plan_create((pangulu_inblock_idx)h->nb, h->B.block_length, sl.size(), sl.data());
        }
        pangulu_sm100_solve_plan_execute(h->solve_plan, (calculate_type*)x.data());
    }
    else
        solve_host(h->B, x);
    // x = Dc P^T x': src/pangulu_reordering.c:699-713
    for (eidx i = 0; i < n; ++i) b[i] = t_mul(x[h->R.metis_perm[i]], h->R.col_scale[i]);
    h->times["gstrs"] = wall_seconds() - t0;
}

void pangulu_gssv(sparse_value_t* rhs, pangulu_gstrf_options* gstrf_options, pangulu_gstrs_options* gstrs_options, void** pangulu_handle)
{
    pangulu_gstrf(gstrf_options, pangulu_handle);
    pangulu_gstrs(rhs, gstrs_options, pangulu_handle);
}

void pangulu_finalize(void** pangulu_handle)
{
    Handle* h = (Handle*)*pangulu_handle;
    if (!h) return;
    if (h->plan) pangulu_sm100_plan_destroy(h->plan);
    if (h->solve_plan) pangulu_sm100_solve_plan_destroy(h->solve_plan);
    for (void* p : h->ipc_opened) pangulu_sm100_ipc_close(p);
    h->ipc_opened.clear();
    if (g_barrier == h->barrier) g_barrier = nullptr;   // (the callback usually belongs to the object that owns this handle)
    if (!h->B.host_only)
        for (eidx k = 0; k < h->B.block_length; ++k)
            if (h->sched.group_count.size() > k && h->sched.group_count[k] > 1 && h->B.slot_of_diag[k] >= 0)
                pangulu_sm100_set_diag_groups(&h->B.slots[h->B.slot_of_diag[k]], 0, nullptr);
    if (h->d_scatter_dst) PSM(free)(h->d_scatter_dst);
    if (h->d_scatter_src) PSM(free)(h->d_scatter_src);
    free_blocks(h->B);
    delete h;
    *pangulu_handle = NULL;
}

// ---- accessors ------------------------------------------------------------------------------------------

double pangulu_sm100_driver_info(void* handle, const char* key)
{
    Handle* h = (Handle*)handle;
    std::string k(key ? key : "");
    if (k == "n") return h->n;
    if (k == "nb") return h->nb;
    if (k == "block_length") return h->B.block_length;
    if (k == "symbolic_nnz") return (double)h->S.symbolic_nnz;
    if (k == "nnz_blocks") return (double)h->B.nnz_total;
    if (k == "num_slots") return (double)h->B.slots.size();
    if (k == "num_own_offdiag") return (double)h->B.n_own_offdiag;
    if (k == "num_own_diag") return (double)h->B.n_own_diag;
    if (k == "num_blocks_global") return (double)h->B.bcol_ptr[h->B.block_length] + h->B.block_length;
    if (k == "nwave") return (double)(h->sched.wave_ptr.size() - 1);
    if (k == "n_split_diag")
    {
        double c = 0;
        for (uint8_t g : h->sched.group_count) c += g > 1;
        return c;
    }
    if (k == "ntask") return (double)h->sched.tasks.size();
    if (k == "n_getrf") return (double)h->sched.n_getrf;
    if (k == "n_tstrf") return (double)h->sched.n_tstrf;
    if (k == "n_gessm") return (double)h->sched.n_gessm;
    if (k == "n_ssssm") return (double)h->sched.n_ssssm;
    if (k == "flop") return h->work.flop_total();
    if (k == "flop_getrf") return h->work.flop_getrf;
    if (k == "flop_tstrf") return h->work.flop_tstrf;
    if (k == "flop_gessm") return h->work.flop_gessm;
    if (k == "flop_ssssm") return h->work.flop_ssssm;
    if (k == "bytes_getrf") return h->work.bytes_getrf;
    if (k == "bytes_tstrf") return h->work.bytes_tstrf;
    if (k == "bytes_gessm") return h->work.bytes_gessm;
    if (k == "bytes_ssssm") return h->work.bytes_ssssm;
    if (k == "value_bytes") return (double)h->B.value_bytes;
    if (k == "own_value_bytes") return (double)h->B.own_value_bytes;
    if (k == "buffer_bytes") return (double)h->B.buf_bytes;
    if (k == "p") return h->B.p;
    if (k == "q") return h->B.q;
    if (k.rfind("plan_", 0) == 0) return h->plan ? pangulu_sm100_plan_stat(h->plan, k.c_str() + 5) : -1.0;
    if (k.rfind("time_", 0) == 0)
    {
        auto it = h->times.find(k.substr(5));
        return it == h->times.end() ? -1.0 : it->second;
    }
    return NAN;
}

pangulu_storage_slot_t* pangulu_sm100_driver_slot(void* handle, long long i)
{
    Handle* h = (Handle*)handle;
    if (i < 0 || (size_t)i >= h->B.slots.size()) return NULL;
    return &h->B.slots[i];
}

const void* pangulu_sm100_driver_perm(void* handle, int which)
{
    Handle* h = (Handle*)handle;
    switch (which)
    {
    case 0: return h->R.row_perm.data();
    case 1: return h->R.col_perm.data();
    case 2: return h->R.metis_perm.data();
    case 3: return h->R.row_scale.data();
    case 4: return h->R.col_scale.data();
    }
    return NULL;
}

const pangulu_task_t* pangulu_sm100_driver_tasks(void* handle, const pangulu_uint64_t** wave_ptr)
{
    Handle* h = (Handle*)handle;
    if (wave_ptr) *wave_ptr = (const pangulu_uint64_t*)h->sched.wave_ptr.data();
    return h->sched.tasks.data();
}

// groups of diagonal block k (schedule.cpp: diag_groups): writes bounds[0..ng] (ng <= 4) and returns ng
int pangulu_sm100_driver_diag_groups(void* handle, long long k, unsigned short* bounds5)
{
    Handle* h = (Handle*)handle;
    if (k < 0 || (size_t)k >= h->sched.group_count.size()) return 0;
    const int ng = h->sched.group_count[k];
    for (int g = 0; g <= ng; ++g) bounds5[g] = h->sched.group_bounds[(size_t)k * (DIAG_MAX_GROUPS + 1) + g];
    return ng;
}

void pangulu_sm100_driver_set_mode(void* handle, int mode) { ((Handle*)handle)->mode = mode; }

// drop the device-resident plan and build it again with the platform options in force now (e.g. "trace")
void pangulu_sm100_driver_rebuild_plan(void* handle)
{
    Handle* h = (Handle*)handle;
    if (h->B.host_only) return;
    if (h->plan) pangulu_sm100_plan_destroy(h->plan);
    h->plan = pangulu_sm100_plan_create((pangulu_inblock_idx)h->nb, h->sched.wave_ptr.size() - 1, (const pangulu_uint64_t*)h->sched.wave_ptr.data(),
                                        h->sched.tasks.data());
    h->wired = false;   // several ranks: the new plan has new flags / compact copies; every rank must disconnect + connect again
}

// close the peer mappings of an earlier connect (before a new round of connect_peer calls, or before finalize)
void pangulu_sm100_driver_disconnect(void* handle)
{
    Handle* h = (Handle*)handle;
    for (void* p : h->ipc_opened) pangulu_sm100_ipc_close(p);
    h->ipc_opened.clear();
    h->wired = false;
}

void pangulu_sm100_driver_reset_values(void* handle, int also_device)
{
    Handle* h = (Handle*)handle;
    scatter_values(h);
    if (also_device && !h->B.host_only) PSM(memcpy)(h->B.d_buf, h->B.h_buf, h->B.value_bytes, 0);
    h->factored = false;
}

// Value scatter on the device (reference: the host loops of src/pangulu_conversion.c:241-350).  The caller passes
// the matrix again in the CSC arrays it gave pangulu_init -- same pattern, new values (a Newton or time-stepping loop).  The
// first call computes where every entry lives in the block value region; every call then moves nnz(A) values to the
// device (instead of nnz(L+U)), clears the region there and scatters.  Returns the seconds spent.
double pangulu_sm100_driver_update_values(void* handle, const sparse_pointer_t* colptr, const sparse_index_t* rowidx, const sparse_value_t* val)
{
    Handle* h = (Handle*)handle;
    BlockStore& B = h->B;
    if (B.host_only)
    {
        fprintf(stderr, "[PanguLU ERROR] update_values needs the GPU (handle was created in host_only mode)\n");
        exit(EXIT_FAILURE);
    }
    const double t0 = wall_seconds();
    const eidx n = h->n;
    const int nb = B.nb;
    const uint64_t nnz = colptr[n];
    const Csc& Ar = h->A_reordered;
    if (!h->d_scatter_dst || h->scatter_nnz != nnz)
    {
        std::vector<unsigned long long> dst(nnz, ~0ull);
        h->pos_in_reordered.assign(nnz, ~0ull);
        auto find_in = [](const iidx* idx, iptr b, iptr e, iidx key) -> int64_t {
            const iidx* it = std::lower_bound(idx + b, idx + e, key);
            if (it == idx + e || *it != key) return -1;
            return (int64_t)(it - idx);
        };
        long long bad = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : bad)
        for (int64_t c0 = 0; c0 < (int64_t)n; ++c0)
        {
            const eidx c = h->R.metis_perm[h->R.col_perm[c0]];
            const eidx J = c / nb;
            const iidx cl = (iidx)(c - J * nb);
            for (uint64_t e = colptr[c0]; e < colptr[c0 + 1]; ++e)
            {
                const eidx r = h->R.metis_perm[h->R.row_perm[rowidx[e]]];
                // position in the reordered matrix (factor check, reset_values)
                {
                    const eidx* b0 = Ar.idx.data() + Ar.ptr[c];
                    const eidx* b1 = Ar.idx.data() + Ar.ptr[c + 1];
                    const eidx* it = std::lower_bound(b0, b1, r);
                    if (it == b1 || *it != r) { bad++; continue; }
                    h->pos_in_reordered[e] = (uint64_t)(it - Ar.idx.data());
                }
                const eidx I = r / nb;
                const iidx rl = (iidx)(r - I * nb);
                const pangulu_storage_slot_t* sl = nullptr;
                int64_t pos = -1;
                if (I == J)
                {
                    const int64_t ds = B.slot_of_diag[J];
                    if (ds < 0) continue;
                    if (r == c) { sl = &B.slots[ds + 1]; pos = cl == 0 ? 0 : sl->columnpointer[cl]; }
                    else if (r > c) { sl = &B.slots[ds]; pos = find_in(sl->rowindex, cl == 0 ? 0 : sl->columnpointer[cl], sl->columnpointer[cl + 1], rl); }
                    else { sl = &B.slots[ds + 1]; pos = find_in(sl->rowindex, rl == 0 ? 0 : sl->columnpointer[rl], sl->columnpointer[rl + 1], cl); }
                }
                else
                {
                    const int64_t gid = B.find_block(I, J);
                    if (gid < 0) { bad++; continue; }
                    const int64_t si = B.slot_of_block[gid];
                    if (si < 0) continue;
                    sl = &B.slots[si];
                    pos = find_in(sl->rowindex, cl == 0 ? 0 : sl->columnpointer[cl], sl->columnpointer[cl + 1], rl);
                }
                if (pos < 0) { bad++; continue; }
                dst[e] = (unsigned long long)(((const char*)(sl->value + pos) - B.h_buf) / sizeof(T));
            }
        }
        if (bad)
        {
            fprintf(stderr, "[PanguLU ERROR] update_values: %lld entries are not in the pattern the handle was built for\n", bad);
            exit(EXIT_FAILURE);
        }
        // diagonal entries pangulu_init inserted (1e-8, src/pangulu_reordering.c:715-796) are not in the caller's arrays
        std::vector<char> seen(Ar.ptr[n], 0);
        for (uint64_t e = 0; e < nnz; ++e) seen[h->pos_in_reordered[e]] = 1;
        std::vector<unsigned long long> extra_dst;
        for (eidx c = 0; c < n; ++c)
            for (eptr e = Ar.ptr[c]; e < Ar.ptr[c + 1]; ++e)
                if (!seen[e] && Ar.idx[e] == c)
                {
                    const int64_t ds = B.slot_of_diag[c / nb];
                    if (ds < 0) continue;
                    const pangulu_storage_slot_t* sl = &B.slots[ds + 1];
                    const iidx cl = (iidx)(c % nb);
                    extra_dst.push_back((unsigned long long)(((const char*)(sl->value + (cl == 0 ? 0 : sl->columnpointer[cl])) - B.h_buf) / sizeof(T)));
                }
        h->scatter_extra = extra_dst.size();
        dst.insert(dst.end(), extra_dst.begin(), extra_dst.end());
        if (h->d_scatter_dst) PSM(free)(h->d_scatter_dst);
        if (h->d_scatter_src) PSM(free)(h->d_scatter_src);
        PSM(malloc)((void**)&h->d_scatter_dst, dst.size() * sizeof(unsigned long long));
        PSM(malloc)((void**)&h->d_scatter_src, dst.size() * sizeof(T));
        PSM(memcpy)(h->d_scatter_dst, dst.data(), dst.size() * sizeof(unsigned long long), 0);
        std::vector<T> tail(h->scatter_extra, t_from_real(1e-8));
        if (h->scatter_extra) PSM(memcpy)(h->d_scatter_src + nnz, tail.data(), tail.size() * sizeof(T), 0);
        h->scatter_nnz = nnz;
    }
    // host copy of the reordered matrix (the factor check and reset_values read it)
    {
        const T* v = (const T*)val;
        T* out = h->A_reordered.val.data();
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < (int64_t)nnz; ++e) out[h->pos_in_reordered[e]] = v[e];
    }
    PSM(memcpy)(h->d_scatter_src, val, nnz * sizeof(T), 0);
    pangulu_sm100_scatter_values(B.d_buf, B.value_bytes, h->d_scatter_dst, h->d_scatter_src, nullptr, nnz + h->scatter_extra);
    PSM(synchronize)();
    h->factored = false;
    return wall_seconds() - t0;
}

// CSR mirrors of the owned strictly-lower blocks rebuilt ON THE DEVICE from their CSC patterns (reference:
// pangulu_transpose_struct_with_valueidx_inblock, src/pangulu_utils.c:479-511).  clear_first = 1 overwrites the three device
// arrays of every such block with 0xff first (they were uploaded from the host-built copies at init).  Returns the number of
// bytes in which the device's index region then differs from the host-built one (0 = bit-identical).
long long pangulu_sm100_driver_device_mirrors(void* handle, int clear_first)
{
    Handle* h = (Handle*)handle;
    BlockStore& B = h->B;
    if (B.host_only) return -1;
    std::vector<pangulu_storage_slot_t*> lower;
    for (pangulu_storage_slot_t& s : B.slots)
        if (s.brow_pos > s.bcol_pos && s.d_rowpointer && s.d_idx_of_csc_value_for_csr) lower.push_back(&s);
    if (clear_first)
        for (pangulu_storage_slot_t* s : lower)
        {
            const size_t nnz = s->columnpointer[B.nb];
            PSM(memset)(s->d_rowpointer, 0xff, sizeof(iptr) * (size_t)(B.nb + 1));
            PSM(memset)(s->d_columnindex, 0xff, sizeof(iidx) * nnz);
            PSM(memset)(s->d_idx_of_csc_value_for_csr, 0xff, sizeof(iptr) * nnz);
        }
    pangulu_sm100_mirror_patterns((pangulu_inblock_idx)B.nb, lower.size(), lower.data());
    const size_t index_bytes = B.buf_bytes - B.value_bytes;
    std::vector<char> got(index_bytes);
    PSM(memcpy)(got.data(), B.d_buf + B.value_bytes, index_bytes, 1);
    long long diff = 0;
    for (size_t i = 0; i < index_bytes; ++i) diff += got[i] != B.h_buf[B.value_bytes + i];
    return diff;
}

// resident mode only: enqueue the whole factorisation on the module's stream and return without
// synchronising (bench.py brackets it with CUDA events on that stream)
void pangulu_sm100_driver_gstrf_enqueue(void* handle)
{
    Handle* h = (Handle*)handle;
    if (h->B.host_only)
    {
        fprintf(stderr, "[PanguLU ERROR] host_only handle: numeric factorisation needs the GPU -- there is no CPU path\n");
        exit(EXIT_FAILURE);
    }
    if (!h->plan)
        h->plan = pangulu_sm100_plan_create((pangulu_inblock_idx)h->nb, h->sched.wave_ptr.size() - 1, (const pangulu_uint64_t*)h->sched.wave_ptr.data(),
                                            h->sched.tasks.data());
    pangulu_sm100_plan_execute(h->plan);
    h->factored = true;
}

// ---- multi-GPU wiring (one rank per GPU; see pangulu_sm100_plan_* in the platform header) ----------------------
void pangulu_sm100_driver_set_barrier(void (*fn)(void)) { g_barrier = fn; }

static void ensure_plan(Handle* h)
{
    if (!h->plan)
        h->plan = pangulu_sm100_plan_create((pangulu_inblock_idx)h->nb, h->sched.wave_ptr.size() - 1, (const pangulu_uint64_t*)h->sched.wave_ptr.data(),
                                            h->sched.tasks.data());
}

// 192 bytes: CUDA IPC handles of this rank's block buffer, of its plan's synchronisation buffer and of the plan's
// compact dense operand copies (zeros when the plan has none)
void pangulu_sm100_driver_ipc_handles(void* handle, unsigned char* out192)
{
    Handle* h = (Handle*)handle;
    ensure_plan(h);
    memset(out192, 0, 192);
    pangulu_sm100_ipc_export(h->B.d_buf, out192);
    size_t bytes = 0;
    pangulu_sm100_ipc_export(pangulu_sm100_plan_sync_base(h->plan, &bytes), out192 + 64);
    void* dense = pangulu_sm100_plan_dense_base(h->plan, &bytes);
    if (dense && bytes) pangulu_sm100_ipc_export(dense, out192 + 128);
}

// One row of 7 int64 per remote block (per independent group, for the halves of a diagonal block) this rank mirrors AND reads: {brow, bcol, kind (0 off-diagonal, 1 lower half,
// 2 upper half of a diagonal block), byte offset of its value array in the block buffer, byte offset of its done flag
// in the plan's synchronisation buffer, byte offset of its compact dense copy in the plan's dense buffer or -1, group or -1}.
// Also tells the plan that these blocks arrive from a peer.
long long pangulu_sm100_driver_mirror_table(void* handle, long long* out, long long cap)
{
    Handle* h = (Handle*)handle;
    BlockStore& B = h->B;
    long long n = 0;
    if (!B.host_only) ensure_plan(h);
    for (const pangulu_storage_slot_t& s : B.slots)
    {
        if (s.bin_id == 0) continue;
        const bool is_diag = s.brow_pos == s.bcol_pos;
        const int ng = is_diag ? h->sched.group_count[s.brow_pos] : 1;
        for (int g = 0; g < ng; ++g)
        {
            long long flag = -1;
            if (!B.host_only)
            {
                flag = pangulu_sm100_plan_flag_offset2(h->plan, s.d_value, g);
                if (flag < 0) continue;  // mirrored but never read by a task of this rank
                pangulu_sm100_plan_expect_remote2(h->plan, s.d_value, g);
            }
            if (out && n < cap)
            {
                long long* r = out + 7 * n;
                r[0] = s.brow_pos;
                r[1] = s.bcol_pos;
                r[2] = !is_diag ? 0 : (s.is_upper ? 2 : 1);
                r[3] = (long long)((char*)s.d_value - B.d_buf);
                r[4] = flag;
                r[5] = (B.host_only || is_diag) ? -1 : pangulu_sm100_plan_dense_offset(h->plan, s.d_value);
                r[6] = is_diag ? g : -1;   // independent group of the diagonal half (-1: a whole off-diagonal block)
            }
            ++n;
        }
    }
    return n;
}

// Register, for every block of `table` (a peer's mirror table) that THIS rank owns, a push into the peer's copy.
// handles128 = the peer's pangulu_sm100_driver_ipc_handles.  Returns the number of pushes registered.
long long pangulu_sm100_driver_connect_peer(void* handle, const unsigned char* handles128, const long long* table, long long n)
{
    Handle* h = (Handle*)handle;
    BlockStore& B = h->B;
    char *peer_buf = nullptr, *peer_sync = nullptr, *peer_dense = nullptr;
    if (!B.host_only)
    {
        ensure_plan(h);
        pangulu_sm100_ipc_open(handles128, (void**)&peer_buf);
        pangulu_sm100_ipc_open(handles128 + 64, (void**)&peer_sync);
        h->ipc_opened.push_back(peer_buf);
        h->ipc_opened.push_back(peer_sync);
        bool any_dense = false;
        for (long long i = 0; i < n; ++i) any_dense = any_dense || table[7 * i + 5] >= 0;
        if (any_dense)
        {
            pangulu_sm100_ipc_open(handles128 + 128, (void**)&peer_dense);
            h->ipc_opened.push_back(peer_dense);
        }
    }
    long long made = 0;
    for (long long i = 0; i < n; ++i)
    {
        const long long* r = table + 7 * i;
        const eidx brow = (eidx)r[0], bcol = (eidx)r[1];
        if (B.owner(brow, bcol) != B.rank) continue;
        int64_t slot = -1;
        if (r[2] == 0)
        {
            const int64_t gid = B.find_block(brow, bcol);
            if (gid >= 0) slot = B.slot_of_block[gid];
        }
        else
        {
            slot = B.slot_of_diag[brow];
            if (slot >= 0 && r[2] == 2) slot += 1;
        }
        if (slot < 0)
        {
            fprintf(stderr, "[PanguLU ERROR] peer asks for block (%u,%u) that rank %d owns but does not hold\n", brow, bcol, B.rank);
            exit(EXIT_FAILURE);
        }
        if (!B.host_only)
        {
            if (r[6] >= 0)
            {
                // one independent group of a diagonal half: its entries are a contiguous range of the value array
                const pangulu_storage_slot_t& sl = B.slots[slot];
                const uint16_t* bd = &h->sched.group_bounds[(size_t)brow * (DIAG_MAX_GROUPS + 1)];
                const int g = (int)r[6];
                const unsigned off = sl.columnpointer[bd[g]], end = sl.columnpointer[bd[g + 1]];
                pangulu_sm100_plan_add_push3(h->plan, sl.d_value, g, off, end - off, peer_buf + r[3], peer_sync + r[4]);
            }
            else
                pangulu_sm100_plan_add_push2(h->plan, B.slots[slot].d_value, peer_buf + r[3], peer_sync + r[4], r[5] >= 0 ? peer_dense + r[5] : nullptr);
        }
        ++made;
    }
    return made;
}

void pangulu_sm100_driver_commit(void* handle)
{
    Handle* h = (Handle*)handle;
    if (!h->B.host_only) pangulu_sm100_plan_commit(h->plan);
    h->barrier = g_barrier;
    h->wired = true;
}

void pangulu_sm100_driver_gstrf_reset(void* handle) { pangulu_sm100_plan_reset(((Handle*)handle)->plan); }
void pangulu_sm100_driver_gstrf_launch(void* handle) { pangulu_sm100_plan_launch(((Handle*)handle)->plan); }

long long pangulu_sm100_driver_trace(void* handle, unsigned long long* out, long long cap)
{
    Handle* h = (Handle*)handle;
    return h->plan ? pangulu_sm100_plan_trace(h->plan, out, cap) : 0;
}

long long pangulu_sm100_driver_task_work(void* handle, float* out, long long cap)
{
    Handle* h = (Handle*)handle;
    return h->plan ? pangulu_sm100_plan_task_work(h->plan, out, cap) : 0;
}

void pangulu_sm100_driver_download(void* handle)
{
    Handle* h = (Handle*)handle;
    PSM(memcpy)(h->B.h_buf, h->B.d_buf, h->B.own_value_bytes, 1);
}

// The three passes of the reference's gstrf self-check (src/pangulu_numeric.c:1082-1341) over the blocks THIS rank owns;
// the caller sums ux, then lux, over the ranks.  Complex types: real parts (the check vectors are real for real input).
void pangulu_sm100_driver_check_ux(void* handle, double* ux)
{
    Handle* h = (Handle*)handle;
    const BlockStore& B = h->B;
    const int nb = B.nb;
    for (const pangulu_storage_slot_t& s : B.slots)
    {
        if (s.bin_id != 0) continue;
        const size_t r0 = (size_t)s.brow_pos * nb;
        if (s.brow_pos == s.bcol_pos && s.is_upper)
        {
            for (int r = 0; r < nb; ++r)
                for (iptr e = s.columnpointer[r]; e < s.columnpointer[r + 1]; ++e) ux[r0 + r] += t_real(((const T*)s.value)[e]);
        }
        else if (s.brow_pos < s.bcol_pos)
        {
            for (int c = 0; c < nb; ++c)
                for (iptr e = s.columnpointer[c]; e < s.columnpointer[c + 1]; ++e) ux[r0 + s.rowindex[e]] += t_real(((const T*)s.value)[e]);
        }
    }
}
void pangulu_sm100_driver_check_lux(void* handle, const double* ux, double* lux)
{
    Handle* h = (Handle*)handle;
    const BlockStore& B = h->B;
    const int nb = B.nb;
    const size_t N = (size_t)B.block_length * nb;
    if (h->rank == 0)
        for (size_t i = 0; i < N; ++i) lux[i] += ux[i];   // unit diagonal of L, counted once
    for (const pangulu_storage_slot_t& s : B.slots)
    {
        if (s.bin_id != 0) continue;
        const size_t r0 = (size_t)s.brow_pos * nb, c0 = (size_t)s.bcol_pos * nb;
        if ((s.brow_pos == s.bcol_pos && !s.is_upper) || s.brow_pos > s.bcol_pos)
            for (int c = 0; c < nb; ++c)
                for (iptr e = s.columnpointer[c]; e < s.columnpointer[c + 1]; ++e) lux[r0 + s.rowindex[e]] += t_real(((const T*)s.value)[e]) * ux[c0 + c];
    }
}
long long pangulu_sm100_driver_check_ax(void* handle, double* ax)
{
    Handle* h = (Handle*)handle;
    const Csc& A = h->A_reordered;
    for (eidx c = 0; c < A.n; ++c)
        for (eptr e = A.ptr[c]; e < A.ptr[c + 1]; ++e) ax[A.idx[e]] += t_real(A.val[e]);
    return (long long)A.n;
}

// ---- factor files ----------------------------------------------------------------------------
// The owned blocks are written one after the other in PanguLU's OWN packed block format -- the image a block has in
// bin 0 and on the wire (src/pangulu_storage.c:295-357; header src/pangulu_communication.c:1929-1942, SURVEY
// Appendix A): 32-byte header {u64 nnz, u32 brow, u32 bcol, u32 is_upper, 12 bytes pad}, value[nnz], u32
// columnpointer[nb+1], u16 rowindex[nnz], pad to 8, and for strictly-lower blocks u32 idx_of_csc_value_for_csr[nnz],
// u32 rowpointer[nb+1], u16 columnindex[nnz], pad to 8 -- behind a file header with the permutations and scalings.
// load_factors puts the values back into a handle initialised on the same matrix (patterns are compared entry by entry):
// pangulu_init + load_factors + pangulu_gstrs solves without factoring again; the files also serve as golden vectors.
namespace {
struct FactorFileHeader
{
    char magic[8];
    uint32_t version, value_type, sizeof_value, nb, n, block_length;
    uint64_t nblocks;
};
struct WireHeader { uint64_t nnz; uint32_t brow, bcol, is_upper; unsigned char pad[12]; };
static_assert(sizeof(WireHeader) == 32, "PanguLU block header");
inline uint32_t value_type_code()
{
#if defined(CALCULATE_TYPE_CR64)
    return 2;
#elif defined(CALCULATE_TYPE_CR32)
    return 3;
#elif defined(CALCULATE_TYPE_R32)
    return 1;
#else
    return 0;
#endif
}
}  // namespace

long long pangulu_sm100_driver_save_factors(void* handle, const char* path)
{
    Handle* h = (Handle*)handle;
    const BlockStore& B = h->B;
    const int nb = B.nb;
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    FactorFileHeader fh;
    memset(&fh, 0, sizeof fh);
    memcpy(fh.magic, "PGLUF01", 8);
    fh.version = 1; fh.value_type = value_type_code(); fh.sizeof_value = (uint32_t)sizeof(T); fh.nb = (uint32_t)nb; fh.n = h->n; fh.block_length = B.block_length;
    for (const pangulu_storage_slot_t& s : B.slots) fh.nblocks += s.bin_id == 0;
    long long bytes = 0;
    auto put = [&](const void* p, size_t n) { if (n) bytes += (long long)fwrite(p, 1, n, f); };
    put(&fh, sizeof fh);
    put(h->R.metis_perm.data(), sizeof(eidx) * h->n);
    put(h->R.row_perm.data(), sizeof(eidx) * h->n);
    put(h->R.col_perm.data(), sizeof(eidx) * h->n);
    put(h->R.row_scale.data(), sizeof(T) * h->n);
    put(h->R.col_scale.data(), sizeof(T) * h->n);
    const unsigned char zeros[8] = {0};
    for (const pangulu_storage_slot_t& s : B.slots)
    {
        if (s.bin_id != 0) continue;
        const size_t nnz = s.columnpointer[nb];
        WireHeader wh;
        memset(&wh, 0, sizeof wh);
        wh.nnz = nnz; wh.brow = s.brow_pos; wh.bcol = s.bcol_pos; wh.is_upper = (uint32_t)s.is_upper;
        put(&wh, sizeof wh);
        put(s.value, sizeof(T) * nnz);
        put(s.columnpointer, sizeof(iptr) * (nb + 1));
        put(s.rowindex, sizeof(iidx) * nnz);
        put(zeros, (8 - (sizeof(iptr) * (nb + 1) + sizeof(iidx) * nnz) % 8) % 8);
        if (s.brow_pos > s.bcol_pos)
        {
            put(s.idx_of_csc_value_for_csr, sizeof(iptr) * nnz);
            put(s.rowpointer, sizeof(iptr) * (nb + 1));
            put(s.columnindex, sizeof(iidx) * nnz);
            put(zeros, (8 - (sizeof(iptr) * nnz + sizeof(iptr) * (nb + 1) + sizeof(iidx) * nnz) % 8) % 8);
        }
    }
    fclose(f);
    return bytes;
}

long long pangulu_sm100_driver_load_factors(void* handle, const char* path)
{
    Handle* h = (Handle*)handle;
    BlockStore& B = h->B;
    const int nb = B.nb;
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    auto fail = [&](const char* what) {
        fprintf(stderr, "[PanguLU ERROR] factor file %s: %s\n", path, what);
        exit(EXIT_FAILURE);
    };
    auto get = [&](void* p, size_t n) { if (n && fread(p, 1, n, f) != n) fail("truncated"); };
    FactorFileHeader fh;
    get(&fh, sizeof fh);
    if (memcmp(fh.magic, "PGLUF01", 8) != 0 || fh.version != 1) fail("not a PanguLU sm100 factor file (version 1)");
    if (fh.value_type != value_type_code() || fh.sizeof_value != sizeof(T)) fail("value type differs from this library's");
    if ((int)fh.nb != nb || fh.n != h->n || fh.block_length != B.block_length) fail("matrix order / block size differ from the handle's");
    std::vector<eidx> perm(h->n);
    get(perm.data(), sizeof(eidx) * h->n);
    if (memcmp(perm.data(), h->R.metis_perm.data(), sizeof(eidx) * h->n) != 0) fail("fill-reducing permutation differs from the handle's");
    get(perm.data(), sizeof(eidx) * h->n);
    get(perm.data(), sizeof(eidx) * h->n);
    std::vector<T> scale(h->n);
    get(scale.data(), sizeof(T) * h->n);
    get(scale.data(), sizeof(T) * h->n);
    std::vector<iptr> ptr(nb + 1);
    std::vector<iidx> idx;
    std::vector<unsigned char> skip;
    long long loaded = 0;
    for (pangulu_storage_slot_t& s : B.slots)
    {
        if (s.bin_id != 0) continue;
        const size_t nnz = s.columnpointer[nb];
        WireHeader wh;
        get(&wh, sizeof wh);
        if (wh.nnz != nnz || wh.brow != s.brow_pos || wh.bcol != s.bcol_pos || (s.brow_pos == s.bcol_pos && wh.is_upper != (uint32_t)s.is_upper))
            fail("block order or size differs from the handle's");
        get(s.value, sizeof(T) * nnz);
        get(ptr.data(), sizeof(iptr) * (nb + 1));
        idx.resize(nnz);
        get(idx.data(), sizeof(iidx) * nnz);
        if (memcmp(ptr.data(), s.columnpointer, sizeof(iptr) * (nb + 1)) != 0 || (nnz && memcmp(idx.data(), s.rowindex, sizeof(iidx) * nnz) != 0))
            fail("nonzero pattern of a block differs from the handle's");
        size_t rest = (8 - (sizeof(iptr) * (nb + 1) + sizeof(iidx) * nnz) % 8) % 8;
        if (s.brow_pos > s.bcol_pos)
            rest += sizeof(iptr) * nnz + sizeof(iptr) * (nb + 1) + sizeof(iidx) * nnz + (8 - (sizeof(iptr) * nnz + sizeof(iptr) * (nb + 1) + sizeof(iidx) * nnz) % 8) % 8;
        skip.resize(rest);
        get(skip.data(), rest);
        ++loaded;
    }
    fclose(f);
    if ((uint64_t)loaded != fh.nblocks) fail("block count differs from the handle's");
    if (!B.host_only) PSM(memcpy)(B.d_buf, B.h_buf, B.own_value_bytes, 0);
    h->factored = true;
    return loaded;
}

// || L (U 1) - A_reordered 1 ||_2 / || A_reordered 1 ||_2 from the host copies of the factors: the reference's
// own gstrf self-check (src/pangulu_numeric.c:1082-1341), single rank.
double pangulu_sm100_driver_factor_check(void* handle)
{
    Handle* h = (Handle*)handle;
    const BlockStore& B = h->B;
    const int nb = B.nb;
    const size_t N = (size_t)B.block_length * nb;
#ifdef PANGULU_COMPLEX
    typedef std::complex<long double> acc_t;
    auto wide = [](const T& v) { return acc_t((long double)v.re, (long double)v.im); };
#else
    typedef long double acc_t;
    auto wide = [](const T& v) { return (long double)v; };
#endif
    std::vector<acc_t> ux(N, acc_t(0)), lux(N, acc_t(0)), ax(N, acc_t(0));
    const Csc& A = h->A_reordered;
    for (eidx c = 0; c < A.n; ++c)
        for (eptr e = A.ptr[c]; e < A.ptr[c + 1]; ++e) ax[A.idx[e]] += wide(A.val[e]);
    // U * 1
    for (const pangulu_storage_slot_t& s : B.slots)
    {
        if (s.bin_id != 0) continue;
        const size_t r0 = (size_t)s.brow_pos * nb;
        if (s.brow_pos == s.bcol_pos && s.is_upper)
        {
            for (int r = 0; r < nb; ++r)
                for (iptr e = s.columnpointer[r]; e < s.columnpointer[r + 1]; ++e) ux[r0 + r] += wide(s.value[e]);
        }
        else if (s.brow_pos < s.bcol_pos)
        {
            for (int c = 0; c < nb; ++c)
                for (iptr e = s.columnpointer[c]; e < s.columnpointer[c + 1]; ++e) ux[r0 + s.rowindex[e]] += wide(s.value[e]);
        }
    }
    // L * (U 1), unit diagonal
    for (size_t i = 0; i < N; ++i) lux[i] = ux[i];
    for (const pangulu_storage_slot_t& s : B.slots)
    {
        if (s.bin_id != 0) continue;
        const size_t r0 = (size_t)s.brow_pos * nb, c0 = (size_t)s.bcol_pos * nb;
        if ((s.brow_pos == s.bcol_pos && !s.is_upper) || s.brow_pos > s.bcol_pos)
        {
            for (int c = 0; c < nb; ++c)
                for (iptr e = s.columnpointer[c]; e < s.columnpointer[c + 1]; ++e) lux[r0 + s.rowindex[e]] += wide(s.value[e]) * ux[c0 + c];
        }
    }
    long double num = 0, den = 0;
    for (eidx i = 0; i < A.n; ++i)
    {
        num += std::norm(acc_t(lux[i] - ax[i]));
        den += std::norm(acc_t(ax[i]));
    }
    return (double)std::sqrt((double)(num / den));
}

}  // extern "C"

namespace {

// re-scatter the reordered matrix into the (zeroed) host value arrays
void scatter_values(Handle* h)
{
    BlockStore& B = h->B;
    const int nb = B.nb;
    const Csc& A = h->A_reordered;
    memset(B.h_buf, 0, B.value_bytes);
    auto find_in = [](const iidx* idx, iptr b, iptr e, iidx key) -> int64_t {
        const iidx* it = std::lower_bound(idx + b, idx + e, key);
        if (it == idx + e || *it != key) return -1;
        return (int64_t)(it - idx);
    };
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t c = 0; c < (int64_t)A.n; ++c)
    {
        const eidx J = (eidx)c / nb;
        const iidx cl = (iidx)((eidx)c - J * nb);
        for (eptr e = A.ptr[c]; e < A.ptr[c + 1]; ++e)
        {
            const eidx r = A.idx[e];
            const eidx I = r / nb;
            const iidx rl = (iidx)(r - I * nb);
            pangulu_storage_slot_t* sl = nullptr;
            int64_t pos = -1;
            if (I == J)
            {
                const int64_t ds = B.slot_of_diag[J];
                if (ds < 0) continue;
                if (r == (eidx)c) { sl = &B.slots[ds + 1]; pos = sl->columnpointer[cl]; }
                else if (r > (eidx)c) { sl = &B.slots[ds]; pos = find_in(sl->rowindex, sl->columnpointer[cl], sl->columnpointer[cl + 1], rl); }
                else { sl = &B.slots[ds + 1]; pos = find_in(sl->rowindex, sl->columnpointer[rl], sl->columnpointer[rl + 1], cl); }
            }
            else
            {
                const int64_t gid = B.find_block(I, J);
                if (gid < 0) continue;
                const int64_t s = B.slot_of_block[gid];
                if (s < 0) continue;
                sl = &B.slots[s];
                pos = find_in(sl->rowindex, sl->columnpointer[cl], sl->columnpointer[cl + 1], rl);
            }
            if (pos >= 0) sl->value[pos] = A.val[e];
        }
    }
}

}  // namespace

namespace pdrv {

// Block forward / backward substitution on the host copies of L and U (reference: src/pangulu_sptrsv.c:24-174
// with the CPU kernels pangulu_platform_0100000.c:435-506; one rank).  x has block_length*nb entries.
void solve_host(const BlockStore& B, std::vector<T>& x)
{
    const int nb = B.nb;
    const eidx bl = B.block_length;
    // forward: for each block column k: x_k <- L_kk^-1 x_k (unit diagonal), then x_i -= L_ik x_k
    for (eidx k = 0; k < bl; ++k)
    {
        T* xk = x.data() + (size_t)k * nb;
        const int64_t ds = B.slot_of_diag[k];
        if (ds < 0) continue;
        const pangulu_storage_slot_t& lo = B.slots[ds];
        for (int c = 0; c < nb; ++c)
        {
            const T xc = xk[c];
            for (iptr e = lo.columnpointer[c]; e < lo.columnpointer[c + 1]; ++e) xk[lo.rowindex[e]] = t_sub(xk[lo.rowindex[e]], t_mul(lo.value[e], xc));
        }
        for (eptr g = B.bcol_ptr[k]; g < B.bcol_ptr[k + 1]; ++g)
        {
            if (B.bcol_row[g] <= k) continue;
            const int64_t s = B.slot_of_block[g];
            if (s < 0) continue;
            const pangulu_storage_slot_t& a = B.slots[s];
            T* xi = x.data() + (size_t)a.brow_pos * nb;
            for (int c = 0; c < nb; ++c)
            {
                const T xc = xk[c];
                for (iptr e = a.columnpointer[c]; e < a.columnpointer[c + 1]; ++e) xi[a.rowindex[e]] = t_sub(xi[a.rowindex[e]], t_mul(a.value[e], xc));
            }
        }
    }
    // backward: for each block column k descending: x_k <- U_kk^-1 x_k, then x_i -= U_ik x_k (i < k)
    for (eidx kk = bl; kk-- > 0;)
    {
        T* xk = x.data() + (size_t)kk * nb;
        const int64_t ds = B.slot_of_diag[kk];
        if (ds < 0) continue;
        const pangulu_storage_slot_t& up = B.slots[ds + 1];
        for (int r = nb - 1; r >= 0; --r)
        {
            const iptr r0 = up.columnpointer[r], r1 = up.columnpointer[r + 1];
            if (r0 == r1) continue;
            T acc = xk[r];
            for (iptr e = r0 + 1; e < r1; ++e) acc = t_sub(acc, t_mul(up.value[e], xk[up.rowindex[e]]));
            T piv = up.value[r0];
            if (!(std::fabs(t_real(piv)) > PANGULU_SPTRSV_TOL)) piv = t_from_real(PANGULU_SPTRSV_TOL);
            xk[r] = t_div(acc, piv);
        }
        for (eptr g = B.bcol_ptr[kk]; g < B.bcol_ptr[kk + 1]; ++g)
        {
            if (B.bcol_row[g] >= kk) break;
            const int64_t s = B.slot_of_block[g];
            if (s < 0) continue;
            const pangulu_storage_slot_t& a = B.slots[s];
            T* xi = x.data() + (size_t)a.brow_pos * nb;
            for (int c = 0; c < nb; ++c)
            {
                const T xc = xk[c];
                for (iptr e = a.columnpointer[c]; e < a.columnpointer[c + 1]; ++e) xi[a.rowindex[e]] = t_sub(xi[a.rowindex[e]], t_mul(a.value[e], xc));
            }
        }
    }
}

}  // namespace pdrv
