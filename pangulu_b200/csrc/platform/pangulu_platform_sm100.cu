// pangulu_platform_sm100.cu -- C ABI of the sm_100a platform module (include/pangulu_platform_0201001.h).
//
// Host side of the module: translates PanguLU's host task arrays (pangulu_task_t with pointers to host
// slot structs) into device-resident descriptor arrays, groups same-destination SSSSM tasks, and
// launches ONE grid per kernel class per batch on the module's stream.  There is no CPU fallback:
// every failure prints "[PanguLU ERROR] ..." and exits, as the reference platform does
// (src/platforms/02_NONSHAREDMEM/01_GPU/000_CUDA/pangulu_platform_0201000.cu:31-50).
#include <cuda_runtime.h>
#include <time.h>
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "pangulu_platform_0201001.h"
#include "sm100_kernels.cuh"
#include "sm100_preprocess.cuh"

// globals of the PanguLU core (src/pangulu.c:7-9); absent when the module is used stand-alone
extern "C" {
extern int pangulu_gpu_kernel_warp_per_block __attribute__((weak));
extern int pangulu_gpu_data_move_warp_per_block __attribute__((weak));
extern int pangulu_gpu_shared_mem_size __attribute__((weak));
}

namespace {

using namespace psm;
typedef calculate_type T;

#define PSM_STR2(x) #x
#define PSM_STR(x) PSM_STR2(x)

// Progress board of the persistent executor (option "progress" = 1): every CTA notes the task it holds in HOST-mapped
// memory, which stays readable after a failed launch -- the dump names the tasks that never finished.
int4* g_progress_host = nullptr;
int g_progress_ctas = 0;
void dump_progress()
{
    if (!g_progress_host) return;
    static const char* kinds[] = {"tile", "sparse", "getrf_smem", "getrf_dense", "getrf_global", "panel", "ctile", "cpanel", "ddens"};
    for (int b = 0; b < g_progress_ctas; ++b)
    {
        const int4 v = g_progress_host[b];
        if (v.w == 2) continue;   // finished its last task
        const int4 q = g_progress_host[512 + b];
        fprintf(stderr, "[PanguLU ERROR]   CTA %3d: task %d kind %s arg %d state %s  set aside: %d %d %d %d\n", b, v.x, (v.y >= 0 && v.y < 9) ? kinds[v.y] : "?", v.z,
                v.w == 0 ? "waiting for dependencies" : "running", q.x, q.y, q.z, q.w);

    }
}

[[noreturn]] void fatal(const char* file, int line, const char* what, const char* detail)
{
    fprintf(stderr, "[PanguLU ERROR] sm100 platform: %s%s%s at %s:%d\n", what, detail ? ": " : "", detail ? detail : "", file, line);
    dump_progress();
    fflush(stderr);
    exit(EXIT_FAILURE);
}
#define CK(call)                                                                       \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) fatal(__FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define CK_LAST() CK(cudaGetLastError())

enum KClass { K_GETRF = 0, K_TSTRF = 1, K_GESSM = 2, K_SSSSM_SPARSE = 3, K_SSSSM_TILE = 4, K_AUX = 5, K_MEGA = 6, K_SSSSM_CTILE = 7, K_CPANEL = 8, K_NCLASS = 9 };
enum { K_GETRF_SMEM = 100, K_GETRF_DENSE = 101 };  // step kinds only; accounted under K_GETRF

struct Options
{
    long long eager_download = 1;
    long long dense_threshold = 120;
    long long ssssm_mode = 0;
    long long profile = 0;
    long long getrf_mode = 0;  // 0 auto, 1 no shared-memory staging, 2 no dense promotion, 3 always dense
    long long tile_variant = 2; // 2 pipelined 512-thread tile kernel, 1 first-generation 256-thread kernel
    long long split_pairs = 1;  // split heavy destinations of thin waves over several CTAs
    long long trace = 0;        // persistent executor: record per-task fetch/start/end timestamps (pangulu_sm100_plan_trace)
    long long executor = 1;     // plans: 1 persistent megakernel (one launch per factorisation), 0 one launch per kernel class per wave
    long long compact = 1;      // plans: SSSSM on compact dense operand copies (sm100_compact.cuh)
    // Every update takes the compact path by default (thresholds 0): measured at 100^3, routing the pairs of nearly empty
    // blocks through the warp-per-column sparse kernel instead cost 9 % of the step (1.41 s vs 1.29 s)
    long long compact_threshold = 0;   // per mille: pairs with sqrt(nnzA nnzB)/nb^2 >= this take the compact path ...
    long long compact_block_threshold = 0;  // ... if both blocks have nnz/nb^2 >= this (per mille): the block-level rule decides who gets a compact copy
    long long cpanel_block_threshold = 1;   // per mille: TSTRF/GESSM of blocks at least this dense run as dense triangular solves (compact panel); sparser ones
                                            // use the warp-per-row/column sweep, which also fills the block's compact copy
    long long cpanel = 1;       // plans: TSTRF/GESSM of blocks with a compact copy as dense triangular solves on the DMMA pipe
    long long dense_all_panels = 0;  // plans: every eligible panel block gets a compact copy (multi-GPU: a peer may need it)
    long long ct_gmax = 4, ct_minst = 2;  // compact tile ring policy: chunks per item (at most), items the ring must hold (at least)
    long long defer = 0;        // persistent executor: tasks a CTA may set aside while their dependencies are outstanding (0 = wait in place)
    long long progress = 0;     // persistent executor: per-CTA progress board in host-mapped memory, dumped when a launch fails (debugging aid)
    long long spin_timeout_ms = 30000;  // persistent executor: a dependency wait longer than this aborts the run with [PanguLU ERROR]
};

// pinned + device slab pair used to ship one batch's descriptors; recycled once its event fired
struct Slab
{
    char* h = nullptr;
    char* d = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    bool busy = false;
};

struct Ctx
{
    std::mutex mu;
    int device = -1;
    cudaStream_t stream = nullptr;
    Options opt;
    std::vector<Slab> slabs;
    long long launches[K_NCLASS] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double ms[K_NCLASS] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int max_smem_optin = 0;
    int sm_count = 0;
};
Ctx g;

// independent groups of diagonal blocks (set by the host driver; keyed by the LOWER-half slot): bounds[0..ng]
struct DiagGroupInfo { int ng; unsigned short bounds[5]; };
std::unordered_map<const pangulu_storage_slot_t*, DiagGroupInfo> g_diag_groups;
inline const DiagGroupInfo* groups_of(const pangulu_storage_slot_t* lower)
{
    auto it = g_diag_groups.find(lower);
    return it == g_diag_groups.end() ? nullptr : &it->second;
}

void ensure_ctx()
{
    if (g.stream) return;
    std::lock_guard<std::mutex> lk(g.mu);
    if (g.stream) return;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        fatal(__FILE__, __LINE__, "no CUDA device visible; this module has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : nullptr);
    CK(cudaGetDevice(&g.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, g.device));
    if (prop.major < 10)
        fatal(__FILE__, __LINE__, "device is not sm_100-class; the module is built for sm_100a only", prop.name);
    g.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    g.sm_count = prop.multiProcessorCount;
    // blocking stream: the PanguLU receiver thread issues H2D copies on the legacy NULL stream
    // (src/pangulu_communication.c:1850,1880); a blocking stream orders against them like the
    // reference's NULL-stream kernels do.
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamDefault));
    CK(cudaEventCreate(&g.ev0));
    CK(cudaEventCreate(&g.ev1));
    g.stream = s;
}

Slab& take_slab(size_t bytes)
{
    for (auto& s : g.slabs)
    {
        if (s.busy && cudaEventQuery(s.done) == cudaSuccess) s.busy = false;
    }
    for (auto& s : g.slabs)
        if (!s.busy && s.cap >= bytes) { s.busy = true; return s; }
    // none free & large enough: if the pool is big, wait for the oldest busy one that fits
    if (g.slabs.size() >= 16)
    {
        for (auto& s : g.slabs)
            if (s.cap >= bytes) { CK(cudaEventSynchronize(s.done)); s.busy = true; return s; }
    }
    Slab s;
    s.cap = std::max<size_t>(bytes, (size_t)1 << 20);
    CK(cudaHostAlloc((void**)&s.h, s.cap, cudaHostAllocDefault));
    CK(cudaMalloc((void**)&s.d, s.cap));
    CK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    s.busy = true;
    g.slabs.push_back(s);
    return g.slabs.back();
}

// ---- slot -> device descriptor ------------------------------------------------------------------
inline u32 slot_nnz(int nb, const pangulu_storage_slot_t* s) { return s->columnpointer[nb]; }

// CSC view (off-diagonal blocks and the LOWER half of a diagonal block)
inline Blk<T> blk_csc(const pangulu_storage_slot_t* s)
{
    Blk<T> b;
    b.val = (T*)s->d_value;
    b.ptr = s->d_columnpointer;
    b.idx = s->d_rowindex;
    const bool mirror = s->brow_pos > s->bcol_pos;
    b.rptr = mirror ? s->d_rowpointer : nullptr;
    b.ridx = mirror ? s->d_columnindex : nullptr;
    b.rmap = mirror ? s->d_idx_of_csc_value_for_csr : nullptr;
    return b;
}
// UPPER half of a diagonal block: CSR held in d_rowpointer/d_columnindex (src/pangulu_storage.c:346-352)
inline Blk<T> blk_upper_diag(const pangulu_storage_slot_t* s)
{
    Blk<T> b;
    b.val = (T*)s->d_value;
    b.ptr = s->d_rowpointer;
    b.idx = s->d_columnindex;
    b.rptr = nullptr; b.ridx = nullptr; b.rmap = nullptr;
    return b;
}
inline BlkRO<T> blk_ro(const pangulu_storage_slot_t* s)
{
    BlkRO<T> b;
    b.val = (const T*)s->d_value;
    b.ptr = s->d_columnpointer;
    b.idx = s->d_rowindex;
    return b;
}
inline void diag_halves(pangulu_storage_slot_t* s, pangulu_storage_slot_t*& lower, pangulu_storage_slot_t*& upper)
{
    if (s->is_upper) { upper = s; lower = s->related_block; }
    else { lower = s; upper = s->related_block; }
    if (!lower || !upper) fatal(__FILE__, __LINE__, "diagonal block without its related half", nullptr);
}

inline size_t getrf_smem_bytes(int nb, size_t nl, size_t nu)
{
    return ((size_t)(4 * nb + nl + nu) * sizeof(T) + 8 * (size_t)(nb + 1) + 2 * (nl + nu) + 15) & ~(size_t)15;
}

// ---- batch program: what one hybrid_batched call / one plan wave turns into --------------------------
struct Step
{
    int kind;          // KClass of the launch, or -1 for a device->host value download
    size_t off;        // first descriptor (index into the class's array)
    size_t count;      // descriptors (GETRF tasks / panel tasks / SSSSM destinations)
    size_t smem = 0;   // dynamic shared memory of the launch where it depends on the batch
    size_t aux_off = 0, aux_count = 0;  // compact-tile steps: range of tile tasks (Program::ctasks)
};
struct Download { void* h; const void* d; size_t bytes; };

struct Program
{
    std::vector<GetrfTask<T>> getrf;
    std::vector<PanelTask<T>> panel;
    std::vector<const pangulu_storage_slot_t*> panel_slot;  // destination slot of each panel task
    std::vector<SsssmDest<T>> dests;
    std::vector<SsssmPair<T>> pairs;
    std::vector<CPair> cpairs;        // operand pairs of the compact-tile destinations
    std::vector<CTileTask> ctasks;    // (destination, column range) tiles of the compact-tile destinations
    std::vector<CPanelTask> cptasks;  // (panel task, dense diagonal, first compact row/column) slices of the compact panel solves
    std::vector<int> getrf_dd;        // per GETRF task: dense triangular copies to refresh (index) or -1
    std::vector<const pangulu_storage_slot_t*> getrf_slot;  // lower-half slot of each GETRF task
    std::vector<const pangulu_storage_slot_t*> panel_diag_slot;  // lower-half slot of the diagonal block of each panel task
    std::vector<Step> steps;
    std::vector<Download> downloads;  // referenced by steps of kind -1 (off/count)
};

struct DevProgram
{
    GetrfTask<T>* getrf = nullptr;
    PanelTask<T>* panel = nullptr;
    SsssmDest<T>* dests = nullptr;
    SsssmPair<T>* pairs = nullptr;
    CPair* cpairs = nullptr;
    DenseOp<T>* ops = nullptr;
    CTileTask* ctasks = nullptr;
    CPanelTask* cptasks = nullptr;
    DiagDense<T>* dds = nullptr;
    int* getrf_dd = nullptr;
};

// Registry of the compact dense operand copies of one plan (sm100_compact.cuh): which blocks have one, their row /
// column lists, and where the copy lives.  Everything is derived from the (immutable) pattern in the host slot.
struct DenseReg
{
    struct Entry
    {
        const pangulu_storage_slot_t* slot;
        u32 cnt, ld, kmask;
        size_t dense_off, dense_len;  // in values
        size_t meta_off;              // in u16: list[cnt], inv[nb], then (B operands) pre[nb+1]
    };
    int nb = 0;
    std::vector<Entry> entries;
    std::unordered_map<const pangulu_storage_slot_t*, int> index;
    std::vector<u16> meta;
    std::unordered_map<std::string, size_t> meta_of_list;   // list content (+ operand side) -> its record in meta
    size_t dense_total = 0;

    int find(const pangulu_storage_slot_t* s) const
    {
        auto it = index.find(s);
        return it == index.end() ? -1 : it->second;
    }
    // as_b = false: block of the L panel used as the left operand (rows compacted); true: block of the U panel
    int get(const pangulu_storage_slot_t* s, bool as_b)
    {
        auto it = index.find(s);
        if (it != index.end()) return it->second;
        Entry e;
        memset(&e, 0, sizeof e);
        e.slot = s;
        const u32* cp = s->columnpointer;
        const u16* ri = s->rowindex;
        const u32 nnz = cp[nb];
        std::vector<u16> list, inv((size_t)nb, (u16)0xffff);
        u32 kmask = 0;
        if (!as_b)
        {
            std::vector<char> present((size_t)nb, 0);
            for (u32 q = 0; q < nnz; ++q) present[ri[q]] = 1;
            for (int r = 0; r < nb; ++r)
                if (present[r]) { inv[r] = (u16)list.size(); list.push_back((u16)r); }
            for (int c = 0; c < nb; ++c)
                if (cp[c + 1] > (c == 0 ? 0u : cp[c])) kmask |= 1u << (c / PSM_CT_KC);
            e.cnt = (u32)list.size();
            e.ld = ((e.cnt + 7u) & ~7u) + 4u;
            e.dense_len = (size_t)nb * e.ld;
        }
        else
        {
            for (int c = 0; c < nb; ++c)
                if (cp[c + 1] > (c == 0 ? 0u : cp[c])) { inv[c] = (u16)list.size(); list.push_back((u16)c); }
            for (u32 q = 0; q < nnz; ++q) kmask |= 1u << (ri[q] / PSM_CT_KC);
            e.cnt = (u32)list.size();
            e.ld = PSM_CT_BLD;
            e.dense_len = (size_t)(nb / PSM_CT_KC) * e.cnt * PSM_CT_BLD;
        }
        e.kmask = kmask;
        e.dense_off = dense_total;
        dense_total += (e.dense_len + 3) & ~(size_t)3;  // keeps every copy 16-byte aligned for all value types
        // Blocks with the same non-empty rows (columns) share ONE list record: the compact tile routine recognises
        // consecutive operand pairs with identical maps by comparing the list pointers and keeps their products in the
        // DMMA registers (sm100_compact.cuh, "runs").
        std::string key(1, as_b ? 'b' : 'a');
        key.append((const char*)list.data(), list.size() * sizeof(u16));
        auto known = meta_of_list.find(key);
        if (known != meta_of_list.end()) e.meta_off = known->second;
        else
        {
            e.meta_off = meta.size();
            meta_of_list.emplace(std::move(key), e.meta_off);
            meta.insert(meta.end(), list.begin(), list.end());
            meta.insert(meta.end(), inv.begin(), inv.end());
            if (as_b)
            {
                u16 run = 0;
                for (int c = 0; c <= nb; ++c)
                {
                    meta.push_back(run);
                    if (c < nb && inv[c] != 0xffff) ++run;
                }
            }
        }
        const int id = (int)entries.size();
        entries.push_back(e);
        index.emplace(s, id);
        return id;
    }
};

// Append the tasks [t0,t1) -- executed in array order like the reference's serial loop
// (pangulu_platform_0201000.cu:880-897) -- as maximal runs of one kernel class.
void append_tasks(Program& P, int nb, const pangulu_task_t* tasks, size_t t0, size_t t1, bool eager_download, DenseReg* reg = nullptr)
{
    size_t i = t0;
    std::vector<size_t> order;
    while (i < t1)
    {
        const int kid = tasks[i].kernel_id;
        if (kid == PANGULU_TASK_SSSSM)
        {
            size_t j = i;
            while (j < t1 && tasks[j].kernel_id == PANGULU_TASK_SSSSM) ++j;
            // group by destination (a diagonal destination is identified by its LOWER half), keeping
            // the array order of the updates of one destination; inside a destination the operand pairs
            // that go to the dense-tile kernel come first
            const size_t nrun = j - i;
            order.resize(nrun);
            std::vector<char> cls(nrun);
            const int mode = (int)g.opt.ssssm_mode;
            const bool tile_ok = nb >= 32 && nb <= 256 && mode != 1;
            const double thr = (double)g.opt.dense_threshold / 1000.0;
            const double nb4thr2 = thr * thr * (double)nb * nb * (double)nb * nb;
            // compact path (plans only): both operands have a compact dense copy.  Whether a block gets one depends on
            // its pattern alone (density), so every rank that holds the block -- its owner and the ranks that mirror it --
            // decides the same way
            const bool compact_ok = reg && g.opt.compact != 0 && nb % 16 == 0 && nb >= 32 && nb <= 256;
            const double cmin = (double)g.opt.compact_block_threshold / 1000.0 * (double)nb * nb;
            const double cthr = (double)g.opt.compact_threshold / 1000.0;
            const double nb4cthr2 = cthr * cthr * (double)nb * nb * (double)nb * nb;
            for (size_t k = 0; k < nrun; ++k)
            {
                order[k] = k;
                const pangulu_task_t& t = tasks[i + k];
                const double na = slot_nnz(nb, t.op1), nbn = slot_nnz(nb, t.op2);
                if (compact_ok && na >= cmin && nbn >= cmin && na * nbn >= nb4cthr2) cls[k] = 2;
                else cls[k] = (tile_ok && (mode == 2 || na * nbn >= nb4thr2)) ? 0 : 1;
            }
            auto key = [&](size_t k) -> const void* {
                pangulu_storage_slot_t* d = tasks[i + k].opdst;
                if (d->brow_pos == d->bcol_pos && d->is_upper) return d->related_block;
                return d;
            };
            // first appearance of each destination in the caller's order = its priority
            std::unordered_map<const void*, size_t> first_seen;
            for (size_t k = 0; k < nrun; ++k) first_seen.emplace(key(k), k);
            std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
                const size_t fa = first_seen[key(a)], fb = first_seen[key(b)];
                return fa != fb ? fa < fb : cls[a] < cls[b];
            });
            std::vector<SsssmDest<T>> by_cls[3];
            std::vector<pangulu_storage_slot_t*> cslot;  // destination slots of by_cls[2]
            const DiagGroupInfo* cur_groups = nullptr;    // groups of the current (diagonal) destination, if it has several
            // groups of a diagonal destination that an update touches = groups in which its left operand has rows (flags bits 8..11)
            auto touched_groups = [&](const pangulu_storage_slot_t* a, const DiagGroupInfo* gi) -> u32 {
                if (!gi || gi->ng <= 1) return 1u;
                u32 m = 0;
                const u32 nnz = a->columnpointer[nb];
                for (u32 q = 0; q < nnz; ++q)
                {
                    const u16 r = a->rowindex[q];
                    int gr = 0;
                    while (gr + 1 < gi->ng && r >= gi->bounds[gr + 1]) ++gr;
                    m |= 1u << gr;
                }
                return m;
            };
            const void* cur = nullptr;
            int cur_cls = -1;
            for (size_t k : order)
            {
                const pangulu_task_t& t = tasks[i + k];
                if (key(k) != cur || cls[k] != cur_cls)
                {
                    cur = key(k);
                    cur_cls = cls[k];
                    SsssmDest<T> D;
                    memset(&D, 0, sizeof D);
                    pangulu_storage_slot_t* d = t.opdst;
                    if (d->brow_pos == d->bcol_pos)
                    {
                        pangulu_storage_slot_t *lo, *up;
                        diag_halves(d, lo, up);
                        D.c = blk_csc(lo);
                        D.c.rptr = nullptr; D.c.ridx = nullptr; D.c.rmap = nullptr;
                        D.u = blk_upper_diag(up);
                        D.nnz_c = slot_nnz(nb, lo) + slot_nnz(nb, up);
                        cur_groups = groups_of(lo);
                    }
                    else
                    {
                        D.c = blk_csc(d);
                        D.nnz_c = slot_nnz(nb, d);
                        cur_groups = nullptr;
                    }
                    D.pair_begin = (u32)(cur_cls == 2 ? P.cpairs.size() : P.pairs.size());
                    D.pair_end = D.pair_begin;
                    by_cls[cur_cls].push_back(D);
                    if (cur_cls == 2) cslot.push_back(d);
                }
                by_cls[cur_cls].back().flags |= touched_groups(t.op1, cur_groups) << 8;
                if (cur_cls == 2)
                {
                    CPair cp;
                    cp.a = reg->get(t.op1, false);
                    cp.b = reg->get(t.op2, true);
                    P.cpairs.push_back(cp);
                    by_cls[cur_cls].back().pair_end = (u32)P.cpairs.size();
                    continue;
                }
                SsssmPair<T> pr;
                pr.a = blk_ro(t.op1);
                pr.b = blk_ro(t.op2);
                P.pairs.push_back(pr);
                by_cls[cur_cls].back().pair_end = (u32)P.pairs.size();
            }
            // Few, heavy destinations (the fronts of the top separators) would leave most SMs idle: split their
            // operand-pair ranges over several CTAs, which then combine with global FP64 reductions (flags bit 0).
            if (!by_cls[0].empty() && g.opt.split_pairs != 0)
            {
                const size_t tiles = (size_t)(nb + PSM_TILE_N - 1) / PSM_TILE_N;
                const size_t target = 2 * (size_t)g.sm_count;
                size_t total_pairs = 0;
                for (auto& D0 : by_cls[0]) total_pairs += D0.pair_end - D0.pair_begin;
                if (by_cls[0].size() * tiles < target && total_pairs > by_cls[0].size())
                {
                    const size_t per = std::max<size_t>(1, (total_pairs * tiles + target - 1) / target);
                    std::vector<SsssmDest<T>> split;
                    for (auto& D0 : by_cls[0])
                    {
                        const u32 np = D0.pair_end - D0.pair_begin;
                        if (np <= per) { split.push_back(D0); continue; }
                        for (u32 b0 = D0.pair_begin; b0 < D0.pair_end; b0 += (u32)per)
                        {
                            SsssmDest<T> piece = D0;
                            piece.pair_begin = b0;
                            piece.pair_end = std::min<u32>(D0.pair_end, b0 + (u32)per);
                            piece.flags |= 1u;
                            split.push_back(piece);
                        }
                    }
                    by_cls[0].swap(split);
                }
            }
            for (int c = 0; c < 3; ++c)
            {
                if (by_cls[c].empty()) continue;
                // per-wave launches: heaviest destinations first (the grid is consumed in order, this trims the tail);
                // persistent executor: keep the caller's priority order
                if (g.opt.executor != 1 && c != 2)  // (class 2 keeps its order: cslot runs parallel to it)
                    std::stable_sort(by_cls[c].begin(), by_cls[c].end(), [](const SsssmDest<T>& x, const SsssmDest<T>& y) {
                        return (x.pair_end - x.pair_begin) > (y.pair_end - y.pair_begin);
                    });
                Step st;
                st.kind = c == 0 ? K_SSSSM_TILE : c == 1 ? K_SSSSM_SPARSE : K_SSSSM_CTILE;
                st.off = P.dests.size();
                st.count = by_cls[c].size();
                if (c == 2 && g.opt.split_pairs != 0 && by_cls[2].size() * 4 < (size_t)4 * g.sm_count)
                {
                    // A thin wave with heavy destinations -- the fronts of the top separators, where each of a few dozen
                    // destinations has dozens of pending operand pairs -- would keep a handful of SMs busy for milliseconds on the
                    // critical path while the rest idles (measured: the last 39 block columns of the 100^3 benchmark took ~4 ms
                    // each at 4-60 % utilisation, profiles/r02_trace_tail_before_split.txt).  Cut the pair lists of such
                    // destinations into pieces that run on different SMs; they combine through the global FP64 reductions all
                    // compact tile tasks end with.
                    size_t total_pairs = 0;
                    for (auto& D0 : by_cls[2]) total_pairs += D0.pair_end - D0.pair_begin;
                    const size_t per = std::max<size_t>(2, (total_pairs * 4 + (size_t)4 * g.sm_count - 1) / ((size_t)4 * g.sm_count));
                    std::vector<SsssmDest<T>> split;
                    std::vector<pangulu_storage_slot_t*> split_slot;
                    for (size_t di = 0; di < by_cls[2].size(); ++di)
                    {
                        const SsssmDest<T>& D0 = by_cls[2][di];
                        const u32 np = D0.pair_end - D0.pair_begin;
                        if (np <= per + per / 2) { split.push_back(D0); split_slot.push_back(cslot[di]); continue; }
                        const u32 npiece = (u32)((np + per - 1) / per);
                        for (u32 q = 0; q < npiece; ++q)
                        {
                            SsssmDest<T> piece = D0;
                            piece.pair_begin = D0.pair_begin + (u32)((size_t)np * q / npiece);
                            piece.pair_end = D0.pair_begin + (u32)((size_t)np * (q + 1) / npiece);
                            piece.flags |= 1u;
                            split.push_back(piece);
                            split_slot.push_back(cslot[di]);
                        }
                    }
                    by_cls[2].swap(split);
                    cslot.swap(split_slot);
                    st.count = by_cls[2].size();
                }
                if (c == 2)
                {
                    // cut every destination into column ranges in which neither the destination nor any of its B operands has
                    // more than TN non-empty columns (a B operand may have columns the destination lacks: their products
                    // are structurally empty)
                    constexpr int TN = ct_traits<T>::TN;
                    st.aux_off = P.ctasks.size();
                    std::vector<u16> cpre((size_t)nb + 1);
                    for (size_t di = 0; di < by_cls[2].size(); ++di)
                    {
                        const SsssmDest<T>& D2 = by_cls[2][di];
                        pangulu_storage_slot_t* ds = cslot[di];
                        pangulu_storage_slot_t *lo = ds, *up = nullptr;
                        if (ds->brow_pos == ds->bcol_pos) diag_halves(ds, lo, up);
                        u16 run = 0;
                        for (int cc = 0; cc <= nb; ++cc)
                        {
                            cpre[cc] = run;
                            if (cc < nb)
                            {
                                bool ne = lo->columnpointer[cc + 1] > (cc ? lo->columnpointer[cc] : 0u);
                                if (up) ne = ne || up->columnpointer[cc + 1] > (cc ? up->columnpointer[cc] : 0u);
                                run += ne;
                            }
                        }
                        int s0 = 0;
                        while (s0 < nb && cpre[nb] > cpre[s0])
                        {
                            // largest end such that every count stays <= TN
                            int e = nb;
                            auto limit = [&](const u16* pre) {
                                // first column c with pre[c+1] - pre[s0] > TN  ->  range must end at c
                                const int target = pre[s0] + TN;
                                if (pre[nb] <= target) return;
                                int a = s0, b = nb;  // smallest c with pre[c] > target
                                while (a < b) { const int m2 = (a + b) >> 1; if (pre[m2] > target) b = m2; else a = m2 + 1; }
                                e = std::min(e, a - 1);
                            };
                            limit(cpre.data());
                            for (u32 pp = D2.pair_begin; pp < D2.pair_end; ++pp)
                            {
                                const DenseReg::Entry& eb = reg->entries[P.cpairs[pp].b];
                                limit(reg->meta.data() + eb.meta_off + eb.cnt + nb);
                            }
                            if (cpre[e] > cpre[s0]) P.ctasks.push_back({(int)(st.off + di), s0, e});
                            s0 = e;
                        }
                    }
                    st.aux_count = P.ctasks.size() - st.aux_off;
                }
                P.dests.insert(P.dests.end(), by_cls[c].begin(), by_cls[c].end());
                P.steps.push_back(st);
            }
            i = j;
        }
        else if (kid == PANGULU_TASK_GETRF)
        {
            size_t j = i;
            // diagonal blocks that fit in shared memory (values + pattern of both halves) take the staged kernel
            std::vector<GetrfTask<T>> fit, big, glob;
            std::vector<const pangulu_storage_slot_t*> fit_slot, big_slot, glob_slot;
            size_t fit_smem = 0;
            Step dl; dl.kind = -1; dl.off = P.downloads.size(); dl.count = 0;
            while (j < t1 && tasks[j].kernel_id == PANGULU_TASK_GETRF)
            {
                pangulu_storage_slot_t *lo, *up;
                diag_halves(tasks[j].opdst, lo, up);
                GetrfTask<T> G;
                G.lower = blk_csc(lo);
                G.upper = blk_upper_diag(up);
                G.c0 = 0; G.c1 = nb; G.grp = 0; G.pad = 0;
                const DiagGroupInfo* gi = groups_of(lo);
                if (gi && gi->ng > 1)
                {
                    G.grp = (int)(tasks[j].compare_flag >> 32);
                    if (G.grp < 0 || G.grp >= gi->ng) fatal(__FILE__, __LINE__, "GETRF task names a group the diagonal block does not have", nullptr);
                    G.c0 = gi->bounds[G.grp];
                    G.c1 = gi->bounds[G.grp + 1];
                }
                const bool grouped = G.c0 != 0 || G.c1 != nb;
                auto cb = [&](const pangulu_storage_slot_t* sl2, int c) { return c == 0 ? 0u : sl2->columnpointer[c]; };
                const size_t nl = lo->columnpointer[G.c1] - cb(lo, G.c0), nu = up->columnpointer[G.c1] - cb(up, G.c0);
                const size_t need = getrf_smem_bytes(nb, nl, nu);
                // getrf_mode: 0 auto, 1 never stage in shared memory, 2 never promote to dense, 3 always dense (nb <= 256)
                const bool want_dense = g.opt.getrf_mode == 3 && nb <= 256;
                if (!want_dense && need + 512 <= (size_t)g.max_smem_optin && g.opt.getrf_mode != 1) { fit.push_back(G); fit_smem = std::max(fit_smem, need); fit_slot.push_back(lo); }
                else if (grouped && (nb > 256 || g.opt.getrf_mode == 2)) { glob.push_back(G); glob_slot.push_back(lo); }
                else { big.push_back(G); big_slot.push_back(lo); }
                if (eager_download)
                {
                    // (a grouped task owns the entries of its index range only: columns [c0,c1) of the lower, rows [c0,c1) of the upper half)
                    P.downloads.push_back({(T*)lo->value + cb(lo, G.c0), (const T*)lo->d_value + cb(lo, G.c0), sizeof(T) * nl});
                    P.downloads.push_back({(T*)up->value + cb(up, G.c0), (const T*)up->d_value + cb(up, G.c0), sizeof(T) * nu});
                    dl.count += 2;
                }
                ++j;
            }
            if (!fit.empty())
            {
                Step st; st.kind = K_GETRF_SMEM; st.off = P.getrf.size(); st.count = fit.size(); st.smem = fit_smem;
                P.getrf.insert(P.getrf.end(), fit.begin(), fit.end());
                P.getrf_slot.insert(P.getrf_slot.end(), fit_slot.begin(), fit_slot.end());
                P.steps.push_back(st);
            }
            if (!big.empty())
            {
                // too full for shared memory: dense-promoted blocked LU (nb <= 256), else the global-memory sparse kernel
                Step st; st.kind = (nb <= 256 && g.opt.getrf_mode != 2) ? (int)K_GETRF_DENSE : (int)K_GETRF; st.off = P.getrf.size(); st.count = big.size(); st.smem = 0;
                P.getrf.insert(P.getrf.end(), big.begin(), big.end());
                P.getrf_slot.insert(P.getrf_slot.end(), big_slot.begin(), big_slot.end());
                P.steps.push_back(st);
            }
            if (!glob.empty())
            {
                Step st; st.kind = (int)K_GETRF; st.off = P.getrf.size(); st.count = glob.size(); st.smem = 0;
                P.getrf.insert(P.getrf.end(), glob.begin(), glob.end());
                P.getrf_slot.insert(P.getrf_slot.end(), glob_slot.begin(), glob_slot.end());
                P.steps.push_back(st);
            }
            if (dl.count) P.steps.push_back(dl);
            i = j;
        }
        else if (kid == PANGULU_TASK_TSTRF || kid == PANGULU_TASK_GESSM)
        {
            size_t j = i;
            Step st; st.kind = K_TSTRF; st.off = P.panel.size(); st.count = 0;
            Step dl; dl.kind = -1; dl.off = P.downloads.size(); dl.count = 0;
            while (j < t1 && (tasks[j].kernel_id == PANGULU_TASK_TSTRF || tasks[j].kernel_id == PANGULU_TASK_GESSM))
            {
                const pangulu_task_t& t = tasks[j];
                pangulu_storage_slot_t *lo, *up;
                diag_halves(t.op1, lo, up);
                PanelTask<T> Q;
                memset(&Q, 0, sizeof Q);
                Q.kind = t.kernel_id;
                Q.dense = -1;
                Q.b = blk_csc(t.opdst);
                {
                    // TSTRF: groups in which the block has columns; GESSM: groups in which it has rows
                    u32 m = 1u;
                    const DiagGroupInfo* gi = groups_of(lo);
                    if (gi && gi->ng > 1)
                    {
                        m = 0;
                        auto grp_of = [&](int x) { int gr = 0; while (gr + 1 < gi->ng && x >= gi->bounds[gr + 1]) ++gr; return gr; };
                        if (t.kernel_id == PANGULU_TASK_TSTRF)
                        {
                            for (int c = 0; c < nb; ++c)
                                if (t.opdst->columnpointer[c + 1] > (c ? t.opdst->columnpointer[c] : 0u)) m |= 1u << grp_of(c);
                        }
                        else
                        {
                            const u32 nnz = t.opdst->columnpointer[nb];
                            for (u32 q = 0; q < nnz; ++q) m |= 1u << grp_of(t.opdst->rowindex[q]);
                        }
                        if (m == 0) m = 1u;
                    }
                    Q.kind |= (int)(m << 8);
                }
                if (t.kernel_id == PANGULU_TASK_TSTRF)
                {
                    if (!Q.b.rptr) fatal(__FILE__, __LINE__, "TSTRF destination has no CSR mirror (must be a strictly-lower block)", nullptr);
                    Q.diag = blk_upper_diag(up);
                }
                else
                {
                    Q.diag = blk_csc(lo);
                }
                P.panel.push_back(Q);
                P.panel_slot.push_back(t.opdst);
                P.panel_diag_slot.push_back(lo);
                st.count++;
                if (eager_download)
                {
                    P.downloads.push_back({t.opdst->value, t.opdst->d_value, sizeof(T) * (size_t)slot_nnz(nb, t.opdst)});
                    dl.count++;
                }
                ++j;
            }
            P.steps.push_back(st);
            if (reg)
            {
                // filled by the plan's post-pass: the panel tasks of this run that take the compact dense path
                Step cs; cs.kind = K_CPANEL; cs.off = st.off; cs.count = 0;
                P.steps.push_back(cs);
            }
            if (dl.count) P.steps.push_back(dl);
            i = j;
        }
        else
        {
            char msg[64];
            snprintf(msg, sizeof msg, "kernel id %d", kid);
            fatal(__FILE__, __LINE__, "invalid kernel id in task array", msg);
        }
    }
}

// ---- launchers ---------------------------------------------------------------------------------------
template <class K> void set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024)
    {
        if ((int)bytes > g.max_smem_optin) fatal(__FILE__, __LINE__, "block size nb needs more shared memory than the device offers; reduce init_options.nb", nullptr);
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
}

// Per-launch device time (option "profile"): two events around every launch, recorded on the module's
// stream and NOT synchronised here -- they are resolved when the statistics are read, so profiling does
// not serialise host and device.
struct EvPair { cudaEvent_t a, b; int cls; };
std::vector<EvPair> g_pending;
std::vector<cudaEvent_t> g_free_events;
cudaEvent_t get_event()
{
    if (!g_free_events.empty()) { cudaEvent_t e = g_free_events.back(); g_free_events.pop_back(); return e; }
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    return e;
}
void resolve_timers()
{
    for (auto& p : g_pending)
    {
        CK(cudaEventSynchronize(p.b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, p.a, p.b));
        g.ms[p.cls] += ms;
        g_free_events.push_back(p.a);
        g_free_events.push_back(p.b);
    }
    g_pending.clear();
}
struct Timer
{
    EvPair p;
    bool on;
    Timer(int c) : on(g.opt.profile != 0)
    {
        if (on) { p.cls = c; p.a = get_event(); p.b = get_event(); CK(cudaEventRecord(p.a, g.stream)); }
    }
    ~Timer()
    {
        if (on) { CK(cudaEventRecord(p.b, g.stream)); g_pending.push_back(p); }
    }
};

template <int WARPS> void launch_ssssm_sparse_w(int nb, const SsssmDest<T>* d, const SsssmPair<T>* p, size_t ndest)
{
    const int groups = (nb + WARPS - 1) / WARPS;
    const size_t smem = (size_t)WARPS * nb * sizeof(T);
    set_smem(ssssm_sparse_kernel<T, WARPS>, smem);
    // grid.x is limited to 2^31-1; batches stay far below (ndest * groups)
    ssssm_sparse_kernel<T, WARPS><<<(unsigned)(ndest * groups), WARPS * PSM_WARP, smem, g.stream>>>(nb, d, p, groups);
    CK_LAST();
}
void launch_ssssm_sparse(int nb, const SsssmDest<T>* d, const SsssmPair<T>* p, size_t ndest)
{
    if (!ndest) return;
    Timer tm(K_SSSSM_SPARSE);
    const size_t per_warp = (size_t)nb * sizeof(T);
    if (8 * per_warp <= (size_t)g.max_smem_optin / 2) launch_ssssm_sparse_w<8>(nb, d, p, ndest);
    else if (4 * per_warp <= (size_t)g.max_smem_optin) launch_ssssm_sparse_w<4>(nb, d, p, ndest);
    else launch_ssssm_sparse_w<1>(nb, d, p, ndest);
    g.launches[K_SSSSM_SPARSE]++;
}

template <int TM> void launch_ssssm_tile_t(int nb, const SsssmDest<T>* d, const SsssmPair<T>* p, size_t ndest)
{
    const int tiles = (nb + PSM_TILE_N - 1) / PSM_TILE_N;
    const size_t a = (size_t)2 * PSM_TILE_K * (TM + 4 + PSM_TILE_N + 4), c = (size_t)PSM_TILE_N * (TM + 4);
    const size_t smem = std::max(a, c) * sizeof(T);
    set_smem(ssssm_tile_kernel<T, TM>, smem);
    ssssm_tile_kernel<T, TM><<<(unsigned)(ndest * tiles), 256, smem, g.stream>>>(nb, d, p, tiles);
    CK_LAST();
}
template <int TM> void launch_ssssm_tile2_t(int nb, const SsssmDest<T>* d, const SsssmPair<T>* p, size_t ndest)
{
    const int tiles = (nb + PSM_TILE_N - 1) / PSM_TILE_N;
    const size_t a = (size_t)2 * PSM_TILE_K * (TM + 4 + PSM_TILE_N + 4), c = (size_t)PSM_TILE_N * (TM + 4);
    const size_t smem = std::max(a, c) * sizeof(T) + 4 * (size_t)(TM + 4) + 4 * (size_t)PSM_TILE_N * (TM / PSM_TILE_K + 1) + 64;
    set_smem(ssssm_tile2_kernel<T, TM>, smem);
    ssssm_tile2_kernel<T, TM><<<(unsigned)(ndest * tiles), 512, smem, g.stream>>>(nb, d, p, tiles);
    CK_LAST();
}
void launch_ssssm_tile(int nb, const SsssmDest<T>* d, const SsssmPair<T>* p, size_t ndest)
{
    if (!ndest) return;
    Timer tm(K_SSSSM_TILE);
    if (g.opt.tile_variant != 1)
    {
        if (nb <= 128) launch_ssssm_tile2_t<128>(nb, d, p, ndest);
        else launch_ssssm_tile2_t<256>(nb, d, p, ndest);
        g.launches[K_SSSSM_TILE]++;
        return;
    }
    if (nb <= 64) launch_ssssm_tile_t<64>(nb, d, p, ndest);
    else if (nb <= 128) launch_ssssm_tile_t<128>(nb, d, p, ndest);
    else launch_ssssm_tile_t<256>(nb, d, p, ndest);
    g.launches[K_SSSSM_TILE]++;
}

void launch_ssssm_ctile(int nb, const SsssmDest<T>* d, const CPair* cp, const DenseOp<T>* ops, const CTileTask* ctasks, size_t ntile)
{
    if (!ntile) return;
    Timer tm(K_SSSSM_CTILE);
    constexpr int TN = ct_traits<T>::TN;
    const size_t smem = CtLayout<T, TN, ct_traits<T>::STAGES>::total;
    set_smem(ssssm_ctile_kernel<T>, smem);
    ssssm_ctile_kernel<T><<<(unsigned)ntile, 512, smem, g.stream>>>(nb, d, cp, ops, ctasks);
    CK_LAST();
    g.launches[K_SSSSM_CTILE]++;
}

void launch_cpanel(int nb, const PanelTask<T>* tasks, const CPanelTask* cpt, size_t n, const DenseOp<T>* ops, const DiagDense<T>* dds)
{
    if (!n) return;
    Timer tm(K_CPANEL);
    const size_t smem = CpLayout<T>::total;
    set_smem(panel_compact_kernel<T>, smem);
    panel_compact_kernel<T><<<(unsigned)n, 512, smem, g.stream>>>(nb, tasks, cpt, ops, dds);
    CK_LAST();
    g.launches[K_CPANEL]++;
}
void launch_diag_densify(int nb, const GetrfTask<T>* t, const int* getrf_dd, const DiagDense<T>* dds, size_t n)
{
    if (!n || !dds) return;
    diag_densify_kernel<T><<<(unsigned)n, 256, 0, g.stream>>>(nb, t, getrf_dd, dds);
    CK_LAST();
    g.launches[K_AUX]++;
}

template <int WARPS> void launch_panel_w(int nb, const PanelTask<T>* t, size_t n, const DenseOp<T>* ops)
{
    const int groups = (nb + WARPS - 1) / WARPS;
    const size_t smem = (size_t)WARPS * nb * sizeof(T);
    set_smem(panel_solve_kernel<T, WARPS>, smem);
    panel_solve_kernel<T, WARPS><<<(unsigned)(n * groups), WARPS * PSM_WARP, smem, g.stream>>>(nb, t, groups, ops);
    CK_LAST();
}
void launch_panel(int nb, const PanelTask<T>* t, size_t n, const DenseOp<T>* ops)
{
    if (!n) return;
    Timer tm(K_TSTRF);
    const size_t per_warp = (size_t)nb * sizeof(T);
    if (8 * per_warp <= (size_t)g.max_smem_optin / 2) launch_panel_w<8>(nb, t, n, ops);
    else if (4 * per_warp <= (size_t)g.max_smem_optin) launch_panel_w<4>(nb, t, n, ops);
    else launch_panel_w<1>(nb, t, n, ops);
    g.launches[K_TSTRF]++;
}

void launch_getrf(int nb, const GetrfTask<T>* t, size_t n)
{
    if (!n) return;
    Timer tm(K_GETRF);
    constexpr int THREADS = 512;
    const size_t smem = (size_t)4 * nb * sizeof(T);
    set_smem(getrf_sparse_kernel<T, THREADS>, smem);
    getrf_sparse_kernel<T, THREADS><<<(unsigned)n, THREADS, smem, g.stream>>>(nb, t);
    CK_LAST();
    g.launches[K_GETRF]++;
}
// L2-resident scratch of the dense-promoted GETRF (grow-only; growing waits for the stream to drain)
T* g_scratch = nullptr;
size_t g_scratch_bytes = 0;
T* ensure_scratch(size_t bytes)
{
    if (bytes > g_scratch_bytes)
    {
        CK(cudaStreamSynchronize(g.stream));
        if (g_scratch) CK(cudaFree(g_scratch));
        g_scratch_bytes = std::max(bytes, (size_t)8 << 20);
        CK(cudaMalloc((void**)&g_scratch, g_scratch_bytes));
    }
    return g_scratch;
}
void launch_getrf_dense(int nb, const GetrfTask<T>* t, size_t n)
{
    if (!n) return;
    const int nbp = (nb + 31) & ~31;
    T* scratch = ensure_scratch(n * (size_t)nbp * nbp * sizeof(T));
    Timer tm(K_GETRF);
    const size_t smem = ((size_t)2 * PSM_GETRF_W * (nbp + 4) + 64) * sizeof(T);
    set_smem(getrf_dense_kernel<T>, smem);
    getrf_dense_kernel<T><<<(unsigned)n, 512, smem, g.stream>>>(nb, t, scratch);
    CK_LAST();
    g.launches[K_GETRF]++;
}
void launch_getrf_smem(int nb, const GetrfTask<T>* t, size_t n, size_t smem)
{
    if (!n) return;
    Timer tm(K_GETRF);
    constexpr int THREADS = 256;
    set_smem(getrf_smem_kernel<T, THREADS>, smem);
    getrf_smem_kernel<T, THREADS><<<(unsigned)n, THREADS, smem, g.stream>>>(nb, t);
    CK_LAST();
    g.launches[K_GETRF]++;
}

void run_steps(int nb, const Program& P, const DevProgram& D, size_t s0, size_t s1)
{
    for (size_t s = s0; s < s1; ++s)
    {
        const Step& st = P.steps[s];
        switch (st.kind)
        {
        case K_SSSSM_SPARSE: launch_ssssm_sparse(nb, D.dests + st.off, D.pairs, st.count); break;
        case K_SSSSM_TILE: launch_ssssm_tile(nb, D.dests + st.off, D.pairs, st.count); break;
        case K_SSSSM_CTILE: launch_ssssm_ctile(nb, D.dests, D.cpairs, D.ops, D.ctasks + st.aux_off, st.aux_count); break;
        case K_GETRF: launch_getrf(nb, D.getrf + st.off, st.count); launch_diag_densify(nb, D.getrf + st.off, D.getrf_dd ? D.getrf_dd + st.off : nullptr, D.dds, st.count); break;
        case K_GETRF_SMEM: launch_getrf_smem(nb, D.getrf + st.off, st.count, st.smem); launch_diag_densify(nb, D.getrf + st.off, D.getrf_dd ? D.getrf_dd + st.off : nullptr, D.dds, st.count); break;
        case K_GETRF_DENSE: launch_getrf_dense(nb, D.getrf + st.off, st.count); launch_diag_densify(nb, D.getrf + st.off, D.getrf_dd ? D.getrf_dd + st.off : nullptr, D.dds, st.count); break;
        case K_CPANEL: launch_cpanel(nb, D.panel, D.cptasks + st.aux_off, st.aux_count, D.ops, D.dds); break;
        case K_TSTRF: launch_panel(nb, D.panel + st.off, st.count, D.ops); break;
        case -1:
            for (size_t k = st.off; k < st.off + st.count; ++k)
                CK(cudaMemcpyAsync(P.downloads[k].h, P.downloads[k].d, P.downloads[k].bytes, cudaMemcpyDeviceToHost, g.stream));
            g.launches[K_AUX] += (long long)st.count;
            break;
        default: fatal(__FILE__, __LINE__, "corrupt step list", nullptr);
        }
    }
}

template <class V> size_t vbytes(const V& v) { return (v.size() * sizeof(typename V::value_type) + 255) & ~(size_t)255; }

inline u32 R_cnt(const DenseReg& R, int op) { return R.entries[op].cnt; }

// one-shot execution of a task array (the hybrid_batched / getrf / tstrf / ... entry points)
void execute_tasks(int nb, size_t ntask, const pangulu_task_t* tasks)
{
    if (!ntask) return;
    ensure_ctx();
    Program P;
    append_tasks(P, nb, tasks, 0, ntask, g.opt.eager_download != 0);
    const size_t b0 = vbytes(P.getrf), b1 = vbytes(P.panel), b2 = vbytes(P.dests), b3 = vbytes(P.pairs);
    Slab& sl = take_slab(b0 + b1 + b2 + b3 + 256);
    char* h = sl.h;
    if (!P.getrf.empty()) memcpy(h, P.getrf.data(), P.getrf.size() * sizeof(GetrfTask<T>));
    if (!P.panel.empty()) memcpy(h + b0, P.panel.data(), P.panel.size() * sizeof(PanelTask<T>));
    if (!P.dests.empty()) memcpy(h + b0 + b1, P.dests.data(), P.dests.size() * sizeof(SsssmDest<T>));
    if (!P.pairs.empty()) memcpy(h + b0 + b1 + b2, P.pairs.data(), P.pairs.size() * sizeof(SsssmPair<T>));
    CK(cudaMemcpyAsync(sl.d, sl.h, b0 + b1 + b2 + b3, cudaMemcpyHostToDevice, g.stream));
    DevProgram D;
    D.getrf = (GetrfTask<T>*)sl.d;
    D.panel = (PanelTask<T>*)(sl.d + b0);
    D.dests = (SsssmDest<T>*)(sl.d + b0 + b1);
    D.pairs = (SsssmPair<T>*)(sl.d + b0 + b1 + b2);
    run_steps(nb, P, D, 0, P.steps.size());
    CK(cudaEventRecord(sl.done, g.stream));
}

void single_task(int nb, int kernel_id, pangulu_storage_slot_t* opdst, pangulu_storage_slot_t* op1, pangulu_storage_slot_t* op2)
{
    pangulu_task_t t;
    memset(&t, 0, sizeof t);
    t.kernel_id = (pangulu_int16_t)kernel_id;
    t.opdst = opdst;
    t.op1 = op1;
    t.op2 = op2;
    t.row = opdst->brow_pos;
    t.col = opdst->bcol_pos;
    execute_tasks(nb, 1, &t);
}

}  // namespace

// the plan type is visible to C callers only as an opaque pointer
struct pangulu_sm100_plan
{
    int nb;
    Program P;
    DevProgram D;
    std::vector<size_t> wave_step;  // wave w -> first step
    void* dev_buf;
    // compact dense operand copies
    DenseReg reg;
    void* dense_buf = nullptr;   // values (zero-filled once; structural zeros stay zero)
    void* dense_meta = nullptr;  // row / column lists and inverse maps
    std::vector<const void*> op_val;  // per compact copy: device value array of the block it mirrors (sync id key)
    void* diag_dense_buf = nullptr;   // dense triangular copies of diagonal blocks (compact panel path)
    std::unordered_map<const void*, size_t> dense_of_val;  // device value array -> compact copy index
    // dense triangular copies of REMOTE diagonal blocks: refreshed on this rank by a task that waits for the mirrored
    // halves (one per independent group)
    struct DDens { int getrf_index, dd, group; bool need_u, need_lt; const pangulu_storage_slot_t* lower; };
    std::vector<DDens> ddens;
    std::unordered_map<const pangulu_storage_slot_t*, std::array<int, 4>> ddens_of;  // lower slot -> ddens index per group
    double stat_ct_pairtiles = 0, stat_ct_chunks = 0, stat_ct_dmma = 0, stat_ct_abytes = 0, stat_ct_tasks = 0, stat_ct_msum = 0, stat_ct_ntsum = 0;
    size_t dense_total_bytes = 0;
    std::vector<float> ct_task_dmma;     // per compact-tile task: DMMA instructions (index = position in Program::ctasks)
    std::vector<float> host_task_work;   // per task of the persistent executor: DMMA instructions (compact tiles), else 0
    // persistent executor
    bool mega = false;
    void* mega_buf = nullptr;
    MegaParams<T> mp;
    int* remaining_init = nullptr;
    int* done_init = nullptr;
    int* pending_init = nullptr;
    size_t nids = 0;
    size_t mega_bytes = 0;
    std::vector<MegaTask> host_tasks;
    // multi-GPU wiring
    std::unordered_map<const void*, int> ids;   // device value array -> sync id
    std::vector<int> h_done_init;
    std::vector<std::vector<MegaPush<T>>> h_push;  // per sync id
    void* push_buf = nullptr;
};

namespace {

// Translate the step list of a plan into the task list of the persistent executor (sm100_kernels.cuh,
// mega_kernel): one task per (destination, 32-column tile) or (destination, 64-column group) of an SSSSM
// step, one per GETRF / TSTRF / GESSM; synchronisation ids are assigned per distinct value array.
void build_mega(pangulu_sm100_plan* pl)
{
    Program& P = pl->P;
    const int nb = pl->nb;
    std::unordered_map<const void*, int>& ids = pl->ids;
    auto id_of = [&](const void* p) {
        auto it = ids.find(p);
        if (it != ids.end()) return it->second;
        const int id = (int)ids.size();
        ids.emplace(p, id);
        return id;
    };
    std::vector<MegaTask> tasks;
    std::vector<int> remaining, pending;
    std::vector<int4> dest_id(P.dests.size(), make_int4(0, -1, -1, -1)), panel_dep(P.panel.size(), make_int4(0, -1, -1, -1));
    std::vector<int2> pair_dep(P.pairs.size()), getrf_id(P.getrf.size()), panel_id(P.panel.size()), cpair_dep(P.cpairs.size());
    // sync id of group gr of a diagonal half: the value pointer plus gr (value arrays are 16-byte aligned, gr < 4)
    auto gkey = [](const void* val, int gr) -> const void* { return (const char*)val + gr; };
    std::vector<char> produced;
    auto touch = [&](int id) {
        if ((size_t)id >= remaining.size()) { remaining.resize(id + 1, 0); produced.resize(id + 1, 0); pending.resize(id + 1, 0); }
    };
    for (size_t p = 0; p < P.pairs.size(); ++p)
    {
        pair_dep[p] = make_int2(id_of(P.pairs[p].a.val), id_of(P.pairs[p].b.val));
        touch(pair_dep[p].x);
        touch(pair_dep[p].y);
    }
    for (size_t p = 0; p < P.cpairs.size(); ++p)
    {
        cpair_dep[p] = make_int2(id_of(pl->op_val[P.cpairs[p].a]), id_of(pl->op_val[P.cpairs[p].b]));
        touch(cpair_dep[p].x);
        touch(cpair_dep[p].y);
    }
    const int tiles = (nb + PSM_TILE_N - 1) / PSM_TILE_N;
    const int sgroup = 64, sgroups = (nb + sgroup - 1) / sgroup;
    for (const Step& st : P.steps)
    {
        switch (st.kind)
        {
        case K_SSSSM_TILE:
        case K_SSSSM_SPARSE:
        case K_SSSSM_CTILE:
            for (size_t d = st.off; d < st.off + st.count; ++d)
            {
                // a diagonal destination counts towards every group its updates touch (flags bits 8..11), any other towards itself
                int ids4[4] = {-1, -1, -1, -1};
                int nid4 = 0;
                const u32 gm = P.dests[d].u.val ? ((P.dests[d].flags >> 8) & 0xfu) : 1u;
                for (int gr = 0; gr < 4; ++gr)
                    if ((gm >> gr) & 1u) { ids4[nid4] = id_of(gkey(P.dests[d].c.val, gr)); touch(ids4[nid4]); ++nid4; }
                if (nid4 == 0) { ids4[0] = id_of(P.dests[d].c.val); touch(ids4[0]); nid4 = 1; }
                dest_id[d] = make_int4(ids4[0], ids4[1], ids4[2], ids4[3]);
                auto count = [&](int n) { for (int q = 0; q < nid4; ++q) remaining[ids4[q]] += n; };
                if (st.kind == K_SSSSM_CTILE)
                    ;  // tile tasks below
                else if (st.kind == K_SSSSM_TILE)
                {
                    for (int t = 0; t < tiles; ++t) tasks.push_back({MEGA_TILE, (int)d, t, (int)(P.dests[d].pair_end - P.dests[d].pair_begin)});
                    count(tiles);
                }
                else
                {
                    for (int t = 0; t < sgroups; ++t) tasks.push_back({MEGA_SPARSE, (int)d, t * sgroup, sgroup});
                    count(sgroups);
                }
            }
            if (st.kind == K_SSSSM_CTILE)
                for (size_t c = st.aux_off; c < st.aux_off + st.aux_count; ++c)
                {
                    const CTileTask& ct = P.ctasks[c];
                    tasks.push_back({MEGA_CTILE, ct.dest, ct.clo | (ct.chi << 16), (int)(P.dests[ct.dest].pair_end - P.dests[ct.dest].pair_begin)});
                    pl->host_task_work.resize(tasks.size(), 0.0f);
                    if (c < pl->ct_task_dmma.size()) pl->host_task_work.back() = pl->ct_task_dmma[c];
                    const int4 di = dest_id[ct.dest];
                    remaining[di.x] += 1;
                    if (di.y >= 0) remaining[di.y] += 1;
                    if (di.z >= 0) remaining[di.z] += 1;
                    if (di.w >= 0) remaining[di.w] += 1;
                }
            break;
        case K_GETRF:
        case K_GETRF_SMEM:
        case K_GETRF_DENSE:
            for (size_t gi = st.off; gi < st.off + st.count; ++gi)
            {
                const int il = id_of(gkey(P.getrf[gi].lower.val, P.getrf[gi].grp)), iu = id_of(gkey(P.getrf[gi].upper.val, P.getrf[gi].grp));
                touch(il);
                touch(iu);
                getrf_id[gi] = make_int2(il, iu);
                produced[il] = produced[iu] = 1;
                tasks.push_back({st.kind == K_GETRF_SMEM ? MEGA_GETRF_SMEM : st.kind == K_GETRF_DENSE ? MEGA_GETRF_DENSE : MEGA_GETRF_GLOBAL, (int)gi, 0, 0});
            }
            break;
        case K_TSTRF:
        case K_CPANEL:
            for (size_t ti = st.off; ti < st.off + st.count; ++ti)
            {
                const int ib = id_of(P.panel[ti].b.val);
                touch(ib);
                int dp[4] = {-1, -1, -1, -1}, ndp = 0;
                const u32 gm = ((u32)P.panel[ti].kind >> 8) & 0xfu;
                for (int gr = 0; gr < 4; ++gr)
                    if ((gm >> gr) & 1u) { dp[ndp] = id_of(gkey(P.panel[ti].diag.val, gr)); touch(dp[ndp]); ++ndp; }
                if (ndp == 0) { dp[0] = id_of(P.panel[ti].diag.val); touch(dp[0]); }
                if (st.kind == K_CPANEL && !P.panel_diag_slot.empty())
                {
                    // the diagonal block lives on another rank: depend on the local refresh of its dense triangular copies
                    auto rd = pl->ddens_of.find(P.panel_diag_slot[ti]);
                    if (rd != pl->ddens_of.end())
                    {
                        ndp = 0;
                        dp[0] = dp[1] = dp[2] = dp[3] = -1;
                        for (int gr = 0; gr < 4; ++gr)
                        {
                            if (!((gm >> gr) & 1u) || rd->second[gr] < 0) continue;
                            pangulu_sm100_plan::DDens& de = pl->ddens[rd->second[gr]];
                            const int ready = id_of((const char*)&de);
                            touch(ready);
                            if (!produced[ready])
                            {
                                produced[ready] = 1;
                                const GetrfTask<T>& Gd = P.getrf[de.getrf_index];
                                const int wl = de.need_lt ? id_of(gkey(Gd.lower.val, gr)) : -1, wu = de.need_u ? id_of(gkey(Gd.upper.val, gr)) : -1;
                                if (wl >= 0) touch(wl);
                                if (wu >= 0) touch(wu);
                                getrf_id[de.getrf_index] = make_int2(wl, wu);
                                tasks.push_back({MEGA_DDENS, de.getrf_index, ready, 0});
                            }
                            dp[ndp++] = ready;
                        }
                    }
                }
                panel_dep[ti] = make_int4(dp[0], dp[1], dp[2], dp[3]);
                panel_id[ti] = make_int2(ib, dp[0]);
                produced[ib] = 1;
                if (st.kind == K_TSTRF)
                    for (int r0 = 0; r0 < nb; r0 += 64) { tasks.push_back({MEGA_PANEL, (int)ti, r0, 0}); pending[ib]++; }
            }
            if (st.kind == K_CPANEL)
                for (size_t c = st.aux_off; c < st.aux_off + st.aux_count; ++c)
                {
                    tasks.push_back({MEGA_CPANEL, (int)c, 0, 0});
                    pending[panel_id[P.cptasks[c].panel].x]++;
                }
            break;
        default: break;  // downloads are not part of a plan
        }
    }
    const size_t nid = remaining.size();
    std::vector<int> done_init(nid);
    for (size_t i = 0; i < nid; ++i) done_init[i] = produced[i] ? 0 : 1;  // blocks no task of this plan produces are final already

    const int nbp = (nb + 31) & ~31;
    const size_t sz_tasks = vbytes(tasks), sz_pd = vbytes(pair_dep), sz_di = vbytes(dest_id), sz_gi = vbytes(getrf_id), sz_pi = vbytes(panel_id), sz_cd = vbytes(cpair_dep), sz_pp = vbytes(panel_dep);
    const size_t sz_ids = (nid * sizeof(int) + 255) & ~(size_t)255;
    const size_t sz_scratch = ((size_t)g.sm_count * nbp * nbp * sizeof(T) + 255) & ~(size_t)255;
    const size_t sz_trace = g.opt.trace ? ((tasks.size() * 3 * sizeof(unsigned long long) + 255) & ~(size_t)255) : 0;
    const size_t total = sz_tasks + sz_pd + sz_di + sz_gi + sz_pi + sz_cd + sz_pp + 6 * sz_ids + 256 + sz_scratch + sz_trace;
    char* d = nullptr;
    CK(cudaMalloc((void**)&d, total));
    pl->mega_buf = d;
    char* q = d;
    auto put = [&](const void* src, size_t bytes, size_t padded) {
        char* at = q;
        if (bytes) CK(cudaMemcpy(at, src, bytes, cudaMemcpyHostToDevice));
        q += padded;
        return at;
    };
    MegaParams<T>& mp = pl->mp;
    mp.nb = nb;
    mp.ntask = (int)tasks.size();
    mp.tasks = (const MegaTask*)put(tasks.data(), tasks.size() * sizeof(MegaTask), sz_tasks);
    mp.pair_dep = (const int2*)put(pair_dep.data(), pair_dep.size() * sizeof(int2), sz_pd);
    mp.dest_id = (const int4*)put(dest_id.data(), dest_id.size() * sizeof(int4), sz_di);
    mp.getrf_id = (const int2*)put(getrf_id.data(), getrf_id.size() * sizeof(int2), sz_gi);
    mp.panel_id = (const int2*)put(panel_id.data(), panel_id.size() * sizeof(int2), sz_pi);
    mp.cpair_dep = (const int2*)put(cpair_dep.data(), cpair_dep.size() * sizeof(int2), sz_cd);
    mp.panel_dep = (const int4*)put(panel_dep.data(), panel_dep.size() * sizeof(int4), sz_pp);
    mp.cpairs = pl->D.cpairs;
    mp.ops = pl->D.ops;
    mp.cptasks = pl->D.cptasks;
    mp.dds = pl->D.dds;
    mp.getrf_dd = pl->D.getrf_dd;
    pl->remaining_init = (int*)put(remaining.data(), nid * sizeof(int), sz_ids);
    pl->done_init = (int*)put(done_init.data(), nid * sizeof(int), sz_ids);
    pl->pending_init = (int*)put(pending.data(), nid * sizeof(int), sz_ids);
    mp.remaining = (int*)q; q += sz_ids;
    mp.done = (int*)q; q += sz_ids;
    mp.pending = (int*)q; q += sz_ids;
    mp.counter = (int*)q; q += 256;
    mp.scratch = (T*)q; q += sz_scratch;
    mp.trace = sz_trace ? (unsigned long long*)q : nullptr;
    pl->host_tasks = tasks;
    pl->host_task_work.resize(tasks.size(), 0.0f);
    mp.dests = pl->D.dests;
    mp.pairs = pl->D.pairs;
    mp.getrf = pl->D.getrf;
    mp.panel = pl->D.panel;
    mp.push_ptr = nullptr;
    mp.pushes = nullptr;
    mp.multi = 0;
    pl->nids = nid;
    pl->mega_bytes = total;
    pl->h_done_init = done_init;
    pl->h_push.assign(nid, {});
    pl->mega = true;
}

void reset_mega(pangulu_sm100_plan* pl)
{
    MegaParams<T>& mp = pl->mp;
    if (pl->nids == 0) return;
    CK(cudaMemcpyAsync(mp.remaining, pl->remaining_init, pl->nids * sizeof(int), cudaMemcpyDeviceToDevice, g.stream));
    CK(cudaMemcpyAsync(mp.done, pl->done_init, pl->nids * sizeof(int), cudaMemcpyDeviceToDevice, g.stream));
    CK(cudaMemcpyAsync(mp.pending, pl->pending_init, pl->nids * sizeof(int), cudaMemcpyDeviceToDevice, g.stream));
    CK(cudaMemsetAsync(mp.counter, 0, sizeof(int), g.stream));
}

void launch_mega(pangulu_sm100_plan* pl)
{
    MegaParams<T>& mp = pl->mp;
    if (mp.ntask == 0) return;
    Timer tm(K_MEGA);
    const size_t smem = (size_t)g.max_smem_optin - 512;  // (static shared memory of the kernel: tickets, mbarriers)
    pl->mp.smem_bytes = smem;  // whole SM: the shared-memory GETRF variant sizes itself against this
    // the sparse SSSSM / panel tasks keep one dense vector of nb values per warp (16 warps) in shared memory
    if ((size_t)16 * pl->nb * sizeof(T) + 4096 > smem)
        fatal(__FILE__, __LINE__, "block size nb is too large for the persistent executor's per-warp shared-memory vectors; reduce init_options.nb or set option executor=0", nullptr);
    const int grid = g.sm_count;
    mp.progress = nullptr;
    if (g.opt.progress)
    {
        if (!g_progress_host)
        {
            CK(cudaHostAlloc((void**)&g_progress_host, sizeof(int4) * 1024, cudaHostAllocMapped));
            g_progress_ctas = grid;
        }
        for (int b = 0; b < 1024; ++b) g_progress_host[b] = b < 512 ? make_int4(-1, -1, 0, 2) : make_int4(-1, -1, -1, -1);
        CK(cudaHostGetDevicePointer((void**)&mp.progress, g_progress_host, 0));
    }
    {
        const unsigned long long lim = (unsigned long long)std::max<long long>(1, g.opt.spin_timeout_ms) * 1000000ull;
        const int zero_flag = 0;
        CK(cudaMemcpyToSymbolAsync(g_spin_limit_ns, &lim, sizeof lim, 0, cudaMemcpyHostToDevice, g.stream));
        CK(cudaMemcpyToSymbolAsync(g_abort, &zero_flag, sizeof zero_flag, 0, cudaMemcpyHostToDevice, g.stream));
        const int gmax = (int)std::min<long long>(16, std::max<long long>(1, g.opt.ct_gmax)), minst = (int)std::min<long long>(8, std::max<long long>(2, g.opt.ct_minst));
        CK(cudaMemcpyToSymbolAsync(g_ct_gmax, &gmax, sizeof gmax, 0, cudaMemcpyHostToDevice, g.stream));
        CK(cudaMemcpyToSymbolAsync(g_ct_minst, &minst, sizeof minst, 0, cudaMemcpyHostToDevice, g.stream));
        const int defer = (int)std::min<long long>(PSM_DEFER_MAX, std::max<long long>(0, g.opt.defer));
        CK(cudaMemcpyToSymbolAsync(g_defer_depth, &defer, sizeof defer, 0, cudaMemcpyHostToDevice, g.stream));
    }
    if (pl->nb <= 128)
    {
        CK(cudaFuncSetAttribute(mega_kernel<T, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mega_kernel<T, 128><<<grid, 512, smem, g.stream>>>(mp);
    }
    else
    {
        CK(cudaFuncSetAttribute(mega_kernel<T, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mega_kernel<T, 256><<<grid, 512, smem, g.stream>>>(mp);
    }
    CK_LAST();
    g.launches[K_MEGA]++;
}

void run_mega(pangulu_sm100_plan* pl)
{
    reset_mega(pl);
    launch_mega(pl);
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------------
// Part 1: PanguLU platform interface
// ------------------------------------------------------------------------------------------------------
void PSM(malloc)(void** platform_address, size_t size) { CK(cudaMalloc(platform_address, size)); }

void PSM(malloc_pinned)(void** platform_address, size_t size) { CK(cudaHostAlloc(platform_address, size, cudaHostAllocDefault)); }

void PSM(synchronize)(void) { CK(cudaDeviceSynchronize()); }

void PSM(memset)(void* s, int c, size_t n) { CK(cudaMemset(s, c, n)); }

void PSM(create_stream)(void** stream) { CK(cudaStreamCreate((cudaStream_t*)stream)); }

static cudaMemcpyKind kind_of(unsigned int kind)
{
    switch (kind)
    {
    case 0: return cudaMemcpyHostToDevice;
    case 1: return cudaMemcpyDeviceToHost;
    case 2: return cudaMemcpyDeviceToDevice;
    }
    fprintf(stderr, "Invalid memcpy kind value\n");
    exit(EXIT_FAILURE);
}

void PSM(memcpy)(void* dst, const void* src, size_t count, unsigned int kind) { CK(cudaMemcpy(dst, src, count, kind_of(kind))); }

void PSM(memcpy_async)(void* dst, const void* src, size_t count, unsigned int kind, void* stream)
{
    CK(cudaMemcpyAsync(dst, src, count, kind_of(kind), (cudaStream_t)stream));
}

void PSM(free)(void* devptr) { CK(cudaFree(devptr)); }

void PSM(get_device_num)(int* device_num) { CK(cudaGetDeviceCount(device_num)); }

void PSM(set_default_device)(int device_num)
{
    CK(cudaSetDevice(device_num));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device_num));
    if (&pangulu_gpu_shared_mem_size) pangulu_gpu_shared_mem_size = (int)prop.sharedMemPerBlock;
}

void PSM(get_device_name)(char* name, int device_num)
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device_num));
    strcpy(name, prop.name);
}

void PSM(get_device_memory_usage)(size_t* used_byte)
{
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    *used_byte = total_b - free_b;
}

void PSM(getrf)(pangulu_inblock_idx nb, pangulu_storage_slot_t* opdst, int tid)
{
    (void)tid;
    single_task(nb, PANGULU_TASK_GETRF, opdst, nullptr, nullptr);
}

void PSM(tstrf)(pangulu_inblock_idx nb, pangulu_storage_slot_t* opdst, pangulu_storage_slot_t* opdiag, int tid)
{
    (void)tid;
    single_task(nb, PANGULU_TASK_TSTRF, opdst, opdiag, nullptr);
}

void PSM(gessm)(pangulu_inblock_idx nb, pangulu_storage_slot_t* opdst, pangulu_storage_slot_t* opdiag, int tid)
{
    (void)tid;
    single_task(nb, PANGULU_TASK_GESSM, opdst, opdiag, nullptr);
}

void PSM(ssssm)(pangulu_inblock_idx nb, pangulu_storage_slot_t* opdst, pangulu_storage_slot_t* op1, pangulu_storage_slot_t* op2, int tid)
{
    (void)tid;
    single_task(nb, PANGULU_TASK_SSSSM, opdst, op1, op2);
}

void PSM(ssssm_batched)(pangulu_inblock_idx nb, pangulu_uint64_t ntask, pangulu_task_t* tasks)
{
    for (pangulu_uint64_t i = 0; i < ntask; ++i)
        if (tasks[i].kernel_id != PANGULU_TASK_SSSSM) tasks[i].kernel_id = PANGULU_TASK_SSSSM;  // the reference treats every entry as SSSSM (…0201000.cu:900-909)
    execute_tasks(nb, ntask, tasks);
}

void PSM(hybrid_batched)(pangulu_inblock_idx nb, pangulu_uint64_t ntask, pangulu_task_t* tasks) { execute_tasks(nb, ntask, tasks); }

void PSM(spmv)(pangulu_inblock_idx nb, pangulu_storage_slot_t* a, calculate_type* x, calculate_type* y)
{
    ensure_ctx();
    Blk<T> A = blk_csc(a);
    const int threads = 256;
    const int blocks = (nb * PSM_WARP + threads - 1) / threads;
    spmv_kernel<T><<<blocks, threads, 0, g.stream>>>(nb, A, (const T*)x, (T*)y);
    CK_LAST();
    g.launches[K_AUX]++;
}

void PSM(vecadd)(pangulu_int64_t length, calculate_type* bval, calculate_type* xval)
{
    ensure_ctx();
    if (length <= 0) return;
    vecadd_kernel<T><<<(unsigned)((length + 255) / 256), 256, 0, g.stream>>>(length, (T*)bval, (const T*)xval);
    CK_LAST();
    g.launches[K_AUX]++;
}

void PSM(sptrsv)(pangulu_inblock_idx nb, pangulu_storage_slot_t* s, calculate_type* xval, pangulu_int64_t uplo)
{
    ensure_ctx();
    // uplo 0: s is a lower half / strictly-lower pattern in CSC; uplo 1: s is the UPPER half of a diagonal block
    Blk<T> S = uplo == 0 ? blk_csc(s) : blk_upper_diag(s);
    constexpr int THREADS = 256;
    // the half's arrays are staged next to x in shared memory when they fit (the kernel decides)
    const size_t smem = std::max((size_t)nb * sizeof(T), (size_t)g.max_smem_optin - 1024);
    set_smem(sptrsv_kernel<T, THREADS>, smem);
    sptrsv_kernel<T, THREADS><<<1, THREADS, smem, g.stream>>>(nb, S, (T*)xval, (int)uplo, (unsigned)smem);
    CK_LAST();
    g.launches[K_AUX]++;
}

// ------------------------------------------------------------------------------------------------------
// Part 2: additive entry points
// ------------------------------------------------------------------------------------------------------
// Block substitution on device-resident factors: a solve plan is the list of off-diagonal blocks of every block row
// (left of the diagonal for the forward sweep, right of it for the backward sweep) and the two diagonal halves.
struct pangulu_sm100_solve_plan
{
    int nb = 0;
    long long block_length = 0;
    std::vector<long long> lptr, uptr;            // per block row: ranges into the device block lists
    std::vector<Blk<T>> dlo, dup;                 // diagonal halves per block row (val == nullptr: not held)
    SolveBlk<T>* d_l = nullptr;
    SolveBlk<T>* d_u = nullptr;
    T* d_x = nullptr;
};

pangulu_sm100_solve_plan* pangulu_sm100_solve_plan_create(pangulu_inblock_idx nb, pangulu_uint64_t block_length, pangulu_uint64_t nslot, pangulu_storage_slot_t** slots)
{
    ensure_ctx();
    pangulu_sm100_solve_plan* sp = new pangulu_sm100_solve_plan();
    sp->nb = nb;
    sp->block_length = (long long)block_length;
    std::vector<std::vector<SolveBlk<T>>> L(block_length), U(block_length);
    Blk<T> none;
    memset(&none, 0, sizeof none);
    sp->dlo.assign(block_length, none);
    sp->dup.assign(block_length, none);
    for (pangulu_uint64_t i = 0; i < nslot; ++i)
    {
        const pangulu_storage_slot_t* s = slots[i];
        if (s->brow_pos >= block_length || s->bcol_pos >= block_length) fatal(__FILE__, __LINE__, "block position outside the matrix in a solve plan", nullptr);
        if (s->brow_pos == s->bcol_pos)
        {
            if (s->is_upper) sp->dup[s->brow_pos] = blk_upper_diag(s);
            else sp->dlo[s->brow_pos] = blk_csc(s);
            continue;
        }
        SolveBlk<T> b;
        b.val = (const T*)s->d_value; b.ptr = s->d_columnpointer; b.idx = s->d_rowindex;
        b.xoff = (long long)s->bcol_pos * nb;
        (s->brow_pos > s->bcol_pos ? L : U)[s->brow_pos].push_back(b);
    }
    std::vector<SolveBlk<T>> fl, fu;
    sp->lptr.assign(block_length + 1, 0);
    sp->uptr.assign(block_length + 1, 0);
    for (pangulu_uint64_t k = 0; k < block_length; ++k)
    {
        fl.insert(fl.end(), L[k].begin(), L[k].end());
        fu.insert(fu.end(), U[k].begin(), U[k].end());
        sp->lptr[k + 1] = (long long)fl.size();
        sp->uptr[k + 1] = (long long)fu.size();
    }
    CK(cudaMalloc((void**)&sp->d_l, fl.size() * sizeof(SolveBlk<T>) + 256));
    CK(cudaMalloc((void**)&sp->d_u, fu.size() * sizeof(SolveBlk<T>) + 256));
    if (!fl.empty()) CK(cudaMemcpy(sp->d_l, fl.data(), fl.size() * sizeof(SolveBlk<T>), cudaMemcpyHostToDevice));
    if (!fu.empty()) CK(cudaMemcpy(sp->d_u, fu.data(), fu.size() * sizeof(SolveBlk<T>), cudaMemcpyHostToDevice));
    CK(cudaMalloc((void**)&sp->d_x, (size_t)block_length * nb * sizeof(T) + 256));
    return sp;
}

// x (host, block_length * nb values, already permuted and scaled) <- U^-1 L^-1 x, all arithmetic on the device
void pangulu_sm100_solve_plan_execute(pangulu_sm100_solve_plan* sp, calculate_type* x_host)
{
    if (!sp) fatal(__FILE__, __LINE__, "null solve plan", nullptr);
    const int nb = sp->nb;
    const size_t bytes = (size_t)sp->block_length * nb * sizeof(T);
    CK(cudaMemcpyAsync(sp->d_x, x_host, bytes, cudaMemcpyHostToDevice, g.stream));
    constexpr int THREADS = 256;
    const size_t smem = std::max((size_t)nb * sizeof(T), (size_t)g.max_smem_optin - 1024);
    set_smem(sptrsv_kernel<T, THREADS>, smem);
    for (long long k = 0; k < sp->block_length; ++k)          // forward: x_k <- L_kk^-1 (x_k - sum_{j<k} L_kj x_j)
    {
        const long long nblk = sp->lptr[k + 1] - sp->lptr[k];
        if (nblk) solve_spmv_row_kernel<T><<<(unsigned)nblk, 256, 0, g.stream>>>(nb, sp->d_l + sp->lptr[k], sp->d_x, k * nb);
        if (sp->dlo[k].val) sptrsv_kernel<T, THREADS><<<1, THREADS, smem, g.stream>>>(nb, sp->dlo[k], sp->d_x + k * nb, 0, (unsigned)smem);
    }
    for (long long k = sp->block_length - 1; k >= 0; --k)      // backward: x_k <- U_kk^-1 (x_k - sum_{j>k} U_kj x_j)
    {
        const long long nblk = sp->uptr[k + 1] - sp->uptr[k];
        if (nblk) solve_spmv_row_kernel<T><<<(unsigned)nblk, 256, 0, g.stream>>>(nb, sp->d_u + sp->uptr[k], sp->d_x, k * nb);
        if (sp->dup[k].val) sptrsv_kernel<T, THREADS><<<1, THREADS, smem, g.stream>>>(nb, sp->dup[k], sp->d_x + k * nb, 1, (unsigned)smem);
    }
    CK_LAST();
    g.launches[K_AUX] += 4 * sp->block_length;
    CK(cudaMemcpyAsync(x_host, sp->d_x, bytes, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
}

void pangulu_sm100_solve_plan_destroy(pangulu_sm100_solve_plan* sp)
{
    if (!sp) return;
    if (sp->d_l) CK(cudaFree(sp->d_l));
    if (sp->d_u) CK(cudaFree(sp->d_u));
    if (sp->d_x) CK(cudaFree(sp->d_x));
    delete sp;
}

static long long* option_slot(const char* key)
{
    if (!key) return nullptr;
    if (!strcmp(key, "eager_download")) return &g.opt.eager_download;
    if (!strcmp(key, "dense_threshold")) return &g.opt.dense_threshold;
    if (!strcmp(key, "ssssm_mode")) return &g.opt.ssssm_mode;
    if (!strcmp(key, "profile")) return &g.opt.profile;
    if (!strcmp(key, "getrf_mode")) return &g.opt.getrf_mode;
    if (!strcmp(key, "tile_variant")) return &g.opt.tile_variant;
    if (!strcmp(key, "split_pairs")) return &g.opt.split_pairs;
    if (!strcmp(key, "executor")) return &g.opt.executor;
    if (!strcmp(key, "trace")) return &g.opt.trace;
    if (!strcmp(key, "compact")) return &g.opt.compact;
    if (!strcmp(key, "compact_threshold")) return &g.opt.compact_threshold;
    if (!strcmp(key, "compact_block_threshold")) return &g.opt.compact_block_threshold;
    if (!strcmp(key, "cpanel")) return &g.opt.cpanel;
    if (!strcmp(key, "cpanel_block_threshold")) return &g.opt.cpanel_block_threshold;
    if (!strcmp(key, "dense_all_panels")) return &g.opt.dense_all_panels;
    if (!strcmp(key, "spin_timeout_ms")) return &g.opt.spin_timeout_ms;
    if (!strcmp(key, "progress")) return &g.opt.progress;
    if (!strcmp(key, "defer")) return &g.opt.defer;
    if (!strcmp(key, "ct_gmax")) return &g.opt.ct_gmax;
    if (!strcmp(key, "ct_minst")) return &g.opt.ct_minst;
    return nullptr;
}
int pangulu_sm100_set_option(const char* key, long long value)
{
    long long* p = option_slot(key);
    if (!p) return -1;
    *p = value;
    return 0;
}
long long pangulu_sm100_get_option(const char* key)
{
    long long* p = option_slot(key);
    return p ? *p : -1;
}

void pangulu_sm100_set_diag_groups(const pangulu_storage_slot_t* lower_half, int ngroups, const unsigned short* bounds)
{
    if (!lower_half) { g_diag_groups.clear(); return; }
    if (ngroups <= 1) { g_diag_groups.erase(lower_half); return; }
    if (ngroups > 4) fatal(__FILE__, __LINE__, "at most 4 groups per diagonal block", nullptr);
    DiagGroupInfo gi;
    gi.ng = ngroups;
    for (int i = 0; i <= ngroups; ++i) gi.bounds[i] = bounds[i];
    g_diag_groups[lower_half] = gi;
}

void pangulu_sm100_free_pinned(void* host_ptr) { if (host_ptr) CK(cudaFreeHost(host_ptr)); }

void pangulu_sm100_download_blocks(pangulu_inblock_idx nb, pangulu_uint64_t n, pangulu_storage_slot_t** slots)
{
    ensure_ctx();
    for (pangulu_uint64_t i = 0; i < n; ++i)
        CK(cudaMemcpyAsync(slots[i]->value, slots[i]->d_value, sizeof(T) * (size_t)slot_nnz(nb, slots[i]), cudaMemcpyDeviceToHost, g.stream));
    g.launches[K_AUX] += (long long)n;
}

void pangulu_sm100_upload_blocks(pangulu_inblock_idx nb, pangulu_uint64_t n, pangulu_storage_slot_t** slots)
{
    ensure_ctx();
    for (pangulu_uint64_t i = 0; i < n; ++i)
        CK(cudaMemcpyAsync(slots[i]->d_value, slots[i]->value, sizeof(T) * (size_t)slot_nnz(nb, slots[i]), cudaMemcpyHostToDevice, g.stream));
    g.launches[K_AUX] += (long long)n;
}

// ---- preprocessing on the device (sm100_preprocess.cuh; reference: src/pangulu_conversion.c:241-350, src/pangulu_utils.c:479-511)
void pangulu_sm100_scatter_values(void* d_region, size_t clear_bytes, const unsigned long long* d_dst, const void* d_src, const void* d_scale,
                                  unsigned long long n)
{
    ensure_ctx();
    if (clear_bytes) CK(cudaMemsetAsync(d_region, 0, clear_bytes, g.stream));
    if (!n) return;
    const unsigned grid = (unsigned)std::min<unsigned long long>((n + 255) / 256, (unsigned long long)g.sm_count * 8);
    scatter_values_kernel<T><<<grid, 256, 0, g.stream>>>((T*)d_region, d_dst, (const T*)d_src, (const T*)d_scale, n);
    CK_LAST();
    g.launches[K_AUX]++;
}

void pangulu_sm100_mirror_patterns(pangulu_inblock_idx nb, pangulu_uint64_t n, pangulu_storage_slot_t** slots)
{
    ensure_ctx();
    if (!n) return;
    const size_t smem = ((size_t)nb * ((nb + 31) / 32) + nb + 1) * sizeof(u32);
    if (smem > (size_t)g.max_smem_optin)
        fatal(__FILE__, __LINE__, "block size nb is too large for the device-side CSR mirror (occupancy bitmap in shared memory)", nullptr);
    std::vector<MirrorTask> tasks((size_t)n);
    for (pangulu_uint64_t i = 0; i < n; ++i)
    {
        const pangulu_storage_slot_t* s = slots[i];
        tasks[i] = {(const u32*)s->d_columnpointer, (const u16*)s->d_rowindex, (u32*)s->d_rowpointer, (u16*)s->d_columnindex,
                    (u32*)s->d_idx_of_csc_value_for_csr};
    }
    MirrorTask* d_tasks = nullptr;
    CK(cudaMalloc(&d_tasks, tasks.size() * sizeof(MirrorTask)));
    CK(cudaMemcpyAsync(d_tasks, tasks.data(), tasks.size() * sizeof(MirrorTask), cudaMemcpyHostToDevice, g.stream));
    set_smem(csr_mirror_kernel, smem);
    csr_mirror_kernel<<<(unsigned)n, 256, smem, g.stream>>>((int)nb, d_tasks);
    CK_LAST();
    CK(cudaStreamSynchronize(g.stream));   // (tasks[] is read by the copy above)
    CK(cudaFree(d_tasks));
    g.launches[K_AUX]++;
}

pangulu_sm100_plan* pangulu_sm100_plan_create(pangulu_inblock_idx nb, pangulu_uint64_t nwave, const pangulu_uint64_t* waves,
                                              const pangulu_task_t* tasks)
{
    ensure_ctx();
    pangulu_sm100_plan* pl = new pangulu_sm100_plan();
    pl->nb = nb;
    pl->dev_buf = nullptr;
    pl->reg.nb = nb;
    for (pangulu_uint64_t w = 0; w < nwave; ++w)
    {
        pl->wave_step.push_back(pl->P.steps.size());
        append_tasks(pl->P, nb, tasks, waves[w], waves[w + 1], false, &pl->reg);
    }
    pl->wave_step.push_back(pl->P.steps.size());
    Program& P = pl->P;
    if (g.opt.dense_all_panels != 0 && g.opt.compact != 0 && nb % 16 == 0 && nb >= 32 && nb <= 256)
    {
        // several ranks: a peer may read an owned block through its compact copy even if no local update does
        const double cmin = (double)g.opt.compact_block_threshold / 1000.0 * (double)nb * nb;
        for (const pangulu_storage_slot_t* sl : P.panel_slot)
            if ((double)slot_nnz(nb, sl) >= cmin) pl->reg.get(sl, sl->brow_pos < sl->bcol_pos);
    }
    // compact dense copies: allocate, zero, describe; tell every panel task which copy (if any) its block feeds
    std::vector<DenseOp<T>> ops(pl->reg.entries.size());
    if (!ops.empty())
    {
        DenseReg& R = pl->reg;
        CK(cudaMalloc(&pl->dense_buf, R.dense_total * sizeof(T) + 256));
        CK(cudaMemset(pl->dense_buf, 0, R.dense_total * sizeof(T) + 256));
        CK(cudaMalloc(&pl->dense_meta, R.meta.size() * sizeof(u16) + 256));
        CK(cudaMemcpy(pl->dense_meta, R.meta.data(), R.meta.size() * sizeof(u16), cudaMemcpyHostToDevice));
        for (size_t i = 0; i < ops.size(); ++i)
        {
            const DenseReg::Entry& e = R.entries[i];
            DenseOp<T>& o = ops[i];
            memset(&o, 0, sizeof o);
            o.dense = (T*)pl->dense_buf + e.dense_off;
            o.list = (const u16*)pl->dense_meta + e.meta_off;
            o.inv = o.list + e.cnt;
            o.cnt = e.cnt;
            o.ld = e.ld;
            o.kmask = e.kmask;
            o.pre = o.inv + pl->reg.nb;
            pl->op_val.push_back(e.slot->d_value);
            pl->dense_of_val.emplace(e.slot->d_value, i);
        }
        pl->dense_total_bytes = R.dense_total * sizeof(T);
        for (size_t i = 0; i < P.panel.size(); ++i) P.panel[i].dense = R.find(P.panel_slot[i]);
    }
    // compact panel path: panel tasks whose block has a compact copy and whose diagonal block is factored on this rank
    std::vector<DiagDense<T>> dds;
    P.getrf_dd.assign(P.getrf.size(), -1);
    if (!ops.empty() && g.opt.cpanel != 0 && sizeof(T) <= 8 && nb % 16 == 0 && nb <= 256)
    {
        struct DiagNeed { bool u = false, lt = false; int index = -1; bool remote = false; };
        std::unordered_map<const pangulu_storage_slot_t*, DiagNeed> need;  // keyed by the lower-half slot
        std::vector<const pangulu_storage_slot_t*> diag_of(P.panel.size(), nullptr);
        std::vector<char> compact(P.panel.size(), 0);
        std::vector<const pangulu_storage_slot_t*> remote_order;
        const double cp_min = (double)g.opt.cpanel_block_threshold / 1000.0 * (double)nb * nb;
        for (size_t i = 0; i < P.panel.size(); ++i)
        {
            if (P.panel[i].dense < 0) continue;
            if ((double)slot_nnz(nb, P.panel_slot[i]) < cp_min) continue;   // nearly empty block: the sparse sweep is cheaper
            const pangulu_storage_slot_t* lo = P.panel_diag_slot[i];
            const bool remote = lo->bin_id != 0;   // diagonal block factored by another rank: its halves arrive in mirror slots
            if (remote && g.opt.executor != 1) continue;
            compact[i] = 1;
            diag_of[i] = lo;
            const bool fresh = need.find(lo) == need.end();
            DiagNeed& dn = need[lo];
            dn.remote = remote;
            if (remote && fresh) remote_order.push_back(lo);
            if ((P.panel[i].kind & 0xff) == PANGULU_TASK_GESSM) dn.lt = true; else dn.u = true;
        }
        const size_t tri = (size_t)(nb / 16) * nb * PSM_CT_BLD;  // values per triangular copy
        const size_t tri_u = tri + (size_t)nb;                   // U copies carry the nb pivot reciprocals behind them
        size_t total = 0;
        for (size_t gi = 0; gi < P.getrf_slot.size(); ++gi)
        {
            auto it = need.find(P.getrf_slot[gi]);
            if (it == need.end()) continue;
            if (it->second.index >= 0) { P.getrf_dd[gi] = it->second.index; continue; }  // another group of the same block
            it->second.index = (int)dds.size();
            P.getrf_dd[gi] = it->second.index;
            DiagDense<T> dd;
            memset(&dd, 0, sizeof dd);
            const pangulu_storage_slot_t* lo = P.getrf_slot[gi];
            const pangulu_storage_slot_t* up = lo->related_block;
            if (it->second.u)
            {
                dd.u = (T*)nullptr + total;  // offset for now
                total += tri_u;
                for (int r = 0; r < nb; ++r)
                    for (u32 e = r ? up->columnpointer[r] : 0u; e < up->columnpointer[r + 1]; ++e) dd.umask[r / 16] |= 1u << (up->rowindex[e] / 16);
            }
            else dd.u = (T*)nullptr - 1;
            if (it->second.lt)
            {
                dd.lt = (T*)nullptr + total;
                total += tri;
                for (int k = 0; k < nb; ++k)
                    for (u32 e = k ? lo->columnpointer[k] : 0u; e < lo->columnpointer[k + 1]; ++e) dd.ltmask[k / 16] |= 1u << (lo->rowindex[e] / 16);
            }
            else dd.lt = (T*)nullptr - 1;
            dds.push_back(dd);
        }
        // remote diagonal blocks: same copies, refreshed by MEGA_DDENS tasks (one pseudo GETRF descriptor per group)
        for (const pangulu_storage_slot_t* lo : remote_order)
        {
            DiagNeed& dn = need[lo];
            const pangulu_storage_slot_t* up = lo->related_block;
            dn.index = (int)dds.size();
            DiagDense<T> dd;
            memset(&dd, 0, sizeof dd);
            if (dn.u)
            {
                dd.u = (T*)nullptr + total;
                total += tri_u;
                for (int r = 0; r < nb; ++r)
                    for (u32 e = r ? up->columnpointer[r] : 0u; e < up->columnpointer[r + 1]; ++e) dd.umask[r / 16] |= 1u << (up->rowindex[e] / 16);
            }
            else dd.u = (T*)nullptr - 1;
            if (dn.lt)
            {
                dd.lt = (T*)nullptr + total;
                total += tri;
                for (int k = 0; k < nb; ++k)
                    for (u32 e = k ? lo->columnpointer[k] : 0u; e < lo->columnpointer[k + 1]; ++e) dd.ltmask[k / 16] |= 1u << (lo->rowindex[e] / 16);
            }
            else dd.lt = (T*)nullptr - 1;
            dds.push_back(dd);
            const DiagGroupInfo* gi = groups_of(lo);
            const int ng = gi ? gi->ng : 1;
            std::array<int, 4> per_group = {-1, -1, -1, -1};
            for (int gr = 0; gr < ng; ++gr)
            {
                GetrfTask<T> G;
                G.lower = blk_csc(lo);
                G.upper = blk_upper_diag(up);
                G.c0 = gi ? gi->bounds[gr] : 0;
                G.c1 = gi ? gi->bounds[gr + 1] : nb;
                G.grp = gr;
                G.pad = 0;
                per_group[gr] = (int)pl->ddens.size();
                pl->ddens.push_back({(int)P.getrf.size(), dn.index, gr, dn.u, dn.lt, lo});
                P.getrf.push_back(G);
                P.getrf_slot.push_back(lo);
                P.getrf_dd.push_back(dn.index);
            }
            pl->ddens_of.emplace(lo, per_group);
        }
        if (!dds.empty())
        {
            CK(cudaMalloc(&pl->diag_dense_buf, total * sizeof(T) + 256));
            CK(cudaMemset(pl->diag_dense_buf, 0, total * sizeof(T) + 256));
            for (DiagDense<T>& dd : dds)
            {
                dd.u = dd.u == (T*)nullptr - 1 ? nullptr : (T*)pl->diag_dense_buf + (dd.u - (T*)nullptr);
                dd.lt = dd.lt == (T*)nullptr - 1 ? nullptr : (T*)pl->diag_dense_buf + (dd.lt - (T*)nullptr);
            }
        }
        // per panel run: compact tasks to the back, described by slices of PSM_CP_ROWS compact rows / columns
        for (size_t si = 0; si + 1 < P.steps.size(); ++si)
        {
            Step& st = P.steps[si];
            Step& cs = P.steps[si + 1];
            if (st.kind != K_TSTRF || cs.kind != K_CPANEL) continue;
            std::vector<size_t> order;
            for (size_t i = st.off; i < st.off + st.count; ++i) if (!compact[i]) order.push_back(i);
            const size_t nplain = order.size();
            for (size_t i = st.off; i < st.off + st.count; ++i) if (compact[i]) order.push_back(i);
            std::vector<PanelTask<T>> tp;
            std::vector<const pangulu_storage_slot_t*> ts, td, tds;
            std::vector<char> tc;
            for (size_t i : order) { tp.push_back(P.panel[i]); ts.push_back(P.panel_slot[i]); td.push_back(diag_of[i]); tc.push_back(compact[i]); tds.push_back(P.panel_diag_slot[i]); }
            for (size_t k = 0; k < order.size(); ++k)
            {
                P.panel[st.off + k] = tp[k]; P.panel_slot[st.off + k] = ts[k]; diag_of[st.off + k] = td[k]; compact[st.off + k] = tc[k];
                P.panel_diag_slot[st.off + k] = tds[k];
            }
            cs.off = st.off + nplain;
            cs.count = st.count - nplain;
            st.count = nplain;
            cs.aux_off = P.cptasks.size();
            for (size_t i = cs.off; i < cs.off + cs.count; ++i)
            {
                const int dd = need[diag_of[i]].index;
                const u32 cnt = R_cnt(pl->reg, P.panel[i].dense);
                for (u32 r0 = 0; r0 < cnt; r0 += PSM_CP_ROWS) P.cptasks.push_back({(int)i, dd, (int)r0});
            }
            cs.aux_count = P.cptasks.size() - cs.aux_off;
        }
    }
    const size_t b0 = vbytes(P.getrf), b1 = vbytes(P.panel), b2 = vbytes(P.dests), b3 = vbytes(P.pairs), b4 = vbytes(P.cpairs), b5 = vbytes(ops), b6 = vbytes(P.ctasks);
    const size_t b7 = vbytes(P.cptasks), b8 = vbytes(dds), b9 = vbytes(P.getrf_dd);
    char* d = nullptr;
    CK(cudaMalloc((void**)&d, b0 + b1 + b2 + b3 + b4 + b5 + b6 + b7 + b8 + b9 + 256));
    pl->dev_buf = d;
    if (!P.cpairs.empty()) CK(cudaMemcpy(d + b0 + b1 + b2 + b3, P.cpairs.data(), P.cpairs.size() * sizeof(CPair), cudaMemcpyHostToDevice));
    if (!ops.empty()) CK(cudaMemcpy(d + b0 + b1 + b2 + b3 + b4, ops.data(), ops.size() * sizeof(DenseOp<T>), cudaMemcpyHostToDevice));
    pl->D.cpairs = (CPair*)(d + b0 + b1 + b2 + b3);
    pl->D.ops = ops.empty() ? nullptr : (DenseOp<T>*)(d + b0 + b1 + b2 + b3 + b4);
    if (!P.ctasks.empty()) CK(cudaMemcpy(d + b0 + b1 + b2 + b3 + b4 + b5, P.ctasks.data(), P.ctasks.size() * sizeof(CTileTask), cudaMemcpyHostToDevice));
    pl->D.ctasks = (CTileTask*)(d + b0 + b1 + b2 + b3 + b4 + b5);
    {
        char* q = d + b0 + b1 + b2 + b3 + b4 + b5 + b6;
        if (!P.cptasks.empty()) CK(cudaMemcpy(q, P.cptasks.data(), P.cptasks.size() * sizeof(CPanelTask), cudaMemcpyHostToDevice));
        pl->D.cptasks = (CPanelTask*)q;
        q += b7;
        if (!dds.empty()) CK(cudaMemcpy(q, dds.data(), dds.size() * sizeof(DiagDense<T>), cudaMemcpyHostToDevice));
        pl->D.dds = dds.empty() ? nullptr : (DiagDense<T>*)q;
        q += b8;
        if (!P.getrf_dd.empty()) CK(cudaMemcpy(q, P.getrf_dd.data(), P.getrf_dd.size() * sizeof(int), cudaMemcpyHostToDevice));
        pl->D.getrf_dd = dds.empty() ? nullptr : (int*)q;
    }
    if (!P.getrf.empty()) CK(cudaMemcpy(d, P.getrf.data(), P.getrf.size() * sizeof(GetrfTask<T>), cudaMemcpyHostToDevice));
    if (!P.panel.empty()) CK(cudaMemcpy(d + b0, P.panel.data(), P.panel.size() * sizeof(PanelTask<T>), cudaMemcpyHostToDevice));
    if (!P.dests.empty()) CK(cudaMemcpy(d + b0 + b1, P.dests.data(), P.dests.size() * sizeof(SsssmDest<T>), cudaMemcpyHostToDevice));
    if (!P.pairs.empty()) CK(cudaMemcpy(d + b0 + b1 + b2, P.pairs.data(), P.pairs.size() * sizeof(SsssmPair<T>), cudaMemcpyHostToDevice));
    pl->D.getrf = (GetrfTask<T>*)d;
    pl->D.panel = (PanelTask<T>*)(d + b0);
    pl->D.dests = (SsssmDest<T>*)(d + b0 + b1);
    pl->D.pairs = (SsssmPair<T>*)(d + b0 + b1 + b2);
    {
        // work statistics of the compact tile path (pangulu_sm100_plan_stat)
        const DenseReg& R = pl->reg;
        for (const CTileTask& ct : P.ctasks)
        {
            const SsssmDest<T>& D2 = P.dests[ct.dest];
            pl->stat_ct_tasks += 1;
            const double dmma_before = pl->stat_ct_dmma;
            for (u32 pp = D2.pair_begin; pp < D2.pair_end; ++pp)
            {
                const DenseReg::Entry& ea = R.entries[P.cpairs[pp].a];
                const DenseReg::Entry& eb = R.entries[P.cpairs[pp].b];
                const u16* pre = R.meta.data() + eb.meta_off + eb.cnt + nb;
                const int nt = pre[ct.chi] - pre[ct.clo];
                if (nt <= 0) continue;
                const int ch = __builtin_popcount(ea.kmask & eb.kmask);
                if (!ch) continue;
                pl->stat_ct_pairtiles += 1;
                pl->stat_ct_chunks += ch;
                pl->stat_ct_dmma += (double)ch * 4.0 * ((ea.cnt + 7) / 8) * ((nt + 7) / 8);
                pl->stat_ct_abytes += (double)ch * (16.0 * ea.ld + 20.0 * nt) * sizeof(T);
                pl->stat_ct_msum += ea.cnt;
                pl->stat_ct_ntsum += nt;
            }
            pl->ct_task_dmma.push_back((float)(pl->stat_ct_dmma - dmma_before));
        }
    }
    if (g.opt.executor == 1) build_mega(pl);
    // the host copies of the descriptor arrays are no longer needed; keep only the step list
    std::vector<GetrfTask<T>>().swap(P.getrf);
    std::vector<PanelTask<T>>().swap(P.panel);
    std::vector<SsssmDest<T>>().swap(P.dests);
    std::vector<SsssmPair<T>>().swap(P.pairs);
    std::vector<CPair>().swap(P.cpairs);
    std::vector<CTileTask>().swap(P.ctasks);
    std::vector<CPanelTask>().swap(P.cptasks);
    std::vector<const pangulu_storage_slot_t*>().swap(P.getrf_slot);
    std::vector<const pangulu_storage_slot_t*>().swap(P.panel_diag_slot);
    std::vector<u16>().swap(pl->reg.meta);
    std::vector<const pangulu_storage_slot_t*>().swap(P.panel_slot);
    return pl;
}

void pangulu_sm100_plan_execute_range(pangulu_sm100_plan* pl, pangulu_uint64_t w0, pangulu_uint64_t w1)
{
    if (!pl) fatal(__FILE__, __LINE__, "null plan", nullptr);
    const size_t nw = pl->wave_step.size() - 1;
    if (w1 > nw) w1 = nw;
    if (w0 >= w1) return;
    run_steps(pl->nb, pl->P, pl->D, pl->wave_step[w0], pl->wave_step[w1]);
}

void pangulu_sm100_plan_execute(pangulu_sm100_plan* pl)
{
    if (!pl) fatal(__FILE__, __LINE__, "null plan", nullptr);
    if (pl->mega && g.opt.executor == 1) { run_mega(pl); return; }
    pangulu_sm100_plan_execute_range(pl, 0, pl->wave_step.size() - 1);
}

// Execute the plan and bring the values of the given blocks to their host arrays WHILE the factorisation runs: a block's
// values are final once its done flag(s) are up, long before the kernel ends (the leaves of the elimination tree finish
// first).  The host polls the flag array through a non-blocking copy stream and enqueues device -> host copies of
// address-contiguous pieces of blocks (the driver's layout: one piece = up to 128 MB of consecutive blocks) as soon as
// every block of a piece is final; what is left goes after the kernel.  The reference downloads every block right after
// the kernel call that finishes it (pangulu_platform_0201000.cu:196-201, 639-640, 680, 714); this is the same overlap without
// a launch per block.  Host arrays must be pinned for the copies to overlap.
void pangulu_sm100_plan_execute_download(pangulu_sm100_plan* pl, pangulu_uint64_t n, pangulu_storage_slot_t** slots)
{
    if (!pl) fatal(__FILE__, __LINE__, "null plan", nullptr);
    const int nb = pl->nb;
    if (!(pl->mega && g.opt.executor == 1) || pl->nids == 0)
    {
        pangulu_sm100_plan_execute(pl);
        pangulu_sm100_download_blocks((pangulu_inblock_idx)nb, n, slots);
        CK(cudaStreamSynchronize(g.stream));
        return;
    }
    struct Piece { char* d; char* h; size_t bytes; std::vector<int> flags; bool issued; };
    std::vector<Piece> pieces;
    const size_t piece_max = (size_t)128 << 20;
    for (pangulu_uint64_t i = 0; i < n; ++i)
    {
        const pangulu_storage_slot_t* sl = slots[i];
        const size_t bytes = sizeof(T) * (size_t)slot_nnz(nb, sl);
        if (!bytes) continue;
        char* d = (char*)sl->d_value;
        char* h = (char*)sl->value;
        // consecutive blocks with the same (small, alignment) gap on both sides extend the piece; the gap bytes travel too
        long long gap = -1;
        if (!pieces.empty())
        {
            const long long gd = d - (pieces.back().d + pieces.back().bytes), gh = h - (pieces.back().h + pieces.back().bytes);
            if (gd == gh && gd >= 0 && gd < 256 && pieces.back().bytes + (size_t)gd + bytes <= piece_max) gap = gd;
        }
        if (gap < 0) { pieces.push_back({d, h, 0, {}, false}); gap = 0; }
        Piece& pc = pieces.back();
        pc.bytes += (size_t)gap + bytes;
        bool any = false;
        for (int gr = 0; gr < 4; ++gr)   // (groups of a diagonal half are keyed value pointer + group)
        {
            auto it = pl->ids.find((const char*)sl->d_value + gr);
            if (it == pl->ids.end()) break;
            pc.flags.push_back(it->second);
            any = true;
        }
        if (!any) pc.flags.push_back(-1);   // no task finalises it: only after the kernel
    }
    static cudaStream_t copy_stream = nullptr;
    static int* h_flags = nullptr;
    static size_t h_flags_cap = 0;
    if (!copy_stream) CK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    if (h_flags_cap < pl->nids)
    {
        if (h_flags) CK(cudaFreeHost(h_flags));
        CK(cudaHostAlloc((void**)&h_flags, pl->nids * sizeof(int), cudaHostAllocDefault));
        h_flags_cap = pl->nids;
    }
    run_mega(pl);
    size_t left = pieces.size();
    while (left)
    {
        const bool finished = cudaStreamQuery(g.stream) == cudaSuccess;
        if (!finished)
        {
            CK(cudaMemcpyAsync(h_flags, pl->mp.done, pl->nids * sizeof(int), cudaMemcpyDeviceToHost, copy_stream));
            CK(cudaStreamSynchronize(copy_stream));
        }
        size_t issued_now = 0;
        for (Piece& pc : pieces)
        {
            if (pc.issued) continue;
            bool ready = finished;
            if (!ready)
            {
                ready = true;
                for (int f : pc.flags)
                    if (f < 0 || h_flags[f] == 0) { ready = false; break; }
            }
            if (!ready) continue;
            CK(cudaMemcpyAsync(pc.h, pc.d, pc.bytes, cudaMemcpyDeviceToHost, copy_stream));
            pc.issued = true;
            ++issued_now;
            --left;
        }
        if (!issued_now && !finished)
        {
            struct timespec ts = {0, 300000};
            nanosleep(&ts, nullptr);
        }
    }
    CK(cudaStreamSynchronize(copy_stream));
    CK(cudaStreamSynchronize(g.stream));
    g.launches[K_AUX] += (long long)pieces.size();
}

// persistent executor trace: for task i, out[4i..4i+3] = {kind, fetch_ns, start_ns, end_ns}; returns the task count
long long pangulu_sm100_plan_trace(pangulu_sm100_plan* pl, unsigned long long* out, long long cap)
{
    if (!pl || !pl->mega || !pl->mp.trace) return 0;
    const long long n = pl->mp.ntask;
    if (!out) return n;
    std::vector<unsigned long long> tmp((size_t)n * 3);
    CK(cudaMemcpy(tmp.data(), pl->mp.trace, tmp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (long long i = 0; i < n && i < cap; ++i)
    {
        // kind in the low byte; for tile tasks the number of operand pairs above it
        out[4 * i] = (unsigned long long)pl->host_tasks[i].kind | ((unsigned long long)((pl->host_tasks[i].kind == MEGA_TILE || pl->host_tasks[i].kind == MEGA_CTILE) ? pl->host_tasks[i].c : 0) << 8);
        out[4 * i + 1] = tmp[3 * i];
        out[4 * i + 2] = tmp[3 * i + 1];
        out[4 * i + 3] = tmp[3 * i + 2];
    }
    return n;
}

// per task of the persistent executor (same order as pangulu_sm100_plan_trace): DMMA instructions it issues (compact
// tile tasks; 0 for the other kinds); returns the task count
long long pangulu_sm100_plan_task_work(pangulu_sm100_plan* pl, float* out, long long cap)
{
    if (!pl || !pl->mega) return 0;
    const long long n = (long long)pl->host_task_work.size();
    if (out)
        for (long long i = 0; i < n && i < cap; ++i) out[i] = pl->host_task_work[i];
    return n;
}

void pangulu_sm100_plan_destroy(pangulu_sm100_plan* pl)
{
    if (!pl) return;
    if (pl->dev_buf) CK(cudaFree(pl->dev_buf));
    if (pl->mega_buf) CK(cudaFree(pl->mega_buf));
    if (pl->push_buf) CK(cudaFree(pl->push_buf));
    if (pl->dense_buf) CK(cudaFree(pl->dense_buf));
    if (pl->dense_meta) CK(cudaFree(pl->dense_meta));
    if (pl->diag_dense_buf) CK(cudaFree(pl->diag_dense_buf));
    delete pl;
}

// ---- multi-GPU wiring of a persistent plan (one rank per GPU, blocks pushed to peers over NVLink) -------------
void* pangulu_sm100_plan_sync_base(pangulu_sm100_plan* pl, size_t* bytes)
{
    if (!pl || !pl->mega) return nullptr;
    if (bytes) *bytes = pl->mega_bytes;
    return pl->mega_buf;
}
long long pangulu_sm100_plan_flag_offset(pangulu_sm100_plan* pl, const void* d_value)
{
    if (!pl || !pl->mega) return -1;
    auto it = pl->ids.find(d_value);
    if (it == pl->ids.end()) return -1;
    return (long long)((char*)(pl->mp.done + it->second) - (char*)pl->mega_buf);
}
long long pangulu_sm100_plan_flag_offset2(pangulu_sm100_plan* pl, const void* d_value, int group)
{
    return pangulu_sm100_plan_flag_offset(pl, (const char*)d_value + group);
}
void pangulu_sm100_plan_expect_remote2(pangulu_sm100_plan* pl, const void* d_value, int group)
{
    pangulu_sm100_plan_expect_remote(pl, (const char*)d_value + group);
}
void pangulu_sm100_plan_expect_remote(pangulu_sm100_plan* pl, const void* d_value)
{
    auto it = pl->ids.find(d_value);
    if (it != pl->ids.end()) pl->h_done_init[it->second] = 0;
}
void pangulu_sm100_plan_add_push2(pangulu_sm100_plan* pl, const void* d_value_local, void* peer_value, void* peer_flag, void* peer_dense)
{
    auto it = pl->ids.find(d_value_local);
    if (it == pl->ids.end()) return;  // the block is not produced or referenced by this plan
    MegaPush<T> e;
    e.peer_val = (T*)peer_value;
    e.peer_flag = (int*)peer_flag;
    e.peer_dense = nullptr;
    e.src_dense = nullptr;
    e.dense_count = 0;
    e.off = 0xffffffffu;
    e.cnt = 0;
    if (peer_dense)
    {
        auto d = pl->dense_of_val.find(d_value_local);
        if (d == pl->dense_of_val.end()) fatal(__FILE__, __LINE__, "a peer reads the compact copy of a block that has none here (set option dense_all_panels on every rank)", nullptr);
        const DenseReg::Entry& en = pl->reg.entries[d->second];
        e.peer_dense = (T*)peer_dense;
        e.src_dense = (const T*)pl->dense_buf + en.dense_off;
        e.dense_count = en.dense_len;
    }
    pl->h_push[it->second].push_back(e);
}
void pangulu_sm100_plan_add_push(pangulu_sm100_plan* pl, const void* d_value_local, void* peer_value, void* peer_flag)
{
    pangulu_sm100_plan_add_push2(pl, d_value_local, peer_value, peer_flag, nullptr);
}
// group `group` of a diagonal half: its entries [off, off+cnt) are copied when the group's GETRF task is done
void pangulu_sm100_plan_add_push3(pangulu_sm100_plan* pl, const void* d_value_local, int group, unsigned int off, unsigned int cnt, void* peer_value,
                                  void* peer_flag)
{
    auto it = pl->ids.find((const char*)d_value_local + group);
    if (it == pl->ids.end()) return;
    MegaPush<T> e;
    memset(&e, 0, sizeof e);
    e.peer_val = (T*)peer_value;
    e.peer_flag = (int*)peer_flag;
    e.off = off;
    e.cnt = cnt;   // (may be 0: nothing to copy, the flag is still raised)
    pl->h_push[it->second].push_back(e);
}
double pangulu_sm100_plan_stat(pangulu_sm100_plan* pl, const char* key)
{
    if (!pl || !key) return -1;
    // "ct_prof_on<n>": profile the compact tile tasks with at least n operand pairs (library built with -DPSM_CT_PROFILE)
    if (!strncmp(key, "ct_prof_max", 11))   // ... and at most n pairs ("ct_prof_max<n>", after "ct_prof_on<n>")
    {
        int mx = key[11] ? atoi(key + 11) : 0x7fffffff;
        CK(cudaMemcpyToSymbol(g_ct_prof_max, &mx, sizeof mx));
        return 0;
    }
    if (!strncmp(key, "ct_prof_on", 10))
    {
        int on = key[10] ? atoi(key + 10) : 1;
        const int mx = 0x7fffffff;
        CK(cudaMemcpyToSymbol(g_ct_prof_on, &on, sizeof on));
        CK(cudaMemcpyToSymbol(g_ct_prof_max, &mx, sizeof mx));
        unsigned long long z[32] = {0};
        CK(cudaMemcpyToSymbol(g_ct_prof, z, sizeof z));
        return 0;
    }
    if (!strncmp(key, "ct_prof", 7) && key[7] >= '0' && key[7] <= '9')
    {
        unsigned long long v[32];
        CK(cudaMemcpyFromSymbol(v, g_ct_prof, sizeof v));
        const int k = atoi(key + 7);
        return k < 32 ? (double)v[k] : -1.0;
    }
    if (!strcmp(key, "ct_tasks")) return pl->stat_ct_tasks;
    if (!strcmp(key, "ct_pairtiles")) return pl->stat_ct_pairtiles;
    if (!strcmp(key, "ct_chunks")) return pl->stat_ct_chunks;
    if (!strcmp(key, "ct_dmma")) return pl->stat_ct_dmma;
    if (!strcmp(key, "ct_bytes")) return pl->stat_ct_abytes;
    if (!strcmp(key, "ct_msum")) return pl->stat_ct_msum;
    if (!strcmp(key, "ct_ntsum")) return pl->stat_ct_ntsum;
    if (!strcmp(key, "dense_bytes")) return (double)pl->dense_total_bytes;
    if (!strcmp(key, "dense_blocks")) return (double)pl->reg.entries.size();
    return -1;
}
void* pangulu_sm100_plan_dense_base(pangulu_sm100_plan* pl, size_t* bytes)
{
    if (bytes) *bytes = pl ? pl->dense_total_bytes : 0;
    return pl ? pl->dense_buf : nullptr;
}
long long pangulu_sm100_plan_dense_offset(pangulu_sm100_plan* pl, const void* d_value)
{
    if (!pl) return -1;
    auto d = pl->dense_of_val.find(d_value);
    if (d == pl->dense_of_val.end()) return -1;
    return (long long)(pl->reg.entries[d->second].dense_off * sizeof(T));
}
void pangulu_sm100_plan_commit(pangulu_sm100_plan* pl)
{
    if (!pl || !pl->mega) return;
    CK(cudaMemcpy(pl->done_init, pl->h_done_init.data(), pl->nids * sizeof(int), cudaMemcpyHostToDevice));
    std::vector<int> ptr(pl->nids + 1, 0);
    std::vector<MegaPush<T>> flat;
    for (size_t i = 0; i < pl->nids; ++i)
    {
        ptr[i] = (int)flat.size();
        flat.insert(flat.end(), pl->h_push[i].begin(), pl->h_push[i].end());
    }
    ptr[pl->nids] = (int)flat.size();
    if (pl->push_buf) CK(cudaFree(pl->push_buf));
    const size_t b0 = (ptr.size() * sizeof(int) + 255) & ~(size_t)255;
    CK(cudaMalloc(&pl->push_buf, b0 + flat.size() * sizeof(MegaPush<T>) + 256));
    CK(cudaMemcpy(pl->push_buf, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!flat.empty()) CK(cudaMemcpy((char*)pl->push_buf + b0, flat.data(), flat.size() * sizeof(MegaPush<T>), cudaMemcpyHostToDevice));
    pl->mp.push_ptr = (const int*)pl->push_buf;
    pl->mp.pushes = (const MegaPush<T>*)((char*)pl->push_buf + b0);
    pl->mp.multi = 1;
}
void pangulu_sm100_plan_reset(pangulu_sm100_plan* pl) { if (pl && pl->mega) reset_mega(pl); }
void pangulu_sm100_plan_launch(pangulu_sm100_plan* pl) { if (pl && pl->mega) launch_mega(pl); }
int pangulu_sm100_ipc_export(void* devptr, void* handle64)
{
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, devptr));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
    memcpy(handle64, &h, 64);
    return 0;
}
int pangulu_sm100_ipc_close(void* devptr)
{
    if (devptr) CK(cudaIpcCloseMemHandle(devptr));
    return 0;
}
int pangulu_sm100_ipc_open(const void* handle64, void** devptr)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CK(cudaIpcOpenMemHandle(devptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

void pangulu_sm100_stats_reset(void)
{
    resolve_timers();
    for (int i = 0; i < K_NCLASS; ++i) { g.launches[i] = 0; g.ms[i] = 0; }
}
long long pangulu_sm100_stats_launches(int cls) { return (cls >= 0 && cls < K_NCLASS) ? g.launches[cls] : -1; }
double pangulu_sm100_stats_ms(int cls)
{
    resolve_timers();
    return (cls >= 0 && cls < K_NCLASS) ? g.ms[cls] : -1.0;
}
void* pangulu_sm100_stream(void) { ensure_ctx(); return (void*)g.stream; }

const char* pangulu_sm100_version(void)
{
#if defined(CALCULATE_TYPE_CR64)
    return "pangulu-sm100 cr64 id=" PSM_STR(PANGULU_SM100_ID) " sm_100a";
#elif defined(CALCULATE_TYPE_CR32)
    return "pangulu-sm100 cr32 id=" PSM_STR(PANGULU_SM100_ID) " sm_100a";
#elif defined(CALCULATE_TYPE_R32)
    return "pangulu-sm100 r32 id=" PSM_STR(PANGULU_SM100_ID) " sm_100a";
#else
    return "pangulu-sm100 r64 id=" PSM_STR(PANGULU_SM100_ID) " sm_100a";
#endif
}

}  // extern "C"
