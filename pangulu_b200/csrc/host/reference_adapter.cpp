// reference_adapter.cpp -- the fast path behind PanguLU's own pangulu_gstrf.
//
// PanguLU's numeric phase (src/pangulu_numeric.c) feeds the platform a few hundred tasks at a time through
// pangulu_hybrid_batched_interface (src/pangulu_kernel_interface.c:256-290); the compact tile / compact panel kernels
// and the persistent executor of this module need the WHOLE task graph (static plan).  This file is the bridge: one
// function that PanguLU's pangulu_gstrf (src/pangulu.c:234-246) calls INSTEAD of pangulu_numeric when the sm_100a
// platform is active and the run has one rank (overlay: INTEGRATION.md, oracle/overlay/apply_overlay.py).  It reads
// nothing but the block slots PanguLU's own storage layer built (bins[0] of pangulu_storage_t,
// src/pangulu_storage.c:239-421: all off-diagonal blocks in block-CSC order, then one (lower, upper) pair per diagonal
// block) -- CSC input, MC64/METIS reordering, symbolic factorisation and block storage stay PanguLU's.
//
//   1. block-level structure and the element-level pattern of L are rebuilt from the slots' patterns,
//   2. the static schedule (independent groups of diagonal blocks, look-ahead windows) is built by schedule.cpp,
//   3. the plan is created once per storage (cached), executed by the persistent kernel on PanguLU's device arrays,
//   4. the factors are copied back to PanguLU's host value arrays (pangulu_gstrs reads them, src/pangulu_sptrsv.c).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>

#include "host_internal.h"

using namespace pdrv;

namespace {

struct Cached
{
    BlockStore B;        // a VIEW: copies of PanguLU's slot structs (same host / device pointers), no buffers of its own
    Schedule sched;
    pangulu_sm100_plan* plan = nullptr;
    std::vector<pangulu_storage_slot_t*> theirs;   // PanguLU's slot of each copy (values are downloaded through the copies: same pointers)
    double t_schedule = 0, t_plan = 0, t_last = 0, t_download = 0;
};
std::map<const void*, std::unique_ptr<Cached>> g_cache;   // keyed by PanguLU's slot array

[[noreturn]] void die(const char* what)
{
    fprintf(stderr, "[PanguLU ERROR] sm100 reference adapter: %s\n", what);
    exit(EXIT_FAILURE);
}

Cached* build(int nb, eidx n, eidx bl, pangulu_storage_slot_t* slots, long long nslot, long long nondiag)
{
    const double t0 = wall_seconds();
    std::unique_ptr<Cached> C(new Cached());
    BlockStore& B = C->B;
    B.nb = nb; B.n = n; B.block_length = bl; B.rank = 0; B.nproc = 1; B.p = 1; B.q = 1;
    B.host_only = false;
    if (nslot != nondiag + 2 * (long long)bl) die("bins[0] does not hold every block (several ranks?): the plan path needs nproc = 1");
    B.slots.assign(slots, slots + nslot);
    C->theirs.resize(nslot);
    for (long long i = 0; i < nslot; ++i) C->theirs[i] = slots + i;
    // block-level CSC of the off-diagonal blocks, in the order PanguLU stores them (block column major, rows ascending)
    B.bcol_ptr.assign((size_t)bl + 1, 0);
    B.bcol_row.resize(nondiag);
    B.bnnz.resize(nondiag);
    B.slot_of_block.resize(nondiag);
    for (long long i = 0; i < nondiag; ++i)
    {
        const pangulu_storage_slot_t& s = B.slots[i];
        if (s.brow_pos == s.bcol_pos || s.bcol_pos >= bl || s.brow_pos >= bl) die("unexpected block position in the off-diagonal part of bins[0]");
        if (i > 0)
        {
            const pangulu_storage_slot_t& p = B.slots[i - 1];
            if (p.bcol_pos > s.bcol_pos || (p.bcol_pos == s.bcol_pos && p.brow_pos >= s.brow_pos)) die("off-diagonal blocks of bins[0] are not in block-CSC order");
        }
        B.bcol_ptr[s.bcol_pos + 1]++;
        B.bcol_row[i] = s.brow_pos;
        B.bnnz[i] = s.columnpointer[nb];
        B.slot_of_block[i] = i;
    }
    for (eidx j = 0; j < bl; ++j) B.bcol_ptr[j + 1] += B.bcol_ptr[j];
    B.slot_of_diag.assign(bl, -1);
    B.diag_nnz_lower.assign(bl, 0);
    B.diag_nnz_upper.assign(bl, 0);
    for (eidx k = 0; k < bl; ++k)
    {
        pangulu_storage_slot_t& lo = B.slots[nondiag + 2 * (long long)k];
        pangulu_storage_slot_t& up = B.slots[nondiag + 2 * (long long)k + 1];
        if (lo.brow_pos != k || lo.bcol_pos != k || up.brow_pos != k || up.bcol_pos != k || lo.is_upper || !up.is_upper)
            die("diagonal part of bins[0] is not (lower, upper) pairs in block-row order");
        lo.related_block = &up;   // the copies point at each other, as PanguLU's slots do
        up.related_block = &lo;
        B.slot_of_diag[k] = nondiag + 2 * (long long)k;
        B.diag_nnz_lower[k] = lo.columnpointer[nb];
        B.diag_nnz_upper[k] = up.columnpointer[nb];
    }
    B.n_own_offdiag = (size_t)nondiag;
    B.n_own_diag = bl;
    // element-level pattern of L (column c: rows >= c ascending) from the lower halves and the strictly-lower blocks
    Symbolic S;
    S.ptr.assign((size_t)n + 1, 0);
    auto col_len = [&](const pangulu_storage_slot_t& s, int cl) { return (eptr)(s.columnpointer[cl + 1] - (cl ? s.columnpointer[cl] : 0u)); };
    for (eidx c = 0; c < n; ++c)
    {
        const eidx J = c / nb;
        const int cl = (int)(c - J * nb);
        eptr cnt = 1 + col_len(B.slots[B.slot_of_diag[J]], cl);
        for (eptr g = B.bcol_ptr[J]; g < B.bcol_ptr[J + 1]; ++g)
            if (B.bcol_row[g] > J) cnt += col_len(B.slots[g], cl);
        S.ptr[c + 1] = S.ptr[c] + cnt;
    }
    S.idx.resize(S.ptr[n]);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t c = 0; c < (int64_t)n; ++c)
    {
        const eidx J = (eidx)c / nb;
        const int cl = (int)((eidx)c - J * nb);
        eidx* out = S.idx.data() + S.ptr[c];
        *out++ = (eidx)c;
        auto emit = [&](const pangulu_storage_slot_t& s, eidx row0) {
            for (iptr e = cl ? s.columnpointer[cl] : 0u; e < s.columnpointer[cl + 1]; ++e) *out++ = row0 + s.rowindex[e];
        };
        emit(B.slots[B.slot_of_diag[J]], J * nb);
        for (eptr g = B.bcol_ptr[J]; g < B.bcol_ptr[J + 1]; ++g)
            if (B.bcol_row[g] > J) emit(B.slots[g], B.bcol_row[g] * nb);
    }
    S.symbolic_nnz = 2 * S.ptr[n] - n;
    C->sched.policy = 2;
    C->sched.lookahead = 2;
    C->sched.split_diag = 1;
    diag_groups(B, S, C->sched);
    build_schedule(B, C->sched);
    for (eidx k = 0; k < bl; ++k)
        if (C->sched.group_count[k] > 1)
            pangulu_sm100_set_diag_groups(&B.slots[B.slot_of_diag[k]], C->sched.group_count[k], &C->sched.group_bounds[(size_t)k * (DIAG_MAX_GROUPS + 1)]);
    C->t_schedule = wall_seconds() - t0;
    const double t1 = wall_seconds();
    pangulu_sm100_set_option("dense_all_panels", 0);
    C->plan = pangulu_sm100_plan_create((pangulu_inblock_idx)nb, C->sched.wave_ptr.size() - 1, (const pangulu_uint64_t*)C->sched.wave_ptr.data(), C->sched.tasks.data());
    C->t_plan = wall_seconds() - t1;
    Cached* raw = C.get();
    g_cache[slots] = std::move(C);
    return raw;
}

// device -> host: PanguLU keeps the off-diagonal blocks of bin 0 in ONE host buffer whose device image is byte-identical
// (src/pangulu_storage.c:297-326, 361); when that holds (checked) the factors of all off-diagonal blocks come back in one
// copy, the diagonal halves one by one.
void download(Cached* C, long long nondiag)
{
    BlockStore& B = C->B;
    const int nb = B.nb;
    long long first_single = 0;
    if (nondiag > 0)
    {
        const char* h0 = (const char*)B.slots[0].value;
        const char* d0 = (const char*)B.slots[0].d_value;
        const pangulu_storage_slot_t& last = B.slots[nondiag - 1];
        const size_t span = (size_t)((const char*)last.value - h0) + (size_t)last.columnpointer[nb] * sizeof(T);
        bool same = true;
        for (long long i = 0; i < nondiag && same; ++i)
            same = ((const char*)B.slots[i].value - h0) == ((const char*)B.slots[i].d_value - d0) && (const char*)B.slots[i].value >= h0;
        if (same)
        {
            PSM(memcpy)((void*)h0, d0, span, 1);
            first_single = nondiag;
        }
    }
    std::vector<pangulu_storage_slot_t*> rest;
    for (long long i = first_single; i < (long long)B.slots.size(); ++i) rest.push_back(&B.slots[i]);
    if (!rest.empty()) pangulu_sm100_download_blocks((pangulu_inblock_idx)nb, rest.size(), rest.data());
    PSM(synchronize)();
}

}  // namespace

extern "C" {

// Called by PanguLU's pangulu_gstrf in place of pangulu_numeric (one rank).  slots = storage->bins[0].slots.
void pangulu_sm100_reference_gstrf(int nb, unsigned int n, unsigned int block_length, pangulu_storage_slot_t* slots, long long slot_count,
                                   long long nondiag_slot_cnt)
{
    auto it = g_cache.find(slots);
    Cached* C = it != g_cache.end() ? it->second.get() : build(nb, n, block_length, slots, slot_count, nondiag_slot_cnt);
    const double t0 = wall_seconds();
    pangulu_sm100_plan_execute(C->plan);
    PSM(synchronize)();
    C->t_last = wall_seconds() - t0;
    const double t1 = wall_seconds();
    download(C, nondiag_slot_cnt);
    C->t_download = wall_seconds() - t1;
}

// seconds of the last call: 0 schedule (first call), 1 plan creation (first call), 2 factorisation on the device, 3 download
double pangulu_sm100_reference_gstrf_time(pangulu_storage_slot_t* slots, int which)
{
    auto it = g_cache.find(slots);
    if (it == g_cache.end()) return -1.0;
    const Cached& C = *it->second;
    return which == 0 ? C.t_schedule : which == 1 ? C.t_plan : which == 2 ? C.t_last : C.t_download;
}

void pangulu_sm100_reference_release(pangulu_storage_slot_t* slots)
{
    auto it = g_cache.find(slots);
    if (it == g_cache.end()) return;
    if (it->second->plan) pangulu_sm100_plan_destroy(it->second->plan);
    pangulu_sm100_set_diag_groups(nullptr, 0, nullptr);
    g_cache.erase(it);
}

}  // extern "C"
