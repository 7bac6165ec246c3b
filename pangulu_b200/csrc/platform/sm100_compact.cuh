// sm100_compact.cuh -- SSSSM on row/column-compacted dense operands ("compact tile" path).
//
// The fill of a nested-dissection LU is not random: inside one nb x nb block of the L panel the entries sit on a
// subset of the ROWS (the separator vertices the eliminated columns reach), inside a block of the U panel on a
// subset of the COLUMNS, and on those subsets the block is close to dense.  Measured on the 3D Poisson benchmark
// matrices, multiplying the operands as dense m' x k' and k' x n' matrices over their non-empty rows / columns /
// k-chunks costs 1.2-1.5x the reference flop model for ALL updates together, where skipping empty 8x8x4 DMMA
// tiles of the uncompacted blocks costs 2.3-3x and dense nb^3 tiles far more.  This is what the reference's CPU
// SSSSM does per task with a gather + cblas_dgemm + scatter (pangulu_platform_0100000.c:211-397); here the gather
// is done ONCE per block, when the block becomes final (the panel-solve warp that finishes a row / column writes
// it into the block's compact dense copy), and every update that reads the block afterwards streams that copy
// with TMA bulk copies straight into the DMMA pipeline.
//
// Layouts of the compact copies (T = value type, nb % 16 == 0, nb <= 256):
//   A operand (block of the L panel, m' non-empty rows):  dense[k][ld], k = 0..nb-1 block column, ld = round8(m') + 4;
//       entry (rows[r'], k) at k*ld + r'.  The 16 columns of a k-chunk are ONE contiguous run of 16*ld values, and
//       ld = 4 mod 8 makes the DMMA fragment loads of the shared-memory image bank-conflict free.
//   B operand (block of the U panel, n' non-empty columns): dense[chunk][c'][20], chunk = k/16; entry (k, cols[c'])
//       at (chunk*n' + c')*20 + k%16 (16 values + 4 padding: conflict-free fragment loads); the piece of one chunk
//       that a 32-column destination tile needs is ONE contiguous run.
// Structural zeros inside the compact copies are exact zeros (written once when the plan is created, never
// touched again: the pattern is immutable), so they only ever contribute 0*x.
//
// One CTA owns a 32-column tile of the destination and applies EVERY queued operand pair to it: per pair the
// product of the compact operands is accumulated in DMMA registers (FP64: mma.sync.m8n8k4 = SASS DMMA.8x8x4, the
// FP64 tensor instruction of sm_100a; other value types: FMA register tile), then added into a full-coordinate
// accumulator tile in shared memory through the pair's row / column lists; after the last pair the tile is
// subtracted from the destination's CSC (and, for a diagonal destination, upper CSR) entries with global FP64
// reductions.  Operand chunks flow through a 3-stage ring of shared-memory buffers filled by cp.async.bulk (TMA)
// and handed over with mbarriers; there is no block-wide barrier inside the pair loop.
#pragma once
#include "sm100_types.cuh"

namespace psm {

#define PSM_CT_KC 16        // k-chunk
#define PSM_CT_BLD 20       // values per (chunk, column) of a B operand
#define PSM_CT_BATCH 32     // operand pairs whose descriptors are staged at once

// compact dense copy of one operand block (A: list = non-empty rows, B: list = non-empty columns)
template <class T>
struct DenseOp
{
    T* dense;
    const u16* list;       // cnt entries, ascending
    const u16* inv;        // nb entries: row / column -> compact index, 0xffff when empty
    u32 cnt;               // m' (A) or n' (B)
    u32 ld;                // A: leading dimension of a k column;  B: unused
    u32 kmask;             // bit c set: k-chunk c has entries
    u32 pad;
    const u16* pre;        // B: nb+1 entries, pre[c] = number of non-empty columns below column c
};
struct CPair { int a, b; };

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity, int site = 0)
{
    const u32 a = smem_u32(bar);
    u32 ok = 0;
#ifdef PSM_MBAR_GUARD
    // debugging aid: name the wait that never completes (site, block, thread, barrier offset in the CTA's shared window)
    unsigned polls = 0;
#endif
    do
    {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
#ifdef PSM_MBAR_GUARD
        if (!ok && ++polls == (1u << 24))
        {
            printf("[PanguLU ERROR] mbarrier wait stuck: site %d block %d thread %d barrier@%u parity %u\n", site, (int)blockIdx.x, (int)threadIdx.x, a, parity);
            __trap();
        }
#endif
    } while (!ok);
    (void)site;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma884c(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// shared-memory footprint of the compact tile routine
#define PSM_CT_NSLOT 8      // mbarrier pairs = maximum pipeline depth
template <class T, int TN, int MAXST>
struct CtLayout
{
    static constexpr int LDC = 258;                      // accumulator tile: [TN][LDC] (nb <= 256; 258 = 2 mod 16: conflict-free adds)
    static constexpr int LDA_MAX = 260;
    static constexpr size_t acc_bytes = (size_t)TN * LDC * sizeof(T);
    static constexpr size_t a_bytes = (size_t)PSM_CT_KC * LDA_MAX * sizeof(T);
    static constexpr size_t b_bytes = (size_t)TN * PSM_CT_BLD * sizeof(T);
    static constexpr size_t ring_bytes = MAXST * (a_bytes + b_bytes);   // MAXST chunks of the largest possible operands
    static constexpr size_t pc_off = acc_bytes + ring_bytes;
    static constexpr size_t pc_bytes = (size_t)PSM_CT_BATCH * 48;
    static constexpr size_t item_off = pc_off + pc_bytes;
    static constexpr size_t rank_off = item_off + (size_t)PSM_CT_BATCH * 16 * 4;
    static constexpr size_t map_off = rank_off + 256 + 2 * TN;          // per warp: 32 row + 32 column indices of its register tile
    static constexpr size_t eptr_off = map_off + 16 * 64 * 2 + 64;     // [nb + 1] destination column pointers of the tile's range (epilogue)
    static constexpr size_t total = eptr_off + 4 * 260;
};

struct CtPairSm   // one staged operand pair (48 bytes)
{
    const void* a_dense;
    const void* b_dense;   // already offset to the first column of this tile: + n0*20
    const u16* rows;
    const u16* cols;       // already offset: + n0
    u16 m, ld, nt, nprime;
    u32 kmask;             // k-chunks both operands have entries in; bit 31: the pair starts a new run (its row / column maps differ from the previous pair's)
    u32 first_item;
};
#define PSM_CT_RUNSTART 0x80000000u

// tile-size traits: columns per destination tile and ring size (in chunks of the largest operands).  A 64-column tile
// halves the operand traffic of a 32-column one (every chunk of the left operand is streamed once per tile).
template <class T> struct ct_traits { static constexpr int TN = 64, STAGES = 2; };                      // double: [64][258] accumulator + 2 x 42.5 KB
template <> struct ct_traits<float> { static constexpr int TN = 64, STAGES = 4; };
template <> struct ct_traits<pangulu_sm100_c32> { static constexpr int TN = 32, STAGES = 4; };
template <> struct ct_traits<pangulu_sm100_c64> { static constexpr int TN = 16, STAGES = 2; };

// DMMA inner products of one k-chunk with compile-time tile counts: the mma.sync instructions sit in straight-line code
// (a predicated mma.sync costs a WARPSYNC each).  The warp's register tile is TM x TNC tiles of 8 x 8; racc layout:
// [(mt*4 + nt2)*2 + h].  Per k4 step TM + TNC fragment loads feed TM * TNC tensor instructions.
template <int TM, int TNC>
__device__ __forceinline__ void ct_chunk(double* racc, const double* Ap, const double* Bp, int ld)
{
#pragma unroll
    for (int k4 = 0; k4 < PSM_CT_KC; k4 += 4)
    {
        double a[TM], b[TNC];
#pragma unroll
        for (int mt = 0; mt < TM; ++mt) a[mt] = Ap[k4 * ld + mt * 8];
#pragma unroll
        for (int nt2 = 0; nt2 < TNC; ++nt2) b[nt2] = Bp[nt2 * 8 * PSM_CT_BLD + k4];
#pragma unroll
        for (int mt = 0; mt < TM; ++mt)
#pragma unroll
            for (int nt2 = 0; nt2 < TNC; ++nt2) dmma884c(racc[(mt * 4 + nt2) * 2], racc[(mt * 4 + nt2) * 2 + 1], a[mt], b[nt2]);
    }
}
template <int TM>
__device__ __forceinline__ void ct_chunk_n(int ntc, double* racc, const double* Ap, const double* Bp, int ld)
{
    switch (ntc)
    {
    case 1: ct_chunk<TM, 1>(racc, Ap, Bp, ld); break;
    case 2: ct_chunk<TM, 2>(racc, Ap, Bp, ld); break;
    case 3: ct_chunk<TM, 3>(racc, Ap, Bp, ld); break;
    default: ct_chunk<TM, 4>(racc, Ap, Bp, ld); break;
    }
}

// Barrier state of a CTA that runs many compact-tile tasks (persistent executor): the mbarriers live in static shared
// memory (other task kinds reuse the dynamic part), are initialised once, and every wait flips the barrier's bit.
struct CtState { u32 full_bits = 0, empty_bits = 0, flush_bit = 0; bool armed = false; };

// optional phase profile of the compact tile routine (warp 0 lane 0 of every CTA, clock64): prologue, staging, wait for
// operands, compute, flush, pair setup, epilogue, produce
__device__ int g_ct_gmax = 4, g_ct_minst = 2;   // ring policy of the compact tile routine (options "ct_gmax", "ct_minst")
__device__ unsigned long long g_ct_prof[32];   // [0,10): thread 0;  [10,20): the producer thread;  [20,28): compact panel routine, thread 0
__device__ int g_ct_prof_on;
__device__ int g_ct_prof_max = 0x7fffffff;     // ... and at most this many pairs

// One CTA, one tile of the destination: the block columns [clo, chi).  The host cuts the tiles so that neither the
// destination nor any B operand queued for it has more than TN NON-EMPTY columns in the range (compact column tiles:
// a block with 40 non-empty columns has 1 tile, not nb/TN).
//
// RUNS.  The product of a pair lives in the warps' DMMA registers in the pair's COMPACT coordinates.  Consecutive pairs
// whose row map (left operand) and column map (right operand) are the same arrays -- the host shares one list between all
// blocks with the same non-empty rows / columns, so equal pointers mean equal maps; in the dense fronts of the big
// separators all maps are the identity -- keep accumulating into the same registers: the tile behaves like one long-K
// GEMM.  Only when the maps change (and after the last pair) are the registers added into the full-coordinate
// accumulator tile in shared memory through the run's row / column lists (one block barrier per RUN, not per pair).
struct CTileTask { int dest, clo, chi; };

template <class T, int TN, int MAXST>
__device__ __noinline__ void ssssm_ctile_cta(unsigned char* smem_raw, int nb, const SsssmDest<T>& D, const CPair* __restrict__ cpairs,
                                                const DenseOp<T>* __restrict__ ops, int clo, int chi, CtState& st)
{
    typedef CtLayout<T, TN, MAXST> Lay;
    constexpr int KC = PSM_CT_KC, BLD = PSM_CT_BLD, LDC = Lay::LDC, NSLOT = PSM_CT_NSLOT;
    constexpr bool kDouble = sizeof(T) == 8 && !is_complex<T>::value;
    static_assert(!kDouble || TN == 64, "DMMA warp grids assume 64-column tiles");
    static_assert(TN % 16 == 0 && TN <= 64, "tile width");
    constexpr int CPT = TN / 16;  // FMA path: columns per thread
    T* acc = reinterpret_cast<T*>(smem_raw);
    unsigned char* ring = smem_raw + Lay::acc_bytes;
    __shared__ __align__(8) unsigned long long bars[2 * NSLOT + 1];  // [0,NSLOT) full, [NSLOT,2*NSLOT) empty, [2*NSLOT] flush hand-over
    unsigned long long* fbar = bars + 2 * NSLOT;
    u32* eptr = reinterpret_cast<u32*>(smem_raw + Lay::eptr_off);
    CtPairSm* pc = reinterpret_cast<CtPairSm*>(smem_raw + Lay::pc_off);
    u32* items = reinterpret_cast<u32*>(smem_raw + Lay::item_off);  // (pair_in_batch << 16) | mask of the item's k-chunks (up to 4)
    unsigned char* crank = smem_raw + Lay::rank_off;                // [nb]: position of a column inside this tile, 0xff otherwise
    u16* ccols = reinterpret_cast<u16*>(smem_raw + Lay::rank_off + 256);  // [TN]: the tile's columns
    __shared__ int s_nitems, s_nst, s_group, s_wcnt[8], s_ntc;

    const int tid = threadIdx.x, warp = tid / PSM_WARP, lane = tid % PSM_WARP;
    u16* wrow = reinterpret_cast<u16*>(smem_raw + Lay::map_off) + warp * 64;   // [32] full rows of this warp's register tile (0xffff: none)
    u16* wcol = wrow + 32;                                                      // [32] tile columns of it
    const bool diag = D.u.val != nullptr;
    // The producer (bulk-copy issue: wait for the slot, ~110 cycles per copy) is lane 0 of the LAST warp, whose piece of the
    // warp grid is the first to be empty.  All threads keep the ring's phase bits in step, so the issue could rotate over the
    // warps (-DPSM_CT_DISTRIBUTED_PRODUCERS: lane 0 of warp i % 16 issues item i); measured, the divergence that puts into
    // every warp costs more than the single producer's delay.
    constexpr int PTID = 480;
    // phase profile: compiled in only with -DPSM_CT_PROFILE (the ten 64-bit counters cost 20 registers per thread)
#ifdef PSM_CT_PROFILE
    // g_ct_prof_on = n: profile the tasks with at least n operand pairs (thread 0 and the producer thread)
    const bool prof = g_ct_prof_on != 0 && (tid == 0 || tid == PTID) && (int)(D.pair_end - D.pair_begin) >= g_ct_prof_on &&
                      (int)(D.pair_end - D.pair_begin) <= g_ct_prof_max;
    long long pt = prof ? clock64() : 0;
    unsigned long long pacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    auto lap = [&](int k) { if (prof) { const long long now = clock64(); pacc[k] += (unsigned long long)(now - pt); pt = now; } };
#else
    auto lap = [](int) {};
#endif

    // ---- task prologue: the tile's columns, a clear accumulator tile, armed barriers
    {
        bool ne = false;
        if (tid >= clo && tid < chi)
        {
            // the pointers of the range's columns stay in shared memory: the epilogue walks the range's entries as one flat list
            const u32 b = col_begin(D.c.ptr, tid), e = D.c.ptr[tid + 1];
            eptr[tid - clo] = b;
            if (tid == chi - 1) eptr[chi - clo] = e;
            ne = e > b;
            if (diag) ne = ne || D.u.ptr[tid + 1] > col_begin(D.u.ptr, tid);
        }
        const unsigned bal = __ballot_sync(PSM_FULL, ne);
        if (tid < 256 && lane == 0) s_wcnt[warp] = __popc(bal);
        for (int i = tid; i < TN * LDC; i += 512) acc[i] = zero<T>();
        const bool arm = !st.armed;
        if (arm)
        {
            if (tid == 0)
            {
                for (int s = 0; s < NSLOT; ++s) { mbar_init(bars + s, 1); mbar_init(bars + NSLOT + s, 16); }
                mbar_init(fbar, 16);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            st.armed = true;
        }
        // The ring may hold generic-proxy writes of earlier tasks of this CTA, and the compact operand copies in global
        // memory were written with generic-proxy stores by the tasks that finished those blocks (possibly on another SM or,
        // over NVLink, another GPU; their completion flags were acquired before this routine was entered): order both
        // before the async-proxy reads of the bulk copies (cross-proxy fence, all state spaces).
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
        if (arm && lane == 0) mbar_arrive(fbar);   // phase 0 of the flush hand-over = "no flush before the first one"
        if (tid < 256)
        {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += s_wcnt[w];
            const int rank = before + __popc(bal & ((1u << lane) - 1u));
            const bool mine = ne && rank < TN;
            crank[tid] = mine ? (unsigned char)rank : (unsigned char)0xff;
            if (mine) ccols[rank] = (u16)tid;
            if (tid == 255) s_ntc = min(TN, before + __popc(bal));
        }
        __syncthreads();
    }
    const int ncol = s_ntc;
    if (ncol == 0) return;
    const int cfirst = clo, clast = chi - 1;

    // ---- thread roles
    const int lr = lane / 4, lk = lane % 4, lc = (lane % 4) * 2;
    const int tx = lane, ty = warp;                   // FMA: rows tx + 32 i, columns ty*CPT ..
    constexpr int NACC = kDouble ? 32 : 8 * CPT;
    T racc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) racc[i] = zero<T>();
    u32 full_bits = st.full_bits, empty_bits = st.empty_bits;

    // state of the current run (lives across descriptor batches)
    int cur_m = 0, cur_nt = 0, cur_ld = 0;
    int m_base = 0, n_base = 0, mtc = 0, ntc = 0;   // DMMA: this warp's piece of the run's compact product
    bool active = false, have_run = false;
    u16 rf8[kDouble ? 1 : 8], cf[kDouble ? 1 : CPT];   // FMA path: this thread's rows / columns
    // staging warp: maps of the last pair with work, carried from batch to batch
    const u16 *carry_rows = nullptr, *carry_cols = nullptr;

    u32 flush_bit = st.flush_bit;
    u32 pend_rc = 0xffffffffu;   // DMMA path: this lane's entry of the run's row map (low half) and column map (high half)
    auto flush = [&]() {
        // registers -> accumulator tile in full row coordinates.  Which warp owns an element of the tile depends on the
        // run's compaction and warp grid, so the adds of consecutive runs must not overlap.  A SPLIT hand-over instead of
        // a block barrier: a warp starts the adds of run i once every warp has ARRIVED after its adds of run i-1 -- which
        // happened a whole run of tensor work ago -- so no warp (in particular not the one that also issues the bulk
        // copies) holds the others up.  Every warp passes here once per run, with or without work.
        mbar_wait(fbar, flush_bit, 1);
        flush_bit ^= 1u;
        if constexpr (kDouble)
        {
            if (active)
            {
                // the run's maps were loaded when the run began (the loads were in flight during its tensor work)
                {
                    const u32 c16 = pend_rc >> 16;
                    u16 v = 0xffff;
                    if (c16 != 0xffffu) { const unsigned char rk = crank[c16]; if (rk != 0xff) v = rk; }
                    wrow[lane] = (u16)(pend_rc & 0xffffu);
                    wcol[lane] = v;
                    __syncwarp();
                }
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
                {
                    if (mt >= mtc) break;   // (warp-uniform: only the tiles the run's grid gave this warp hold anything)
                    const u16 r = wrow[mt * 8 + lr];
#pragma unroll
                    for (int nt2 = 0; nt2 < 4; ++nt2)
                    {
                        if (nt2 >= ntc) break;
                        const u32 c2 = *reinterpret_cast<const u32*>(wcol + nt2 * 8 + lc);   // two consecutive columns
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                        {
                            const u32 c = h ? (c2 >> 16) : (c2 & 0xffffu);
                            T& v = racc[(mt * 4 + nt2) * 2 + h];
                            if (r != 0xffff && c != 0xffffu) acc[(int)c * LDC + r] += v;
                            v = zero<T>();
                        }
                    }
                }
            }
        }
        else
        {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                {
                    const u16 r = rf8[i], c = cf[j];
                    T& v = racc[i * CPT + j];
                    if (r != 0xffff && c != 0xffff) { T tv = acc[(int)c * LDC + r]; vadd(tv, v); acc[(int)c * LDC + r] = tv; }
                    v = zero<T>();
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(fbar);
    };
    // warp grid gm x gn (16 warps) of a run by its tile counts: fewest tensor instructions on the busiest warp, then fewest
    // fragment loads; at most 4 x 4 tiles of 8 x 8 per warp.  The row groups map onto the four SM sub-partitions (warp % 4),
    // so narrow pairs still keep all four tensor pipes busy.  Returns this warp's piece.
    auto grid_of = [&](int m, int nt, int& mb, int& nbs, int& mc, int& nc) {
        const int MT = (m + 7) >> 3, NT = (nt + 7) >> 3;
        int best = 1 << 30, glog = 3;
#pragma unroll
        for (int g = 4; g >= 1; --g)   // gm = 16, 8, 4, 2
        {
            const int gm = 1 << g, gn = 16 >> g;
            const int tm = (MT + gm - 1) >> g, tn = (NT + gn - 1) / gn;
            const int cost = (tm <= 4 && tn <= 4) ? tm * tn * 16 + tm + tn : (1 << 30);
            if (cost < best) { best = cost; glog = g; }
        }
        const int gm = 1 << glog, gn = 16 >> glog;
        const int tm = (MT + gm - 1) >> glog, tn = (NT + gn - 1) / gn;
        const int mi = warp & (gm - 1), ni = warp >> glog;
        mb = mi * tm * 8; nbs = ni * tn * 8;
        mc = min(tm, max(0, MT - mi * tm));
        nc = min(tn, max(0, NT - ni * tn));
    };
    // a new run: warp grid, this warp's piece, the row / column lists of its register tile
    auto begin_run = [&](const CtPairSm& e) {
        cur_m = e.m; cur_nt = e.nt; cur_ld = e.ld;
        if constexpr (kDouble)
        {
            grid_of(cur_m, cur_nt, m_base, n_base, mtc, ntc);
            // make the tile counts provably warp-uniform (vote results live in uniform registers)
            mtc = __popc(__ballot_sync(PSM_FULL, lane < mtc));
            ntc = __popc(__ballot_sync(PSM_FULL, lane < ntc));
            active = mtc > 0 && ntc > 0;
            pend_rc = 0xffffffffu;
            if (active)
            {
                // issued here, consumed by the run's flush: the two L2 round trips overlap the run's tensor work
                const int r = m_base + lane, c = n_base + lane;
                const u32 r16 = (lane < mtc * 8 && r < cur_m) ? (u32)e.rows[r] : 0xffffu;
                const u32 c16 = (lane < ntc * 8 && c < cur_nt) ? (u32)e.cols[c] : 0xffffu;
                pend_rc = r16 | (c16 << 16);
            }
        }
        else
        {
            active = (cur_m > tx) && (ty * CPT < cur_nt);
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2)
            {
                const int r = tx + 32 * i2;
                rf8[i2] = (active && r < cur_m) ? e.rows[r] : (u16)0xffff;
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j)
            {
                const int c = ty * CPT + j;
                u16 v = 0xffff;
                if (active && c < cur_nt) { const unsigned char rk = crank[e.cols[c]]; if (rk != 0xff) v = rk; }
                cf[j] = v;
            }
        }
    };

    lap(0);
    for (u32 pb = D.pair_begin; pb < D.pair_end; pb += PSM_CT_BATCH)
    {
        const int npb = (int)min((u32)PSM_CT_BATCH, D.pair_end - pb);
        __syncthreads();  // every warp is done with the previous batch's descriptors
        // ---- stage the descriptors of this batch and enumerate its (pair, chunk) items
        if (warp == 0)
        {
            u32 km = 0;
            int bytes = 0;
            CtPairSm e;
            e.rows = nullptr; e.cols = nullptr;
            if (lane < npb)
            {
                const CPair cp = cpairs[pb + lane];
                const DenseOp<T>& A = ops[cp.a];
                const DenseOp<T>& B = ops[cp.b];
                const u32 n0 = B.pre[cfirst], n1 = B.pre[clast + 1];
                e.a_dense = A.dense;
                e.b_dense = B.dense + (size_t)n0 * BLD;
                e.rows = A.list;
                e.cols = B.list + n0;
                e.m = (u16)A.cnt;
                e.ld = (u16)A.ld;
                e.nt = (u16)min(n1 - n0, (u32)TN);
                e.nprime = (u16)B.cnt;
                km = (n1 > n0) ? (A.kmask & B.kmask) : 0u;
                if (km) bytes = (int)((KC * A.ld + e.nt * BLD) * sizeof(T));
            }
            // run boundaries: a pair with work continues the run of the previous pair with work iff both maps are the
            // same arrays (and the same piece of the column list)
            {
                const unsigned wm = __ballot_sync(PSM_FULL, km != 0);
                const unsigned below = wm & ((1u << lane) - 1u);
                const int src = below ? 31 - __clz(below) : 0;
                const unsigned long long pr = __shfl_sync(PSM_FULL, (unsigned long long)e.rows, src);
                const unsigned long long pcl = __shfl_sync(PSM_FULL, (unsigned long long)e.cols, src);
                const u16* prow = below ? (const u16*)pr : carry_rows;
                const u16* pcol = below ? (const u16*)pcl : carry_cols;
                const bool cont = km != 0 && prow == e.rows && pcol == e.cols;
                if (wm)
                {
                    const int last = 31 - __clz(wm);
                    carry_rows = (const u16*)__shfl_sync(PSM_FULL, (unsigned long long)e.rows, last);
                    carry_cols = (const u16*)__shfl_sync(PSM_FULL, (unsigned long long)e.cols, last);
                }
                e.kmask = km | ((km != 0 && !cont) ? PSM_CT_RUNSTART : 0u);
            }
            // an item = up to G k-chunks of one pair (one mbarrier hand-over, 2 G bulk copies).  The ring is cut into 8, 4 or 2
            // slots by the largest chunk of the batch.
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) bytes = max(bytes, __shfl_xor_sync(PSM_FULL, bytes, o));
            const int chunk_bytes = max(128, (bytes + 127) & ~127);
            // Hand-overs cost a few hundred cycles each (mbarrier round trip, dispatch): as many chunks per item as the ring
            // allows (at most g_ct_gmax) while it still holds g_ct_minst items; then as many slots as fit (at most 8).
            int G = g_ct_gmax;
            while (G > 1 && (int)Lay::ring_bytes / (G * chunk_bytes) < g_ct_minst) --G;
            const int nst = min(NSLOT, max(2, (int)Lay::ring_bytes / (G * chunk_bytes)));
            const int cnt = (__popc(km) + G - 1) / G;   // items of this lane's pair
            int off = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int v = __shfl_up_sync(PSM_FULL, off, o);
                if (lane >= o) off += v;
            }
            if (lane == 31) { s_nitems = off; s_nst = nst; s_group = G; }
            off -= cnt;
            if (lane < npb)
            {
                e.first_item = (u32)off;
                pc[lane] = e;
                u32 k = km;
                while (k)
                {
                    u32 m = 0;
                    for (int q = 0; q < G && k; ++q) { m |= k & (0u - k); k &= k - 1; }
                    items[off++] = ((u32)lane << 16) | m;
                }
            }
        }
        __syncthreads();
        lap(1);
        const int nitems = s_nitems;
        if (nitems == 0) continue;
        const int nst = s_nst, G = s_group;
        const int slot_bytes = (int)(Lay::ring_bytes / nst) & ~127;
        // look-ahead nst-2, not nst-1: the slot a new item overwrites was released one full item earlier, so the producer
        // (a lane of a warp that also computes) rarely waits for the slowest warp of the item just consumed
        const int LA = nst > 2 ? nst - 2 : 1;

        // producer step: bulk copies of item i into its slot.  Called by EVERY thread for every item, in item order; `mine`
        // selects the one lane that issues.
        auto produce = [&](int i) {
            const int s = i % nst;
#ifdef PSM_CT_DISTRIBUTED_PRODUCERS
            const bool mine = lane == 0 && (i & 15) == warp;   // measured at 100^3: 1.283 s against 1.261 s with one producer
#else
            const bool mine = tid == PTID;
#endif
            lap(7);
            if (i >= nst)
            {
                if (mine) mbar_wait(bars + NSLOT + s, (empty_bits >> s) & 1u, 2);   // consumers released the slot
                empty_bits ^= 1u << s;
            }
            lap(8);
            if (!mine) return;
            const u32 it = items[i];
            const CtPairSm& e = pc[it >> 16];
            u32 m = it & 0xffffu;
            // slot image: the item's A chunks, then its B chunks.  Consecutive k-chunks are contiguous in the A copy, and in
            // the B copy when the tile covers all of the operand's columns: one bulk copy per RUN of chunks (a copy costs
            // ~110 issue cycles whatever its size, profiles/r02_bulk_issue_microbench.txt).
            const u32 abytes = (u32)(KC * e.ld * sizeof(T)), bbytes = (u32)(e.nt * BLD * sizeof(T));
            const u32 nchk = (u32)__popc(m);
            unsigned char* sa = ring + (size_t)s * slot_bytes;
            unsigned char* sbb = sa + nchk * abytes;
            const bool bwhole = e.nt == e.nprime;
            mbar_expect_tx(bars + s, nchk * (abytes + bbytes));
            while (m)
            {
                const int ch = __ffs(m) - 1;
                const int r = __ffs(~(m >> ch)) - 1;   // chunks ch .. ch + r - 1 are all present
                m &= ~(((1u << r) - 1u) << ch);
                bulk_g2s(sa, reinterpret_cast<const T*>(e.a_dense) + (size_t)ch * KC * e.ld, (u32)r * abytes, bars + s);
                sa += (u32)r * abytes;
                if (bwhole)
                {
                    bulk_g2s(sbb, reinterpret_cast<const T*>(e.b_dense) + (size_t)ch * e.nprime * BLD, (u32)r * bbytes, bars + s);
                    sbb += (u32)r * bbytes;
                }
                else
                    for (int q = 0; q < r; ++q, sbb += bbytes)
                        bulk_g2s(sbb, reinterpret_cast<const T*>(e.b_dense) + (size_t)(ch + q) * e.nprime * BLD, bbytes, bars + s);
            }
            lap(9);
        };
        for (int i = 0; i < min(LA, nitems); ++i) produce(i);
        lap(7);

        // ---- pairs of the batch; items of a pair are consecutive in the item list
        int ci = 0;
        for (int pl = 0; pl < npb; ++pl)
        {
            const CtPairSm& e = pc[pl];
            const u32 ekm = e.kmask;
            const int nchunk = __popc(ekm & 0xffffu);
            const int cnt = (nchunk + G - 1) / G;   // items of this pair
            if (cnt == 0) continue;
            if (ekm & PSM_CT_RUNSTART)
            {
                lap(3);
                if (have_run) flush();
                lap(4);
                have_run = true;
                begin_run(e);
                lap(5);
            }
            const bool warp_active = __any_sync(PSM_FULL, active);
            for (int j = 0; j < cnt; ++j, ++ci)
            {
                if (ci + LA < nitems) produce(ci + LA);
                lap(7);
                const int s = ci % nst;
                mbar_wait(bars + s, (full_bits >> s) & 1u, 3);
                full_bits ^= 1u << s;
                lap(2);
                const int nch = min(G, nchunk - j * G);   // chunks in this item
                if (warp_active)
                  for (int q = 0; q < nch; ++q)
                  {
                    const T* As = reinterpret_cast<const T*>(ring + (size_t)s * slot_bytes) + (size_t)q * (KC * cur_ld);
                    const T* Bs = reinterpret_cast<const T*>(ring + (size_t)s * slot_bytes) + (size_t)nch * (KC * cur_ld) + (size_t)q * (cur_nt * BLD);
                    if constexpr (kDouble)
                    {
                        const double* Ap = As + lk * cur_ld + m_base + lr;
                        const double* Bp = Bs + (n_base + lr) * BLD + lk;
                        double* ra = reinterpret_cast<double*>(racc);
                        switch (mtc)
                        {
                        case 1: ct_chunk_n<1>(ntc, ra, Ap, Bp, cur_ld); break;
                        case 2: ct_chunk_n<2>(ntc, ra, Ap, Bp, cur_ld); break;
                        case 3: ct_chunk_n<3>(ntc, ra, Ap, Bp, cur_ld); break;
                        default: ct_chunk_n<4>(ntc, ra, Ap, Bp, cur_ld); break;
                        }
                    }
                    else
                    {
                        const int mi = min(8, (cur_m - tx + 31) / 32);  // row slices of this lane that exist
                        if (active)
                        {
#pragma unroll 4
                            for (int kk = 0; kk < KC; ++kk)
                            {
                                T bv[CPT];
#pragma unroll
                                for (int j2 = 0; j2 < CPT; ++j2) bv[j2] = Bs[(ty * CPT + j2) * BLD + kk];
#pragma unroll
                                for (int i2 = 0; i2 < 8; ++i2)
                                    if (i2 < mi)
                                    {
                                        const T av = As[kk * cur_ld + tx + 32 * i2];
#pragma unroll
                                        for (int j2 = 0; j2 < CPT; ++j2) fma_acc(racc[i2 * CPT + j2], av, bv[j2]);
                                    }
                            }
                        }
                    }
                  }
                // EVERY lane must have seen the item's phase of the full barrier before the warp releases the slot: a lane still
                // polling when the slot is refilled would find the barrier one phase further -- same parity as the phase it
                // waits for -- and spin forever.  (Lanes of a warp drift apart when one of them issues the bulk copies.)
                __syncwarp();
                if (lane == 0) mbar_arrive(bars + NSLOT + s);
                lap(3);
            }
        }
        // the releases of the last nst items are never waited for: their phases are complete once every warp has passed the
        // block barrier that precedes the next use of the ring (top of the next batch / the task's epilogue); account for them
        for (int j = max(0, nitems - nst); j < nitems; ++j) empty_bits ^= 1u << (j % nst);
    }
    if (have_run) flush();
    st.full_bits = full_bits;
    st.empty_bits = empty_bits;
    st.flush_bit = flush_bit;
    lap(4);

    // ---- epilogue: accumulator tile -> destination pattern (other CTAs may update the same block: reductions)
    __syncthreads();
    {
        // The entries of the columns [clo, chi) are ONE contiguous piece of the CSC arrays: a flat loop with four independent
        // index loads in flight per thread (a loop over columns is a chain of dependent L2 round trips -- ~15 us per task,
        // which is all a small task costs); the column of an entry by bisection over the staged pointers.
        const int nrange = chi - clo;
        const u32 e_lo = eptr[0], e_hi = eptr[nrange];
        for (u32 base = e_lo + tid; base < e_hi; base += 4 * 512)
        {
            u16 ri[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                const u32 e = base + u * 512;
                ri[u] = e < e_hi ? D.c.idx[e] : (u16)0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                const u32 e = base + u * 512;
                if (e < e_hi)
                {
                    int lo = 0, hi = nrange;   // largest column with eptr[column] <= e (skips empty columns)
                    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (eptr[mid] <= e) lo = mid; else hi = mid; }
                    const unsigned char rk = crank[clo + lo];
                    if (rk != 0xff)
                    {
                        const T v = acc[(int)rk * LDC + ri[u]];
                        if (!is_zero(v)) red_sub(&D.c.val[e], v);
                    }
                }
            }
        }
    }
    if (diag)
    {
        // upper CSR half: row r holds columns >= r ascending; the tile's columns lie in [cfirst, clast]
        for (int r = tid; r <= min(nb - 1, clast); r += 512)
        {
            const u32 r0 = col_begin(D.u.ptr, r), r1 = D.u.ptr[r + 1];
            u32 lo = r0, hi = r1;
            while (lo < hi)
            {
                const u32 mid = (lo + hi) >> 1;
                if ((int)D.u.idx[mid] < cfirst) lo = mid + 1;
                else hi = mid;
            }
            for (u32 e = lo; e < r1; ++e)
            {
                const int c = D.u.idx[e];
                if (c > clast) break;
                const unsigned char rk = crank[c];
                if (rk == 0xff) continue;
                const T v = acc[(int)rk * LDC + r];
                if (!is_zero(v)) red_sub(&D.u.val[e], v);
            }
        }
    }
    __syncthreads();  // the accumulator tile is free again
    lap(6);
#ifdef PSM_CT_PROFILE
    if (prof)
        for (int k = 0; k < 10; ++k) atomicAdd(&g_ct_prof[k + (tid == PTID ? 10 : 0)], pacc[k]);
#endif
}

// stand-alone launch (per-wave executor): one CTA per tile task
template <class T>
__global__ void __launch_bounds__(512, 1)
ssssm_ctile_kernel(int nb, const SsssmDest<T>* __restrict__ dests, const CPair* __restrict__ cpairs, const DenseOp<T>* __restrict__ ops,
                   const CTileTask* __restrict__ ctasks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TN = ct_traits<T>::TN;
    const CTileTask ct = ctasks[blockIdx.x];
    const SsssmDest<T> D = dests[ct.dest];
    CtState st;
    ssssm_ctile_cta<T, TN, ct_traits<T>::STAGES>(smem_raw, nb, D, cpairs, ops, ct.clo, ct.chi, st);
}

// Write one finished value of a block into its compact dense copy (called by the panel-solve warps).
//   A operand (block of the L panel):  entry (row r, column k)
//   B operand (block of the U panel):  entry (row k, column c)
template <class T>
__device__ __forceinline__ void dense_store_a(const DenseOp<T>& A, int r, int k, T v) { A.dense[(size_t)k * A.ld + A.inv[r]] = v; }
template <class T>
__device__ __forceinline__ void dense_store_b(const DenseOp<T>& B, int k, int cprime, T v)
{
    B.dense[((size_t)(k / PSM_CT_KC) * B.cnt + cprime) * PSM_CT_BLD + (k % PSM_CT_KC)] = v;
}


// ================================================================================================
// TSTRF / GESSM on compact dense data ("compact panel" path).
//
// A panel block that has a compact dense copy is solved as a dense triangular system over its non-empty
// rows (TSTRF: X U = B) or columns (GESSM: L X = B, i.e. X^T L^T = B^T).  With
//     X[k][s] = B(rows[r0+s], k)  (TSTRF)   /   B(k, cols[r0+s])  (GESSM),   s < 64,  k < nb
//     Tm      = U                 (TSTRF)   /   L^T                 (GESSM)
// both are   X[c][s] = (X[c][s] - sum_{k<c} X[k][s] Tm(k,c)) / Tm(c,c)   (no division for GESSM: unit diagonal).
// One CTA takes 64 compact rows/columns.  The k range is processed in chunks of 16 (only chunks in which the block
// has entries): the contribution of all earlier chunks is a dense 64 x 16 x (16 K) product on the DMMA pipe (X is kept
// in shared memory in the A-operand layout [k][68], Tm arrives by TMA in the B-operand layout [chunk][column][20]
// that the GETRF task wrote), followed by a 16-column substitution with one thread per row -- the same operations
// on every entry as the reference's row/column sweeps (pangulu_platform_0100000.c:137-209), in blocked order.
// Results go to the block's CSC value array and to its compact dense copy.
// ================================================================================================
template <class T>
struct DiagDense
{
    T* u;            // U   (upper half incl. diagonal) in the B-operand layout [chunk][nb][20], followed by the nb RECIPROCALS of the
                     // (clamped) pivots, or null
    T* lt;           // L^T (strict lower half, transposed), same layout, or null
    u32 umask[16];   // umask[K] bit J: sub-block (rows 16K.., columns 16J..) of U has entries
    u32 ltmask[16];
};

#define PSM_CP_ROWS 64
#define PSM_CP_LDX 68

template <class T>
struct CpLayout
{
    static constexpr size_t x_bytes = (size_t)256 * PSM_CP_LDX * sizeof(T);
    static constexpr size_t sub_bytes = (size_t)16 * PSM_CT_BLD * sizeof(T);     // one 16 x 16 sub-block of Tm
    static constexpr size_t tbuf_bytes = 16 * sub_bytes + 16 * sizeof(T);           // one column block (all K) + the 16 pivot reciprocals of the chunk
    static constexpr size_t total = x_bytes + 2 * tbuf_bytes + 4 * (260 + 64) + 2 * 256 + 64;  // + pointer / prefix / inverse-map staging
};

struct CpState { u32 bits = 0; bool armed = false; };

// GETRF epilogue: scatter the finished halves of a diagonal block into its dense triangular copies
template <class T>
__device__ __forceinline__ void diag_densify(int nb, const GetrfTask<T>& G, const DiagDense<T>& DD)
{
    const int warp = threadIdx.x / PSM_WARP, lane = threadIdx.x % PSM_WARP, nw = blockDim.x / PSM_WARP;
    (void)nb;
    if (DD.u)
    {
        // The compact panel solves multiply by the reciprocal of the pivot (one division per pivot HERE instead of a chain of
        // 16 dependent divisions per row and k-chunk there; a result differs from the reference's x / pivot by one rounding).
        T* rcp = DD.u + (size_t)(nb / 16) * nb * PSM_CT_BLD;
        for (int r = G.c0 + warp; r < G.c1; r += nw)   // row r of U (CSR, diagonal first); only the task's group
        {
            const u32 e0 = col_begin(G.upper.ptr, r), e1 = G.upper.ptr[r + 1];
            for (u32 e = e0 + lane; e < e1; e += PSM_WARP)
                DD.u[((size_t)(r / 16) * nb + G.upper.idx[e]) * PSM_CT_BLD + (r % 16)] = G.upper.val[e];
            if (lane == 0 && e1 > e0) rcp[r] = vrcp(clamp_pivot(G.upper.val[e0]));
        }
    }
    if (DD.lt)
        for (int k = G.c0 + warp; k < G.c1; k += nw)   // column k of L (CSC): L(c,k) = L^T(k,c)
            for (u32 e = col_begin(G.lower.ptr, k) + lane; e < G.lower.ptr[k + 1]; e += PSM_WARP)
                DD.lt[((size_t)(k / 16) * nb + G.lower.idx[e]) * PSM_CT_BLD + (k % 16)] = G.lower.val[e];
}

template <class T>
__device__ __noinline__ void panel_compact_cta(unsigned char* smem_raw, int nb, const PanelTask<T>& Q, const DenseOp<T>& Dn, const DiagDense<T>& DD,
                                                  int r0, CpState& st)
{
    typedef CpLayout<T> Lay;
    constexpr int LDX = PSM_CP_LDX, BLD = PSM_CT_BLD, R = PSM_CP_ROWS;
    constexpr bool kDouble = sizeof(T) == 8 && !is_complex<T>::value;
    T* X = reinterpret_cast<T*>(smem_raw);                               // [nb][LDX]
    unsigned char* tb0 = smem_raw + Lay::x_bytes;                         // 2 x [16 K][16 cols][20]
    __shared__ __align__(8) unsigned long long cbars[2];
    const int tid = threadIdx.x, warp = tid / PSM_WARP, lane = tid % PSM_WARP;
    const bool gessm = (Q.kind & 0xff) == PANGULU_TASK_GESSM;
    const T* Tm = gessm ? DD.lt : DD.u;
    const u32* tmask = gessm ? DD.ltmask : DD.umask;
    const u32 kmask = Dn.kmask;
    const int nrows = min(R, (int)Dn.cnt - r0);   // compact rows / columns of this slice
    const Blk<T> B = Q.b;
#ifdef PSM_CT_PROFILE
    // phases (thread 0): 0 clear + staging, 1 gather, 2 wait for the triangular factor, 3 tensor update, 4 substitution, 5 scatter, 6 dense copy
    const bool cprof = g_ct_prof_on != 0 && tid == 0;
    long long cpt = cprof ? clock64() : 0;
    unsigned long long cpacc[7] = {0, 0, 0, 0, 0, 0, 0};
    auto clap = [&](int k) { if (cprof) { const long long now = clock64(); cpacc[k] += (unsigned long long)(now - cpt); cpt = now; } };
#else
    auto clap = [](int) {};
#endif

    if (!st.armed)
    {
        if (tid == 0)
        {
            mbar_init(cbars + 0, 1);
            mbar_init(cbars + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        st.armed = true;
    }
    // ---- X <- 0 on the block's k-chunks, then gather the slice from the CSC arrays
    for (int ch = 0; ch < nb / 16; ++ch)
        if ((kmask >> ch) & 1u)
            for (int i = tid; i < 16 * LDX; i += 512) X[ch * 16 * LDX + i] = zero<T>();
    // generic-proxy writes (X zeroing here; the dense triangular copies written by the GETRF / refresh task whose flag
    // was acquired before this routine) before the async-proxy bulk reads
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    // producer: the column block of Tm for chunk J -> buffer `par`; one bulk copy per needed 16 x 16 sub-block
    auto issue = [&](int J, int par) {
        unsigned char* tb = tb0 + (size_t)par * Lay::tbuf_bytes;
        u32 need = 0;
        for (int K = 0; K <= J; ++K)
            if (((kmask >> K) & 1u) && (K == J || ((tmask[K] >> J) & 1u))) need |= 1u << K;
        mbar_expect_tx(cbars + par, (u32)(__popc(need) * Lay::sub_bytes + (gessm ? 0 : 16 * sizeof(T))));
        for (int K = 0; K <= J; ++K)
            if ((need >> K) & 1u)
                bulk_g2s(tb + (size_t)K * Lay::sub_bytes, Tm + ((size_t)K * nb + 16 * J) * BLD, (u32)Lay::sub_bytes, cbars + par);
        if (!gessm) bulk_g2s(tb + 16 * Lay::sub_bytes, Tm + (size_t)(nb / 16) * nb * BLD + 16 * J, (u32)(16 * sizeof(T)), cbars + par);   // pivot reciprocals
    };
    // the copies are issued by a lane of a warp that has no part in the substitution (the issue costs ~110 cycles per copy)
    constexpr int ISSUER = 256;
    int Jcur = kmask ? (__ffs(kmask) - 1) : -1;
    int par = 0;
    if (tid == ISSUER && Jcur >= 0) issue(Jcur, 0);
    // Gather (and, at the end, scatter) run as FLAT loops over the slice's entries with the block's pointer array and
    // inverse row map staged in shared memory: many independent loads in flight instead of a chain of dependent ones
    // per column.
    u32* sptr = reinterpret_cast<u32*>(smem_raw + Lay::x_bytes + 2 * Lay::tbuf_bytes);   // [nb+1] column pointers (TSTRF) / [65] slice prefix (GESSM)
    u32* sbeg = sptr + 260;                                                               // GESSM: first entry of each slice column
    u16* sinv = reinterpret_cast<u16*>(sbeg + 64);                                        // TSTRF: [nb] row -> compact row
    u32 nflat = 0;
    if (gessm)
    {
        if (warp == 0)
        {
            u32 run = 0;
            for (int base = 0; base < R; base += PSM_WARP)
            {
                const int sidx = base + lane;
                u32 b = 0, len = 0;
                if (sidx < nrows)
                {
                    const int c = Dn.list[r0 + sidx];
                    b = col_begin(B.ptr, c);
                    len = B.ptr[c + 1] - b;
                }
                u32 incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1)
                {
                    const u32 v = __shfl_up_sync(PSM_FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                sptr[sidx] = run + incl - len;
                sbeg[sidx] = b;
                run += __shfl_sync(PSM_FULL, incl, 31);
            }
            if (lane == 0) sptr[R] = run;
        }
    }
    else
    {
        for (int i = tid; i <= nb; i += 512) sptr[i] = i == 0 ? 0u : B.ptr[i];
        for (int i = tid; i < nb; i += 512) sinv[i] = Dn.inv[i];
    }
    __syncthreads();
    nflat = gessm ? sptr[R] : sptr[nb];
    // Each warp walks a contiguous 1/16 of the flat entry range, four 32-entry groups at a time; the column (TSTRF) or
    // slice column (GESSM) of an entry is found by advancing a cursor over the staged pointers (monotone per lane).
    const int nseg = gessm ? R : nb;
    const u32 per_warp = (nflat + 15) / 16;
    const u32 jb = min(nflat, (u32)warp * per_warp), je = min(nflat, jb + per_warp);
    int seg0 = 0;
    {
        int lo = 0, hi = nseg;   // largest segment with sptr[seg] <= jb
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sptr[mid] <= jb) lo = mid; else hi = mid; }
        seg0 = lo;
    }
    // pass = 0: X <- block values;  pass = 1: block values <- X
    auto sweep = [&](int pass) {
        int seg = seg0;
        for (u32 j0 = jb + lane; j0 < je; j0 += 4 * PSM_WARP)
        {
            u32 e[4];
            int kk[4], sl[4];
            u16 ri[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                const u32 j = j0 + u * PSM_WARP;
                kk[u] = -1;
                if (j < je)
                {
                    while (j >= sptr[seg + 1]) ++seg;
                    kk[u] = seg;
                    e[u] = gessm ? sbeg[seg] + (j - sptr[seg]) : j;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (kk[u] >= 0) ri[u] = B.idx[e[u]];
#pragma unroll
            for (int u = 0; u < 4; ++u)
            {
                sl[u] = -1;
                if (kk[u] >= 0)
                {
                    if (gessm) { sl[u] = kk[u]; kk[u] = ri[u]; }   // segment = slice column, entry row = k
                    else
                    {
                        const int sc = (int)sinv[ri[u]] - r0;
                        sl[u] = (sc >= 0 && sc < R) ? sc : -1;
                    }
                }
            }
            if (pass == 0)
            {
                T v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (sl[u] >= 0) v[u] = B.val[e[u]];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (sl[u] >= 0) X[kk[u] * LDX + sl[u]] = v[u];
            }
            else
            {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (sl[u] >= 0) B.val[e[u]] = X[kk[u] * LDX + sl[u]];
            }
        }
    };
    clap(0);
    sweep(0);
    __syncthreads();
    clap(1);

    const int lr = lane / 4, lk = lane % 4, lc = (lane % 4) * 2;
    const int mtile = warp & 7, ntile = warp >> 3;
    u32 bits = st.bits;
    while (Jcur >= 0)
    {
        const u32 rest = kmask & ~((2u << Jcur) - 1u);
        const int Jnext = rest ? (__ffs(rest) - 1) : -1;
#ifdef PSM_CP_ISSUE_EARLY
        if (tid == ISSUER && Jnext >= 0) issue(Jnext, par ^ 1);
#endif
        mbar_wait(cbars + par, (bits >> par) & 1u, 4);
        bits ^= 1u << par;
        clap(2);
        const T* Tb = reinterpret_cast<const T*>(tb0 + (size_t)par * Lay::tbuf_bytes);
        // ---- X[:, J] -= X[:, <J] Tm(<J, J)
        u32 need = 0;
        for (int K = 0; K < Jcur; ++K)
            if (((kmask >> K) & 1u) && ((tmask[K] >> Jcur) & 1u)) need |= 1u << K;
        if (need)
        {
            if constexpr (kDouble)
            {
                double a0 = 0, a1 = 0, b0 = 0, b1 = 0;   // two accumulator pairs (k4 parity) for DMMA latency
                u32 nk = need;
                while (nk)
                {
                    const int K = __ffs(nk) - 1;
                    nk &= nk - 1;
                    const double* Ap = X + (K * 16 + lk) * LDX + mtile * 8 + lr;
                    const double* Bp = Tb + (size_t)K * 16 * BLD + (ntile * 8 + lr) * BLD + lk;
                    dmma884c(a0, a1, Ap[0], Bp[0]);
                    dmma884c(b0, b1, Ap[4 * LDX], Bp[4]);
                    dmma884c(a0, a1, Ap[8 * LDX], Bp[8]);
                    dmma884c(b0, b1, Ap[12 * LDX], Bp[12]);
                }
                double* Xo = X + (Jcur * 16 + ntile * 8 + lc) * LDX + mtile * 8 + lr;
                Xo[0] -= a0 + b0;
                Xo[LDX] -= a1 + b1;
            }
            else
            {
                // FMA: thread = (row s = tid % 64, column pair cg = tid / 64)
                const int s = tid % 64, cg = tid / 64;
                T acc0 = zero<T>(), acc1 = zero<T>();
                u32 nk = need;
                while (nk)
                {
                    const int K = __ffs(nk) - 1;
                    nk &= nk - 1;
                    const T* Tk = Tb + (size_t)K * 16 * BLD;
#pragma unroll 4
                    for (int kk = 0; kk < 16; ++kk)
                    {
                        const T xv = X[(K * 16 + kk) * LDX + s];
                        fma_acc(acc0, xv, Tk[(cg * 2) * BLD + kk]);
                        fma_acc(acc1, xv, Tk[(cg * 2 + 1) * BLD + kk]);
                    }
                }
                vsub(X[(Jcur * 16 + cg * 2) * LDX + s], acc0);
                vsub(X[(Jcur * 16 + cg * 2 + 1) * LDX + s], acc1);
            }
            __syncthreads();
        }
        clap(3);
        // prefetch the next column block into the other buffer (its readers finished before the barrier that ended the previous
        // chunk) while the substitution runs
#ifndef PSM_CP_ISSUE_EARLY
        if (tid == ISSUER && Jnext >= 0) issue(Jnext, par ^ 1);
#endif
        // ---- 16-column substitution against the diagonal sub-block, one thread per row
        if (tid < R)
        {
            const T* Tjj = Tb + (size_t)Jcur * 16 * BLD;   // Tm(16J + k, 16J + c) at [c*20 + k]
            const T* rcp = reinterpret_cast<const T*>(tb0 + (size_t)par * Lay::tbuf_bytes + 16 * Lay::sub_bytes);
            T x[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = X[(Jcur * 16 + c) * LDX + tid];
            // right-looking inside the 16 columns: the 15-k updates after a division are independent (same operations on
            // every entry, in the same order, as the column-by-column form)
#pragma unroll
            for (int k = 0; k < 16; ++k)
            {
#ifdef PSM_CP_DIVIDE
                if (!gessm) x[k] = vdiv(x[k], clamp_pivot(Tjj[k * BLD + k]));
#else
                if (!gessm) x[k] = vmul(x[k], rcp[k]);
#endif
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c > k) fnma(x[c], x[k], Tjj[c * BLD + k]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) X[(Jcur * 16 + c) * LDX + tid] = x[c];
        }
        __syncthreads();
        clap(4);
        Jcur = Jnext;
        par ^= 1;
    }
    st.bits = bits;

    // ---- write back: CSC values and the compact dense copy
    sweep(1);
    clap(5);
    if (gessm)
    {
        // dense copy (B-operand layout): (k, c') at ((k/16) n' + c') 20 + k%16 -- 16 consecutive k are contiguous
        const int nchunk = nb / 16;
        for (int i = tid; i < nchunk * nrows * 16; i += 512)
        {
            const int kk = i % 16, s2 = (i / 16) % nrows, ch = i / (16 * nrows);
            if ((kmask >> ch) & 1u) Dn.dense[((size_t)ch * Dn.cnt + r0 + s2) * BLD + kk] = X[(ch * 16 + kk) * LDX + s2];
        }
    }
    else
    {
        // dense copy (A-operand layout): (r', k) at k ld + r'
        for (int i = tid; i < nb * nrows; i += 512)
        {
            const int s2 = i % nrows, k = i / nrows;
            if ((kmask >> (k / 16)) & 1u) Dn.dense[(size_t)k * Dn.ld + r0 + s2] = X[k * LDX + s2];
        }
    }
    __syncthreads();
    clap(6);
#ifdef PSM_CT_PROFILE
    if (cprof)
        for (int k = 0; k < 7; ++k) atomicAdd(&g_ct_prof[20 + k], cpacc[k]);
#endif
}

template <class T>
__global__ void diag_densify_kernel(int nb, const GetrfTask<T>* __restrict__ tasks, const int* __restrict__ getrf_dd, const DiagDense<T>* __restrict__ dds)
{
    const int dd = getrf_dd ? getrf_dd[blockIdx.x] : -1;
    if (dd < 0) return;
    diag_densify<T>(nb, tasks[blockIdx.x], dds[dd]);
}

// stand-alone launch (per-wave executor)
struct CPanelTask { int panel, dd, r0; };
template <class T>
__global__ void __launch_bounds__(512, 1)
panel_compact_kernel(int nb, const PanelTask<T>* __restrict__ tasks, const CPanelTask* __restrict__ cpt, const DenseOp<T>* __restrict__ ops,
                     const DiagDense<T>* __restrict__ dds)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const CPanelTask ct = cpt[blockIdx.x];
    const PanelTask<T> Q = tasks[ct.panel];
    CpState st;
    panel_compact_cta<T>(smem_raw, nb, Q, ops[Q.dense], dds[ct.dd], ct.r0, st);
}

}  // namespace psm
