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
    u16 tile_start[18];    // B: number of non-empty columns below column 16*t, t = 0..nb/16
    u16 pad2[2];
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
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity)
{
    const u32 a = smem_u32(bar);
    u32 ok = 0;
    do
    {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
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
template <class T, int TN, int STAGES>
struct CtLayout
{
    static constexpr int LDC = 258;                      // accumulator tile: [TN][LDC] (nb <= 256; 258 = 2 mod 16: conflict-free adds)
    static constexpr int LDA_MAX = 260;
    static constexpr size_t acc_bytes = (size_t)TN * LDC * sizeof(T);
    static constexpr size_t a_bytes = (size_t)PSM_CT_KC * LDA_MAX * sizeof(T);
    static constexpr size_t b_bytes = (size_t)TN * PSM_CT_BLD * sizeof(T);
    static constexpr size_t stage_bytes = a_bytes + b_bytes;
    static constexpr size_t pc_off = acc_bytes + STAGES * stage_bytes;
    static constexpr size_t pc_bytes = (size_t)PSM_CT_BATCH * 48;
    static constexpr size_t item_off = pc_off + pc_bytes;
    static constexpr size_t total = item_off + (size_t)PSM_CT_BATCH * 16 * 2 + 64;
};

struct CtPairSm   // one staged operand pair (48 bytes)
{
    const void* a_dense;
    const void* b_dense;   // already offset to the first column of this tile: + n0*20
    const u16* rows;
    const u16* cols;       // already offset: + n0
    u16 m, ld, nt, nprime;
    u32 kmask;
    u32 first_item;
};

// tile-size traits: FP64 real uses the DMMA warp grid (8 x 2 warps of 32 x 16), everything else an FMA tile
template <class T> struct ct_traits { static constexpr int TN = 32, STAGES = 3; };
template <> struct ct_traits<pangulu_sm100_c64> { static constexpr int TN = 16, STAGES = 2; };

// Barrier state of a CTA that runs many compact-tile tasks (persistent executor): the mbarriers live in static shared
// memory (other task kinds reuse the dynamic part), are initialised once, and their phases advance with `gcount`.
struct CtState { u32 gcount = 0; bool armed = false; };

template <class T, int TN, int STAGES>
__device__ __forceinline__ void ssssm_ctile_cta(unsigned char* smem_raw, int nb, const SsssmDest<T>& D, const CPair* __restrict__ cpairs,
                                                const DenseOp<T>* __restrict__ ops, int c0, CtState& st)
{
    typedef CtLayout<T, TN, STAGES> Lay;
    constexpr int KC = PSM_CT_KC, BLD = PSM_CT_BLD, LDC = Lay::LDC;
    constexpr bool kDouble = sizeof(T) == 8 && !is_complex<T>::value;
    static_assert(!kDouble || TN == 32, "DMMA warp grid is 8 x 2 warps of 32 x 16");
    constexpr int CPT = TN / 16;  // FMA path: columns per thread
    T* acc = reinterpret_cast<T*>(smem_raw);
    __shared__ __align__(8) unsigned long long bars[2 * STAGES];  // [0,STAGES) full, [STAGES,2*STAGES) empty
    CtPairSm* pc = reinterpret_cast<CtPairSm*>(smem_raw + Lay::pc_off);
    u16* items = reinterpret_cast<u16*>(smem_raw + Lay::item_off);  // (pair_in_batch << 4) | chunk
    __shared__ int s_nitems;

    const int tid = threadIdx.x, warp = tid / PSM_WARP, lane = tid % PSM_WARP;
    const bool diag = D.u.val != nullptr;
    const int ncol = min(TN, nb - c0);
    if (!diag && col_begin(D.c.ptr, c0) == D.c.ptr[c0 + ncol]) return;

    // ---- task prologue: clear the accumulator tile, arm the barriers on first use
    for (int i = tid; i < TN * LDC; i += 512) acc[i] = zero<T>();
    if (!st.armed)
    {
        if (tid == 0)
        {
            for (int s = 0; s < STAGES; ++s) { mbar_init(bars + s, 1); mbar_init(bars + STAGES + s, 16); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        st.armed = true;
    }
    // the stage buffers may hold generic-proxy writes of earlier tasks of this CTA; order them before the bulk copies
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // warp tiling
    const int wn = warp & 1, wm = warp >> 1;          // DMMA: rows wm*32.., columns wn*16..
    const int m_base = wm * 32, n_base = wn * 16;
    const int lr = lane / 4, lk = lane % 4, lc = (lane % 4) * 2;
    const int tx = lane, ty = warp;                   // FMA: rows tx + 32 i, columns ty*CPT ..
    constexpr int NACC = kDouble ? 16 : 8 * CPT;
    T racc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) racc[i] = zero<T>();

    u32 gcount = st.gcount;  // items consumed so far by this CTA (stage = gcount % STAGES, phase = (gcount / STAGES) & 1)
    for (u32 pb = D.pair_begin; pb < D.pair_end; pb += PSM_CT_BATCH)
    {
        const int npb = (int)min((u32)PSM_CT_BATCH, D.pair_end - pb);
        __syncthreads();  // every warp is done with the previous batch's descriptors
        // ---- stage the descriptors of this batch and enumerate its (pair, chunk) items
        if (warp == 0)
        {
            u32 km = 0;
            CtPairSm e;
            if (lane < npb)
            {
                const CPair cp = cpairs[pb + lane];
                const DenseOp<T>& A = ops[cp.a];
                const DenseOp<T>& B = ops[cp.b];
                const u32 n0 = B.tile_start[c0 / 16], n1 = B.tile_start[(c0 + ncol + 15) / 16];
                e.a_dense = A.dense;
                e.b_dense = B.dense + (size_t)n0 * BLD;
                e.rows = A.list;
                e.cols = B.list + n0;
                e.m = (u16)A.cnt;
                e.ld = (u16)A.ld;
                e.nt = (u16)(n1 - n0);
                e.nprime = (u16)B.cnt;
                km = (n1 > n0) ? (A.kmask & B.kmask) : 0u;
                e.kmask = km;
            }
            const int cnt = __popc(km);
            int off = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int v = __shfl_up_sync(PSM_FULL, off, o);
                if (lane >= o) off += v;
            }
            if (lane == 31) s_nitems = off;
            off -= cnt;
            if (lane < npb)
            {
                e.first_item = (u32)off;
                pc[lane] = e;
                u32 k = km;
                while (k)
                {
                    const int c = __ffs(k) - 1;
                    k &= k - 1;
                    items[off++] = (u16)((lane << 4) | c);
                }
            }
        }
        __syncthreads();
        const int nitems = s_nitems;
        if (nitems == 0) continue;

        // ---- software pipeline: iteration i issues the bulk copies of item i and consumes item i - (STAGES-1)
        int cur_pair = -1;
        u16 rfull[4] = {0xffff, 0xffff, 0xffff, 0xffff}, cfull[4] = {0xffff, 0xffff, 0xffff, 0xffff};
        u16 rf8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rf8[i] = 0xffff;
        int cur_m = 0, cur_nt = 0, cur_ld = 0;

        auto flush = [&]() {
            // registers -> accumulator tile in full coordinates.  Which warp owns an element of the tile depends on the
            // pair's compaction, and warps drift apart by up to STAGES-1 items: the barrier keeps the adds of
            // consecutive pairs apart (every warp passes here once per pair, with or without work).
            __syncthreads();
            if constexpr (kDouble)
            {
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt2 = 0; nt2 < 2; ++nt2)
                    {
                        const u16 r = rfull[mt];
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                        {
                            const u16 c = cfull[nt2 * 2 + h];
                            T& v = racc[(mt * 2 + nt2) * 2 + h];
                            if (r != 0xffff && c != 0xffff) acc[(int)c * LDC + r] += v;
                            v = zero<T>();
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
                        const u16 r = rf8[i], c = cfull[j];
                        T& v = racc[i * CPT + j];
                        if (r != 0xffff && c != 0xffff) { T t = acc[(int)c * LDC + r]; vadd(t, v); acc[(int)c * LDC + r] = t; }
                        v = zero<T>();
                    }
            }
        };

        for (int i = 0; i < nitems + STAGES - 1; ++i)
        {
            if (tid == 0 && i < nitems)
            {
                const u32 g = gcount + (u32)i;
                const int s = (int)(g % STAGES);
                if (g >= (u32)STAGES) mbar_wait(bars + STAGES + s, ((g / STAGES) - 1) & 1);  // consumers released the stage
                const u16 it = items[i];
                const CtPairSm& e = pc[it >> 4];
                const int ch = it & 15;
                unsigned char* sb = smem_raw + Lay::acc_bytes + (size_t)s * Lay::stage_bytes;
                const u32 abytes = (u32)(KC * e.ld * sizeof(T)), bbytes = (u32)(e.nt * BLD * sizeof(T));
                mbar_expect_tx(bars + s, abytes + bbytes);
                bulk_g2s(sb, reinterpret_cast<const T*>(e.a_dense) + (size_t)ch * KC * e.ld, abytes, bars + s);
                bulk_g2s(sb + Lay::a_bytes, reinterpret_cast<const T*>(e.b_dense) + (size_t)ch * e.nprime * BLD, bbytes, bars + s);
            }
            const int ci = i - (STAGES - 1);
            if (ci < 0) continue;
            const u16 it = items[ci];
            const int pl = it >> 4;
            if (pl != cur_pair)
            {
                if (cur_pair >= 0) flush();
                cur_pair = pl;
                const CtPairSm& e = pc[pl];
                cur_m = e.m; cur_nt = e.nt; cur_ld = e.ld;
                if constexpr (kDouble)
                {
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
                    {
                        const int r = m_base + mt * 8 + lr;
                        rfull[mt] = r < cur_m ? e.rows[r] : (u16)0xffff;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                    {
                        const int c = n_base + (q >> 1) * 8 + lc + (q & 1);
                        cfull[q] = c < cur_nt ? (u16)(e.cols[c] - c0) : (u16)0xffff;
                    }
                }
                else
                {
#pragma unroll
                    for (int i2 = 0; i2 < 8; ++i2)
                    {
                        const int r = tx + 32 * i2;
                        rf8[i2] = r < cur_m ? e.rows[r] : (u16)0xffff;
                    }
#pragma unroll
                    for (int j = 0; j < CPT; ++j)
                    {
                        const int c = ty * CPT + j;
                        cfull[j] = c < cur_nt ? (u16)(e.cols[c] - c0) : (u16)0xffff;
                    }
                }
            }
            const u32 g = gcount + (u32)ci;
            const int s = (int)(g % STAGES);
            mbar_wait(bars + s, (g / STAGES) & 1);
            const T* As = reinterpret_cast<const T*>(smem_raw + Lay::acc_bytes + (size_t)s * Lay::stage_bytes);
            const T* Bs = reinterpret_cast<const T*>(smem_raw + Lay::acc_bytes + (size_t)s * Lay::stage_bytes + Lay::a_bytes);
            if constexpr (kDouble)
            {
                const int mpad = (cur_m + 7) & ~7;
                const int mtc = min(4, max(0, (mpad - m_base) / 8));
                const int ntc = min(2, max(0, (cur_nt + 7 - n_base) / 8));
                if (mtc > 0 && ntc > 0)
                {
                    const double* Ap = As + lk * cur_ld + m_base + lr;
                    const double* Bp = Bs + (n_base + lr) * BLD + lk;
#pragma unroll
                    for (int k4 = 0; k4 < KC; k4 += 4)
                    {
                        double a[4], b[2];
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) a[mt] = mt < mtc ? Ap[k4 * cur_ld + mt * 8] : 0.0;
                        b[0] = Bp[k4];
                        b[1] = ntc > 1 ? Bp[8 * BLD + k4] : 0.0;
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt)
                            if (mt < mtc)
                            {
                                dmma884c(racc[(mt * 2) * 2], racc[(mt * 2) * 2 + 1], a[mt], b[0]);
                                if (ntc > 1) dmma884c(racc[(mt * 2 + 1) * 2], racc[(mt * 2 + 1) * 2 + 1], a[mt], b[1]);
                            }
                    }
                }
            }
            else
            {
                const int mi = min(8, (cur_m - tx + 31) / 32);  // row slices of this lane that exist
                if (mi > 0 && ty * CPT < cur_nt)
                {
#pragma unroll 4
                    for (int kk = 0; kk < KC; ++kk)
                    {
                        T bv[CPT];
#pragma unroll
                        for (int j = 0; j < CPT; ++j) bv[j] = Bs[(ty * CPT + j) * BLD + kk];
#pragma unroll
                        for (int i2 = 0; i2 < 8; ++i2)
                            if (i2 < mi)
                            {
                                const T av = As[kk * cur_ld + tx + 32 * i2];
#pragma unroll
                                for (int j = 0; j < CPT; ++j) fma_acc(racc[i2 * CPT + j], av, bv[j]);
                            }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + STAGES + s);
        }
        if (cur_pair >= 0) flush();
        gcount += (u32)nitems;
    }

    st.gcount = gcount;

    // ---- epilogue: accumulator tile -> destination pattern (other CTAs may update the same block: reductions)
    __syncthreads();
    for (int jj = warp; jj < ncol; jj += 16)
    {
        const int j = c0 + jj;
        const u32 e0 = col_begin(D.c.ptr, j), e1 = D.c.ptr[j + 1];
        const T* Cj = acc + jj * LDC;
        for (u32 e = e0 + lane; e < e1; e += PSM_WARP)
        {
            const T v = Cj[D.c.idx[e]];
            if (!is_zero(v)) red_sub(&D.c.val[e], v);
        }
    }
    if (diag)
    {
        for (int r = tid; r < min(nb, c0 + ncol); r += 512)
        {
            const u32 r0 = col_begin(D.u.ptr, r), r1 = D.u.ptr[r + 1];
            u32 lo = r0, hi = r1;
            while (lo < hi)
            {
                const u32 mid = (lo + hi) >> 1;
                if ((int)D.u.idx[mid] < c0) lo = mid + 1;
                else hi = mid;
            }
            for (u32 e = lo; e < r1; ++e)
            {
                const int c = D.u.idx[e];
                if (c >= c0 + ncol) break;
                const T v = acc[(c - c0) * LDC + r];
                if (!is_zero(v)) red_sub(&D.u.val[e], v);
            }
        }
    }
    __syncthreads();  // the accumulator tile is free again
}

template <class T>
__global__ void __launch_bounds__(512, 1)
ssssm_ctile_kernel(int nb, const SsssmDest<T>* __restrict__ dests, const CPair* __restrict__ cpairs, const DenseOp<T>* __restrict__ ops, int tiles_per_dest)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TN = ct_traits<T>::TN;
    const int d = blockIdx.x / tiles_per_dest;
    const int c0 = (blockIdx.x % tiles_per_dest) * TN;
    const SsssmDest<T> D = dests[d];
    CtState st;
    ssssm_ctile_cta<T, TN, ct_traits<T>::STAGES>(smem_raw, nb, D, cpairs, ops, c0, st);
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

}  // namespace psm
