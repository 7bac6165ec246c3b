// sm100_kernels.cuh -- batched sparse block kernels of PanguLU's numeric factorisation for sm_100a.
//
// All four kernel classes take an ARRAY of task descriptors resident in HBM and run one grid over the
// whole batch (the reference launches one kernel per task and synchronises the device in between,
// src/platforms/02_NONSHAREDMEM/01_GPU/000_CUDA/pangulu_platform_0201000.cu:875-909).
//
// Numerical semantics follow the reference CPU platform (the runnable oracle):
//   GETRF  pangulu_platform_0100000.c:57-135   right-looking, no pivoting, pivot clamp
//   TSTRF  pangulu_platform_0100000.c:137-175  B <- B U^-1 row by row through B's CSR mirror
//   GESSM  pangulu_platform_0100000.c:178-209  B <- L^-1 B column by column (unit diagonal)
//   SSSSM  pangulu_platform_0100000.c:211-397  C <- C - A B, products outside C's pattern dropped;
//                                              diagonal destination: both halves updated
// The nonzero pattern of every block is immutable; only value arrays are written.
#pragma once
#include "sm100_types.cuh"

namespace psm {

#define PSM_WARP 32
#define PSM_FULL 0xffffffffu

__device__ __forceinline__ u32 col_begin(const u32* ptr, int j) { return j == 0 ? 0u : ptr[j]; }

// x -= v with a global-memory reduction (several CTAs updating one destination)
// Fire-and-forget global reductions (SASS REDG).  atomicAdd on a pointer whose state space the compiler cannot prove (the
// value arrays are reached through descriptors in memory) compiles to a GENERIC returning atomic plus a shared-memory CAS
// fallback, and the thread waits an L2 round trip for the predicate of every single add.
__device__ __forceinline__ void red_add_global(double* a, double v)
{
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(a)), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add_global(float* a, float v)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(a)), "f"(v) : "memory");
}
__device__ __forceinline__ void red_sub(double* a, double v) { red_add_global(a, -v); }
__device__ __forceinline__ void red_sub(float* a, float v) { red_add_global(a, -v); }
__device__ __forceinline__ void red_sub(pangulu_sm100_c64* a, pangulu_sm100_c64 v) { red_add_global(&a->re, -v.re); red_add_global(&a->im, -v.im); }
__device__ __forceinline__ void red_sub(pangulu_sm100_c32* a, pangulu_sm100_c32 v) { red_add_global(&a->re, -v.re); red_add_global(&a->im, -v.im); }

__device__ inline void vadd(double& x, double a) { x += a; }
__device__ inline void vadd(float& x, float a) { x += a; }
__device__ inline void vadd(pangulu_sm100_c64& x, pangulu_sm100_c64 a) { x.re += a.re; x.im += a.im; }
__device__ inline void vadd(pangulu_sm100_c32& x, pangulu_sm100_c32 a) { x.re += a.re; x.im += a.im; }

}  // namespace psm
#include "sm100_compact.cuh"
namespace psm {

// ------------------------------------------------------------------------------------------------
// SSSSM, sparse path.  grid = (#destinations x column groups); one warp owns one column j of the
// destination and applies EVERY operand pair queued for that destination (ascending elimination
// level) into a dense per-warp accumulator in shared memory, then subtracts it once from C(:,j).
// No atomics: inside one pair/k step the lanes touch distinct rows of A(:,k).
//   smem = warps * nb * sizeof(T)
// Algorithmic HBM bytes per pair: nnzA(sv+si) + nnzB(sv+si) + nnzC(2sv+si) + 3P; applying
// all pairs of a destination in one pass reads/writes C once per batch instead of once per pair.
// ------------------------------------------------------------------------------------------------
// warp shuffles of a value of type T
__device__ __forceinline__ double shfl_v(double v, int src) { return __shfl_sync(PSM_FULL, v, src); }
__device__ __forceinline__ float shfl_v(float v, int src) { return __shfl_sync(PSM_FULL, v, src); }
__device__ __forceinline__ pangulu_sm100_c64 shfl_v(pangulu_sm100_c64 v, int src)
{ pangulu_sm100_c64 r; r.re = __shfl_sync(PSM_FULL, v.re, src); r.im = __shfl_sync(PSM_FULL, v.im, src); return r; }
__device__ __forceinline__ pangulu_sm100_c32 shfl_v(pangulu_sm100_c32 v, int src)
{ pangulu_sm100_c32 r; r.re = __shfl_sync(PSM_FULL, v.re, src); r.im = __shfl_sync(PSM_FULL, v.im, src); return r; }

// One warp, one destination column j: acc (nb entries of shared memory, owned by the warp) <- sum over the
// pairs of A * B(:,j), then C(:,j) -= acc on the pattern.  The B entries of the column are read 32 at a time
// (index, value and the extent of the matching A column per lane) and walked with shuffles, and the first
// 32 entries of the NEXT A column are loaded before the current one is accumulated, so a step costs one
// shared-memory round trip instead of a chain of dependent L2 loads.
template <class T>
__device__ __forceinline__ void ssssm_sparse_column(int nb, const SsssmDest<T>& D, const SsssmPair<T>* __restrict__ pairs, int j, T* acc, bool atomic)
{
    const int lane = threadIdx.x % PSM_WARP;
    const u32 c0 = col_begin(D.c.ptr, j), c1 = D.c.ptr[j + 1];
    const bool diag = D.u.val != nullptr;
    if (!diag && c0 == c1) return;  // empty destination column (padding, or nothing stored there)

    for (int i = lane; i < nb; i += PSM_WARP) acc[i] = zero<T>();
    __syncwarp();

    for (u32 p = D.pair_begin; p < D.pair_end; ++p)
    {
        const BlkRO<T> A = pairs[p].a, B = pairs[p].b;
        const u32 b0 = col_begin(B.ptr, j), b1 = B.ptr[j + 1];
        for (u32 eb = b0; eb < b1; eb += PSM_WARP)
        {
            const u32 e = eb + lane;
            T bv = zero<T>();
            u32 a0 = 0, a1 = 0;
            if (e < b1)
            {
                const int k = B.idx[e];
                bv = B.val[e];
                a0 = col_begin(A.ptr, k);
                a1 = A.ptr[k + 1];
            }
            const int cnt = min((u32)PSM_WARP, b1 - eb);
            // software pipeline over the B entries of this batch: (ci, cv) = this lane's first entry of the current A column
            u32 s0 = __shfl_sync(PSM_FULL, a0, 0), s1 = __shfl_sync(PSM_FULL, a1, 0);
            int ci = -1;
            T cv = zero<T>();
            if (s0 + lane < s1) { ci = A.idx[s0 + lane]; cv = A.val[s0 + lane]; }
            for (int sidx = 0; sidx < cnt; ++sidx)
            {
                const T bvs = shfl_v(bv, sidx);
                u32 n0 = 0, n1 = 0;
                int ni = -1;
                T nv = zero<T>();
                if (sidx + 1 < cnt)
                {
                    n0 = __shfl_sync(PSM_FULL, a0, sidx + 1);
                    n1 = __shfl_sync(PSM_FULL, a1, sidx + 1);
                    if (n0 + lane < n1) { ni = A.idx[n0 + lane]; nv = A.val[n0 + lane]; }
                }
                if (ci >= 0) fma_acc(acc[ci], cv, bvs);
                for (u32 t = s0 + PSM_WARP + lane; t < s1; t += PSM_WARP) fma_acc(acc[A.idx[t]], A.val[t], bvs);
                __syncwarp();
                s0 = n0; s1 = n1; ci = ni; cv = nv;
            }
        }
    }

    // C(:,j) -= acc  (for a diagonal destination D.c is the strict-lower half: rows > j)
    if (atomic)
        for (u32 e = c0 + lane; e < c1; e += PSM_WARP) red_sub(&D.c.val[e], acc[D.c.idx[e]]);
    else
        for (u32 e = c0 + lane; e < c1; e += PSM_WARP) vsub(D.c.val[e], acc[D.c.idx[e]]);
    if (diag)
    {
        // upper half is CSR with ascending columns (diagonal first): entry (r, j), r <= j, sits in row r
        for (int r = lane; r <= j; r += PSM_WARP)
        {
            const T v = acc[r];
            if (is_zero(v)) continue;
            u32 lo = col_begin(D.u.ptr, r), hi = D.u.ptr[r + 1];
            while (lo < hi)
            {
                const u32 mid = (lo + hi) >> 1;
                const int c = D.u.idx[mid];
                if (c < j) lo = mid + 1;
                else hi = mid;
            }
            if (lo < D.u.ptr[r + 1] && D.u.idx[lo] == j)
            {
                if (atomic) red_sub(&D.u.val[lo], v);
                else vsub(D.u.val[lo], v);
            }
        }
    }
    __syncwarp();
}

template <class T, int WARPS>
__global__ void __launch_bounds__(WARPS * PSM_WARP)
ssssm_sparse_kernel(int nb, const SsssmDest<T>* __restrict__ dests, const SsssmPair<T>* __restrict__ pairs, int groups_per_dest)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* acc_all = reinterpret_cast<T*>(smem_raw);
    const int warp = threadIdx.x / PSM_WARP;
    const int d = blockIdx.x / groups_per_dest;
    const int j = (blockIdx.x % groups_per_dest) * WARPS + warp;
    if (j >= nb) return;
    const SsssmDest<T> D = dests[d];
    ssssm_sparse_column<T>(nb, D, pairs, j, acc_all + (size_t)warp * nb, (D.flags & 1u) != 0);
}

// ------------------------------------------------------------------------------------------------
// GESSM / TSTRF (sparse triangular panel solves).  grid = (#tasks x row/column groups); one warp
// owns one column (GESSM) or one row (TSTRF) of B: gather it into a dense shared-memory vector,
// eliminate entry by entry in index order (lanes sweep the matching column of L / row of U), scatter
// back.  kind = PANGULU_TASK_GESSM or PANGULU_TASK_TSTRF.  As in the SSSSM column routine the per-entry
// metadata (extent of the L column / U row, pivot) is fetched 32 entries at a time and the first 32
// elements of the next sweep are loaded ahead.
//   smem = warps * nb * sizeof(T)
// ------------------------------------------------------------------------------------------------
// dop (optional): compact dense copy of the block (sm100_compact.cuh); the finished row / column is written into it too
template <class T>
__device__ __forceinline__ void panel_solve_line(int nb, const PanelTask<T>& P, int rc, T* x, const DenseOp<T>* dop = nullptr)
{
    const int lane = threadIdx.x % PSM_WARP;
    const Blk<T> B = P.b, Dg = P.diag;
    const bool gessm = (P.kind & 0xff) == PANGULU_TASK_GESSM;
    // GESSM: column rc of B (CSC), L = strict-lower CSC of the diagonal block (unit diagonal)
    // TSTRF: row rc of B through its CSR mirror, U = upper CSR of the diagonal block (diagonal first)
    const u32* bptr = gessm ? B.ptr : B.rptr;
    const u16* bidx = gessm ? B.idx : B.ridx;
    const u32 b0 = col_begin(bptr, rc), b1 = bptr[rc + 1];
    if (b0 == b1) return;
    for (int i = lane; i < nb; i += PSM_WARP) x[i] = zero<T>();
    __syncwarp();
    if (gessm)
        for (u32 e = b0 + lane; e < b1; e += PSM_WARP) x[bidx[e]] = B.val[e];
    else
        for (u32 e = b0 + lane; e < b1; e += PSM_WARP) x[bidx[e]] = B.val[B.rmap[e]];
    __syncwarp();
    const int skip = gessm ? 0 : 1;  // the U row starts with the pivot
    for (u32 eb = b0; eb < b1; eb += PSM_WARP)
    {
        const u32 e = eb + lane;
        int r = 0;
        u32 d0 = 0, d1 = 0;
        T piv = zero<T>();
        if (e < b1)
        {
            r = bidx[e];
            d0 = col_begin(Dg.ptr, r);
            d1 = Dg.ptr[r + 1];
            if (!gessm) piv = clamp_pivot(Dg.val[d0]);
        }
        const int cnt = min((u32)PSM_WARP, b1 - eb);
        u32 s0 = __shfl_sync(PSM_FULL, d0, 0) + skip, s1 = __shfl_sync(PSM_FULL, d1, 0);
        int ci = -1;
        T cv = zero<T>();
        if (s0 + lane < s1) { ci = Dg.idx[s0 + lane]; cv = Dg.val[s0 + lane]; }
        for (int sidx = 0; sidx < cnt; ++sidx)
        {
            const int rs = __shfl_sync(PSM_FULL, r, sidx);
            u32 n0 = 0, n1 = 0;
            int ni = -1;
            T nv = zero<T>();
            if (sidx + 1 < cnt)
            {
                n0 = __shfl_sync(PSM_FULL, d0, sidx + 1) + skip;
                n1 = __shfl_sync(PSM_FULL, d1, sidx + 1);
                if (n0 + lane < n1) { ni = Dg.idx[n0 + lane]; nv = Dg.val[n0 + lane]; }
            }
            T v = x[rs];
            if (!gessm)
            {
                v = vdiv(v, shfl_v(piv, sidx));
                __syncwarp();
                if (lane == 0) x[rs] = v;
            }
            if (ci >= 0) fnma(x[ci], v, cv);
            for (u32 q = s0 + PSM_WARP + lane; q < s1; q += PSM_WARP) fnma(x[Dg.idx[q]], v, Dg.val[q]);
            __syncwarp();
            s0 = n0; s1 = n1; ci = ni; cv = nv;
        }
    }
    if (gessm)
        for (u32 e = b0 + lane; e < b1; e += PSM_WARP) B.val[e] = x[bidx[e]];
    else
        for (u32 e = b0 + lane; e < b1; e += PSM_WARP) B.val[B.rmap[e]] = x[bidx[e]];
    if (dop)
    {
        const DenseOp<T> Dn = *dop;
        const int cidx = Dn.inv[rc];  // compact column (GESSM: block of the U panel) / compact row (TSTRF: block of the L panel)
        if (gessm)
            for (u32 e = b0 + lane; e < b1; e += PSM_WARP) dense_store_b(Dn, bidx[e], cidx, x[bidx[e]]);
        else
            for (u32 e = b0 + lane; e < b1; e += PSM_WARP) Dn.dense[(size_t)bidx[e] * Dn.ld + cidx] = x[bidx[e]];
    }
    __syncwarp();
}

template <class T, int WARPS>
__global__ void __launch_bounds__(WARPS * PSM_WARP)
panel_solve_kernel(int nb, const PanelTask<T>* __restrict__ tasks, int groups_per_task, const DenseOp<T>* __restrict__ ops)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* x_all = reinterpret_cast<T*>(smem_raw);
    const int warp = threadIdx.x / PSM_WARP;
    const int t = blockIdx.x / groups_per_task;
    const int rc = (blockIdx.x % groups_per_task) * WARPS + warp;
    if (rc >= nb) return;
    const PanelTask<T> P = tasks[t];
    panel_solve_line<T>(nb, P, rc, x_all + (size_t)warp * nb, (ops && P.dense >= 0) ? ops + P.dense : nullptr);
}

// ------------------------------------------------------------------------------------------------
// GETRF, sparse right-looking LU of one diagonal block per CTA (no pivoting).
// Level l: scale L(:,l) by the clamped pivot and scatter it -- and U(l,:) -- into two dense
// shared-memory vectors; then every warp takes whole rows r in L(:,l) (U(r,:) -= L(r,l) U(l,:)) and
// whole columns c in U(l,:) (L(:,c) -= L(:,l) U(l,c)) and sweeps them coalesced against the dense
// vectors (entries outside the level's pattern meet a zero).  The vectors are double-buffered so a
// level needs two block barriers.
//   smem = 4 * nb * sizeof(T)
// ------------------------------------------------------------------------------------------------
// The level loop is software-pipelined so that a level costs ONE block barrier: while warps 1.. apply level l's
// rank-1 update to the trailing rows / columns, warp 0 applies it to row and column `ln` (the next level that has
// work) first and then already prepares that level (divides L(:,ln) by the clamped pivot, scatters it and U(ln,:)
// into the other pair of dense vectors).  Every entry still sees the same operations in the same order.
template <class T, int THREADS>
__device__ __forceinline__ void getrf_body(int nb, T* lval, const u32* lptr, const u16* lidx, T* uval, const u32* uptr, const u16* uidx, T* buf, int c0 = 0, int c1 = -1)
{
    const int tid = threadIdx.x, warp = tid / PSM_WARP, lane = tid % PSM_WARP;
    constexpr int NW = THREADS / PSM_WARP;
    if (c1 < 0) c1 = nb;   // levels [c0, c1): one independent group of the block (rows / columns outside it are not touched)
    // next level >= from that has something to do: a non-empty U row (rows beyond n are padding,
    // pangulu_platform_0100000.c:74-77) and more than the pivot alone
    auto next_level = [&](int from) {
        for (int l = from; l < c1; ++l)
        {
            const u32 u0 = col_begin(uptr, l), u1 = uptr[l + 1];
            if (u0 == u1) continue;
            if (col_begin(lptr, l) == lptr[l + 1] && u1 - u0 == 1) continue;
            return l;
        }
        return c1;
    };
    // warp 0: scale L(:,l), scatter it and U(l,:) into the vectors of parity `par`, after clearing what level `old`
    // left there
    auto prepare = [&](int l, int par, int old) {
        T* lcol = buf + (size_t)(par * 2) * nb;
        T* urow = lcol + nb;
        if (old >= 0)
        {
            for (u32 e = col_begin(lptr, old) + lane; e < lptr[old + 1]; e += PSM_WARP) lcol[lidx[e]] = zero<T>();
            for (u32 e = col_begin(uptr, old) + 1 + lane; e < uptr[old + 1]; e += PSM_WARP) urow[uidx[e]] = zero<T>();
            __syncwarp();
        }
        const u32 u0 = col_begin(uptr, l), u1 = uptr[l + 1];
        const T piv = clamp_pivot(uval[u0]);
        for (u32 e = col_begin(lptr, l) + lane; e < lptr[l + 1]; e += PSM_WARP)
        {
            const T v = vdiv(lval[e], piv);
            lval[e] = v;
            lcol[lidx[e]] = v;
        }
        for (u32 e = u0 + 1 + lane; e < u1; e += PSM_WARP) urow[uidx[e]] = uval[e];
    };
    auto update_row = [&](int r, const T* lcol, const T* urow) {   // U(r,:) -= L(r,l) U(l,:)
        const T lr = lcol[r];
        const u32 r0 = col_begin(uptr, r), r1 = uptr[r + 1];
        for (u32 q = r0 + lane; q < r1; q += PSM_WARP)
        {
            T v = uval[q];
            fnma(v, lr, urow[uidx[q]]);
            uval[q] = v;
        }
    };
    auto update_col = [&](int c, const T* lcol, const T* urow) {   // L(:,c) -= L(:,l) U(l,c)
        const T uc = urow[c];
        const u32 q0 = col_begin(lptr, c), q1 = lptr[c + 1];
        for (u32 q = q0 + lane; q < q1; q += PSM_WARP)
        {
            T v = lval[q];
            fnma(v, lcol[lidx[q]], uc);
            lval[q] = v;
        }
    };

    int level = next_level(c0), prev_level = -1, par = 0;
    if (level >= c1) return;
    if (warp == 0) prepare(level, 0, -1);
    __syncthreads();
    while (level < c1)
    {
        const int ln = next_level(level + 1);
        const T* lcol = buf + (size_t)(par * 2) * nb;
        const T* urow = lcol + nb;
        const u32 l0 = col_begin(lptr, level), l1 = lptr[level + 1];
        const u32 u0 = col_begin(uptr, level), u1 = uptr[level + 1];
        const int nrow = (int)(l1 - l0), ncol = (int)(u1 - u0 - 1);
        if (warp == 0)
        {
            // critical path: row ln and column ln first (if level `level` touches them), then prepare level ln
            if (ln < c1)
            {
                bool hit = false;
                for (u32 e = l0 + lane; e < l1 && !hit; e += PSM_WARP) hit = lidx[e] == ln;
                if (__any_sync(PSM_FULL, hit)) update_row(ln, lcol, urow);
                hit = false;
                for (u32 e = u0 + 1 + lane; e < u1 && !hit; e += PSM_WARP) hit = uidx[e] == ln;
                if (__any_sync(PSM_FULL, hit)) update_col(ln, lcol, urow);
                __syncwarp();
                prepare(ln, par ^ 1, prev_level);
            }
        }
        else
        {
            for (int it = warp - 1; it < nrow + ncol; it += NW - 1)
            {
                if (it < nrow)
                {
                    const int r = lidx[l0 + it];
                    if (r != ln) update_row(r, lcol, urow);
                }
                else
                {
                    const int c = uidx[u0 + 1 + (it - nrow)];
                    if (c != ln) update_col(c, lcol, urow);
                }
            }
        }
        __syncthreads();
        prev_level = level;
        level = ln;
        par ^= 1;
    }
}

// variant 1: the block stays in global memory (L2); used when it does not fit in shared memory
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
getrf_sparse_kernel(int nb, const GetrfTask<T>* __restrict__ tasks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* buf = reinterpret_cast<T*>(smem_raw);  // [2][2][nb] : parity x {lcol, urow}
    const GetrfTask<T> G = tasks[blockIdx.x];
    for (int i = threadIdx.x; i < 4 * nb; i += THREADS) buf[i] = zero<T>();
    __syncthreads();
    getrf_body<T, THREADS>(nb, G.lower.val, G.lower.ptr, G.lower.idx, G.upper.val, G.upper.ptr, G.upper.idx, buf, G.c0, G.c1);
}

// variant 2: the whole block (values, pointers, indices of both halves) is staged in shared memory, so a
// level costs shared-memory latency instead of L2 round trips -- the diagonal blocks sit on the critical
// path of the factorisation (one GETRF per block row, chained).
//   smem = (4 nb + nnzL + nnzU) sizeof(T) + 8 (nb+1) + 2 (nnzL + nnzU), rounded; chosen per launch as the
//   maximum over the batch
template <class T, int THREADS>
__device__ __forceinline__ void getrf_smem_cta(unsigned char* smem_raw, int nb, const GetrfTask<T>& G)
{
    const int tid = threadIdx.x;
    // only the entries of the group [c0, c1) are staged (columns c0..c1-1 of the lower half, rows c0..c1-1 of the upper
    // half: the block is block diagonal over its groups); the shared-memory arrays are addressed with GLOBAL entry numbers
    const int c0 = G.c0, c1 = G.c1;
    const u32 lb = col_begin(G.lower.ptr, c0), le = G.lower.ptr[c1], ub = col_begin(G.upper.ptr, c0), ue = G.upper.ptr[c1];
    const u32 nl = le - lb, nu = ue - ub;
    T* buf = reinterpret_cast<T*>(smem_raw);
    T* lval = buf + 4 * nb;
    T* uval = lval + nl;
    u32* lptr = reinterpret_cast<u32*>(uval + nu);
    u32* uptr = lptr + (nb + 1);
    u16* lidx = reinterpret_cast<u16*>(uptr + (nb + 1));
    u16* uidx = lidx + nl;
    for (int i = tid; i < 4 * nb; i += THREADS) buf[i] = zero<T>();
    for (u32 i = tid; i < nl; i += THREADS) { lval[i] = G.lower.val[lb + i]; lidx[i] = G.lower.idx[lb + i]; }
    for (u32 i = tid; i < nu; i += THREADS) { uval[i] = G.upper.val[ub + i]; uidx[i] = G.upper.idx[ub + i]; }
    for (int i = tid; i <= nb; i += THREADS) { lptr[i] = i == 0 ? 0u : G.lower.ptr[i]; uptr[i] = i == 0 ? 0u : G.upper.ptr[i]; }
    __syncthreads();
    getrf_body<T, THREADS>(nb, lval - lb, lptr, lidx - lb, uval - ub, uptr, uidx - ub, buf, c0, c1);
    __syncthreads();
    for (u32 i = tid; i < nl; i += THREADS) G.lower.val[lb + i] = lval[i];
    for (u32 i = tid; i < nu; i += THREADS) G.upper.val[ub + i] = uval[i];
}

template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
getrf_smem_kernel(int nb, const GetrfTask<T>* __restrict__ tasks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GetrfTask<T> G = tasks[blockIdx.x];
    getrf_smem_cta<T, THREADS>(smem_raw, nb, G);
}

// variant 3: dense-promoted GETRF for diagonal blocks too full for variant 2 (the separators' blocks).
// The block is expanded into a column-major nb x nb tile in an L2-resident global scratch, factored by ONE
// CTA with a blocked right-looking LU (panel width 16: panel columns live in registers one row per thread,
// one barrier per eliminated column; U12 by forward substitution one column per thread; trailing update
// with a 4x4 register tile per thread from shared-memory copies of the panel), and the entries of the
// pattern are gathered back.  Structural zeros stay exact zeros (they only ever receive 0*x), and every
// entry sees the same sequence "divide by the clamped pivot, then x -= l*u in ascending level order" as in
// the sparse variants, so the three GETRF kernels agree bit for bit.
//   smem = 2 * 16 * nbp * sizeof(T) + 64 * sizeof(T), nbp = nb rounded up to 32;  scratch = nbp*nbp*sizeof(T) per task
#define PSM_GETRF_W 16
template <class T, int THREADS>
__device__ __forceinline__ void getrf_dense_cta(unsigned char* smem_raw, int nb_block, const GetrfTask<T>& G, T* __restrict__ Tm)
{
    constexpr int W = PSM_GETRF_W;
    constexpr int NW = THREADS / PSM_WARP;
    // the task's group [c0, c1) of the block is factored as a dense nb x nb matrix in LOCAL coordinates (index - off)
    (void)nb_block;
    const int off = G.c0, nb = G.c1 - G.c0;
    const int nbp = (nb + 31) & ~31;
    const int ldp = nbp + 4;                  // panel stride = 4 mod 8: conflict-free DMMA fragment loads
    T* Ls = reinterpret_cast<T*>(smem_raw);   // [W][ldp]  L panel (rows >= j0)
    T* Us = Ls + W * ldp;                     // [W][ldp]  U12 rows of the panel
    T* urow = Us + W * ldp;                   // [2][W+...] pivot row of the current column, double buffered
    constexpr bool kDouble = sizeof(T) == 8 && !is_complex<T>::value;
    const Blk<T> Lo = G.lower, Up = G.upper;  // Tm: column major, leading dimension nbp
    const int tid = threadIdx.x, warp = tid / PSM_WARP, lane = tid % PSM_WARP;

    // ---- expand: zero, then scatter the strict-lower CSC half and the upper CSR half
    for (int i = tid; i < nbp * nbp; i += THREADS) Tm[i] = zero<T>();
    __syncthreads();
    for (int c = warp; c < nb; c += NW)
    {
        for (u32 e = col_begin(Lo.ptr, off + c) + lane; e < Lo.ptr[off + c + 1]; e += PSM_WARP) Tm[(size_t)c * nbp + (Lo.idx[e] - off)] = Lo.val[e];
        for (u32 e = col_begin(Up.ptr, off + c) + lane; e < Up.ptr[off + c + 1]; e += PSM_WARP) Tm[(size_t)(Up.idx[e] - off) * nbp + c] = Up.val[e];  // row c of U
    }
    __syncthreads();

    const int r = tid;  // this thread's row in the panel phase
    for (int j0 = 0; j0 < nb; j0 += W)
    {
        const int w = min(W, nb - j0);
        // ---- panel: columns j0 .. j0+w-1, rows j0 .. nb-1, one row per thread in registers
        T p[W];
#pragma unroll
        for (int q = 0; q < W; ++q) p[q] = (q < w && r >= j0 && r < nb) ? Tm[(size_t)(j0 + q) * nbp + r] : zero<T>();
#pragma unroll
        for (int jj = 0; jj < W; ++jj)
        {
            if (jj < w)
            {
                const int l = j0 + jj;
                T* ur = urow + (jj & 1) * W;
                if (r == l)
                {
#pragma unroll
                    for (int q = 0; q < W; ++q) ur[q] = p[q];
                }
                __syncthreads();
                if (r > l && r < nb)
                {
                    const T piv = clamp_pivot(ur[jj]);
                    p[jj] = vdiv(p[jj], piv);
#pragma unroll
                    for (int q = 0; q < W; ++q)
                        if (q > jj) fnma(p[q], p[jj], ur[q]);
                }
            }
        }
        // publish the panel (L part below the diagonal of each column, U part above) to memory and to Ls
        if (r >= j0 && r < nb)
        {
#pragma unroll
            for (int q = 0; q < W; ++q)
                if (q < w) { Tm[(size_t)(j0 + q) * nbp + r] = p[q]; Ls[q * ldp + r] = p[q]; }
        }
        __syncthreads();
        const int t0 = j0 + w;  // first trailing row/column
        if (t0 >= nb) break;
        // ---- U12 = L11^-1 T12: one trailing column per thread, forward substitution over the w panel rows
        for (int c = t0 + tid; c < nb; c += THREADS)
        {
            T u[W];
#pragma unroll
            for (int q = 0; q < W; ++q) u[q] = q < w ? Tm[(size_t)c * nbp + j0 + q] : zero<T>();
#pragma unroll
            for (int q = 1; q < W; ++q)
#pragma unroll
                for (int s2 = 0; s2 < W; ++s2)
                    if (s2 < q && q < w) fnma(u[q], Ls[s2 * ldp + j0 + q], u[s2]);
#pragma unroll
            for (int q = 0; q < W; ++q)
                if (q < w) { Tm[(size_t)c * nbp + j0 + q] = u[q]; Us[q * ldp + c] = u[q]; }
        }
        __syncthreads();
        // ---- trailing update T22 -= L21 U12
        if constexpr (kDouble)
        {
            // FP64 tensor cores: every warp takes 16 x 16 pieces of the trailing matrix (2 x 2 DMMA tiles), reads them from the
            // L2-resident scratch, subtracts the rank-16 product of the two panels (fragments from shared memory, A negated)
            // and writes them back.  Rows / columns between nb and its multiple of 16 carry stale panel data; they only ever
            // meet themselves and are not stored.
            const int lr = lane / 4, lk = lane % 4;
            const int ng = (nb - t0 + 15) / 16;
            for (int gidx = warp; gidx < ng * ng; gidx += NW)
            {
                const int r0 = t0 + (gidx % ng) * 16, c0 = t0 + (gidx / ng) * 16;
                double cf[2][2][2];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                        {
                            const int rr2 = r0 + 8 * i + lr, cc2 = c0 + 8 * j + 2 * lk + h;
                            cf[i][j][h] = (rr2 < nb && cc2 < nb) ? Tm[(size_t)cc2 * nbp + rr2] : 0.0;
                        }
#pragma unroll
                for (int k4 = 0; k4 < W; k4 += 4)
                {
                    double a[2], b2[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) a[i] = -Ls[(k4 + lk) * ldp + r0 + 8 * i + lr];
#pragma unroll
                    for (int j = 0; j < 2; ++j) b2[j] = Us[(k4 + lk) * ldp + c0 + 8 * j + lr];
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) dmma884c(cf[i][j][0], cf[i][j][1], a[i], b2[j]);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                        {
                            const int rr2 = r0 + 8 * i + lr, cc2 = c0 + 8 * j + 2 * lk + h;
                            if (rr2 < nb && cc2 < nb) Tm[(size_t)cc2 * nbp + rr2] = cf[i][j][h];
                        }
            }
        }
        else
        // 4x4 register tile per thread: rows lane + 32 i, columns warp + NW j
        for (int cb = t0; cb < nb; cb += 4 * NW)
            for (int rb = t0; rb < nb; rb += 128)
            {
                T acc[4][4];
                int rr[4], cc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) rr[i] = rb + lane + 32 * i;
#pragma unroll
                for (int j = 0; j < 4; ++j) cc[j] = cb + warp + NW * j;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = (rr[i] < nb && cc[j] < nb) ? Tm[(size_t)cc[j] * nbp + rr[i]] : zero<T>();
#pragma unroll
                for (int q = 0; q < W; ++q)
                {
                    if (q < w)
                    {
                        T lv[4], uv[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) lv[i] = rr[i] < nbp ? Ls[q * ldp + rr[i]] : zero<T>();
#pragma unroll
                        for (int j = 0; j < 4; ++j) uv[j] = cc[j] < nbp ? Us[q * ldp + cc[j]] : zero<T>();
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) fnma(acc[i][j], lv[i], uv[j]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (rr[i] < nb && cc[j] < nb) Tm[(size_t)cc[j] * nbp + rr[i]] = acc[i][j];
            }
        __syncthreads();
    }
    __syncthreads();
    // ---- gather the pattern back
    for (int c = warp; c < nb; c += NW)
    {
        for (u32 e = col_begin(Lo.ptr, off + c) + lane; e < Lo.ptr[off + c + 1]; e += PSM_WARP) Lo.val[e] = Tm[(size_t)c * nbp + (Lo.idx[e] - off)];
        for (u32 e = col_begin(Up.ptr, off + c) + lane; e < Up.ptr[off + c + 1]; e += PSM_WARP) Up.val[e] = Tm[(size_t)(Up.idx[e] - off) * nbp + c];
    }
}

template <class T>
__global__ void __launch_bounds__(512)
getrf_dense_kernel(int nb, const GetrfTask<T>* __restrict__ tasks, T* __restrict__ scratch_all)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbp = (nb + 31) & ~31;
    const GetrfTask<T> G = tasks[blockIdx.x];
    getrf_dense_cta<T, 512>(smem_raw, nb, G, scratch_all + (size_t)blockIdx.x * nbp * nbp);
}

// ------------------------------------------------------------------------------------------------
// solve-phase helpers (pangulu_platform_0100000.c:435-506) on device arrays
// ------------------------------------------------------------------------------------------------
template <class T>
__device__ inline void atomic_sub(T* addr, T v);
template <> __device__ inline void atomic_sub<double>(double* a, double v) { atomicAdd(a, -v); }
template <> __device__ inline void atomic_sub<float>(float* a, float v) { atomicAdd(a, -v); }
template <> __device__ inline void atomic_sub<pangulu_sm100_c64>(pangulu_sm100_c64* a, pangulu_sm100_c64 v)
{ atomicAdd(&a->re, -v.re); atomicAdd(&a->im, -v.im); }
template <> __device__ inline void atomic_sub<pangulu_sm100_c32>(pangulu_sm100_c32* a, pangulu_sm100_c32 v)
{ atomicAdd(&a->re, -v.re); atomicAdd(&a->im, -v.im); }

// (vmul: sm100_types.cuh)

// y -= A x, A in CSC: one warp per column
template <class T>
__global__ void spmv_kernel(int nb, Blk<T> A, const T* __restrict__ x, T* __restrict__ y)
{
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) / PSM_WARP, lane = threadIdx.x % PSM_WARP;
    if (col >= nb) return;
    const T xv = x[col];
    for (u32 e = col_begin(A.ptr, col) + lane; e < A.ptr[col + 1]; e += PSM_WARP) atomic_sub(&y[A.idx[e]], vmul(A.val[e], xv));
}

template <class T> __device__ inline T warp_sum(T s)
{
    for (int o = PSM_WARP / 2; o > 0; o >>= 1)
    {
        if constexpr (is_complex<T>::value)
        {
            s.re += __shfl_down_sync(PSM_FULL, s.re, o);
            s.im += __shfl_down_sync(PSM_FULL, s.im, o);
        }
        else
            s += __shfl_down_sync(PSM_FULL, s, o);
    }
    return s;
}

// b += x
template <class T>
__global__ void vecadd_kernel(long long n, T* __restrict__ b, const T* __restrict__ x)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { T v = b[i]; vadd(v, x[i]); b[i] = v; }
}

// in-block triangular solve with one CTA: uplo 0 = unit-lower CSC forward (column sweeps),
// 1 = upper CSR backward (row dot products, diagonal first in each row, pivot clamp of
// pangulu_platform_0100000.c:500-503)
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS) sptrsv_kernel(int nb, Blk<T> S, T* __restrict__ xg, int uplo, unsigned smem_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* x = reinterpret_cast<T*>(smem_raw);
    __shared__ T part[THREADS / PSM_WARP];
    for (int i = threadIdx.x; i < nb; i += THREADS) x[i] = xg[i];
    // The substitution is a chain of nb dependent steps; with the half's arrays in global memory every step pays an L2 round
    // trip (~130 us per 256 x 256 half).  Stage pointers, indices and values in shared memory when they fit.
    {
        const u32 nnz = S.ptr[nb];
        const size_t off_ptr = ((size_t)nb * sizeof(T) + 15) & ~(size_t)15;
        const size_t off_val = (off_ptr + 4 * (size_t)(nb + 1) + 15) & ~(size_t)15;
        const size_t off_idx = off_val + (size_t)nnz * sizeof(T);
        if (off_idx + 2 * (size_t)nnz <= smem_bytes)
        {
            u32* sptr = reinterpret_cast<u32*>(smem_raw + off_ptr);
            T* sval = reinterpret_cast<T*>(smem_raw + off_val);
            u16* sidx = reinterpret_cast<u16*>(smem_raw + off_idx);
            for (int i = threadIdx.x; i <= nb; i += THREADS) sptr[i] = i == 0 ? 0u : S.ptr[i];
            for (u32 e = threadIdx.x; e < nnz; e += THREADS) { sval[e] = S.val[e]; sidx[e] = S.idx[e]; }
            S.ptr = sptr; S.val = sval; S.idx = sidx;
        }
    }
    __syncthreads();
    if (uplo == 0)
    {
        for (int i = 0; i < nb; ++i)
        {
            const u32 e0 = col_begin(S.ptr, i), e1 = S.ptr[i + 1];
            if (e0 == e1) continue;   // (uniform: nothing to apply, nothing to wait for)
            const T xi = x[i];         // final: every earlier column was applied before the barrier that ended its step
            for (u32 e = e0 + threadIdx.x; e < e1; e += THREADS) fnma(x[S.idx[e]], S.val[e], xi);
            __syncthreads();
        }
    }
    else
    {
        for (int i = nb - 1; i >= 0; --i)
        {
            const u32 r0 = col_begin(S.ptr, i), r1 = S.ptr[i + 1];
            if (r0 == r1) continue;
            if (r1 - r0 <= 1 + PSM_WARP)
            {
                // short row: one warp, no block barrier for the partial sums
                if (threadIdx.x < PSM_WARP)
                {
                    T s = zero<T>();
                    const u32 e = r0 + 1 + threadIdx.x;
                    if (e < r1) fma_acc(s, S.val[e], x[S.idx[e]]);
                    s = warp_sum(s);
                    if (threadIdx.x == 0)
                    {
                        T xi = x[i];
                        vsub(xi, s);
                        x[i] = vdiv(xi, clamp_pivot(S.val[r0]));
                    }
                }
                __syncthreads();
                continue;
            }
            T s = zero<T>();
            for (u32 e = r0 + 1 + threadIdx.x; e < r1; e += THREADS) fma_acc(s, S.val[e], x[S.idx[e]]);
            s = warp_sum(s);
            if (threadIdx.x % PSM_WARP == 0) part[threadIdx.x / PSM_WARP] = s;
            __syncthreads();
            if (threadIdx.x == 0)
            {
                T tot = zero<T>();
                for (int w = 0; w < THREADS / PSM_WARP; ++w) vadd(tot, part[w]);
                T xi = x[i];
                vsub(xi, tot);
                x[i] = vdiv(xi, clamp_pivot(S.val[r0]));
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < nb; i += THREADS) xg[i] = x[i];
}

// ------------------------------------------------------------------------------------------------
// Block forward / backward substitution with the factors resident in HBM (pangulu_gstrs, src/pangulu_sptrsv.c:24-174 with
// the block kernels above).  One launch per block row and phase instead of one per block:
//   solve_spmv_row_kernel : x[row] -= sum over the row's off-diagonal blocks  A(row, j) x[j]   (CTA per block, warp per
//                           column, FP reductions into the row's nb entries -- the sum order is free, as in the reference
//                           with several ranks);
//   sptrsv_kernel         : the row's diagonal half in place.
// ------------------------------------------------------------------------------------------------
template <class T> struct SolveBlk { const T* val; const u32* ptr; const u16* idx; long long xoff; };

template <class T>
__global__ void __launch_bounds__(256) solve_spmv_row_kernel(int nb, const SolveBlk<T>* __restrict__ blks, T* __restrict__ x, long long yoff)
{
    const SolveBlk<T> b = blks[blockIdx.x];
    const int warp = threadIdx.x / PSM_WARP, lane = threadIdx.x % PSM_WARP;
    T* y = x + yoff;
    for (int col = warp; col < nb; col += 8)
    {
        const u32 e0 = col_begin(b.ptr, col), e1 = b.ptr[col + 1];
        if (e0 == e1) continue;
        const T xv = x[b.xoff + col];
        if (is_zero(xv)) continue;
        for (u32 e = e0 + lane; e < e1; e += PSM_WARP) atomic_sub(&y[b.idx[e]], vmul(b.val[e], xv));
    }
}

// ================================================================================================
// SSSSM, dense-tile path ("promoted" blocks: the only genuinely dense contraction of the factorisation).
//
// One CTA owns a TM x 32 tile of the destination (all rows of the block, 32 columns) and keeps it as
// an FP64 accumulator in registers across EVERY operand pair queued for that destination.  Per pair
// the k dimension is swept in chunks of 16: the 16 columns of A and the 16 x 32 piece of B are
// scattered from their CSC arrays into zero-filled shared-memory tiles (k-major, row stride TM+4 /
// 36 doubles = conflict-free for both the scatter and the fragment loads) and multiplied with
// DMMA.8x8x4 (mma.sync.m8n8k4.f64 -- the FP64 tensor-core instruction of sm_100a; FP32 uses an FFMA
// register tile on the same shared-memory layout, no TF32).  Chunks whose A columns are all empty
// are skipped.  The epilogue spills the accumulator through shared memory and subtracts it from the
// destination's CSC entries (and, for a diagonal destination, from the upper CSR half), so the
// nonzero pattern stays untouched.  Double-buffered tiles, two block barriers per chunk.
//   smem = max(2 * 16 * (TM+4 + 36), 32 * (TM+4)) * sizeof(T)
// ================================================================================================
#define PSM_TILE_N 32
#define PSM_TILE_K 16

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <class T, int TM>
__global__ void __launch_bounds__(256, 2)
ssssm_tile_kernel(int nb, const SsssmDest<T>* __restrict__ dests, const SsssmPair<T>* __restrict__ pairs, int tiles_per_dest)
{
    static_assert(TM % 32 == 0 && TM >= 64 && TM <= 256, "tile height");
    constexpr int TN = PSM_TILE_N, KC = PSM_TILE_K;
    constexpr int LDA = TM + 4, LDB = TN + 4;
    constexpr int WM = TM / 32;          // warps along M
    constexpr int WN = 8 / WM;           // warps along N
    constexpr int WTN = TN / WN;         // columns per warp
    constexpr int NT = WTN / 8;          // 8-column DMMA tiles per warp (double path)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sm = reinterpret_cast<T*>(smem_raw);
    T* As0 = sm;                          // [2][KC][LDA]
    T* Bs0 = sm + 2 * KC * LDA;           // [2][KC][LDB]
    T* Cs = sm;                           // [TN][LDA], epilogue only

    const int tid = threadIdx.x, warp = tid / PSM_WARP, lane = tid % PSM_WARP;
    const int d = blockIdx.x / tiles_per_dest;
    const int c0 = (blockIdx.x % tiles_per_dest) * TN;
    const SsssmDest<T> D = dests[d];
    const bool diag = D.u.val != nullptr;
    const int ncol = min(TN, nb - c0);
    // nothing to update in these columns?
    if (!diag && col_begin(D.c.ptr, c0) == D.c.ptr[c0 + ncol]) return;

    // accumulators
    constexpr bool kDouble = sizeof(T) == 8 && !is_complex<T>::value;
    constexpr int NACC = kDouble ? (4 * NT * 2) : (TM / 32) * 4;  // per thread
    T acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = zero<T>();
    const int wm = warp % WM, wn = warp / WM;
    const int m_base = wm * 32, n_base = wn * WTN;

    // zero both tile buffers once
    for (int i = tid; i < 2 * KC * (LDA + LDB); i += 256) sm[i] = zero<T>();
    __syncthreads();

    // loader roles
    const int a_kk = tid / 16, a_l = tid % 16;   // 16 threads per A column
    const int b_jj = tid / 8, b_l = tid % 8;     // 8 threads per B column
    const int b_col = c0 + b_jj;
    const unsigned grp_shift = (lane / 8) * 8;

    int it = 0;
    for (u32 p = D.pair_begin; p < D.pair_end; ++p)
    {
        const BlkRO<T> A = pairs[p].a, B = pairs[p].b;
        u32 cur = 0, bend = 0;
        if (b_jj < ncol) { cur = col_begin(B.ptr, b_col); bend = B.ptr[b_col + 1]; }
        for (int k0 = 0; k0 < nb; k0 += KC)
        {
            const int kc = min(KC, nb - k0);
            const bool a_has = A.ptr[k0 + kc] > col_begin(A.ptr, k0);
            const int buf = it & 1;
            T* As = As0 + buf * KC * LDA;
            T* Bs = Bs0 + buf * KC * LDB;
            // B piece: rows [k0, k0+kc) of the 32 columns, and advance the per-column cursor
            {
                bool in1 = false, in2 = false;
                u32 e1 = cur + b_l, e2 = cur + 8 + b_l;
                int r1 = 0, r2 = 0;
                if (e1 < bend) { r1 = B.idx[e1]; in1 = r1 < k0 + kc; }
                if (e2 < bend) { r2 = B.idx[e2]; in2 = r2 < k0 + kc; }
                if (a_has)
                {
                    if (in1) Bs[(r1 - k0) * LDB + b_jj] = B.val[e1];
                    if (in2) Bs[(r2 - k0) * LDB + b_jj] = B.val[e2];
                }
                const unsigned m1 = __ballot_sync(PSM_FULL, in1), m2 = __ballot_sync(PSM_FULL, in2);
                cur += __popc((m1 >> grp_shift) & 0xffu) + __popc((m2 >> grp_shift) & 0xffu);
            }
            if (!a_has) continue;  // uniform: nothing to multiply in this chunk
            // A piece: columns k0 .. k0+kc-1, all rows
            if (a_kk < kc)
            {
                const int k = k0 + a_kk;
                const u32 a0 = col_begin(A.ptr, k), a1 = A.ptr[k + 1];
                T* Ak = As + a_kk * LDA;
                for (u32 e = a0 + a_l; e < a1; e += 16) Ak[A.idx[e]] = A.val[e];
            }
            __syncthreads();
            // clear the other buffer for the next chunk (its last reader finished before the barrier above)
            {
                T* Ao = As0 + (buf ^ 1) * KC * LDA;
                T* Bo = Bs0 + (buf ^ 1) * KC * LDB;
                for (int i = tid; i < KC * LDA; i += 256) Ao[i] = zero<T>();
                for (int i = tid; i < KC * LDB; i += 256) Bo[i] = zero<T>();
            }
            if constexpr (kDouble)
            {
                const int lr = lane / 4, lk = lane % 4;
#pragma unroll
                for (int k4 = 0; k4 < KC; k4 += 4)
                {
                    double a[4], b[NT];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) a[mt] = As[(k4 + lk) * LDA + m_base + mt * 8 + lr];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) b[nt] = Bs[(k4 + lk) * LDB + n_base + nt * 8 + lr];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) dmma884(acc[(mt * NT + nt) * 2], acc[(mt * NT + nt) * 2 + 1], a[mt], b[nt]);
                }
            }
            else
            {
                // FMA register tile: thread (tx, ty) owns rows tx + 32*i (i < TM/32) and columns ty*4 .. ty*4+3
                const int tx = tid % 32, ty = tid / 32;
#pragma unroll 4
                for (int kk = 0; kk < KC; ++kk)
                {
                    T bv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) bv[j] = Bs[kk * LDB + ty * 4 + j];
#pragma unroll
                    for (int i = 0; i < TM / 32; ++i)
                    {
                        const T av = As[kk * LDA + tx + 32 * i];
#pragma unroll
                        for (int j = 0; j < 4; ++j) fma_acc(acc[i * 4 + j], av, bv[j]);
                    }
                }
            }
            __syncthreads();
            ++it;
        }
    }

    // epilogue: accumulator -> shared memory (column major [TN][LDA]) -> destination pattern
    __syncthreads();
    if constexpr (kDouble)
    {
        const int lr = lane / 4, lc = (lane % 4) * 2;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
            {
                const int r = m_base + mt * 8 + lr, c = n_base + nt * 8 + lc;
                Cs[c * LDA + r] = acc[(mt * NT + nt) * 2];
                Cs[(c + 1) * LDA + r] = acc[(mt * NT + nt) * 2 + 1];
            }
    }
    else
    {
        const int tx = tid % 32, ty = tid / 32;
#pragma unroll
        for (int i = 0; i < TM / 32; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) Cs[(ty * 4 + j) * LDA + tx + 32 * i] = acc[i * 4 + j];
    }
    __syncthreads();
    for (int jj = warp; jj < ncol; jj += 8)
    {
        const int j = c0 + jj;
        const u32 e0 = col_begin(D.c.ptr, j), e1 = D.c.ptr[j + 1];
        const T* Cj = Cs + jj * LDA;
        for (u32 e = e0 + lane; e < e1; e += PSM_WARP) vsub(D.c.val[e], Cj[D.c.idx[e]]);
    }
    if (diag)
    {
        // upper CSR half: row r holds columns >= r ascending; the tile covers columns [c0, c0+ncol)
        for (int r = tid; r < min(nb, c0 + ncol); r += 256)
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
                vsub(D.u.val[e], Cs[(c - c0) * LDA + r]);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// SSSSM dense-tile path, pipelined variant (the default).  Same tiling idea as ssssm_tile_kernel, but:
//  * 512 threads (16 warps): one warp gathers one of the 16 A columns of a chunk, 16 threads gather one
//    of the 32 B columns; warp grid (TM/32) x (16/(TM/32)) for the DMMA phase;
//  * the CSC entries of chunk i+1 are fetched into REGISTERS before the DMMA phase of chunk i and
//    scattered into the other shared-memory buffer afterwards, so L2 latency hides behind the tensor
//    work instead of sitting between two barriers;
//  * a destination may be split over several CTAs by operand-pair range (flags & 1): those CTAs
//    subtract their partial product with FP64 global reductions instead of plain read-modify-write.
//   smem = max(2 * 16 * (TM+4 + 36), 32 * (TM+4)) * sizeof(T)
// ------------------------------------------------------------------------------------------------
template <class T, int TM>
__device__ __forceinline__ void ssssm_tile2_cta(unsigned char* smem_raw, int nb, const SsssmDest<T>& D, const SsssmPair<T>* __restrict__ pairs, int c0, bool force_atomic)
{
    static_assert(TM == 128 || TM == 256, "tile height");
    constexpr int TN = PSM_TILE_N, KC = PSM_TILE_K;
    constexpr int LDA = TM + 4, LDB = TN + 4;
    constexpr int WM = TM / 32, WN = 16 / WM, WTN = TN / WN, NT = WTN / 8;
    constexpr int AR = TM / 32;  // A entries per lane per chunk column
    constexpr bool kDouble = sizeof(T) == 8 && !is_complex<T>::value;
    T* sm = reinterpret_cast<T*>(smem_raw);
    T* As0 = sm;
    T* Bs0 = sm + 2 * KC * LDA;
    T* Cs = sm;

    const int tid = threadIdx.x, warp = tid / PSM_WARP, lane = tid % PSM_WARP;
    const bool diag = D.u.val != nullptr;
    const int ncol = min(TN, nb - c0);
    if (!diag && col_begin(D.c.ptr, c0) == D.c.ptr[c0 + ncol]) return;

    constexpr int NACC = kDouble ? (4 * NT * 2) : (TM / 32) * 2;
    T acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = zero<T>();
    const int wm = warp % WM, wn = warp / WM;
    const int m_base = wm * 32, n_base = wn * WTN;

    for (int i = tid; i < 2 * KC * (LDA + LDB); i += 512) sm[i] = zero<T>();

    // ---- fetch stage state -----------------------------------------------------------------------------
    // Per operand pair a short prologue puts A's column pointers and, for each of the 32 B columns, the offset of
    // the first entry of every 16-row chunk into shared memory.  After that a chunk's loads need no dependent
    // global reads: they are issued right after the barrier and land while the tensor phase runs.
    constexpr size_t kBig = (2 * KC * (LDA + LDB) > TN * LDA) ? 2 * KC * (LDA + LDB) : TN * LDA;
    u32* sAptr = reinterpret_cast<u32*>(sm + kBig);          // [TM + 1]
    int* sBoff = reinterpret_cast<int*>(sAptr + (TM + 4));    // [TN][KCH + 1], KCH = TM / KC chunks
    constexpr int KCH = TM / KC;
    const int b_jj = tid / 16, b_l = tid % 16;
    const int b_col = c0 + b_jj;
    int k0 = 0;
    BlkRO<T> A, B;
    u32 bbeg = 0;
    u16 aidx[AR];
    T aval[AR];
    int bidx = -1;
    T bval = zero<T>();
    // next chunk of the CURRENT pair whose A columns are not all empty -> registers; false at the end of the pair
    auto fetch = [&]() -> bool {
        for (; k0 < nb; k0 += KC)
        {
            const int kc = min(KC, nb - k0);
            if (sAptr[k0 + kc] == sAptr[k0]) continue;
            const int ch = k0 / KC;
            const int o0 = sBoff[b_jj * (KCH + 1) + ch], o1 = sBoff[b_jj * (KCH + 1) + ch + 1];
            bidx = -1;
            if (o0 + b_l < o1) { bidx = B.idx[bbeg + o0 + b_l]; bval = B.val[bbeg + o0 + b_l]; }
            u32 a0 = 0, a1 = 0;
            if (warp < kc) { a0 = sAptr[k0 + warp]; a1 = sAptr[k0 + warp + 1]; }
#pragma unroll
            for (int i = 0; i < AR; ++i)
            {
                const u32 e = a0 + lane + 32 * i;
                aidx[i] = 0xffff;
                if (e < a1) { aidx[i] = A.idx[e]; aval[i] = A.val[e]; }
            }
            k0 += KC;
            return true;
        }
        return false;
    };

    int it = 0;
    for (u32 p = D.pair_begin; p < D.pair_end; ++p)
    {
        A = pairs[p].a;
        B = pairs[p].b;
        // ---- pair prologue
        __syncthreads();  // previous pair's chunks are done with sAptr / sBoff
        for (int i = tid; i <= nb; i += 512) sAptr[i] = i == 0 ? 0u : A.ptr[i];
        for (int i = tid; i < TN * (KCH + 1); i += 512) sBoff[i] = 0;
        u32 bend = 0;
        bbeg = 0;
        if (b_jj < ncol) { bbeg = col_begin(B.ptr, b_col); bend = B.ptr[b_col + 1]; }
        u16 rr[KCH];
#pragma unroll
        for (int i = 0; i < KCH; ++i)
        {
            const u32 e = bbeg + b_l + 16 * i;
            rr[i] = 0xffff;
            if (e < bend) rr[i] = B.idx[e];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < KCH; ++i)
            if (rr[i] != 0xffff) atomicAdd(&sBoff[b_jj * (KCH + 1) + (rr[i] / KC) + 1], 1);
        __syncthreads();
        if (b_l == 0)
        {
            int run = 0;
            for (int c = 1; c <= KCH; ++c) { run += sBoff[b_jj * (KCH + 1) + c]; sBoff[b_jj * (KCH + 1) + c] = run; }
        }
        __syncthreads();
        k0 = 0;
        bool more = fetch();
        while (more)
        {
            const int buf = it & 1;
            T* As = As0 + buf * KC * LDA;
            T* Bs = Bs0 + buf * KC * LDB;
            {
                T* Ak = As + warp * LDA;
#pragma unroll
                for (int i = 0; i < AR; ++i)
                    if (aidx[i] != 0xffff) Ak[aidx[i]] = aval[i];
                if (bidx >= 0) Bs[(bidx % KC) * LDB + b_jj] = bval;
            }
            __syncthreads();
            more = fetch();  // loads of the next chunk stay in flight during the tensor phase below
            {
                T* Ao = As0 + (buf ^ 1) * KC * LDA;
                T* Bo = Bs0 + (buf ^ 1) * KC * LDB;
                for (int i = tid; i < KC * LDA; i += 512) Ao[i] = zero<T>();
                for (int i = tid; i < KC * LDB; i += 512) Bo[i] = zero<T>();
            }
            if constexpr (kDouble)
            {
                const int lr = lane / 4, lk = lane % 4;
#pragma unroll
                for (int k4 = 0; k4 < KC; k4 += 4)
                {
                    double a[4], b[NT];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) a[mt] = As[(k4 + lk) * LDA + m_base + mt * 8 + lr];
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) b[nt] = Bs[(k4 + lk) * LDB + n_base + nt * 8 + lr];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = 0; nt < NT; ++nt) dmma884(acc[(mt * NT + nt) * 2], acc[(mt * NT + nt) * 2 + 1], a[mt], b[nt]);
                }
            }
            else
            {
                // FMA register tile: thread (tx, ty) owns rows tx + 32 i (i < TM/32) and columns 2 ty, 2 ty + 1
                const int tx = tid % 32, ty = tid / 32;
#pragma unroll 4
                for (int kk = 0; kk < KC; ++kk)
                {
                    const T b0 = Bs[kk * LDB + ty * 2], b1 = Bs[kk * LDB + ty * 2 + 1];
#pragma unroll
                    for (int i = 0; i < TM / 32; ++i)
                    {
                        const T av = As[kk * LDA + tx + 32 * i];
                        fma_acc(acc[i * 2], av, b0);
                        fma_acc(acc[i * 2 + 1], av, b1);
                    }
                }
            }
            __syncthreads();
            ++it;
        }
    }

    // epilogue: accumulator -> shared memory (column major [TN][LDA]) -> destination pattern
    __syncthreads();
    if constexpr (kDouble)
    {
        const int lr = lane / 4, lc = (lane % 4) * 2;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
            {
                const int r = m_base + mt * 8 + lr, c = n_base + nt * 8 + lc;
                Cs[c * LDA + r] = acc[(mt * NT + nt) * 2];
                Cs[(c + 1) * LDA + r] = acc[(mt * NT + nt) * 2 + 1];
            }
    }
    else
    {
        const int tx = tid % 32, ty = tid / 32;
#pragma unroll
        for (int i = 0; i < TM / 32; ++i)
        {
            Cs[(ty * 2) * LDA + tx + 32 * i] = acc[i * 2];
            Cs[(ty * 2 + 1) * LDA + tx + 32 * i] = acc[i * 2 + 1];
        }
    }
    __syncthreads();
    const bool shared_dest = force_atomic || (D.flags & 1u) != 0;
    for (int jj = warp; jj < ncol; jj += 16)
    {
        const int j = c0 + jj;
        const u32 e0 = col_begin(D.c.ptr, j), e1 = D.c.ptr[j + 1];
        const T* Cj = Cs + jj * LDA;
        if (shared_dest)
            for (u32 e = e0 + lane; e < e1; e += PSM_WARP) red_sub(&D.c.val[e], Cj[D.c.idx[e]]);
        else
            for (u32 e = e0 + lane; e < e1; e += PSM_WARP) vsub(D.c.val[e], Cj[D.c.idx[e]]);
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
                if (shared_dest) red_sub(&D.u.val[e], Cs[(c - c0) * LDA + r]);
                else vsub(D.u.val[e], Cs[(c - c0) * LDA + r]);
            }
        }
    }
}

template <class T, int TM>
__global__ void __launch_bounds__(512, 1)
ssssm_tile2_kernel(int nb, const SsssmDest<T>* __restrict__ dests, const SsssmPair<T>* __restrict__ pairs, int tiles_per_dest)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = blockIdx.x / tiles_per_dest;
    const int c0 = (blockIdx.x % tiles_per_dest) * PSM_TILE_N;
    const SsssmDest<T> D = dests[d];
    ssssm_tile2_cta<T, TM>(smem_raw, nb, D, pairs, c0, false);
}

// ================================================================================================
// Persistent executor ("megakernel"): the whole factorisation of one rank in ONE launch.
//
// PanguLU's scheduler is synchronisation-free on the host: per-block dependency counters, tasks fire when
// their counter reaches zero (src/pangulu_numeric.c).  This kernel moves that idea onto the device.  One
// resident CTA per SM pulls tasks from a list in topological order (atomic ticket), waits -- spinning on
// per-block flags in global memory, acquire loads -- until the blocks it reads are final and (for
// GETRF/TSTRF/GESSM) until every update of its destination has landed, runs the same CTA-level routines
// the per-wave launches use, and publishes completion with a release.  No launch gaps, no wave barriers:
// bulk SSSSM updates overlap the GETRF -> TSTRF/GESSM chain of the critical path.
// Because updates of one destination now run concurrently, they are always combined with FP64 global
// reductions.  A task only depends on tasks earlier in the list, and every fetched task is owned by a
// running CTA, so the spin waits cannot deadlock.
// ================================================================================================
struct MegaTask { int kind, a, b, c; };
// off/cnt: entry range of the value array to copy (off = 0xffffffff: the whole array, length given by the pushing task)
template <class T> struct MegaPush { T* peer_val; int* peer_flag; T* peer_dense; const T* src_dense; size_t dense_count; u32 off, cnt; };
enum { MEGA_TILE = 0, MEGA_SPARSE = 1, MEGA_GETRF_SMEM = 2, MEGA_GETRF_DENSE = 3, MEGA_GETRF_GLOBAL = 4, MEGA_PANEL = 5, MEGA_CTILE = 6, MEGA_CPANEL = 7, MEGA_DDENS = 8 };

template <class T>
struct MegaParams
{
    int nb, ntask;
    const MegaTask* tasks;
    const SsssmDest<T>* dests;
    const SsssmPair<T>* pairs;
    const int2* pair_dep;   // per pair: sync ids of A and B
    const CPair* cpairs;    // operand pairs of the compact-tile destinations (indices into ops)
    const int2* cpair_dep;  // per compact pair: sync ids of A and B
    const DenseOp<T>* ops;  // compact dense copies of operand blocks
    const CPanelTask* cptasks;  // slices of the compact panel solves
    const DiagDense<T>* dds;    // dense triangular copies of diagonal blocks
    const int* getrf_dd;        // per GETRF task: index into dds or -1
    const int4* dest_id;    // per destination descriptor: sync ids of the destination (a diagonal block: one per group the updates touch; unused: -1)
    const GetrfTask<T>* getrf;
    const int2* getrf_id;   // sync ids of the lower and the upper half
    const PanelTask<T>* panel;
    const int2* panel_id;   // x: sync id of the panel block
    const int4* panel_dep;  // sync ids of the diagonal block's groups the solve depends on (unused: -1)
    int* remaining;         // per sync id: SSSSM sub-tasks still to land on the block
    int* done;              // per sync id: 1 once the block is final
    int* pending;           // per sync id: TSTRF/GESSM sub-tasks (16 rows/columns each) still running
    int* counter;           // ticket
    T* scratch;             // dense GETRF tile per CTA
    size_t smem_bytes;      // dynamic shared memory the kernel was launched with
    // multi-GPU: blocks finalised here that other ranks read are written straight into the peers' mirror copies
    // over NVLink (peer-mapped pointers), followed by a system-scope release of the peer's done flag
    const int* push_ptr;    // per sync id: range into pushes[] (null when single-GPU)
    const MegaPush<T>* pushes;
    int multi;              // 1: flags may be written by peers -> system-scope acquire
    unsigned long long* trace;  // optional: per task {fetch, start, end} in ns (globaltimer), else null
    int4* progress;             // optional (host-mapped): per CTA {task, kind, first argument, 0 waiting / 1 running / 2 done}
};
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// Spin waits are bounded: a dependency that does not arrive within g_spin_limit_ns (a dead peer rank, CTAs that are
// not co-resident under MPS/MIG, a scheduling bug) ends the run with the reference's error convention
// (pangulu_platform_0201000.cu:31-50: print "[PanguLU ERROR]" and exit) instead of hanging every GPU: the waiting
// thread prints, raises g_abort and traps, the launch fails and the host's CK() prints and exits.
__device__ unsigned long long g_spin_limit_ns = 30ull * 1000ull * 1000ull * 1000ull;
__device__ int g_abort = 0;
__device__ __noinline__ void spin_timeout(const int* flag, int value)
{
    printf("[PanguLU ERROR] sm100 persistent executor: block %d waited more than %llu ms for a dependency (flag %p = %d); a peer rank died or the "
           "executor's CTAs are not co-resident\n", (int)blockIdx.x, g_spin_limit_ns / 1000000ull, (const void*)flag, value);
    atomicExch(&g_abort, 1);
    __threadfence_system();
    __trap();
}
// every 4096 polls (>= 1.6 ms once the back-off has reached 400 ns): compare the global timer with the first reading
#define PSM_SPIN_GUARD(flag, value)                                                   \
    if ((++polls & 4095u) == 0)                                                       \
    {                                                                                 \
        const unsigned long long now = global_ns();                                   \
        if (t_first == 0) t_first = now;                                              \
        else if (now - t_first > g_spin_limit_ns || g_abort) spin_timeout(flag, value); \
    }
__device__ __forceinline__ void spin_until_set(const int* flag, int multi = 0)
{
    unsigned ns = 20, polls = 0;
    unsigned long long t_first = 0;
    if (multi)
        while (ld_acquire_sys(flag) == 0) { __nanosleep(ns); if (ns < 400) ns *= 2; PSM_SPIN_GUARD(flag, 0) }
    else
        while (ld_acquire(flag) == 0) { __nanosleep(ns); if (ns < 400) ns *= 2; PSM_SPIN_GUARD(flag, 0) }
}
// Whole CTA: `bytes` bytes from src to the same offset of up to 8 peer destinations (NVLink-mapped pointers), 128-bit
// accesses where source and destinations are 16-byte aligned alike (block value arrays and compact copies are), every
// source vector read ONCE and stored to all peers back to back so that the links to all peers are busy at the same time.
__device__ __forceinline__ void peer_copy(char* const* dst, int nd, const char* src, size_t bytes)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    size_t head = (16 - ((size_t)src & 15)) & 15;
    bool vec = true;
    for (int d = 0; d < nd; ++d) vec = vec && (((size_t)dst[d] & 15) == ((size_t)src & 15));
    if (!vec || head > bytes) head = bytes;   // no common alignment: word copies only
    const size_t words_head = head / 4;       // (value types are multiples of 4 bytes)
    for (size_t i = tid; i < words_head; i += nthr)
    {
        const unsigned v = reinterpret_cast<const unsigned*>(src)[i];
        for (int d = 0; d < nd; ++d) reinterpret_cast<unsigned*>(dst[d])[i] = v;
    }
    const size_t nvec = (bytes - head) / 16;
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    for (size_t i = tid; i < nvec; i += nthr)
    {
        const uint4 v = s4[i];
        for (int d = 0; d < nd; ++d) reinterpret_cast<uint4*>(dst[d] + head)[i] = v;
    }
    const size_t done = head + nvec * 16;
    for (size_t i = done / 4 + tid; i < bytes / 4; i += nthr)
    {
        const unsigned v = reinterpret_cast<const unsigned*>(src)[i];
        for (int d = 0; d < nd; ++d) reinterpret_cast<unsigned*>(dst[d])[i] = v;
    }
}
// Whole CTA: copy the n finished values of a block (and its compact dense copy, where a peer reads the block through it)
// into every peer mirror registered for sync id `id`, then raise the peers' done flags.  All pushes of one id cover the
// same entry range (the whole block, or one independent group of a diagonal half).
template <class T>
__device__ __forceinline__ void push_block(const int* push_ptr, const MegaPush<T>* pushes, int id, const T* src, u32 n)
{
    if (!push_ptr) return;
    const int e0 = push_ptr[id], e1 = push_ptr[id + 1];
    if (e0 == e1) return;
    for (int eb = e0; eb < e1; eb += 8)
    {
        const int nd = min(8, e1 - eb);
        const bool whole = pushes[eb].off == 0xffffffffu;
        const u32 o = whole ? 0u : pushes[eb].off, cnt = whole ? n : pushes[eb].cnt;
        char* dst[8];
        for (int d = 0; d < nd; ++d) dst[d] = reinterpret_cast<char*>(pushes[eb + d].peer_val + o);
        peer_copy(dst, nd, reinterpret_cast<const char*>(src + o), (size_t)cnt * sizeof(T));
        int ndn = 0;
        const T* sd = nullptr;
        size_t nden = 0;
        for (int d = 0; d < nd; ++d)
            if (pushes[eb + d].peer_dense) { dst[ndn++] = reinterpret_cast<char*>(pushes[eb + d].peer_dense); sd = pushes[eb + d].src_dense; nden = pushes[eb + d].dense_count; }
        // the consumers read the block through its compact dense copy: same layout on all ranks
        if (ndn) peer_copy(dst, ndn, reinterpret_cast<const char*>(sd), nden * sizeof(T));
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        for (int e = e0; e < e1; ++e) st_release_sys(pushes[e].peer_flag, 1);
}
__device__ __forceinline__ void spin_until_zero(const int* cnt)
{
    unsigned ns = 20, polls = 0;
    unsigned long long t_first = 0;
    int v;
    while ((v = ld_acquire(cnt)) != 0) { __nanosleep(ns); if (ns < 400) ns *= 2; PSM_SPIN_GUARD(cnt, v) }
}

// cooperative global -> shared copy of `bytes` bytes (both 16-byte aligned at the start), 128-bit accesses, 4 in flight
__device__ __forceinline__ void stage_copy(void* dst, const void* src, size_t bytes, int tid, int nthreads)
{
    const size_t nvec = bytes / 16;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    size_t i = tid;
    for (; i + 3 * (size_t)nthreads < nvec; i += 4 * (size_t)nthreads)
    {
        const uint4 a = s4[i], b = s4[i + nthreads], c = s4[i + 2 * (size_t)nthreads], d = s4[i + 3 * (size_t)nthreads];
        d4[i] = a; d4[i + nthreads] = b; d4[i + 2 * (size_t)nthreads] = c; d4[i + 3 * (size_t)nthreads] = d;
    }
    for (; i < nvec; i += nthreads) d4[i] = s4[i];
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(src);
    unsigned char* db = reinterpret_cast<unsigned char*>(dst);
    for (size_t k = nvec * 16 + tid; k < bytes; k += nthreads) db[k] = sb[k];
}

// Have the dependencies of task t arrived?  One warp, non-blocking (the same flags and counters the task's own blocking
// waits look at; those waits still run afterwards and carry the ordering).
template <class T>
__device__ __forceinline__ bool task_ready(const MegaParams<T>& P, int t, int lane)
{
    const MegaTask mt = P.tasks[t];
    bool ok = true;
    auto flag = [&](int id) { return id < 0 || (P.multi ? ld_acquire_sys(P.done + id) : ld_acquire(P.done + id)) != 0; };
    switch (mt.kind)
    {
    case MEGA_TILE:
    case MEGA_SPARSE:
    case MEGA_CTILE:
    {
        const SsssmDest<T>* D = P.dests + mt.a;
        const u32 pb = D->pair_begin, pe = D->pair_end;
        const int2* deps = mt.kind == MEGA_CTILE ? P.cpair_dep : P.pair_dep;
        for (u32 p = pb + lane; p < pe && ok; p += PSM_WARP)
        {
            const int2 dep = deps[p];
            ok = flag(dep.x) && flag(dep.y);
        }
        break;
    }
    case MEGA_GETRF_SMEM:
    case MEGA_GETRF_DENSE:
    case MEGA_GETRF_GLOBAL:
        ok = ld_acquire(P.remaining + P.getrf_id[mt.a].x) == 0;
        break;
    case MEGA_DDENS:
    {
        const int2 id = P.getrf_id[mt.a];
        ok = flag(id.x) && flag(id.y);
        break;
    }
    case MEGA_CPANEL:
    case MEGA_PANEL:
    {
        const int panel = mt.kind == MEGA_CPANEL ? P.cptasks[mt.a].panel : mt.a;
        const int4 dp = P.panel_dep[panel];
        ok = flag(dp.x) && flag(dp.y) && flag(dp.z) && flag(dp.w) && ld_acquire(P.remaining + P.panel_id[panel].x) == 0;
        break;
    }
    }
    return __all_sync(PSM_FULL, ok);
}

#define PSM_DEFER_MAX 8
__device__ int g_defer_depth = 0;   // option "defer": tasks a CTA may set aside while their dependencies are outstanding

template <class T, int TM>
__global__ void __launch_bounds__(512, 1)
mega_kernel(const MegaParams<T> P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ticket;
    __shared__ int s_last;
    __shared__ int s_def[PSM_DEFER_MAX], s_ndef, s_exhausted;
    const int tid = threadIdx.x, warp = tid / PSM_WARP;
    const int nb = P.nb;
    CtState ct_state;
    CpState cp_state;
    const int defer_depth = min(PSM_DEFER_MAX, max(0, g_defer_depth));
    if (tid == 0) { s_ndef = 0; s_exhausted = 0; }
    __syncthreads();
    for (;;)
    {
        if (defer_depth == 0)
        {
            if (tid == 0) s_ticket = atomicAdd(P.counter, 1);
        }
        else if (warp == 0)
        {
            // Choose the next task.  A task whose dependencies are outstanding is SET ASIDE (at most defer_depth of them) and the
            // CTA goes on with the list, instead of holding its SM while it waits -- with several ranks a third of the SM time
            // went into such waits at the head of the list.  Progress: the CTA polls what it set aside before every new
            // ticket; with a full set-aside list it takes the OLDEST entry and waits for it in place (the globally oldest
            // unfinished task is always ready and is either running or the oldest entry of the CTA that holds it).
            // The list lives in shared memory, its LENGTH and the exhausted flag in registers of every lane (all lanes take
            // the same branches; branching on shared state that lane 0 is changing would let the lanes disagree).
            const int lane = tid;
            int ndef = s_ndef, exhausted = s_exhausted;
            __syncwarp();
            auto drop = [&](int i) {   // remove entry i (all lanes call it)
                __syncwarp();
                if (lane == 0) for (int j = i; j + 1 < ndef; ++j) s_def[j] = s_def[j + 1];
                --ndef;
                __syncwarp();
            };
            int pick = -1;
            for (int i = 0; i < ndef && pick < 0; ++i)
            {
                const int cand = s_def[i];
                if (task_ready<T>(P, cand, lane)) { pick = cand; drop(i); }
            }
            while (pick < 0)
            {
                if (!exhausted)
                {
                    int t = 0;
                    if (lane == 0) t = atomicAdd(P.counter, 1);
                    t = __shfl_sync(PSM_FULL, t, 0);
                    if (t >= P.ntask) { exhausted = 1; continue; }
                    if (P.trace && lane == 0) P.trace[3 * (size_t)t] = global_ns();
                    if (task_ready<T>(P, t, lane)) pick = t;
                    else if (ndef < defer_depth)
                    {
                        __syncwarp();
                        if (lane == 0) s_def[ndef] = t;
                        ++ndef;
                        __syncwarp();
                    }
                    else
                    {
                        // full: run the oldest entry (its waits block in place), the new task goes to the back of the list
                        pick = s_def[0];
                        drop(0);
                        if (lane == 0) s_def[ndef] = t;
                        ++ndef;
                        __syncwarp();
                    }
                }
                else if (ndef == 0) pick = P.ntask;   // nothing left anywhere
                else { pick = s_def[0]; drop(0); }
            }
            if (lane == 0) { s_ticket = pick; s_ndef = ndef; s_exhausted = exhausted; }
            if (P.progress && lane == 0)
            {
                volatile int* pd = reinterpret_cast<volatile int*>(P.progress + 512 + blockIdx.x);
                for (int j = 0; j < 4; ++j) pd[j] = j < ndef ? s_def[j] : -1;
            }
        }
        __syncthreads();
        const int t = s_ticket;
        if (t >= P.ntask) break;
        const MegaTask mt = P.tasks[t];
        if (P.trace && tid == 0 && defer_depth == 0) P.trace[3 * (size_t)t] = global_ns();
        if (P.progress && tid == 0)
        {
            volatile int* pg = reinterpret_cast<volatile int*>(P.progress + blockIdx.x);
            pg[0] = t; pg[1] = mt.kind; pg[2] = mt.a; pg[3] = 0;
        }
        switch (mt.kind)
        {
        case MEGA_TILE:
        case MEGA_SPARSE:
        case MEGA_CTILE:
        {
            const SsssmDest<T> D = P.dests[mt.a];
            if (warp == 0)
            {
                // operand blocks final?  One pair per lane: the acquire loads of a destination's pairs overlap instead of
                // forming one chain (the block barrier below carries their ordering to the other warps)
                const int2* deps = mt.kind == MEGA_CTILE ? P.cpair_dep : P.pair_dep;
                for (u32 p = D.pair_begin + (tid % PSM_WARP); p < D.pair_end; p += PSM_WARP)
                {
                    const int2 dep = deps[p];
                    spin_until_set(P.done + dep.x, P.multi);
                    spin_until_set(P.done + dep.y, P.multi);
                }
            }
            __syncthreads();
            if (P.trace && tid == 0) P.trace[3 * (size_t)t + 1] = global_ns();
            if (P.progress && tid == 0) ((volatile int*)(P.progress + blockIdx.x))[3] = 1;
            if (mt.kind == MEGA_CTILE)
                ssssm_ctile_cta<T, ct_traits<T>::TN, ct_traits<T>::STAGES>(smem_raw, nb, D, P.cpairs, P.ops, mt.b & 0xffff, mt.b >> 16, ct_state);
            else if (mt.kind == MEGA_TILE)
                ssssm_tile2_cta<T, TM>(smem_raw, nb, D, P.pairs, mt.b * PSM_TILE_N, true);
            else
            {
                T* acc = reinterpret_cast<T*>(smem_raw) + (size_t)warp * nb;
                for (int j = mt.b + warp; j < mt.b + mt.c && j < nb; j += 16) ssssm_sparse_column<T>(nb, D, P.pairs, j, acc, true);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0)
            {
                const int4 di = P.dest_id[mt.a];
                atomicSub(P.remaining + di.x, 1);
                if (di.y >= 0) atomicSub(P.remaining + di.y, 1);
                if (di.z >= 0) atomicSub(P.remaining + di.z, 1);
                if (di.w >= 0) atomicSub(P.remaining + di.w, 1);
            }
            break;
        }
        case MEGA_GETRF_SMEM:
        case MEGA_GETRF_DENSE:
        case MEGA_GETRF_GLOBAL:
        {
            const GetrfTask<T> G = P.getrf[mt.a];
            const int2 id = P.getrf_id[mt.a];
            if (tid == 0) spin_until_zero(P.remaining + id.x);
            __syncthreads();
            if (P.trace && tid == 0) P.trace[3 * (size_t)t + 1] = global_ns();
            if (P.progress && tid == 0) ((volatile int*)(P.progress + blockIdx.x))[3] = 1;
            if (mt.kind == MEGA_GETRF_SMEM)
                getrf_smem_cta<T, 512>(smem_raw, nb, G);
            else if (mt.kind == MEGA_GETRF_DENSE)
            {
                const int nbp = (nb + 31) & ~31;
                getrf_dense_cta<T, 512>(smem_raw, nb, G, P.scratch + (size_t)blockIdx.x * nbp * nbp);
            }
            else
            {
                T* buf = reinterpret_cast<T*>(smem_raw);
                for (int i = tid; i < 4 * nb; i += 512) buf[i] = zero<T>();
                __syncthreads();
                getrf_body<T, 512>(nb, G.lower.val, G.lower.ptr, G.lower.idx, G.upper.val, G.upper.ptr, G.upper.idx, buf, G.c0, G.c1);
            }
            if (P.getrf_dd && P.getrf_dd[mt.a] >= 0)
            {
                __syncthreads();  // (the factors were written by other threads of this CTA)
                diag_densify<T>(nb, G, P.dds[P.getrf_dd[mt.a]]);
            }
            asm volatile("fence.proxy.async;" ::: "memory");   // the dense triangular copies are read with bulk copies (async proxy)
            __threadfence();
            __syncthreads();
            if (tid == 0) { st_release(P.done + id.x, 1); st_release(P.done + id.y, 1); }
            // (two halves with one sync id = an empty lower half sharing its value pointer with the upper one: a single
            // push, after which the flag may be raised)
            if (id.x != id.y) push_block<T>(P.push_ptr, P.pushes, id.x, G.lower.val, G.lower.ptr[nb]);
            push_block<T>(P.push_ptr, P.pushes, id.y, G.upper.val, G.upper.ptr[nb]);
            break;
        }
        case MEGA_DDENS:
        {
            // refresh the dense triangular copies of a diagonal block factored by ANOTHER rank, once its mirrored halves
            // (one independent group of them) have arrived; mt.b = sync id the compact panel solves wait for
            const GetrfTask<T> G = P.getrf[mt.a];
            const int2 id = P.getrf_id[mt.a];
            if (tid == 0)
            {
                if (id.x >= 0) spin_until_set(P.done + id.x, P.multi);
                if (id.y >= 0) spin_until_set(P.done + id.y, P.multi);
            }
            __syncthreads();
            if (P.trace && tid == 0) P.trace[3 * (size_t)t + 1] = global_ns();
            if (P.progress && tid == 0) ((volatile int*)(P.progress + blockIdx.x))[3] = 1;
            diag_densify<T>(nb, G, P.dds[P.getrf_dd[mt.a]]);
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release(P.done + mt.b, 1);
            break;
        }
        case MEGA_CPANEL:
        {
            const CPanelTask cpt = P.cptasks[mt.a];
            const PanelTask<T> Q = P.panel[cpt.panel];
            const int2 id = P.panel_id[cpt.panel];
            if (tid == 0)
            {
                const int4 dp = P.panel_dep[cpt.panel];
                spin_until_set(P.done + dp.x, P.multi);
                if (dp.y >= 0) spin_until_set(P.done + dp.y, P.multi);
                if (dp.z >= 0) spin_until_set(P.done + dp.z, P.multi);
                if (dp.w >= 0) spin_until_set(P.done + dp.w, P.multi);
                spin_until_zero(P.remaining + id.x);
            }
            __syncthreads();
            if (P.trace && tid == 0) P.trace[3 * (size_t)t + 1] = global_ns();
            if (P.progress && tid == 0) ((volatile int*)(P.progress + blockIdx.x))[3] = 1;
            panel_compact_cta<T>(smem_raw, nb, Q, P.ops[Q.dense], P.dds[cpt.dd], cpt.r0, cp_state);
            asm volatile("fence.proxy.async;" ::: "memory");   // the compact copy just written is read with bulk copies
            __threadfence();
            __syncthreads();
            if (tid == 0)
            {
                s_last = atomicSub(P.pending + id.x, 1) == 1;
                if (s_last) { __threadfence(); st_release(P.done + id.x, 1); }
            }
            __syncthreads();
            if (s_last) push_block<T>(P.push_ptr, P.pushes, id.x, Q.b.val, Q.b.ptr[nb]);
            break;
        }
        case MEGA_PANEL:
        {
            const PanelTask<T> Q = P.panel[mt.a];
            const int2 id = P.panel_id[mt.a];
            if (tid == 0)
            {
                const int4 dp = P.panel_dep[mt.a];
                spin_until_set(P.done + dp.x, P.multi);
                if (dp.y >= 0) spin_until_set(P.done + dp.y, P.multi);
                if (dp.z >= 0) spin_until_set(P.done + dp.z, P.multi);
                if (dp.w >= 0) spin_until_set(P.done + dp.w, P.multi);
                spin_until_zero(P.remaining + id.x);
            }
            __syncthreads();
            if (P.trace && tid == 0) P.trace[3 * (size_t)t + 1] = global_ns();
            if (P.progress && tid == 0) ((volatile int*)(P.progress + blockIdx.x))[3] = 1;
            // one task = 64 consecutive rows/columns of the block (four per warp); the diagonal half they are solved
            // against is staged in shared memory when it fits.  The last sub-task of a block to finish publishes it.
            T* x = reinterpret_cast<T*>(smem_raw) + (size_t)warp * nb;
            PanelTask<T> Q2 = Q;
            {
                unsigned char* stage = smem_raw + (size_t)16 * nb * sizeof(T);
                const size_t vec_bytes = (size_t)16 * nb * sizeof(T);
                const size_t budget = P.smem_bytes > vec_bytes ? P.smem_bytes - vec_bytes : 0;   // (launch_mega refuses nb that does not fit)
                const u32 nd = Q.diag.ptr[nb];
                const size_t need = (size_t)nd * (sizeof(T) + 2) + 4 * (size_t)(nb + 1) + 64;
                if (need <= budget)
                {
                    T* sval = reinterpret_cast<T*>(stage);
                    u32* sptr = reinterpret_cast<u32*>(stage + (((size_t)nd * sizeof(T) + 15) & ~(size_t)15));
                    u16* sidx = reinterpret_cast<u16*>(reinterpret_cast<unsigned char*>(sptr) + ((4 * (size_t)(nb + 1) + 15) & ~(size_t)15));
                    if ((reinterpret_cast<size_t>(Q.diag.val) & 15) == 0 && (reinterpret_cast<size_t>(Q.diag.idx) & 15) == 0)
                    {
                        stage_copy(sval, Q.diag.val, (size_t)nd * sizeof(T), tid, 512);
                        stage_copy(sidx, Q.diag.idx, (size_t)nd * 2, tid, 512);
                    }
                    else
                        for (u32 i = tid; i < nd; i += 512) { sval[i] = Q.diag.val[i]; sidx[i] = Q.diag.idx[i]; }
                    for (int i = tid; i <= nb; i += 512) sptr[i] = i == 0 ? 0u : Q.diag.ptr[i];
                    Q2.diag.val = sval; Q2.diag.ptr = sptr; Q2.diag.idx = sidx;
                }
                __syncthreads();
            }
            const DenseOp<T>* dop = (P.ops && Q.dense >= 0) ? P.ops + Q.dense : nullptr;
            for (int rc = mt.b + warp; rc < mt.b + 64 && rc < nb; rc += 16) panel_solve_line<T>(nb, Q2, rc, x, dop);
            asm volatile("fence.proxy.async;" ::: "memory");
            __threadfence();
            __syncthreads();
            if (tid == 0)
            {
                s_last = atomicSub(P.pending + id.x, 1) == 1;
                if (s_last) { __threadfence(); st_release(P.done + id.x, 1); }
            }
            __syncthreads();
            if (s_last) push_block<T>(P.push_ptr, P.pushes, id.x, Q.b.val, Q.b.ptr[nb]);
            break;
        }
        }
        __syncthreads();
        if (P.trace && tid == 0) P.trace[3 * (size_t)t + 2] = global_ns();
        if (P.progress && tid == 0) ((volatile int*)(P.progress + blockIdx.x))[3] = 2;
    }
}

}  // namespace psm
