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

// ------------------------------------------------------------------------------------------------
// SSSSM, sparse path.  grid = (#destinations x column groups); one warp owns one column j of the
// destination and applies EVERY operand pair queued for that destination (ascending elimination
// level) into a dense per-warp accumulator in shared memory, then subtracts it once from C(:,j).
// No atomics: inside one pair/k step the lanes touch distinct rows of A(:,k).
//   smem = warps * nb * sizeof(T)
// Algorithmic HBM bytes per pair: nnzA(sv+si) + nnzB(sv+si) + nnzC(2sv+si) + 3P; applying
// all pairs of a destination in one pass reads/writes C once per batch instead of once per pair.
// ------------------------------------------------------------------------------------------------
template <class T, int WARPS>
__global__ void __launch_bounds__(WARPS * PSM_WARP)
ssssm_sparse_kernel(int nb, const SsssmDest<T>* __restrict__ dests, const SsssmPair<T>* __restrict__ pairs, int groups_per_dest)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* acc_all = reinterpret_cast<T*>(smem_raw);
    const int warp = threadIdx.x / PSM_WARP, lane = threadIdx.x % PSM_WARP;
    const int d = blockIdx.x / groups_per_dest;
    const int j = (blockIdx.x % groups_per_dest) * WARPS + warp;
    if (j >= nb) return;
    T* acc = acc_all + (size_t)warp * nb;
    const SsssmDest<T> D = dests[d];

    // nothing to do for an empty destination column (a padded edge column, or a column of C that
    // holds no entry in either half)
    const u32 c0 = col_begin(D.c.ptr, j), c1 = D.c.ptr[j + 1];
    const bool diag = D.u.val != nullptr;
    if (!diag && c0 == c1) return;

    for (int i = lane; i < nb; i += PSM_WARP) acc[i] = zero<T>();
    __syncwarp();

    for (u32 p = D.pair_begin; p < D.pair_end; ++p)
    {
        const BlkRO<T> A = pairs[p].a, B = pairs[p].b;
        const u32 b0 = col_begin(B.ptr, j), b1 = B.ptr[j + 1];
        for (u32 e = b0; e < b1; ++e)
        {
            const int k = B.idx[e];
            const T bv = B.val[e];
            const u32 a0 = col_begin(A.ptr, k), a1 = A.ptr[k + 1];
            for (u32 t = a0 + lane; t < a1; t += PSM_WARP) fma_acc(acc[A.idx[t]], A.val[t], bv);
            __syncwarp();
        }
    }

    // C(:,j) -= acc  (for a diagonal destination D.c is the strict-lower half: rows > j)
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
            if (lo < D.u.ptr[r + 1] && D.u.idx[lo] == j) vsub(D.u.val[lo], v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// GESSM / TSTRF (sparse triangular panel solves).  grid = (#tasks x row/column groups); one warp
// owns one column (GESSM) or one row (TSTRF) of B: gather it into a dense shared-memory vector,
// eliminate entry by entry in index order (lanes sweep the matching column of L / row of U), scatter
// back.  kind = PANGULU_TASK_GESSM or PANGULU_TASK_TSTRF.
//   smem = warps * nb * sizeof(T)
// ------------------------------------------------------------------------------------------------
template <class T, int WARPS>
__global__ void __launch_bounds__(WARPS * PSM_WARP)
panel_solve_kernel(int nb, const PanelTask<T>* __restrict__ tasks, int groups_per_task)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* x_all = reinterpret_cast<T*>(smem_raw);
    const int warp = threadIdx.x / PSM_WARP, lane = threadIdx.x % PSM_WARP;
    const int t = blockIdx.x / groups_per_task;
    const int rc = (blockIdx.x % groups_per_task) * WARPS + warp;
    if (rc >= nb) return;
    T* x = x_all + (size_t)warp * nb;
    const PanelTask<T> P = tasks[t];
    const Blk<T> B = P.b, Dg = P.diag;

    if (P.kind == PANGULU_TASK_GESSM)
    {
        // column rc of B (CSC); L = strict-lower CSC of the diagonal block, unit diagonal
        const u32 b0 = col_begin(B.ptr, rc), b1 = B.ptr[rc + 1];
        if (b0 == b1) return;
        for (int i = lane; i < nb; i += PSM_WARP) x[i] = zero<T>();
        __syncwarp();
        for (u32 e = b0 + lane; e < b1; e += PSM_WARP) x[B.idx[e]] = B.val[e];
        __syncwarp();
        for (u32 e = b0; e < b1; ++e)
        {
            const int r = B.idx[e];
            const T v = x[r];
            const u32 l0 = col_begin(Dg.ptr, r), l1 = Dg.ptr[r + 1];
            for (u32 q = l0 + lane; q < l1; q += PSM_WARP) fnma(x[Dg.idx[q]], v, Dg.val[q]);
            __syncwarp();
        }
        for (u32 e = b0 + lane; e < b1; e += PSM_WARP) B.val[e] = x[B.idx[e]];
    }
    else
    {
        // row rc of B through its CSR mirror; U = upper CSR of the diagonal block, diagonal first
        const u32 b0 = col_begin(B.rptr, rc), b1 = B.rptr[rc + 1];
        if (b0 == b1) return;
        for (int i = lane; i < nb; i += PSM_WARP) x[i] = zero<T>();
        __syncwarp();
        for (u32 e = b0 + lane; e < b1; e += PSM_WARP) x[B.ridx[e]] = B.val[B.rmap[e]];
        __syncwarp();
        for (u32 e = b0; e < b1; ++e)
        {
            const int c = B.ridx[e];
            const u32 u0 = col_begin(Dg.ptr, c), u1 = Dg.ptr[c + 1];
            const T piv = clamp_pivot(Dg.val[u0]);
            const T v = vdiv(x[c], piv);
            __syncwarp();
            if (lane == 0) x[c] = v;
            for (u32 q = u0 + 1 + lane; q < u1; q += PSM_WARP) fnma(x[Dg.idx[q]], v, Dg.val[q]);
            __syncwarp();
        }
        for (u32 e = b0 + lane; e < b1; e += PSM_WARP) B.val[B.rmap[e]] = x[B.ridx[e]];
    }
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
template <class T, int THREADS>
__global__ void __launch_bounds__(THREADS)
getrf_sparse_kernel(int nb, const GetrfTask<T>* __restrict__ tasks)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* buf = reinterpret_cast<T*>(smem_raw);  // [2][2][nb] : parity x {lcol, urow}
    const GetrfTask<T> G = tasks[blockIdx.x];
    const Blk<T> Lo = G.lower, Up = G.upper;
    const int tid = threadIdx.x, warp = tid / PSM_WARP, lane = tid % PSM_WARP;
    constexpr int NW = THREADS / PSM_WARP;

    for (int i = tid; i < 4 * nb; i += THREADS) buf[i] = zero<T>();
    __syncthreads();

    int prev_level = -1, step = 0;
    for (int level = 0; level < nb; ++level)
    {
        const u32 u0 = col_begin(Up.ptr, level), u1 = Up.ptr[level + 1];
        if (u0 == u1) continue;  // padded row beyond n (pangulu_platform_0100000.c:74-77)
        const u32 l0 = col_begin(Lo.ptr, level), l1 = Lo.ptr[level + 1];
        const int par = (step++) & 1;  // parity of the processed-level count: skipped rows do not flip it
        T* lcol = buf + (size_t)(par * 2) * nb;
        T* urow = lcol + nb;
        const T piv = clamp_pivot(Up.val[u0]);
        for (u32 e = l0 + tid; e < l1; e += THREADS)
        {
            const T v = vdiv(Lo.val[e], piv);
            Lo.val[e] = v;
            lcol[Lo.idx[e]] = v;
        }
        for (u32 e = u0 + 1 + tid; e < u1; e += THREADS) urow[Up.idx[e]] = Up.val[e];
        // clear the OTHER parity's vectors (used by prev_level; all its readers passed the barrier
        // that ended that level)
        if (prev_level >= 0)
        {
            T* plcol = buf + (size_t)((par ^ 1) * 2) * nb;
            T* purow = plcol + nb;
            const u32 pl0 = col_begin(Lo.ptr, prev_level), pl1 = Lo.ptr[prev_level + 1];
            const u32 pu0 = col_begin(Up.ptr, prev_level), pu1 = Up.ptr[prev_level + 1];
            for (u32 e = pl0 + tid; e < pl1; e += THREADS) plcol[Lo.idx[e]] = zero<T>();
            for (u32 e = pu0 + 1 + tid; e < pu1; e += THREADS) purow[Up.idx[e]] = zero<T>();
        }
        __syncthreads();

        const int nrow = (int)(l1 - l0), ncol = (int)(u1 - u0 - 1);
        for (int it = warp; it < nrow + ncol; it += NW)
        {
            if (it < nrow)
            {
                const int r = Lo.idx[l0 + it];
                const T lr = lcol[r];
                const u32 r0 = col_begin(Up.ptr, r), r1 = Up.ptr[r + 1];
                for (u32 q = r0 + lane; q < r1; q += PSM_WARP)
                {
                    T v = Up.val[q];
                    fnma(v, lr, urow[Up.idx[q]]);
                    Up.val[q] = v;
                }
            }
            else
            {
                const u32 eu = u0 + 1 + (it - nrow);
                const int c = Up.idx[eu];
                const T uc = urow[c];
                const u32 q0 = col_begin(Lo.ptr, c), q1 = Lo.ptr[c + 1];
                for (u32 q = q0 + lane; q < q1; q += PSM_WARP)
                {
                    T v = Lo.val[q];
                    fnma(v, lcol[Lo.idx[q]], uc);
                    Lo.val[q] = v;
                }
            }
        }
        __syncthreads();
        prev_level = level;
    }
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

__device__ inline double vmul(double a, double b) { return a * b; }
__device__ inline float vmul(float a, float b) { return a * b; }
__device__ inline pangulu_sm100_c64 vmul(pangulu_sm100_c64 a, pangulu_sm100_c64 b)
{ pangulu_sm100_c64 r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }
__device__ inline pangulu_sm100_c32 vmul(pangulu_sm100_c32 a, pangulu_sm100_c32 b)
{ pangulu_sm100_c32 r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }

// y -= A x, A in CSC: one warp per column
template <class T>
__global__ void spmv_kernel(int nb, Blk<T> A, const T* __restrict__ x, T* __restrict__ y)
{
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) / PSM_WARP, lane = threadIdx.x % PSM_WARP;
    if (col >= nb) return;
    const T xv = x[col];
    for (u32 e = col_begin(A.ptr, col) + lane; e < A.ptr[col + 1]; e += PSM_WARP) atomic_sub(&y[A.idx[e]], vmul(A.val[e], xv));
}

__device__ inline void vadd(double& x, double a) { x += a; }
__device__ inline void vadd(float& x, float a) { x += a; }
__device__ inline void vadd(pangulu_sm100_c64& x, pangulu_sm100_c64 a) { x.re += a.re; x.im += a.im; }
__device__ inline void vadd(pangulu_sm100_c32& x, pangulu_sm100_c32 a) { x.re += a.re; x.im += a.im; }

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
__global__ void __launch_bounds__(THREADS) sptrsv_kernel(int nb, Blk<T> S, T* __restrict__ xg, int uplo)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* x = reinterpret_cast<T*>(smem_raw);
    __shared__ T part[THREADS / PSM_WARP];
    for (int i = threadIdx.x; i < nb; i += THREADS) x[i] = xg[i];
    __syncthreads();
    if (uplo == 0)
    {
        for (int i = 0; i < nb; ++i)
        {
            const T xi = x[i];
            __syncthreads();
            for (u32 e = col_begin(S.ptr, i) + threadIdx.x; e < S.ptr[i + 1]; e += THREADS) fnma(x[S.idx[e]], S.val[e], xi);
            __syncthreads();
        }
    }
    else
    {
        for (int i = nb - 1; i >= 0; --i)
        {
            const u32 r0 = col_begin(S.ptr, i), r1 = S.ptr[i + 1];
            if (r0 == r1) continue;
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

}  // namespace psm
