// sm100_preprocess.cuh -- preprocessing on the GPU: the two per-entry passes PanguLU runs on the host between
// the symbolic factorisation and the numeric one.
//
//   1. value scatter into the filled structure (src/pangulu_conversion.c:241-350: the reordered, scaled matrix entries are
//      written into the block value arrays, fill-in stays zero).  Here: the value region of the rank's block buffer is
//      cleared on the device and one thread per matrix entry stores  scale[i] * a[i]  at the entry's position, which the
//      host computed ONCE (the pattern does not change between factorisations of a Newton / time-stepping loop), so a
//      re-factorisation uploads nnz(A) values instead of nnz(L+U).
//   2. CSR mirror of the strictly-lower blocks (pangulu_transpose_struct_with_valueidx_inblock, src/pangulu_utils.c:479-511:
//      rowpointer / columnindex / idx_of_csc_value_for_csr from columnpointer / rowindex).  Integer work, bit-exact: one
//      CTA per block builds an nb x nb occupancy bitmap in shared memory (row r, bit c); the row pointers are the prefix
//      sum of the rows' popcounts, and the CSR position of entry (r, c) is  rowpointer[r] + popcount(bits of row r below c)
//      -- the stable column order of the host's counting sort without any serial pass.
#pragma once
#include "sm100_types.cuh"

namespace psm {

#define PSM_NO_DST 0xffffffffffffffffull

template <class T>
__global__ void __launch_bounds__(256)
scatter_values_kernel(T* __restrict__ region, const unsigned long long* __restrict__ dst, const T* __restrict__ src, const T* __restrict__ scale,
                      unsigned long long n)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        const unsigned long long d = dst[i];
        if (d == PSM_NO_DST) continue;   // entry of a block another rank owns
        region[d] = scale ? vmul(src[i], scale[i]) : src[i];
    }
}

struct MirrorTask
{
    const u32* colptr;   // nb + 1 (entry 0 is read as 0)
    const u16* rowidx;
    u32* rowptr;         // out: nb + 1
    u16* colidx;         // out: nnz
    u32* csr2csc;        // out: nnz, CSR position -> index into the value array
};

// shared memory: nb * words bitmap words + (nb + 1) row offsets, words = ceil(nb / 32)
__global__ void __launch_bounds__(256)
csr_mirror_kernel(int nb, const MirrorTask* __restrict__ tasks)
{
    extern __shared__ __align__(16) unsigned char pre_smem[];
    const MirrorTask t = tasks[blockIdx.x];
    const int words = (nb + 31) / 32;
    u32* bits = reinterpret_cast<u32*>(pre_smem);         // [nb][words]
    u32* roff = bits + (size_t)nb * words;                 // [nb + 1]
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, nwarp = blockDim.x / 32;
    for (int i = tid; i < nb * words; i += blockDim.x) bits[i] = 0u;
    __syncthreads();
    for (int c = warp; c < nb; c += nwarp)
    {
        const u32 e0 = c == 0 ? 0u : t.colptr[c], e1 = t.colptr[c + 1];
        for (u32 e = e0 + lane; e < e1; e += 32) atomicOr(&bits[(size_t)t.rowidx[e] * words + (c >> 5)], 1u << (c & 31));
    }
    __syncthreads();
    // row counts, then an exclusive prefix sum by one warp (nb <= 2048: at most 64 rounds)
    for (int r = tid; r < nb; r += blockDim.x)
    {
        u32 cnt = 0;
        for (int w = 0; w < words; ++w) cnt += __popc(bits[(size_t)r * words + w]);
        roff[r + 1] = cnt;
    }
    if (tid == 0) roff[0] = 0;
    __syncthreads();
    if (warp == 0)
    {
        u32 run = 0;
        for (int base = 0; base < nb; base += 32)
        {
            const int r = base + lane;
            u32 v = r < nb ? roff[r + 1] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const u32 up = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += up;
            }
            if (r < nb) roff[r + 1] = run + v;
            run += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    for (int r = tid; r <= nb; r += blockDim.x) t.rowptr[r] = roff[r];
    for (int c = warp; c < nb; c += nwarp)
    {
        const u32 e0 = c == 0 ? 0u : t.colptr[c], e1 = t.colptr[c + 1];
        for (u32 e = e0 + lane; e < e1; e += 32)
        {
            const int r = t.rowidx[e];
            const u32* row = bits + (size_t)r * words;
            u32 below = 0;
            for (int w = 0; w < (c >> 5); ++w) below += __popc(row[w]);
            below += __popc(row[c >> 5] & ((1u << (c & 31)) - 1u));
            const u32 pos = roff[r] + below;
            t.colidx[pos] = (u16)c;
            t.csr2csc[pos] = e;
        }
    }
}

}  // namespace psm
