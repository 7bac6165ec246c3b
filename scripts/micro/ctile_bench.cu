// Micro-benchmark: the compact tile routine (ssssm_ctile_kernel<double>, sm100_compact.cuh) on synthetic destinations of
// controlled shape: m' non-empty rows of the left operands, nt non-empty columns of the right operands inside the 64-column
// tile, k-chunks per pair, length of the runs of pairs that share their maps, size of the operand pool (L2-resident or not).
// One CTA per SM, P pairs per CTA.  Reports the share of the FP64 tensor peak (16.1 cycles per DMMA and sub-partition).
// Development probe only.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "sm100_kernels.cuh"
using namespace psm;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
typedef double T;
static const int nb = 256;

template <class V> V* up(const std::vector<V>& h)
{
    V* d; CK(cudaMalloc(&d, h.size() * sizeof(V) + 256));
    CK(cudaMemcpy(d, h.data(), h.size() * sizeof(V), cudaMemcpyHostToDevice));
    return d;
}

struct Case { const char* name; int m, nt, chunks, runlen, pool, pairs; };

int run(const Case& c, int sms)
{
    // row lists: variant v uses rows [v*shift, v*shift + m) ; column lists: columns [0, nt) (variant-independent unless runlen == 1 and alt cols)
    const int nvar = c.runlen == 0 ? 1 : 2;   // runlen 0: all pairs share one map (a single run)
    std::vector<u16> meta;
    std::vector<size_t> rows_off(nvar), cols_off(nvar);
    for (int v = 0; v < nvar; ++v)
    {
        rows_off[v] = meta.size();
        const int r0 = v * (nb - c.m);
        for (int r = 0; r < c.m; ++r) meta.push_back((u16)(r0 + r));
        std::vector<u16> inv(nb, 0xffff);
        for (int r = 0; r < c.m; ++r) inv[r0 + r] = (u16)r;
        meta.insert(meta.end(), inv.begin(), inv.end());
        while (meta.size() % 8) meta.push_back(0);
    }
    {
        cols_off[0] = meta.size();
        for (int q = 0; q < c.nt; ++q) meta.push_back((u16)q);
        std::vector<u16> inv(nb, 0xffff);
        for (int q = 0; q < c.nt; ++q) inv[q] = (u16)q;
        meta.insert(meta.end(), inv.begin(), inv.end());
        for (int q = 0; q <= nb; ++q) meta.push_back((u16)std::min(q, c.nt));
        while (meta.size() % 8) meta.push_back(0);
    }
    u16* d_meta = up(meta);
    const int ld = ((c.m + 7) & ~7) + 4;
    const size_t a_len = (size_t)nb * ld, b_len = (size_t)(nb / 16) * c.nt * PSM_CT_BLD;
    const size_t a_stride = (a_len + 31) & ~(size_t)31, b_stride = (b_len + 31) & ~(size_t)31;
    T* d_dense; CK(cudaMalloc(&d_dense, (size_t)c.pool * (a_stride + b_stride) * sizeof(T)));
    {
        std::vector<T> h((size_t)c.pool * (a_stride + b_stride));
        for (size_t i = 0; i < h.size(); ++i) h[i] = 1e-3 * (double)((i * 2654435761u) % 1000);
        CK(cudaMemcpy(d_dense, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
    const u32 kmask = c.chunks >= 16 ? 0xffffu : ((1u << c.chunks) - 1u);
    std::vector<DenseOp<T>> ops;
    for (int p = 0; p < c.pool; ++p)
        for (int v = 0; v < nvar; ++v)
        {
            DenseOp<T> a; memset(&a, 0, sizeof a);
            a.dense = d_dense + (size_t)p * (a_stride + b_stride);
            a.list = d_meta + rows_off[v]; a.inv = a.list + c.m; a.cnt = c.m; a.ld = ld; a.kmask = kmask; a.pre = nullptr;
            ops.push_back(a);
        }
    const int b_first = (int)ops.size();
    for (int p = 0; p < c.pool; ++p)
    {
        DenseOp<T> b; memset(&b, 0, sizeof b);
        b.dense = d_dense + (size_t)p * (a_stride + b_stride) + a_stride;
        b.list = d_meta + cols_off[0]; b.inv = b.list + c.nt; b.cnt = c.nt; b.ld = PSM_CT_BLD; b.kmask = kmask; b.pre = b.inv + nb;
        ops.push_back(b);
    }
    DenseOp<T>* d_ops = up(ops);
    // destination: one dense 256 x 256 block per CTA would be 512 KB each; all CTAs share 4 destinations (reductions at the end only)
    std::vector<u32> ptr(nb + 1); std::vector<u16> idx((size_t)nb * nb);
    for (int j = 0; j <= nb; ++j) ptr[j] = (u32)(j * nb);
    for (int j = 0; j < nb; ++j) for (int r = 0; r < nb; ++r) idx[(size_t)j * nb + r] = (u16)r;
    u32* d_ptr = up(ptr); u16* d_idx = up(idx);
    T* d_val; CK(cudaMalloc(&d_val, (size_t)4 * nb * nb * sizeof(T))); CK(cudaMemset(d_val, 0, (size_t)4 * nb * nb * sizeof(T)));
    std::vector<SsssmDest<T>> dests(sms);
    std::vector<CPair> cpairs;
    std::vector<CTileTask> tasks(sms);
    unsigned rng = 12345;
    for (int s = 0; s < sms; ++s)
    {
        SsssmDest<T>& D = dests[s]; memset(&D, 0, sizeof D);
        D.c.val = d_val + (size_t)(s % 4) * nb * nb; D.c.ptr = d_ptr; D.c.idx = d_idx;
        D.pair_begin = (u32)cpairs.size();
        for (int p = 0; p < c.pairs; ++p)
        {
            rng = rng * 1664525u + 1013904223u;
            const int pa = (rng >> 8) % c.pool;
            rng = rng * 1664525u + 1013904223u;
            const int pb2 = (rng >> 8) % c.pool;
            const int v = c.runlen == 0 ? 0 : (p / c.runlen) % 2;
            cpairs.push_back({pa * nvar + v, b_first + pb2});
        }
        D.pair_end = (u32)cpairs.size();
        D.nnz_c = nb * nb;
        tasks[s] = {s, 0, 64};
    }
    SsssmDest<T>* d_dests = up(dests); CPair* d_cp = up(cpairs); CTileTask* d_tasks = up(tasks);
    constexpr int TN = ct_traits<T>::TN;
    const size_t smem = CtLayout<T, TN, ct_traits<T>::STAGES>::total;
    CK(cudaFuncSetAttribute(ssssm_ctile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep)
    {
        CK(cudaEventRecord(e0));
        ssssm_ctile_kernel<T><<<sms, 512, smem>>>(nb, d_dests, d_cp, d_ops, d_tasks);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    const double dmma = (double)sms * c.pairs * c.chunks * 4.0 * ((c.m + 7) / 8) * ((c.nt + 7) / 8);
    const double floor_ms = dmma / (sms * 4.0) * 16.1 / 1.965e9 * 1e3;
    const double bytes = (double)sms * c.pairs * c.chunks * (16.0 * ld + 20.0 * c.nt) * 8;
    printf("%-34s m' %3d nt %2d chunks %2d runlen %3d pool %4d (%6.1f MB): %8.3f ms  floor %8.3f ms  eff %5.3f  operand %6.1f GB/s\n", c.name, c.m, c.nt, c.chunks, c.runlen,
           c.pool, (double)c.pool * (a_stride + b_stride) * 8 / 1e6, best, floor_ms, floor_ms / best, bytes / (best * 1e-3) / 1e9);
    cudaFree(d_meta); cudaFree(d_dense); cudaFree(d_ops); cudaFree(d_ptr); cudaFree(d_idx); cudaFree(d_val); cudaFree(d_dests); cudaFree(d_cp); cudaFree(d_tasks);
    return 0;
}

int main()
{
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const Case cases[] = {
        {"dense, one run, L2-resident", 256, 64, 16, 0, 8, 128},
        {"dense, one run, big pool", 256, 64, 16, 0, 600, 128},
        {"dense, runs of 8", 256, 64, 16, 8, 600, 128},
        {"dense, runs of 1", 256, 64, 16, 1, 600, 128},
        {"m' 224 (28 tiles), one run", 224, 64, 16, 0, 600, 128},
        {"m' 216 nt 56", 216, 56, 16, 0, 600, 128},
        {"m' 128 nt 64", 128, 64, 16, 0, 600, 128},
        {"m' 128 nt 64 runs of 1", 128, 64, 16, 1, 600, 128},
        {"m' 64 nt 64", 64, 64, 16, 0, 600, 256},
        {"m' 64 nt 64 runs of 1", 64, 64, 16, 1, 600, 256},
        {"m' 256 nt 32", 256, 32, 16, 0, 600, 128},
        {"m' 256 nt 16", 256, 16, 16, 0, 600, 128},
        {"m' 100 nt 24, 8 chunks, runs of 1", 100, 24, 8, 1, 600, 256},
        {"m' 48 nt 16, 6 chunks, runs of 1", 48, 16, 6, 1, 600, 512},
        {"m' 24 nt 24, 4 chunks, runs of 1", 24, 24, 4, 1, 600, 512},
        {"m' 8 nt 8, 2 chunks, runs of 1", 8, 8, 2, 1, 600, 1024},
    };
    for (const Case& c : cases) run(c, sms);
    return 0;
}
