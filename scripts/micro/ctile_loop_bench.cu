// Micro-benchmark: ceiling of the compact tile routine's tensor loop -- DMMA.8x8x4 fed by shared-memory fragment loads, the
// exact inner code of sm100_compact.cuh (ct_chunk<TM,TNC>), no TMA, no barriers.  Development probe only.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DCALCULATE_TYPE_R64 -I include -I pangulu_b200/csrc/platform ...
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_kernels.cuh"
using namespace psm;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int TM, int TNC, int GLOG>
__global__ void __launch_bounds__(512, 1) k_loop(double* out, int iters, int m, int nt, int nch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    const int ld = ((m + 7) & ~7) + 4;
    const int chunk = PSM_CT_KC * ld + nt * PSM_CT_BLD;
    for (int i = threadIdx.x; i < nch * chunk; i += 512) sm[i] = 1.0 + (i % 7) * 1e-3;
    __syncthreads();
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, lr = lane / 4, lk = lane % 4;
    constexpr int gm = 1 << GLOG;
    const int mi = warp & (gm - 1), ni = warp >> GLOG;
    const int m_base = mi * TM * 8, n_base = ni * TNC * 8;
    double racc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) racc[i] = 0.0;
    for (int it = 0; it < iters; ++it)
        for (int q = 0; q < nch; ++q)
        {
            const double* As = sm + (size_t)q * chunk;
            const double* Bs = As + PSM_CT_KC * ld;
            ct_chunk<TM, TNC>(racc, As + lk * ld + m_base + lr, Bs + (n_base + lr) * PSM_CT_BLD + lk, ld);
        }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += racc[i];
    if (s == 12345.678) out[threadIdx.x] = s;
}

template <int TM, int TNC, int GLOG>
int run(const char* name)
{
    constexpr int gm = 1 << GLOG, gn = 16 >> GLOG;
    const int m = gm * TM * 8, nt = gn * TNC * 8;
    const int ld = m + 4;
    const int nch = 2;
    const size_t smem = (size_t)nch * (PSM_CT_KC * ld + nt * PSM_CT_BLD) * 8;
    double* out; CK(cudaMalloc(&out, 4096 * 8));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(k_loop<TM, TNC, GLOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 4000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep)
    {
        CK(cudaEventRecord(e0));
        k_loop<TM, TNC, GLOG><<<sms, 512, smem>>>(out, iters, m, nt, nch);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double dmma_per_warp = (double)iters * nch * 4 * TM * TNC;
    const double tf = dmma_per_warp * 16 * sms * 512 / (ms * 1e-3) / 1e12;
    printf("%-28s tile %3d x %2d, warp grid %2d x %d, reg tile %d x %d: %8.3f ms  %6.2f TFLOP/s  %5.2f cycles/DMMA/subpartition at 1965 MHz\n", name, m, nt, gm, gn, TM, TNC, ms,
           tf, (ms * 1e-3) * 1.965e9 / (dmma_per_warp * 4));
    CK(cudaFree(out));
    return 0;
}

int main()
{
    run<4, 4, 3>("dense 256x64 (new)");
    run<4, 2, 3>("256x32 (round-1 tall grid)");
    run<4, 1, 2>("128x32 small grid");
    run<2, 2, 3>("128x32");
    run<2, 1, 4>("256x8");
    run<1, 2, 4>("128x16");
    run<1, 1, 4>("128x8");
    run<3, 3, 3>("192x48");
    run<4, 3, 3>("256x48");
    return 0;
}
