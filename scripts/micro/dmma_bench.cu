// Micro-benchmark: FP64 tensor (DMMA) shapes and DFMA issue rates on sm_100a.  Development probe only.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int SHAPE, int NACC>
__global__ void __launch_bounds__(512, 1) k_dmma(double* out, int iters, double a0, double b0)
{
    double acc[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = a0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; ++i) b[i] = b0 + threadIdx.x * 1e-9 - i;
    for (int it = 0; it < iters; ++it)
    {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
        {
            if constexpr (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a[0]), "d"(b[0]));
            else if constexpr (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5},{%6},{%0,%1,%2,%3};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if constexpr (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else if constexpr (SHAPE == 3)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7,%8,%9,%10,%11},{%12,%13,%14,%15},{%0,%1,%2,%3};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            else
            {
                // DFMA: 4 independent FMAs per "instruction group"
                acc[i][0] = fma(a[0], b[0], acc[i][0]); acc[i][1] = fma(a[1], b[1], acc[i][1]);
                acc[i][2] = fma(a[2], b[2], acc[i][2]); acc[i][3] = fma(a[3], b[3], acc[i][3]);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    if (s == 12345.678) out[threadIdx.x] = s;
}

template <int SHAPE, int NACC>
int run(const char* name, double flop_per_inst, int threads)
{
    double* out; CK(cudaMalloc(&out, 4096 * 8));
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int iters = 20000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep)
    {
        CK(cudaEventRecord(e0));
        k_dmma<SHAPE, NACC><<<sms, threads>>>(out, iters, 1.0, 2.0);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double warps = (double)sms * threads / 32;
    const double insts = warps * iters * NACC;
    const double tf = insts * flop_per_inst / (ms * 1e-3) / 1e12;
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("%-14s threads %4d nacc %2d: %8.3f ms  %7.2f TFLOP/s  (%.2f cycles/inst/subpartition at %d MHz nominal)\n", name, threads, NACC, ms, tf,
           (ms * 1e-3) * clk * 1e3 / (iters * NACC * (threads / 32 / 4.0)), clk / 1000);
    CK(cudaFree(out));
    return 0;
}

int main()
{
    for (int t : {128, 256, 512})
    {
        if (t == 128) { run<0, 8>("m8n8k4", 512, 128); run<1, 8>("m16n8k4", 1024, 128); run<2, 8>("m16n8k8", 2048, 128); run<3, 8>("m16n8k16", 4096, 128); run<4, 8>("dfma x4", 4 * 2 * 32, 128); }
        if (t == 256) { run<0, 8>("m8n8k4", 512, 256); run<1, 8>("m16n8k4", 1024, 256); run<2, 8>("m16n8k8", 2048, 256); run<3, 8>("m16n8k16", 4096, 256); run<4, 8>("dfma x4", 4 * 2 * 32, 256); }
        if (t == 512) { run<0, 8>("m8n8k4", 512, 512); run<1, 8>("m16n8k4", 1024, 512); run<2, 8>("m16n8k8", 2048, 512); run<3, 8>("m16n8k16", 4096, 512); run<4, 8>("dfma x4", 4 * 2 * 32, 512); }
    }
    run<0, 2>("m8n8k4 dep2", 512, 512);
    run<3, 2>("m16n8k16 dep2", 4096, 512);
    run<3, 4>("m16n8k16 dep4", 4096, 512);
    return 0;
}
