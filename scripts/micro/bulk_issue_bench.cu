// Micro-benchmark: what does ISSUING a cp.async.bulk (UBLKCP, global -> shared) cost the issuing thread, and when does the
// data land, as a function of the copy size?  One CTA per SM, one thread issues.  Development probe only.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// pieces: the S bytes are issued as `pieces` copies by ONE thread (lanes = 1) or spread over `lanes` threads of the warp
__global__ void __launch_bounds__(128, 1) k_bulk(const char* src, size_t stride, int S, int pieces, int lanes, int reps, long long* out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    const char* base = src + (size_t)blockIdx.x * stride;
    long long t_issue = 0, t_done = 0;
    unsigned parity = 0;
    const int piece = S / pieces;
    for (int r = 0; r < reps; ++r)
    {
        __syncthreads();
        long long t0 = clock64();
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(S) : "memory");
        __syncwarp();
        if (tid < lanes)
            for (int p = tid; p < pieces; p += lanes)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (size_t)p * piece)),
                             "l"(base + (size_t)((r * 7) % 3) * S + (size_t)p * piece), "r"(piece), "r"(smem_u32(&bar))
                             : "memory");
        long long t1 = clock64();
        if (tid == 0)
        {
            unsigned ok = 0;
            do { asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory"); } while (!ok);
        }
        long long t2 = clock64();
        parity ^= 1;
        if (tid == 0 && r > 0) { t_issue += t1 - t0; t_done += t2 - t0; }
    }
    if (tid == 0) { out[2 * blockIdx.x] = t_issue / (reps - 1); out[2 * blockIdx.x + 1] = t_done / (reps - 1); }
}

int main()
{
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t stride = 3 * 65536;
    char* src; CK(cudaMalloc(&src, stride * sms)); CK(cudaMemset(src, 1, stride * sms));
    long long* out; CK(cudaMalloc(&out, sms * 16));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    long long h[2 * 256];
    printf("cp.async.bulk global->shared, one CTA per SM (%d SMs), source L2-resident; cycles are per-SM means\n", sms);
    printf("%8s %7s %6s | %12s %12s | %10s\n", "bytes", "pieces", "lanes", "issue cyc", "landed cyc", "B/clk/SM");
    for (int S : {1024, 4096, 16384, 32768, 65536})
        for (int cfg = 0; cfg < 5; ++cfg)
        {
            int pieces = 1, lanes = 1;
            if (cfg == 1) { pieces = 4; lanes = 1; }
            if (cfg == 2) { pieces = 16; lanes = 1; }
            if (cfg == 3) { pieces = 32; lanes = 32; }
            if (cfg == 4) { pieces = 4; lanes = 4; }
            if (S / pieces < 16) continue;
            k_bulk<<<sms, 128, 65536>>>(src, stride, S, pieces, lanes, 50, out);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, out, sms * 16, cudaMemcpyDeviceToHost));
            double a = 0, b = 0;
            for (int i = 0; i < sms; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
            a /= sms; b /= sms;
            printf("%8d %7d %6d | %12.0f %12.0f | %10.1f\n", S, pieces, lanes, a, b, S / b);
        }
    return 0;
}
