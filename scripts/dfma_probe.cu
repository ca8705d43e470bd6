// Development aid: DFMA issue/latency on B200 as a function of resident
// warps per SM and independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 1.2345) out[0] = s;
}
template <int ILP>
void run(int warps_per_sm)
{
    int dev_sms = 148;
    double *d; cudaMalloc(&d, 8);
    int threads = 128, ctas_per_sm = warps_per_sm / 4;
    int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // dynamic smem forces the residency we want
    size_t smem = (size_t) (220 * 1024 / ctas_per_sm) - 2048;
    cudaFuncSetAttribute(k<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    k<ILP><<<dev_sms * ctas_per_sm, threads, smem>>>(d, 10, 0.999, 1e-9);
    cudaEventRecord(e0);
    k<ILP><<<dev_sms * ctas_per_sm, threads, smem>>>(d, iters, 0.999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double) dev_sms * ctas_per_sm * threads * iters * 16.0 * ILP;
    double per_smsp_cycle = fmas / 32.0 / (dev_sms * 4) / (ms * 1e-3 * 1.965e9);
    printf("warps/SM %2d ILP %d: %.2f TFLOP/s, %.3f warp-DFMA/cycle/SMSP\n", warps_per_sm, ILP,
           2 * fmas / (ms * 1e-3) / 1e12, per_smsp_cycle);
    cudaFree(d);
}
int main()
{
    for (int w : {4, 8, 16, 32}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
    return 0;
}
