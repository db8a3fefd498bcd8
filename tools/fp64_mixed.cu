// Micro-benchmark: do DMMA (mma.sync.m8n8k4.f64) and DFMA share one FP64 pipe on this GPU?  Three kernels with the same
// instruction counts per warp: DMMA only, DFMA only, and both interleaved in the same warp.  If the mixed kernel takes
// ~max(t_dmma, t_dfma) the two run on separate pipes (the contraction and the FP64 epilogue overlap, and a contraction can
// be split across both); if it takes ~t_dmma + t_dfma they share the pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mixed tools/fp64_mixed.cu && ./fp64_mixed
#include <cstdio>
#include <cuda_runtime.h>

#define DMMA(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b))

template <int kMode>   // 0 dmma, 1 dfma, 2 both
__global__ void mixed_kernel(double* out, int iters)
{
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
    double f0 = threadIdx.x * 1e-3, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
    double a = threadIdx.x * 1e-3, b = 1.0000001;
    const double m = 1.0000001, k = 1e-9;
    for (int i = 0; i < iters; ++i) {
        if (kMode != 1) { DMMA(c0); DMMA(c1); DMMA(c2); DMMA(c3); }
        if (kMode != 0) {
            // 4 DMMA occupy the DMMA path for as long as 32 DFMA occupy the FMA path (8x the flops per instruction)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                f0 = fma(f0, m, k); f1 = fma(f1, m, k); f2 = fma(f2, m, k); f3 = fma(f3, m, k);
                f4 = fma(f4, m, k); f5 = fma(f5, m, k); f6 = fma(f6, m, k); f7 = fma(f7, m, k);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1] + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms[3] = {0, 0, 0};
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); mixed_kernel<0><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms[0], e0, e1);
        cudaEventRecord(e0); mixed_kernel<1><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms[1], e0, e1);
        cudaEventRecord(e0); mixed_kernel<2><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms[2], e0, e1);
    }
    const double warps = (double)blocks * (threads / 32);
    const double dmma_flops = 2.0 * 256 * 4 * iters * warps, dfma_flops = 2.0 * 32 * 32 * iters * warps;
    printf("{\"gpu\": \"%s\", \"dmma_only_ms\": %.3f, \"dfma_only_ms\": %.3f, \"mixed_ms\": %.3f, \"dmma_tflops\": %.2f, \"dfma_tflops\": %.2f, "
           "\"mixed_total_tflops\": %.2f, \"verdict\": \"%s\"}\n",
           prop.name, ms[0], ms[1], ms[2], dmma_flops / (ms[0] * 1e-3) / 1e12, dfma_flops / (ms[1] * 1e-3) / 1e12,
           (dmma_flops + dfma_flops) / (ms[2] * 1e-3) / 1e12,
           ms[2] < 0.75 * (ms[0] + ms[1]) ? "separate pipes: DMMA and DFMA overlap" : "shared pipe: DMMA and DFMA serialise");
    return 0;
}
