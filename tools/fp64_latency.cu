// Micro-benchmark: dependent-issue latency and throughput of FP64 add / mul / fma on this GPU.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain_kernel(double* out, long long* cycles, int iters, double seed)
{
    double a = seed + threadIdx.x * 1e-9;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (OP == 0) a = __fma_rn(a, b, c);
            if (OP == 1) a = __dadd_rn(a, c);
            if (OP == 2) a = __dmul_rn(a, b);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
__global__ void tput_kernel(double* out, int iters, double seed)
{
    double a[8];
    for (int u = 0; u < 8; ++u) a[u] = seed + u + threadIdx.x * 1e-9;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) a[u] = __fma_rn(a[u], b, c);
            if (OP == 1) a[u] = __dadd_rn(a[u], c);
            if (OP == 2) a[u] = __dmul_rn(a[u], b);
        }
    }
    double s = 0;
    for (int u = 0; u < 8; ++u) s += a[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    cudaMallocManaged(&cyc, sizeof(long long));
    const char* names[3] = {"dfma", "dadd", "dmul"};
    printf("{");
    for (int op = 0; op < 3; ++op) {
        const int iters = 2000;
        if (op == 0) chain_kernel<0><<<1, 32>>>(out, cyc, iters, 1.0);
        if (op == 1) chain_kernel<1><<<1, 32>>>(out, cyc, iters, 1.0);
        if (op == 2) chain_kernel<2><<<1, 32>>>(out, cyc, iters, 1.0);
        cudaDeviceSynchronize();
        printf("\"%s_latency_cycles\": %.2f, ", names[op], (double)*cyc / (iters * 16.0));
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        const int blocks = prop.multiProcessorCount * 4, threads = 512, it2 = 20000;
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            if (op == 0) tput_kernel<0><<<blocks, threads>>>(out, it2, 1.0);
            if (op == 1) tput_kernel<1><<<blocks, threads>>>(out, it2, 1.0);
            if (op == 2) tput_kernel<2><<<blocks, threads>>>(out, it2, 1.0);
            cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        }
        printf("\"%s_tops\": %.2f, ", names[op], 8.0 * it2 * (double)blocks * threads / (ms * 1e-3) / 1e12);
    }
    printf("\"sm_clock_khz\": %d}\n", prop.clockRate);
    return 0;
}
