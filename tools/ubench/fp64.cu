// FP64 pipe microbenchmarks for sm_100a: latency of dependent DFMA / rsqrt / division chains, throughput of
// independent DFMA streams per SM, and DMMA (mma.sync m8n8k4 f64) throughput.  Development aid.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dep_dfma(double* out, double a, double b, int iters, long long* cyc)
{
    double x = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x = x * a + b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void dep_rsqrt(double* out, int iters, long long* cyc)
{
    double x = out[threadIdx.x] + 2.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) x = rsqrt(x) + 1.5;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void dep_div(double* out, int iters, long long* cyc)
{
    double x = out[threadIdx.x] + 2.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) x = 3.0 / x + 1.5;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

// 8 independent chains per thread
__global__ void thr_dfma(double* out, double a, double b, int iters)
{
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = out[threadIdx.x] + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = x[k] * a + b;
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void thr_dmma(double* out, int iters)
{
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) { c[k][0] = threadIdx.x; c[k][1] = k; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dep_dmma(double* out, int iters, long long* cyc)
{
    double c0 = threadIdx.x, c1 = 1.0, a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
    double* d; long long* dc; long long hc;
    cudaMalloc(&d, sizeof(double) * 148 * 1024 * 4); cudaMalloc(&dc, 8);
    cudaMemset(d, 0, sizeof(double) * 148 * 1024 * 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const int it = 2000;
    dep_dfma<<<1, 32>>>(d, 0.999, 0.001, it, dc); cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency        : %.1f cycles\n", (double)hc / (it * 16));
    dep_rsqrt<<<1, 32>>>(d, it, dc); cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    printf("dependent rsqrt(double)+add   : %.1f cycles\n", (double)hc / (it * 4));
    dep_div<<<1, 32>>>(d, it, dc); cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    printf("dependent division+add        : %.1f cycles\n", (double)hc / (it * 4));
    dep_dmma<<<1, 32>>>(d, it, dc); cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DMMA m8n8k4 latency : %.1f cycles\n", (double)hc / (it * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int warps = 1; warps <= 32; warps *= 2) {
        thr_dfma<<<p.multiProcessorCount, warps * 32>>>(d, 0.999, 0.001, it);
        cudaEventRecord(e0);
        thr_dfma<<<p.multiProcessorCount, warps * 32>>>(d, 0.999, 0.001, it);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)p.multiProcessorCount * warps * 32 * 8.0 * it;
        printf("DFMA throughput, %2d warps/SM : %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", warps, 2 * fma / ms / 1e9,
               fma / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
    }
    for (int warps = 1; warps <= 32; warps *= 2) {
        thr_dmma<<<p.multiProcessorCount, warps * 32>>>(d, it);
        cudaEventRecord(e0);
        thr_dmma<<<p.multiProcessorCount, warps * 32>>>(d, it);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)p.multiProcessorCount * warps * 8.0 * it * 256.0;
        printf("DMMA throughput, %2d warps/SM : %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", warps, 2 * fma / ms / 1e9,
               fma / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
