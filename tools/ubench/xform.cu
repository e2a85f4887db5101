// How long does the in-load ensemble transform of route_window_kernel (transform_p_rows: 16 rows x 64 members times a
// 64 x 64 matrix on the FP64 tensor cores, operands in shared memory) take per task, with 1 .. 14 warps of an SM doing it
// at the same time?  Development aid.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned t_swz(int k, int c) { return (unsigned)(k * 64 + ((((c >> 3) ^ ((k >> 1) & 3)) << 3) | (c & 7))); }

template <int VARIANT>
__device__ __forceinline__ void transform(double* P, const double* T, int len, int lane)
{
    const int g = lane >> 2, t = lane & 3;
    for (int r0 = 0; r0 < len; r0 += 8) {
        double* row = P + (r0 + g) * 64;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { const double2 v = *reinterpret_cast<double2*>(row + 8 * j + 2 * t); s += v.x + v.y; }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const double mu = s * (1.0 / 64.0);
        double acc[8][2];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = 0.0;
        if (VARIANT == 0) {
#pragma unroll 2
            for (int ks = 0; ks < 16; ++ks) {
                const int k = 8 * (ks >> 1) + 2 * t + (ks & 1);
                const double av = row[k] - mu;
                double b[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = T[t_swz(k, 8 * j + g)];
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma8x8x4(acc[j][0], acc[j][1], av, b[j]);
            }
        } else {
            // A fragments first (vector loads), then the k loop
            double av[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const double2 v = *reinterpret_cast<double2*>(row + 8 * j + 2 * t); av[2 * j] = v.x - mu; av[2 * j + 1] = v.y - mu; }
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) {
                const int k = 8 * (ks >> 1) + 2 * t + (ks & 1);
                double b[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = T[t_swz(k, 8 * j + g)];
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma8x8x4(acc[j][0], acc[j][1], av[ks], b[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double2 v = *reinterpret_cast<double2*>(row + 8 * j + 2 * t);
            v.x += acc[j][0]; v.y += acc[j][1];
            *reinterpret_cast<double2*>(row + 8 * j + 2 * t) = v;
        }
    }
}


// both 8-row tiles of a task at once: every B fragment feeds two tensor-core operations
__device__ __forceinline__ void transform2(double* P, const double* T, int len, int lane)
{
    const int g = lane >> 2, t = lane & 3;
    double* row0 = P + g * 64;
    double* row1 = P + (8 + g) * 64;
    double mu[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double* row = i ? row1 : row0;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { const double2 v = *reinterpret_cast<double2*>(row + 8 * j + 2 * t); s += v.x + v.y; }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        mu[i] = s * (1.0 / 64.0);
    }
    double acc[2][8][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 2
    for (int ks = 0; ks < 16; ++ks) {
        const int k = 8 * (ks >> 1) + 2 * t + (ks & 1);
        const double a0 = row0[k] - mu[0], a1 = row1[k] - mu[1];
        double b[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = T[t_swz(k, 8 * j + g)];
#pragma unroll
        for (int j = 0; j < 8; ++j) { dmma8x8x4(acc[0][j][0], acc[0][j][1], a0, b[j]); dmma8x8x4(acc[1][j][0], acc[1][j][1], a1, b[j]); }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double* row = i ? row1 : row0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            double2 v = *reinterpret_cast<double2*>(row + 8 * j + 2 * t);
            v.x += acc[i][j][0]; v.y += acc[i][j][1];
            *reinterpret_cast<double2*>(row + 8 * j + 2 * t) = v;
        }
    }
}

template <int VARIANT>
__global__ void __launch_bounds__(448, 1) k(int nactive, int iters, long long* cyc)
{
    extern __shared__ double sm[];
    double* T = sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* P = sm + 4096 + warp * 16 * 64;
    for (int e = threadIdx.x; e < 4096; e += blockDim.x) T[e] = 1e-3 * (e % 17);
    for (int e = lane; e < 1024; e += 32) P[e] = 1.0 + 1e-3 * e;
    __syncthreads();
    if (warp >= nactive) return;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { if (VARIANT == 2) transform2(P, T, 16, lane); else transform<VARIANT>(P, T, 16, lane); __syncwarp(); }
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) cyc[warp] = t1 - t0;
}

int main()
{
    long long* dc; long long hc[16];
    cudaMalloc(&dc, 128);
    const int smem = (4096 + 14 * 1024) * 8;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int it = 200;
    for (int v = 0; v < 3; ++v)
        for (int na : {1, 2, 4, 8, 14}) {
            if (v == 0) k<0><<<148, 448, smem>>>(na, it, dc); else if (v == 1) k<1><<<148, 448, smem>>>(na, it, dc); else k<2><<<148, 448, smem>>>(na, it, dc);
            cudaMemcpy(hc, dc, 128, cudaMemcpyDeviceToHost);
            printf("variant %d, %2d warps: %.0f cycles per 16-row task (256 DMMA) = %.2f us, %.1f cycles per DMMA\n", v, na,
                   (double)hc[0] / it, (double)hc[0] / it / 1965.0, (double)hc[0] / it / 256);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
