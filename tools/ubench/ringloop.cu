// micro-benchmark of the CHAIN-style reach loop: cp.async ring + LDS + FMA + STG, one warp per CTA
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int kRing = 12;
__device__ __forceinline__ void cp_row(unsigned saddr, const double* g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ double2 lds_row(unsigned saddr) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(saddr)); return v; }
__device__ __forceinline__ double lds_f64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(unsigned a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

// MODE bit0: stores, bit1: metadata LDS, bit2: use wait_group<kRing-1> (else wait_all each iter)
template <int MODE>
__global__ void k(double* O, double* I, long long* out, int len, int ld, int reps)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned ring = sbase + 4096 + lane * 16u, ringO = ring + kRing * 512u;
    // fake metadata
    for (int i = lane; i < 512; i += 32) { ((double*)smem)[i] = 0.5; }
    __syncwarp();
    double* Or = O + (size_t)blockIdx.x * 64 * ld + lane * 2;
    double* Ir = I + (size_t)blockIdx.x * 64 * ld + lane * 2;
    long long t0 = clock64();
    double2 o = make_double2(0, 0);
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int j = 0; j < kRing; ++j) {
            if (j < len) { cp_row(ring + j * 512u, Ir + (size_t)j * ld); cp_row(ringO + j * 512u, Or + (size_t)j * ld); }
            cp_async_commit();
        }
        int sl = 0;
        for (int i = 0; i < len; ++i) {
            if (MODE & 4) cp_async_wait_group<kRing - 1>(); else cp_async_wait_group<0>();
            const double2 side = lds_row(ring + sl * 512u), b = lds_row(ringO + sl * 512u);
            if (i + kRing < len) { cp_row(ring + sl * 512u, Ir + (size_t)(i + kRing) * ld); cp_row(ringO + sl * 512u, Or + (size_t)(i + kRing) * ld); }
            cp_async_commit();
            sl = sl + 1 == kRing ? 0 : sl + 1;
            double al = 0.5; uint32_t h = 1;
            if (MODE & 2) { h = lds_u32(sbase + 2048 + 4 * i); al = lds_f64(sbase + 32 * i); }
            double2 inflow = (h & 1) ? o : make_double2(0, 0);
            double2 on, it;
            on.x = al * inflow.x + b.x; on.y = al * inflow.y + b.y;
            it.x = inflow.x + side.x; it.y = inflow.y + side.y;
            if (MODE & 1) { __stcg((double2*)(Ir + (size_t)i * ld), it); __stcg((double2*)(Or + (size_t)i * ld), on); }
            o = on;
        }
        cp_async_wait_group<0>();
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
    if (o.x == 123.456) out[0] = 0;
}
template <int MODE> void run(const char* name, double* O, double* I, long long* out, int grid)
{
    const int len = 32, ld = 64, reps = 50;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480);
    k<MODE><<<grid, 32, 20480>>>(O, I, out, len, ld, reps);
    k<MODE><<<grid, 32, 20480>>>(O, I, out, len, ld, reps);
    long long h[1]; cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s grid %4d: %.1f cyc/reach\n", name, grid, (double)h[0] / (len * reps));
}
int main()
{
    double *O, *I; long long* out;
    size_t n = (size_t)2048 * 64 * 64;
    cudaMalloc(&O, n * 8); cudaMalloc(&I, n * 8); cudaMemset(O, 0, n * 8); cudaMemset(I, 0, n * 8);
    cudaMalloc(&out, 8 * 2048);
    for (int grid : {1, 444, 1776}) {
        run<4>("ring only (wait_group 11)", O, I, out, grid);
        run<0>("ring only (wait_group 0 each iter)", O, I, out, grid);
        run<5>("ring + stores", O, I, out, grid);
        run<6>("ring + metadata LDS", O, I, out, grid);
        run<7>("ring + stores + metadata (chain loop)", O, I, out, grid);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
