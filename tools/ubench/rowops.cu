// micro-benchmark: cost of store->load sequences with strong (.cg) vs weak row ops, one warp
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* buf, long long* out, int iters, int ld)
{
    const int lane = threadIdx.x;
    double2 acc = make_double2(0, 0);
    double* p = buf + lane * 2;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        double2 v;
        const double* src = p + (size_t)(i + 8) * ld;
        double* dst = p + (size_t)i * ld + (size_t)4096 * ld;
        if (MODE == 0) { v = __ldcg((const double2*)src); }
        else { asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(src)); }
        acc.x += v.x; acc.y += v.y;
        if (MODE == 0) __stcg((double2*)dst, acc);
        else asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(dst), "d"(acc.x), "d"(acc.y) : "memory");
    }
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    if (acc.x == 123.0) out[1] = 1;
}
// prefetch distance 8 variant: load i+8 issued, consume i
template <int MODE>
__global__ void kp(double* buf, long long* out, int iters, int ld)
{
    const int lane = threadIdx.x;
    double2 acc = make_double2(0, 0);
    double* p = buf + lane * 2;
    double2 r[8];
    for (int j = 0; j < 8; ++j) r[j] = __ldcg((const double2*)(p + (size_t)j * ld));
    long long t0 = clock64();
    for (int i0 = 0; i0 < iters; i0 += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + j;
            double2 v = r[j];
            const double* src = p + (size_t)(i + 8) * ld;
            if (MODE == 0) r[j] = __ldcg((const double2*)src);
            else asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r[j].x), "=d"(r[j].y) : "l"(src));
            acc.x = acc.x * 0.5 + v.x; acc.y = acc.y * 0.5 + v.y;
            double* dst = p + (size_t)i * ld + (size_t)8192 * ld;
            if (MODE == 0) __stcg((double2*)dst, acc);
            else asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(dst), "d"(acc.x), "d"(acc.y) : "memory");
        }
    }
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    if (acc.x == 123.0) out[1] = 1;
}
int main()
{
    const int ld = 64, iters = 4096;
    double* buf; long long* out; long long h[2];
    cudaMalloc(&buf, sizeof(double) * ld * 20000); cudaMemset(buf, 0, sizeof(double) * ld * 20000);
    cudaMalloc(&out, 16);
    for (int rep = 0; rep < 2; ++rep) {
        k<0><<<1, 32>>>(buf, out, iters, ld); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("dependent ld->st strong(.cg): %.1f cyc/iter\n", (double)h[0] / iters);
        k<1><<<1, 32>>>(buf, out, iters, ld); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("dependent ld->st weak       : %.1f cyc/iter\n", (double)h[0] / iters);
        kp<0><<<1, 32>>>(buf, out, iters, ld); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("prefetch-8 strong(.cg)      : %.1f cyc/iter\n", (double)h[0] / iters);
        kp<1><<<1, 32>>>(buf, out, iters, ld); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost); printf("prefetch-8 weak             : %.1f cyc/iter\n", (double)h[0] / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
