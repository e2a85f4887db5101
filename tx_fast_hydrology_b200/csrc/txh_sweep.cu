// txh_sweep.cu -- one routing step of a SMALL network, latency first (sm_100a).
//
// The reference's operational models are sub-basins of 10^2..10^4 reaches (app/app.py:121-166), stepped one
// step at a time with a dense Kalman filter after every step; its covariance propagation `_aqat_par`
// (nutils.py:194-214) routes the n columns of P.  At that size the critical path of a task-parallel launch
// is a few hundred cross-SM hand-overs (~160 us measured for n = 1000), more than walking every reach in
// order.  So: ONE warp per 32 member columns walks the whole network depth first (Sweep, txh_topology.hpp),
// columns are independent, nothing is exchanged between warps.  Per row the dependent chain is one add and
// one FMA -- o' = alpha (acc + parked) + (beta i + chi o + gamma q) -- with the outflow of the previous row
// in registers and the other upstream reaches parked in shared-memory slots.  Rows and their 48-byte records
// {hdr, row, alpha, beta, chi, gamma} stream in through a cp.async ring kSwDepth rows ahead (records
// 2*kSwDepth ahead, because the address of a state row comes out of its record).
//
// APPLY = true is the homogeneous operator of nutils.py:148-154: i_prev is rebuilt from the OLD upstream
// outflows (self-loop included, numba_init_inflows has no guard) on the fly, and only X is written.
#include <cuda_runtime.h>

#include <cstdint>

#include "txh_kernels.cuh"
#include "txh_topology.hpp"

namespace txh {

void count_launch();

namespace {

constexpr int kSwDepth = 16;
constexpr unsigned kSwRec = 48u;

__device__ __forceinline__ void cp_async8(unsigned sa, const void* g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(unsigned sa, const void* g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ double lds_f64(unsigned sa)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa));
    return v;
}
__device__ __forceinline__ double2 lds_v2(unsigned sa)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(sa));
    return v;
}
__device__ __forceinline__ void sts_v2(unsigned sa, double2 v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(sa), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ uint2 lds_u2(unsigned sa)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(sa));
    return v;
}

// records[k] = {hdr[k], row[k], alpha, beta, chi, gamma of that row, pad}; coef is [n][4] in row (schedule) order
__global__ void __launch_bounds__(256)
sweep_records_kernel(const uint32_t* __restrict__ hdr, const int32_t* __restrict__ row, const double* __restrict__ coef,
                     int n, unsigned char* __restrict__ rec)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int r = row[k];
    unsigned char* p = rec + (size_t)k * kSwRec;
    *reinterpret_cast<uint2*>(p) = make_uint2(hdr[k], (uint32_t)r);
    const double2 ab = *reinterpret_cast<const double2*>(coef + 4 * (size_t)r);
    const double2 cg = *reinterpret_cast<const double2*>(coef + 4 * (size_t)r + 2);
    *reinterpret_cast<double*>(p + 8) = ab.x;
    *reinterpret_cast<double2*>(p + 16) = make_double2(ab.y, cg.x);
    *reinterpret_cast<double2*>(p + 32) = make_double2(cg.y, 0.0);
}

struct SwRow {
    uint32_t h;
    int row;
    double al, be, ch, ga, x, i, q;
};

template <bool APPLY>
__global__ void __launch_bounds__(32)
route_sweep_kernel(const unsigned char* __restrict__ rec, int n, double* __restrict__ X, double* __restrict__ I,
                   const double* __restrict__ q, int ld, int M)
{
    extern __shared__ __align__(16) unsigned char sw_smem[];
    constexpr int D = kSwDepth;
    const int lane = threadIdx.x;
    const int col = blockIdx.x * 32 + lane;
    const bool active = col < M;
    const unsigned sRec = (unsigned)__cvta_generic_to_shared(sw_smem);          // [2D][48]
    const unsigned sX = sRec + 2u * D * kSwRec;                                  // [D][32] doubles
    const unsigned sI = sX + D * 256u;                                           // [D][32] doubles
    const unsigned sQ = sI + D * 256u;                                           // [D] doubles
    const unsigned sSlot = sQ + D * 8u;                                          // [slots][32] {old, new}

    // rows of sweep row kx, issued as one group (possibly empty, so that the group count stays in step)
    auto issue_rows = [&](int kx) {
        if (kx < n) {
            const int row = (int)lds_u2(sRec + (unsigned)(kx % (2 * D)) * kSwRec).y;
            const unsigned slot = (unsigned)(kx % D);
            if (active) {
                cp_async8(sX + slot * 256u + lane * 8u, X + (size_t)row * ld + col);
                if (!APPLY) cp_async8(sI + slot * 256u + lane * 8u, I + (size_t)row * ld + col);
            }
            if (!APPLY && q != nullptr && lane == 0) cp_async8(sQ + slot * 8u, q + row);
        }
    };
    auto issue_rec = [&](int kr) {
        if (kr < n && lane < 3) cp_async16(sRec + (unsigned)(kr % (2 * D)) * kSwRec + lane * 16u, rec + (size_t)kr * kSwRec + lane * 16u);
    };
    auto load = [&](int k) {
        SwRow r;
        const unsigned ra = sRec + (unsigned)(k % (2 * D)) * kSwRec;
        const uint2 hr = lds_u2(ra);
        r.h = hr.x; r.row = (int)hr.y;
        r.al = lds_f64(ra + 8u);
        const double2 bc = lds_v2(ra + 16u);
        r.be = bc.x; r.ch = bc.y;
        r.ga = lds_f64(ra + 32u);
        const unsigned slot = (unsigned)(k % D);
        r.x = lds_f64(sX + slot * 256u + lane * 8u);
        r.i = APPLY ? 0.0 : lds_f64(sI + slot * 256u + lane * 8u);
        r.q = (!APPLY && q != nullptr) ? lds_f64(sQ + slot * 8u) : 0.0;
        return r;
    };

    for (int kr = 0; kr < 2 * D; ++kr) issue_rec(kr);
    cp_async_commit();
    cp_async_wait_all();
    __syncwarp();
    for (int kx = 0; kx < D; ++kx) { issue_rows(kx); cp_async_commit(); }
    cp_async_wait_group<D - 1>();                                                // row 0 has landed
    __syncwarp();
    SwRow nx = load(0);
    double acc_old = 0.0, acc_new = 0.0;
    for (int k = 0; k < n; ++k) {
        const SwRow r = nx;
        __syncwarp();                                                            // every lane has read row k's ring cells
        // group k: record k + 2D, rows of k + D (their address is in record k + D, which came with group k - D)
        issue_rec(k + 2 * D);
        issue_rows(k + D);
        cp_async_commit();
        cp_async_wait_group<D - 1>();                                            // row k + 1 has landed
        __syncwarp();
        if (k + 1 < n) nx = load(k + 1);
        double in_new = (r.h & 1u) ? acc_new : 0.0;
        double in_old = (r.h & 1u) ? acc_old : 0.0;
        unsigned sa = sSlot + ((r.h >> 8) & 127u) * 512u + lane * 16u;
        for (uint32_t c = (r.h >> 15) & 255u; c > 0; --c) {
            const double2 v = lds_v2(sa);
            if (APPLY) in_old += v.x;
            in_new += v.y;
            sa += 512u;
        }
        double rest;
        if (APPLY) {
            if (r.h & SWEEP_OUTLET) in_old += r.x;                               // nutils.py:136-141: no self-loop guard
            rest = r.be * in_old + r.ch * r.x;
        } else {
            rest = r.be * r.i + (r.ch * r.x + r.ga * r.q);
        }
        const double on = r.al * in_new + rest;
        if (active) {
            X[(size_t)r.row * ld + col] = on;
            if (!APPLY) I[(size_t)r.row * ld + col] = in_new;
        }
        const uint32_t ps = (r.h >> 1) & 127u;
        if (ps) sts_v2(sSlot + (ps - 1u) * 512u + lane * 16u, make_double2(r.x, on));
        acc_old = r.x; acc_new = on;
    }
    cp_async_wait_all();
}

}  // namespace

size_t sweep_record_bytes(int64_t n) { return (size_t)n * kSwRec; }

cudaError_t launch_sweep_records(const uint32_t* hdr, const int32_t* row, const double* coef, int n, unsigned char* rec,
                                 cudaStream_t st)
{
    sweep_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hdr, row, coef, n, rec);
    count_launch();
    return cudaGetLastError();
}

size_t sweep_smem_bytes(int slots)
{
    return (size_t)2 * kSwDepth * kSwRec + (size_t)kSwDepth * (256 + 256 + 8) + (size_t)slots * 512;
}

cudaError_t launch_route_sweep(const unsigned char* rec, int n, int slots, double* X, double* I, const double* q, int ld,
                               int M, bool apply, cudaStream_t st)
{
    const size_t smem = sweep_smem_bytes(slots);
    const unsigned grid = (unsigned)((M + 31) / 32);
    cudaError_t e;
    if (apply) {
        if ((e = cudaFuncSetAttribute(route_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        route_sweep_kernel<true><<<grid, 32, smem, st>>>(rec, n, X, nullptr, nullptr, ld, M);
    } else {
        if ((e = cudaFuncSetAttribute(route_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        route_sweep_kernel<false><<<grid, 32, smem, st>>>(rec, n, X, I, q, ld, M);
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace txh
