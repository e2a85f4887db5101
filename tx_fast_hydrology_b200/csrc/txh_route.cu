// txh_route.cu -- routing kernels for sm_100a.
//
// The per-timestep Muskingum update  o = alpha*i_next + beta*i_prev + chi*o_prev + gamma*q,
// i_next[down] += o  (tx_fast_hydrology/nutils.py:64-89) is a sparse unit-lower-triangular
// solve per step.  route_dataflow_kernel evaluates `nsteps` of them in ONE persistent
// launch: the network is cut into tasks (pure-chain spine segments and bundled side
// subtrees, txh_topology.cpp); warps claim (step, task, member-block) tickets in a
// topological, critical-path-first order and synchronise point-to-point through
// per-task step counters in global memory -- no grid-wide barrier, so consecutive
// timesteps overlap along the network (time skewing) while every dependency of the
// reference's walk is honoured.  Ensemble members are the SIMD axis: a lane owns two
// adjacent member columns (one 128-bit access per row), so a thread only ever
// reads values of its own columns and intra-task hand-offs need no synchronisation.
#include "txh_kernels.cuh"

#include <atomic>

namespace txh {

static std::atomic<int64_t> g_launches{0};
int64_t launch_count() { return g_launches.load(); }

namespace {

__device__ __forceinline__ int ld_acquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed(const int* p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// state rows are produced and consumed by different SMs inside one launch: bypass L1
__device__ __forceinline__ double2 ld_row(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void st_row(double* p, double2 v) { __stcg(reinterpret_cast<double2*>(p), v); }


// Spin until *flag >= need.  Returns false if the launch was poisoned (another warp
// hit the watchdog) or this wait itself ran out of time.
__device__ __forceinline__ bool wait_flag(const int* flag, int need, int* status, unsigned long long limit_ns)
{
    if (ld_acquire(flag) >= need) return true;
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire(flag) < need) {
        ++spins;
        if ((spins & 255u) == 0) {
            if (ld_relaxed(status) != 0) return false;
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > limit_ns) { atomicExch(status, 1); return false; }
        }
        __nanosleep(40);
    }
    return true;
}

template <bool HAS_F, bool HAS_W, bool REC>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
route_dataflow_kernel(const RouteArgs a)
{
    extern __shared__ double2 scratch_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double2* scratch = scratch_all + (size_t)warp * a.slots * 32 + lane;   // this lane's column

    const long long per_step = (long long)a.n_tasks * a.n_mblocks;
    const long long total = per_step * a.nsteps;
    const int ld = a.ld;

    if (ld_relaxed(a.status) != 0) return;          // a previous launch on this handle was poisoned

    for (;;) {
        unsigned long long tk = 0;
        if (lane == 0) tk = atomicAdd(a.ticket, 1ull);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if ((long long)tk >= total) break;
        const int s = (int)(tk / per_step);
        const int rem = (int)(tk - (long long)s * per_step);
        const int task = rem / a.n_mblocks;
        const int mb = rem - task * a.n_mblocks;
        const TaskDesc td = a.tasks[task];

        // ---- dependencies: producers done with step s, consumers and self done with s-1 ----
        bool ok = true;
        const int ndep = td.n_raw + td.n_war + 1;
        for (int d = lane; d < ndep; d += 32) {
            int idx, need;
            if (d < td.n_raw) { idx = a.deps[td.dep_off + d]; need = s + 1; }
            else if (d < td.n_raw + td.n_war) { idx = a.deps[td.dep_off + d]; need = s; }
            else { idx = task; need = s; }
            ok = wait_flag(a.done + (size_t)idx * a.n_mblocks + mb, need, a.status, a.watchdog_ns) && ok;
        }
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;

        const int col = mb * kMemberBlock + lane * 2;
        const bool active = col < ld;
        const int ccol = active ? col : 0;

        double w0 = 0.0, w1 = 0.0;
        double2 wm0 = make_double2(0.0, 0.0), wm1 = wm0;
        const double* F0 = nullptr;
        const double* F1 = nullptr;
        if (HAS_F) {
            const StepInterp si = a.steps[s];
            F0 = a.F + (size_t)si.r0 * a.n;
            F1 = a.F + (size_t)si.r1 * a.n;
            w0 = si.w0; w1 = si.w1;
            if (HAS_W) {
                // member m sees  (w0*mul[r0][m])*F[r0] + (w1*mul[r1][m])*F[r1]
                const int c0 = min(col, a.wm_ld - 1), c1 = min(col + 1, a.wm_ld - 1);
                const double* m0 = a.Wmul + (size_t)si.r0 * a.wm_ld;
                const double* m1 = a.Wmul + (size_t)si.r1 * a.wm_ld;
                wm0 = make_double2(w0 * __ldg(m0 + c0), w0 * __ldg(m0 + c1));
                wm1 = make_double2(w1 * __ldg(m1 + c0), w1 * __ldg(m1 + c1));
            }
        }

        const uint32_t* inp = a.inw + td.in_off;
        double* Orow = a.O + (size_t)td.begin * ld + ccol;
        double* Irow = a.I + (size_t)td.begin * ld + ccol;
        const double* O0 = a.O + ccol;
        double2 acc = make_double2(0.0, 0.0);
        double2 io_n = make_double2(0.0, 0.0), oo_n = io_n;
        if (active) { io_n = ld_row(Irow); oo_n = ld_row(Orow); }

        for (int k = td.begin; k < td.begin + td.len; ++k, Orow += ld, Irow += ld) {
            const double2 io = io_n, oo = oo_n;
            if (active && k + 1 < td.begin + td.len) { io_n = ld_row(Irow + ld); oo_n = ld_row(Orow + ld); }
            const uint32_t h = __ldg(a.hdr + k);
            const double2 c01 = __ldg(reinterpret_cast<const double2*>(a.coef + 4 * (size_t)k));
            const double2 c23 = __ldg(reinterpret_cast<const double2*>(a.coef + 4 * (size_t)k) + 1);
            double2 inflow = (h & HDR_ACC) ? acc : make_double2(0.0, 0.0);
            const int nin = (int)(h >> 6);
            for (int t = 0; t < nin; ++t) {
                const uint32_t w = __ldg(inp++);
                double2 v;
                if (w & INW_ROW) v = active ? ld_row(O0 + (size_t)(w & ~INW_ROW) * ld) : make_double2(0.0, 0.0);
                else v = scratch[w * 32];
                inflow.x += v.x; inflow.y += v.y;
            }
            double2 r;
            r.x = c01.y * io.x + c23.x * oo.x;                 // beta*i_prev + chi*o_prev
            r.y = c01.y * io.y + c23.x * oo.y;
            if (HAS_F) {
                const double f0 = __ldg(F0 + k), f1 = __ldg(F1 + k);
                double2 q;
                if (HAS_W) { q.x = wm0.x * f0 + wm1.x * f1; q.y = wm0.y * f0 + wm1.y * f1; }
                else { q.x = w0 * f0 + w1 * f1; q.y = q.x; }
                r.x += c23.y * q.x;                            // + gamma*q
                r.y += c23.y * q.y;
            }
            double2 on;
            on.x = c01.x * inflow.x + r.x;                     // alpha*i_next + ...
            on.y = c01.x * inflow.y + r.y;
            if (active) { st_row(Irow, inflow); st_row(Orow, on); }
            const uint32_t slot = (h >> 1) & 31u;
            if (slot) scratch[(slot - 1) * 32] = on;
            acc = on;
            if (REC) {
                const int rs = a.rec_slot[k];
                if (rs >= 0 && ((s + 1) % a.rec_every) == 0 && col < a.M) {
                    double* dst = a.rec_out + ((size_t)((s + 1) / a.rec_every - 1) * a.rec_count + rs) * a.M + col;
                    dst[0] = on.x;
                    if (col + 1 < a.M) dst[1] = on.y;
                }
            }
        }

        // ---- publish: every lane's rows are visible before the step counter moves ----
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            st_release(a.done + (size_t)task * a.n_mblocks + mb, s + 1);
        }
    }
}

// One topological level of one step (independent second device path).
__global__ void __launch_bounds__(256) route_level_kernel(const LevelArgs a)
{
    const int half = (a.ld + 1) / 2;                          // double2 columns per row
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)a.count * half) return;
    const int r = (int)(gid / half);
    const int col = (int)(gid - (long long)r * half) * 2;
    const int k = a.lvl_pos[r];
    double2 inflow = make_double2(0.0, 0.0);
    for (int u = a.up_off[k]; u < a.up_off[k + 1]; ++u) {
        const double2 v = ld_row(a.O + (size_t)a.up_pos[u] * a.ld + col);
        inflow.x += v.x; inflow.y += v.y;
    }
    const double al = a.coef[4 * (size_t)k], be = a.coef[4 * (size_t)k + 1], ch = a.coef[4 * (size_t)k + 2],
                 ga = a.coef[4 * (size_t)k + 3];
    double* Orow = a.O + (size_t)k * a.ld + col;
    double* Irow = a.I + (size_t)k * a.ld + col;
    const double2 io = ld_row(Irow), oo = ld_row(Orow);
    const double q = a.q ? a.q[k] : 0.0;
    double2 on;
    on.x = al * inflow.x + (be * io.x + ch * oo.x + ga * q);
    on.y = al * inflow.y + (be * io.y + ch * oo.y + ga * q);
    st_row(Irow, inflow);
    st_row(Orow, on);
}

__global__ void __launch_bounds__(256)
init_inflows_kernel(const int32_t* up_off, const int32_t* up_pos, const uint8_t* is_outlet,
                    const double* O, double* I, long long n, int ld)
{
    const int half = (ld + 1) / 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * half) return;
    const long long k = gid / half;
    const int col = (int)(gid - k * half) * 2;
    double2 s = make_double2(0.0, 0.0);
    for (int u = up_off[k]; u < up_off[k + 1]; ++u) {
        const double2 v = *reinterpret_cast<const double2*>(O + (size_t)up_pos[u] * ld + col);
        s.x += v.x; s.y += v.y;
    }
    if (is_outlet[k]) {      // muskingum.py:417 / nutils.py:136-141: no self-loop guard
        const double2 v = *reinterpret_cast<const double2*>(O + (size_t)k * ld + col);
        s.x += v.x; s.y += v.y;
    }
    *reinterpret_cast<double2*>(I + (size_t)k * ld + col) = s;
}

__global__ void __launch_bounds__(256)
apply_gain_kernel(const int32_t* up_off, const int32_t* up_pos, const double* G, double* O, double* I,
                  long long n, int ld)
{
    const int half = (ld + 1) / 2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * half) return;
    const long long k = gid / half;
    const int col = (int)(gid - k * half) * 2;
    double2 s = make_double2(0.0, 0.0);
    for (int u = up_off[k]; u < up_off[k + 1]; ++u) {
        const double2 v = *reinterpret_cast<const double2*>(G + (size_t)up_pos[u] * ld + col);
        s.x += v.x; s.y += v.y;
    }
    const double2 g = *reinterpret_cast<const double2*>(G + (size_t)k * ld + col);
    double2* op = reinterpret_cast<double2*>(O + (size_t)k * ld + col);
    double2* ip = reinterpret_cast<double2*>(I + (size_t)k * ld + col);
    double2 o = *op, i = *ip;
    o.x += g.x; o.y += g.y;
    i.x += s.x; i.y += s.y;
    *op = o; *ip = i;
}

__global__ void __launch_bounds__(256)
pack_kernel(const int32_t* reach_of_pos, const double* src, double* dst, long long n, int M, int ld,
            int src_member_major)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * ld) return;
    const long long k = gid / ld;
    const int m = (int)(gid - k * ld);
    double v = 0.0;
    if (m < M) {
        const long long j = reach_of_pos[k];
        v = src_member_major ? src[(size_t)m * n + j] : src[(size_t)j * M + m];
    }
    dst[gid] = v;
}

__global__ void __launch_bounds__(256)
unpack_kernel(const int32_t* reach_of_pos, const double* src, double* dst, long long n, int M, int ld,
              int dst_member_major)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * M) return;
    const long long k = gid / M;
    const int m = (int)(gid - k * M);
    const long long j = reach_of_pos[k];
    const double v = src[(size_t)k * ld + m];
    if (dst_member_major) dst[(size_t)m * n + j] = v; else dst[(size_t)j * M + m] = v;
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const int32_t* pos, long long count, const double* X, int ld, int M, double* out)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= count * M) return;
    const long long k = gid / M;
    const int m = (int)(gid - k * M);
    out[gid] = X[(size_t)pos[k] * ld + m];
}

__global__ void __launch_bounds__(256)
permute_vec_kernel(const int32_t* reach_of_pos, const double* src, double* dst, long long n)
{
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < n) dst[gid] = src[reach_of_pos[gid]];
}

inline unsigned blocks_for(long long work, int threads) { return (unsigned)((work + threads - 1) / threads); }

}  // namespace

cudaError_t launch_route_dataflow(const RouteArgs& a, int num_sms, cudaStream_t st)
{
    const size_t smem = (size_t)kWarpsPerCta * a.slots * 32 * sizeof(double2);
    void (*kern)(const RouteArgs) = nullptr;
    const bool f = a.F != nullptr, w = a.Wmul != nullptr, r = a.rec_slot != nullptr;
    if (!f) kern = r ? route_dataflow_kernel<false, false, true> : route_dataflow_kernel<false, false, false>;
    else if (!w) kern = r ? route_dataflow_kernel<true, false, true> : route_dataflow_kernel<true, false, false>;
    else kern = r ? route_dataflow_kernel<true, true, true> : route_dataflow_kernel<true, true, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsPerCta * 32, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    const long long tickets = (long long)a.n_tasks * a.n_mblocks * a.nsteps;
    long long want = (tickets + kWarpsPerCta - 1) / kWarpsPerCta;
    long long cap = (long long)num_sms * occ;
    const unsigned grid = (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
    kern<<<grid, kWarpsPerCta * 32, smem, st>>>(a);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_route_level(const LevelArgs& a, cudaStream_t st)
{
    const int half = (a.ld + 1) / 2;
    route_level_kernel<<<blocks_for((long long)a.count * half, 256), 256, 0, st>>>(a);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_init_inflows(const int32_t* up_off, const int32_t* up_pos, const uint8_t* is_outlet,
                                const double* O, double* I, int64_t n, int ld, int M, cudaStream_t st)
{
    (void)M;
    init_inflows_kernel<<<blocks_for(n * ((ld + 1) / 2), 256), 256, 0, st>>>(up_off, up_pos, is_outlet, O, I, n, ld);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_apply_gain(const int32_t* up_off, const int32_t* up_pos, const double* G, double* O,
                              double* I, int64_t n, int ld, int M, cudaStream_t st)
{
    (void)M;
    apply_gain_kernel<<<blocks_for(n * ((ld + 1) / 2), 256), 256, 0, st>>>(up_off, up_pos, G, O, I, n, ld);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_pack(const int32_t* reach_of_pos, const double* src, double* dst, int64_t n, int M,
                        int ld, int src_member_major, cudaStream_t st)
{
    pack_kernel<<<blocks_for(n * ld, 256), 256, 0, st>>>(reach_of_pos, src, dst, n, M, ld, src_member_major);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_unpack(const int32_t* reach_of_pos, const double* src, double* dst, int64_t n, int M,
                          int ld, int dst_member_major, cudaStream_t st)
{
    unpack_kernel<<<blocks_for(n * M, 256), 256, 0, st>>>(reach_of_pos, src, dst, n, M, ld, dst_member_major);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_gather_rows(const int32_t* pos, int64_t count, const double* X, int ld, int M,
                               double* out, cudaStream_t st)
{
    gather_rows_kernel<<<blocks_for(count * M, 256), 256, 0, st>>>(pos, count, X, ld, M, out);
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_permute_vec(const int32_t* reach_of_pos, const double* src, double* dst, int64_t n,
                               cudaStream_t st)
{
    permute_vec_kernel<<<blocks_for(n, 256), 256, 0, st>>>(reach_of_pos, src, dst, n);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace txh
