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


constexpr int kPF = 4;      // row prefetch distance (reaches ahead) inside a task

struct StepCtx {
    double w0, w1;
    double2 wm0, wm1;
    const double* F0;
    const double* F1;
};

template <bool HAS_F, bool HAS_W>
__device__ __forceinline__ double2 forcing_q(const StepCtx& c, int k)
{
    double2 q = make_double2(0.0, 0.0);
    if (HAS_F) {
        const double f0 = __ldg(c.F0 + k), f1 = __ldg(c.F1 + k);
        if (HAS_W) { q.x = c.wm0.x * f0 + c.wm1.x * f1; q.y = c.wm0.y * f0 + c.wm1.y * f1; }
        else { q.x = c.w0 * f0 + c.w1 * f1; q.y = q.x; }
    }
    return q;
}

template <bool REC>
__device__ __forceinline__ void record(const RouteArgs& a, int k, int s, int col, double2 on)
{
    if (REC) {
        const int rs = a.rec_slot[k];
        if (rs >= 0 && ((s + 1) % a.rec_every) == 0 && col < a.M) {
            double* dst = a.rec_out + ((size_t)((s + 1) / a.rec_every - 1) * a.rec_count + rs) * a.M + col;
            dst[0] = on.x;
            if (col + 1 < a.M) dst[1] = on.y;
        }
    }
}

// POCKET: a bundle of side subtrees evaluated depth-first by one warp.  Intermediate
// outflows travel through the accumulator or this lane's scratch column.
template <bool HAS_F, bool HAS_W, bool REC>
__device__ __forceinline__ void run_pocket(const RouteArgs& a, const TaskDesc& td, int s, int col, bool active,
                                           double2* scratch, const StepCtx& sc)
{
    const int ld = a.ld, begin = td.begin, end = td.begin + td.len;
    const int ccol = active ? col : 0;
    double* Ob = a.O + ccol;
    double* Ib = a.I + ccol;
    const uint32_t* inp = a.inw + td.in_off;
    double2 acc = make_double2(0.0, 0.0);
    double2 pi[kPF], po[kPF];
#pragma unroll
    for (int j = 0; j < kPF; ++j) {
        pi[j] = po[j] = make_double2(0.0, 0.0);
        if (active && begin + j < end) { pi[j] = ld_row(Ib + (size_t)(begin + j) * ld); po[j] = ld_row(Ob + (size_t)(begin + j) * ld); }
    }
    for (int k0 = begin; k0 < end; k0 += kPF) {
#pragma unroll
        for (int j = 0; j < kPF; ++j) {
            const int k = k0 + j;
            if (k < end) {
                const double2 io = pi[j], oo = po[j];
                if (active && k + kPF < end) { pi[j] = ld_row(Ib + (size_t)(k + kPF) * ld); po[j] = ld_row(Ob + (size_t)(k + kPF) * ld); }
                const uint32_t h = __ldg(a.hdr + k);
                const double2 c01 = __ldg(reinterpret_cast<const double2*>(a.coef + 4 * (size_t)k));
                const double2 c23 = __ldg(reinterpret_cast<const double2*>(a.coef + 4 * (size_t)k) + 1);
                double2 inflow = (h & HDR_ACC) ? acc : make_double2(0.0, 0.0);
                const int nin = (int)(h >> 6);
                for (int t = 0; t < nin; ++t) {
                    const uint32_t w = __ldg(inp++);
                    double2 v;
                    if (w & INW_ROW) v = active ? ld_row(Ob + (size_t)(w & ~INW_ROW) * ld) : make_double2(0.0, 0.0);
                    else v = scratch[w * 32];
                    inflow.x += v.x; inflow.y += v.y;
                }
                const double2 q = forcing_q<HAS_F, HAS_W>(sc, k);
                double2 on;
                on.x = c01.x * inflow.x + (c01.y * io.x + c23.x * oo.x + c23.y * q.x);
                on.y = c01.x * inflow.y + (c01.y * io.y + c23.x * oo.y + c23.y * q.y);
                if (active) { st_row(Ib + (size_t)k * ld, inflow); st_row(Ob + (size_t)k * ld, on); }
                const uint32_t slot = (h >> 1) & 31u;
                if (slot) scratch[(slot - 1) * 32] = on;
                acc = on;
                record<REC>(a, k, s, col, on);
            }
        }
    }
}

// PRE: every reach of a spine segment independently -- gathers the side inflow from the pocket
// roots, folds the old state and the forcing into  b = alpha*side + beta*i_prev + chi*o_prev + gamma*q
// and parks (side, b) in the segment's own I / O rows for the CHAIN task.
template <bool HAS_F, bool HAS_W>
__device__ __forceinline__ void run_pre(const RouteArgs& a, const TaskDesc& td, int col, bool active,
                                        const StepCtx& sc)
{
    const int ld = a.ld, begin = td.begin, end = td.begin + td.len;
    const int ccol = active ? col : 0;
    double* Ob = a.O + ccol;
    double* Ib = a.I + ccol;
    const uint32_t* inp = a.inw + td.in_off;
    double2 pi[kPF], po[kPF];
#pragma unroll
    for (int j = 0; j < kPF; ++j) {
        pi[j] = po[j] = make_double2(0.0, 0.0);
        if (active && begin + j < end) { pi[j] = ld_row(Ib + (size_t)(begin + j) * ld); po[j] = ld_row(Ob + (size_t)(begin + j) * ld); }
    }
    for (int k0 = begin; k0 < end; k0 += kPF) {
#pragma unroll
        for (int j = 0; j < kPF; ++j) {
            const int k = k0 + j;
            if (k < end) {
                const double2 io = pi[j], oo = po[j];
                if (active && k + kPF < end) { pi[j] = ld_row(Ib + (size_t)(k + kPF) * ld); po[j] = ld_row(Ob + (size_t)(k + kPF) * ld); }
                const uint32_t h = __ldg(a.hdr + k);
                const double2 c01 = __ldg(reinterpret_cast<const double2*>(a.coef + 4 * (size_t)k));
                const double2 c23 = __ldg(reinterpret_cast<const double2*>(a.coef + 4 * (size_t)k) + 1);
                double2 side = make_double2(0.0, 0.0);
                const int nin = (int)((h >> 6) & 0x1fffu);
                for (int t = 0; t < nin; ++t) {
                    const uint32_t w = __ldg(inp++);
                    if (active) {
                        const double2 v = ld_row(Ob + (size_t)(w & ~INW_ROW) * ld);
                        side.x += v.x; side.y += v.y;
                    }
                }
                const double2 q = forcing_q<HAS_F, HAS_W>(sc, k);
                double2 b;
                b.x = c01.x * side.x + (c01.y * io.x + c23.x * oo.x + c23.y * q.x);
                b.y = c01.x * side.y + (c01.y * io.y + c23.x * oo.y + c23.y * q.y);
                if (active) { st_row(Ib + (size_t)k * ld, side); st_row(Ob + (size_t)k * ld, b); }
            }
        }
    }
}

// CHAIN: the first-order recurrence down the segment,  o_k = alpha_k * (o_{k-1} + late_k) + b_k,
// i_k = o_{k-1} + late_k + side_k, where late_k are outflows of other spine segments (the
// upstream segment of the same path, long tributaries).  One FMA per reach on the critical path.
template <bool REC>
__device__ __forceinline__ void run_chain(const RouteArgs& a, const TaskDesc& td, int s, int col, bool active)
{
    const int ld = a.ld, begin = td.begin, end = td.begin + td.len;
    const int ccol = active ? col : 0;
    double* Ob = a.O + ccol;
    double* Ib = a.I + ccol;
    const uint32_t* inp = a.inw + td.in_off;
    double2 o = make_double2(0.0, 0.0);
    double2 ps[kPF], pb[kPF];
#pragma unroll
    for (int j = 0; j < kPF; ++j) {
        ps[j] = pb[j] = make_double2(0.0, 0.0);
        if (active && begin + j < end) { ps[j] = ld_row(Ib + (size_t)(begin + j) * ld); pb[j] = ld_row(Ob + (size_t)(begin + j) * ld); }
    }
    for (int k0 = begin; k0 < end; k0 += kPF) {
#pragma unroll
        for (int j = 0; j < kPF; ++j) {
            const int k = k0 + j;
            if (k < end) {
                const double2 side = ps[j], b = pb[j];
                if (active && k + kPF < end) { ps[j] = ld_row(Ib + (size_t)(k + kPF) * ld); pb[j] = ld_row(Ob + (size_t)(k + kPF) * ld); }
                const uint32_t h = __ldg(a.hdr + k);
                const double al = __ldg(a.coef + 4 * (size_t)k);
                double2 inflow = (h & HDR_ACC) ? o : make_double2(0.0, 0.0);
                const int nlate = (int)(h >> 19);
                for (int t = 0; t < nlate; ++t) {
                    const uint32_t w = __ldg(inp++);
                    if (active) {
                        const double2 v = ld_row(Ob + (size_t)(w & ~INW_ROW) * ld);
                        inflow.x += v.x; inflow.y += v.y;
                    }
                }
                double2 on, it;
                on.x = al * inflow.x + b.x;
                on.y = al * inflow.y + b.y;
                it.x = inflow.x + side.x;
                it.y = inflow.y + side.y;
                if (active) { st_row(Ib + (size_t)k * ld, it); st_row(Ob + (size_t)k * ld, on); }
                o = on;
                record<REC>(a, k, s, col, on);
            }
        }
    }
}

// Prepares the dataflow runtime for one launch: dependency counters, step counters and the
// ready queue seeded with every task that has no same-step producer.
__global__ void __launch_bounds__(256) dataflow_init_kernel(const InitArgs a)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int pairs = a.n_tasks * a.n_mblocks;
    if (gid < pairs) {
        a.pending[gid] = a.tasks[gid / a.n_mblocks].need0;
        a.stepno[gid] = 0;
    }
    if (gid < a.n_init * a.n_mblocks) {
        const int t = a.init_ready[gid / a.n_mblocks];
        a.queue[gid] = (uint32_t)(t * a.n_mblocks + gid % a.n_mblocks) + 1u;
    }
    if (gid == 0) {
        a.q_head[0] = 0ull;
        a.q_head[1] = (unsigned long long)a.n_init * a.n_mblocks;
        a.q_head[2] = 0ull;     // completed tasks
    }
}

template <bool HAS_F, bool HAS_W, bool REC>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
route_dataflow_kernel(const RouteArgs a)
{
    extern __shared__ double2 scratch_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double2* scratch = scratch_all + (size_t)warp * a.slots * 32 + lane;   // this lane's column
    const int nmb = a.n_mblocks;
    unsigned long long* q_tail = a.q_head + 1;
    unsigned long long* n_done = a.q_head + 2;

    if (ld_relaxed(a.status) != 0) return;          // a previous launch on this handle was poisoned

    int next_pair = -1;                             // task made ready by this warp: run it without queueing
    for (;;) {
        int pair = next_pair;
        next_pair = -1;
        if (pair < 0) {
            // ---- pop: claim a queue position, wait until its producer has published it ----
            if (lane == 0) {
                const unsigned long long idx = atomicAdd(a.q_head, 1ull);
                pair = -2;
                if ((long long)idx < a.total) {
                    const uint32_t* slot = a.queue + idx;
                    unsigned spins = 0;
                    unsigned long long t0 = 0;
                    for (;;) {
                        uint32_t v;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
                        if (v != 0u) { pair = (int)(v - 1u); break; }
                        if ((++spins & 63u) == 0) {
                            unsigned long long fin;
                            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(fin) : "l"(n_done) : "memory");
                            if ((long long)fin >= a.total) break;           // everything ran (inline continuations skip the queue)
                            if (ld_relaxed(a.status) != 0) break;
                            const unsigned long long now = globaltimer_ns();
                            if (t0 == 0) t0 = now;
                            else if (now - t0 > a.watchdog_ns) { atomicExch(a.status, 1); break; }
                        }
                        __nanosleep(32);
                    }
                }
            }
            pair = __shfl_sync(0xffffffffu, pair, 0);
            if (pair < 0) break;
        }
        const int task = pair / nmb;
        const int mb = pair - task * nmb;
        const TaskDesc td = a.tasks[task];
        const int s = ld_relaxed(a.stepno + pair);  // written by the pair's previous run (another SM): L1 must not serve it

        const int col = mb * kMemberBlock + lane * 2;
        const bool active = col < a.ld;

        if (td.kind == TASK_CHAIN) {
            run_chain<REC>(a, td, s, col, active);
        } else {
            StepCtx sc;
            sc.w0 = sc.w1 = 0.0; sc.wm0 = sc.wm1 = make_double2(0.0, 0.0); sc.F0 = sc.F1 = nullptr;
            if (HAS_F) {
                const StepInterp si = a.steps[s];
                sc.F0 = a.F + (size_t)si.r0 * a.n;
                sc.F1 = a.F + (size_t)si.r1 * a.n;
                sc.w0 = si.w0; sc.w1 = si.w1;
                if (HAS_W) {
                    // member m sees  (w0*mul[r0][m])*F[r0] + (w1*mul[r1][m])*F[r1]
                    const int c0 = min(col, a.wm_ld - 1), c1 = min(col + 1, a.wm_ld - 1);
                    const double* m0 = a.Wmul + (size_t)si.r0 * a.wm_ld;
                    const double* m1 = a.Wmul + (size_t)si.r1 * a.wm_ld;
                    sc.wm0 = make_double2(si.w0 * __ldg(m0 + c0), si.w0 * __ldg(m0 + c1));
                    sc.wm1 = make_double2(si.w1 * __ldg(m1 + c0), si.w1 * __ldg(m1 + c1));
                }
            }
            if (td.kind == TASK_PRE) run_pre<HAS_F, HAS_W>(a, td, col, active, sc);
            else run_pocket<HAS_F, HAS_W, REC>(a, td, s, col, active, scratch, sc);
        }

        // ---- completion: re-arm, publish, notify dependants --------------------------------
        __syncwarp();
        const bool more = s + 1 < a.nsteps;
        if (lane == 0) {
            a.stepno[pair] = s + 1;
            if (more) a.pending[pair] = td.need;    // nobody can signal step s+1 before the notifications below
            __threadfence();
        }
        __syncwarp();
        const int nn = td.n_same + (more ? td.n_next : 0);
        for (int d0 = 0; d0 < nn; d0 += 32) {
            const int d = d0 + lane;
            int tgt = -1;
            bool ready = false;
            if (d < nn) {
                tgt = a.notify[td.nfy_off + d] * nmb + mb;
                ready = atomicSub(a.pending + tgt, 1) == 1;
                if (ready) __threadfence();
            }
            unsigned mask = __ballot_sync(0xffffffffu, ready);
            if (mask != 0u && next_pair < 0) {
                // keep one ready dependant for this warp (a CHAIN if there is one): no queue round trip
                const bool is_chain = ready && a.tasks[tgt / nmb].kind == TASK_CHAIN;
                const unsigned cm = __ballot_sync(0xffffffffu, is_chain);
                const int keep = __ffs(cm ? cm : mask) - 1;
                next_pair = __shfl_sync(0xffffffffu, tgt, keep);
                if (lane == keep) ready = false;
            }
            if (ready) {
                const unsigned long long idx = atomicAdd(q_tail, 1ull);
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.queue + idx), "r"((uint32_t)tgt + 1u) : "memory");
            }
        }
        if (lane == 0) atomicAdd(n_done, 1ull);
        __syncwarp();
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

cudaError_t launch_dataflow_init(const InitArgs& a, cudaStream_t st)
{
    const int pairs = a.n_tasks * a.n_mblocks;
    dataflow_init_kernel<<<blocks_for(pairs, 256), 256, 0, st>>>(a);
    g_launches++;
    return cudaGetLastError();
}

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
    const long long pairs = (long long)a.n_tasks * a.n_mblocks;
    long long want = (pairs + kWarpsPerCta - 1) / kWarpsPerCta;       // one step's worth of warps is plenty
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
