// txh_window.cu -- window-resident routing kernel for sm_100a.
//
// route_window_kernel evaluates `nsteps` timesteps of the Muskingum update (tx_fast_hydrology/nutils.py:64-89,
// as called from muskingum.py:527-533) in ONE launch with the network state resident in shared memory:
// the network is cut into tasks of a few rows each (pockets = bundled side subtrees, segments = pieces of
// the long paths; txh_topology.cpp), a warp claims a task, loads its rows of O and I once, runs EVERY
// step of the launch on them in shared memory, and writes them back once.  HBM/L2 traffic per launch is
// one read and one write of the state plus the single rows tasks hand to each other, instead of a read
// and a write per step.
//
// Tasks exchange rows through a slot ring in global memory, ring[step][slot][ld]: a pocket publishes the
// outflow of its root, a segment the outflow of its last reach, per step.  A per-task progress word
// (steps published, release/acquire) orders them; tasks are claimed in a topological order, so a task
// only ever waits for tasks claimed before it, which are running or done -- every CTA is resident.
// Along a long path the step is the affine recurrence o_k = alpha_k (o_{k-1} + side_k) + r_k, so a
// segment first evaluates it with nothing entering (PRE), then needs ONE fused multiply-add per hop,
// out = A_last * o_in + B_last, before handing on -- consecutive steps ripple down the path as a
// systolic wavefront -- and fixes its interior rows up afterwards (FIX).
// Ensemble members are the SIMD axis: a lane owns two adjacent member columns (128-bit accesses).
#include "txh_kernels.cuh"

namespace txh {

void count_launch();

namespace {

__device__ __forceinline__ int ld_relaxed_s32(const int* p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double2 ld_row(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }
__device__ __forceinline__ void st_row(double* p, double2 v) { __stcg(reinterpret_cast<double2*>(p), v); }
__device__ __forceinline__ void cp_async4(unsigned sa, const void* g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(g) : "memory");
}
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
__device__ __forceinline__ double2 lds_row(unsigned sa)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(sa));
    return v;
}
__device__ __forceinline__ void sts_row(unsigned sa, double2 v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(sa), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double lds_f64(unsigned sa)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sa));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(unsigned sa)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa));
    return v;
}
__device__ __forceinline__ void sts_u32(unsigned sa, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sa), "r"(v) : "memory");
}

constexpr int kInRing = 8;                 // rows of other tasks in flight per warp (cp.async ring)
constexpr uint32_t kIdMask = 0x3fffffffu;

// Waits until every listed producer has published at least `need` steps; returns the smallest progress seen
// (>= need), or -1 when the launch is being abandoned (watchdog / poisoned handle).
__device__ __forceinline__ int wait_producers(const WinArgs& a, unsigned list_sa, int cnt, int mb, int need, int lane)
{
    unsigned spins = 0, nap = 32;
    unsigned long long t0 = 0;
    for (;;) {
        int mn = 0x7fffffff;
        for (int i = lane; i < cnt; i += 32) {
            const int p = (int)lds_u32(list_sa + 4u * i);
            mn = min(mn, ld_relaxed_s32(a.prog + (size_t)p * a.n_mblocks + mb));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if (mn >= need) { fence_acq_rel(); return mn; }          // acquire: the producers' rows are visible
        if ((++spins & 15u) == 0) {
            if (ld_relaxed_s32(a.status) != 0) return -1;
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > a.watchdog_ns) { if (lane == 0) atomicExch(a.status, 1); return -1; }
        }
        __nanosleep(nap);
        if (nap < 256u) nap <<= 1;
    }
}

struct FCtx {
    double w0, w1;
    double2 wm0, wm1;
};

template <bool HAS_F, bool HAS_W>
__device__ __forceinline__ double2 forcing_q(const FCtx& c, unsigned f0, unsigned f1, int r)
{
    double2 q = make_double2(0.0, 0.0);
    if (HAS_F) {
        const double a0 = lds_f64(f0 + 8u * r), a1 = lds_f64(f1 + 8u * r);
        if (HAS_W) { q.x = c.wm0.x * a0 + c.wm1.x * a1; q.y = c.wm0.y * a0 + c.wm1.y * a1; }
        else { q.x = c.w0 * a0 + c.w1 * a1; q.y = q.x; }
    }
    return q;
}

template <bool HAS_F, bool HAS_W>
__global__ void __launch_bounds__(256, 1)
route_window_kernel(const WinArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(smem_all + (size_t)warp * a.smem_per_warp);
    const unsigned sI = sb + lane * 16u, sO = sb + a.off_O + lane * 16u;
    const unsigned sScr = sb + a.off_scr + lane * 16u, sIn = sb + a.off_in + lane * 16u;
    const unsigned sCoef = sb + a.off_coef, sCum = sb + a.off_cum, sF0 = sb + a.off_f0, sF1 = sb + a.off_f1;
    const unsigned sHdr = sb + a.off_hdr, sWords = sb + a.off_words, sProd = sb + a.off_prod, sList = sb + a.off_list;
    const int nmb = a.n_mblocks, ld = a.ld;
    const long long total = (long long)a.n_tasks * nmb;

    for (;;) {
        long long t = 0;
        if (lane == 0) t = (long long)atomicAdd(a.ticket, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= total) break;
        if (ld_relaxed_s32(a.status) != 0) break;               // a launch on this handle was poisoned
        const int task = (int)(t / nmb);
        const int mb = (int)(t - (long long)task * nmb);
        const WTaskDesc td = a.tasks[task];
        const int len = td.len;
        const int col = mb * kMemberBlock + lane * 2;
        const bool active = col < ld;
        const int ccol = active ? col : 0;
        double* Og = a.O + (size_t)td.begin * ld + ccol;
        double* Ig = a.I + (size_t)td.begin * ld + ccol;

        // ---- load the task: its rows of I and O, per-row metadata, input stream, producer list ----
        if (active)
            for (int r = 0; r < len; ++r) {
                cp_async16(sI + r * 512u, Ig + (size_t)r * ld);
                cp_async16(sO + r * 512u, Og + (size_t)r * ld);
            }
        {
            const double* gc = a.coef + 4 * (size_t)td.begin;
            for (int i = lane; i < 2 * len; i += 32) cp_async16(sCoef + 16u * i, gc + 2 * i);
            for (int i = lane; i < len; i += 32) {
                cp_async8(sCum + 8u * i, a.cumA + td.begin + i);
                cp_async4(sHdr + 4u * i, a.hdr + td.begin + i);
            }
            for (int i = lane; i < td.n_words; i += 32) cp_async4(sWords + 4u * i, a.inw + td.in_off + i);
            for (int i = lane; i < td.n_prod; i += 32) cp_async4(sProd + 4u * i, a.prod + td.prod_off + i);
        }
        cp_async_commit();
        cp_async_wait_all();
        __syncwarp();
        // rows of other tasks consumed through the input ring, in consumption order
        int nL = 0;
        for (int base = 0; base < td.n_words; base += 32) {
            const int i = base + lane;
            const uint32_t w = i < td.n_words ? lds_u32(sWords + 4u * i) : 0u;
            const bool is = (w & WIN_SLOT) != 0u && i >= td.n_in;
            const unsigned m = __ballot_sync(0xffffffffu, is);
            if (is) sts_u32(sList + 4u * (nL + __popc(m & ((1u << lane) - 1u))), w & kIdMask);
            nL += __popc(m);
        }
        __syncwarp();
        const int n_side_prod = td.n_prod - td.n_in;
        int have = n_side_prod > 0 ? 0 : 0x7fffffff;            // side producers are known to have published `have` steps
        int have_in = td.n_in > 0 ? 0 : 0x7fffffff;
        int cur_r0 = -1, cur_r1 = -1;
        bool dead = false;

        for (int s = 0; s < a.nsteps && !dead; ++s) {
            double* ringS = a.ring + (size_t)s * a.n_slots * ld + ccol;
            FCtx fc;
            fc.w0 = fc.w1 = 0.0; fc.wm0 = fc.wm1 = make_double2(0.0, 0.0);
            if (HAS_F) {
                const StepInterp si = a.steps[s];
                if (si.r0 != cur_r0 || si.r1 != cur_r1) {      // a new bracket of the forcing table
                    const double* F0 = a.F + (size_t)si.r0 * a.n + td.begin;
                    const double* F1 = a.F + (size_t)si.r1 * a.n + td.begin;
                    __syncwarp();
                    for (int i = lane; i < len; i += 32) { cp_async8(sF0 + 8u * i, F0 + i); cp_async8(sF1 + 8u * i, F1 + i); }
                    cur_r0 = si.r0; cur_r1 = si.r1;
                }
                fc.w0 = si.w0; fc.w1 = si.w1;
                if (HAS_W) {
                    // member m sees (w0*mul[r0][m]) * F[r0] + (w1*mul[r1][m]) * F[r1]
                    const int c0 = min(col, a.wm_ld - 1), c1 = min(col + 1, a.wm_ld - 1);
                    const double* m0 = a.Wmul + (size_t)si.r0 * a.wm_ld;
                    const double* m1 = a.Wmul + (size_t)si.r1 * a.wm_ld;
                    fc.wm0 = make_double2(si.w0 * __ldg(m0 + c0), si.w0 * __ldg(m0 + c1));
                    fc.wm1 = make_double2(si.w1 * __ldg(m1 + c0), si.w1 * __ldg(m1 + c1));
                }
            }
            cp_async_commit();
            if (s >= have) {
                have = wait_producers(a, sProd + 4u * td.n_in, n_side_prod, mb, s + 1, lane);
                if (have < 0) { dead = true; break; }
            }
            // input ring prologue: the first kInRing rows of this step, one cp.async group per row
            int li = 0, ci = 0;
#pragma unroll
            for (int j = 0; j < kInRing; ++j) {
                if (li < nL) {
                    if (active) cp_async16(sIn + (li & (kInRing - 1)) * 512u, ringS + (size_t)lds_u32(sList + 4u * li) * ld);
                    ++li;
                }
                cp_async_commit();
            }
            cp_async_wait_group<kInRing>();                       // the forcing rows (older than the ring groups)
            __syncwarp();
            auto next_input = [&]() -> double2 {
                cp_async_wait_group<kInRing - 1>();
                const double2 v = lds_row(sIn + (ci & (kInRing - 1)) * 512u);
                ++ci;
                if (li < nL) {
                    if (active) cp_async16(sIn + (li & (kInRing - 1)) * 512u, ringS + (size_t)lds_u32(sList + 4u * li) * ld);
                    ++li;
                }
                cp_async_commit();
                return v;
            };

            if (td.kind == WTASK_POCKET) {
                int wi = 0;
                double2 acc = make_double2(0.0, 0.0);
                for (int r = 0; r < len; ++r) {
                    const uint32_t h = lds_u32(sHdr + 4u * r);
                    const double2 io = lds_row(sI + r * 512u), oo = lds_row(sO + r * 512u);
                    const double al = lds_f64(sCoef + 32u * r), be = lds_f64(sCoef + 32u * r + 8u);
                    const double ch = lds_f64(sCoef + 32u * r + 16u), ga = lds_f64(sCoef + 32u * r + 24u);
                    double2 inflow = (h & HDR_ACC) ? acc : make_double2(0.0, 0.0);
                    const int nin = (int)((h >> 6) & 0x1ffffffu);
                    for (int k = 0; k < nin; ++k) {
                        const uint32_t w = lds_u32(sWords + 4u * wi++);
                        double2 v;
                        if (w & WIN_SLOT) v = next_input();
                        else if (w & WIN_OWN) v = lds_row(sO + (w & kIdMask) * 512u);
                        else v = lds_row(sScr + w * 512u);
                        inflow.x += v.x; inflow.y += v.y;
                    }
                    const double2 q = forcing_q<HAS_F, HAS_W>(fc, sF0, sF1, r);
                    double2 on;
                    on.x = al * inflow.x + (be * io.x + ch * oo.x + ga * q.x);
                    on.y = al * inflow.y + (be * io.y + ch * oo.y + ga * q.y);
                    sts_row(sI + r * 512u, inflow);
                    sts_row(sO + r * 512u, on);
                    if (h & HDR_PUSH) {                              // read by another task: publish
                        const uint32_t slot = lds_u32(sWords + 4u * wi++);
                        if (active) st_row(ringS + (size_t)slot * ld, on);
                    }
                    const uint32_t sl = (h >> 1) & 31u;
                    if (sl) sts_row(sScr + (sl - 1) * 512u, on);
                    acc = on;
                }
                __syncwarp();
                if (lane == 0) st_release_s32(a.prog + t, s + 1);
            } else {
                // PRE: side_k = pocket roots joining reach k; r_k from the old state; B_k = the recurrence with
                // nothing entering the segment.  (side_k, B_k) replace (i_k, o_k) in shared memory.
                double2 B = make_double2(0.0, 0.0);
                for (int r = 0; r < len; ++r) {
                    const uint32_t h = lds_u32(sHdr + 4u * r);
                    const double2 io = lds_row(sI + r * 512u), oo = lds_row(sO + r * 512u);
                    const double al = lds_f64(sCoef + 32u * r), be = lds_f64(sCoef + 32u * r + 8u);
                    const double ch = lds_f64(sCoef + 32u * r + 16u), ga = lds_f64(sCoef + 32u * r + 24u);
                    const int nin = (int)((h >> 6) & 0x1fffu);
                    double2 side = make_double2(0.0, 0.0);
                    for (int k = 0; k < nin; ++k) { const double2 v = next_input(); side.x += v.x; side.y += v.y; }
                    const double2 q = forcing_q<HAS_F, HAS_W>(fc, sF0, sF1, r);
                    double2 inflow = side;
                    if (h & HDR_ACC) { inflow.x += B.x; inflow.y += B.y; }
                    B.x = al * inflow.x + (be * io.x + ch * oo.x + ga * q.x);
                    B.y = al * inflow.y + (be * io.y + ch * oo.y + ga * q.y);
                    sts_row(sI + r * 512u, side);
                    sts_row(sO + r * 512u, B);
                }
                // hop: out = A_last * o_in + B_last, o_in = the rows entering the segment
                double2 oin = make_double2(0.0, 0.0);
                if (td.n_in > 0) {
                    if (s >= have_in) {
                        have_in = wait_producers(a, sProd, td.n_in, mb, s + 1, lane);
                        if (have_in < 0) { dead = true; break; }
                    }
                    for (int k = 0; k < td.n_in; ++k) {
                        const double2 v = active ? ld_row(ringS + (size_t)(lds_u32(sWords + 4u * k) & kIdMask) * ld)
                                                 : make_double2(0.0, 0.0);
                        oin.x += v.x; oin.y += v.y;
                    }
                }
                const double Al = lds_f64(sCum + 8u * (len - 1));
                double2 out;
                out.x = Al * oin.x + B.x;
                out.y = Al * oin.y + B.y;
                if (active) st_row(ringS + (size_t)td.out_slot * ld, out);
                __syncwarp();
                if (lane == 0) st_release_s32(a.prog + t, s + 1);
                // FIX: o_k = B_k + A_k o_in, i_k = o_{k-1} + side_k
                double2 op = oin;
                for (int r = 0; r < len; ++r) {
                    const double2 side = lds_row(sI + r * 512u), Bk = lds_row(sO + r * 512u);
                    double2 on = out;
                    if (r + 1 < len) {
                        const double A = lds_f64(sCum + 8u * r);
                        on.x = A * oin.x + Bk.x;
                        on.y = A * oin.y + Bk.y;
                    }
                    double2 it;
                    it.x = op.x + side.x;
                    it.y = op.y + side.y;
                    sts_row(sO + r * 512u, on);
                    sts_row(sI + r * 512u, it);
                    op = on;
                }
            }
            cp_async_wait_all();
        }
        cp_async_wait_all();
        // ---- write the task's rows back ----
        if (active && !dead)
            for (int r = 0; r < len; ++r) {
                st_row(Ig + (size_t)r * ld, lds_row(sI + r * 512u));
                st_row(Og + (size_t)r * ld, lds_row(sO + r * 512u));
            }
        __syncwarp();
    }
}

// ticket, progress words and the forcing interpolation of the launch's steps (see dataflow_init_kernel)
__global__ void __launch_bounds__(256) window_init_kernel(const InitArgs a, int32_t* prog, long long n_prog,
                                                          unsigned long long* ticket)
{
    const long long gid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = gid0; g < n_prog; g += stride) prog[g] = 0;
    if (gid0 == 0) *ticket = 0ull;
    if (a.times) {
        for (long long s = gid0; s < a.nsteps; s += stride) {
            const double x = (double)(a.t0_ns + (a.step_base + s + 1) * a.dt_ns);
            int lo = 0, hi = a.R;                                   // np.searchsorted(xp, x), side='left'
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (a.times[mid] < x) lo = mid + 1; else hi = mid;
            }
            StepInterp si;
            if (lo == 0) { si.r0 = 0; si.r1 = 0; si.w0 = 1.0; si.w1 = 0.0; }
            else if (lo >= a.R) { si.r0 = a.R - 1; si.r1 = a.R - 1; si.w0 = 1.0; si.w1 = 0.0; }
            else {
                const double dx_0 = __dsub_rn(x, a.times[lo - 1]), dx_1 = __dsub_rn(a.times[lo], x);
                if (a.method == 1) {
                    const double frac = __ddiv_rn(dx_0, __dadd_rn(dx_0, dx_1));
                    si.r0 = lo - 1; si.r1 = lo; si.w0 = __dsub_rn(1.0, frac); si.w1 = frac;
                } else {
                    const int r = fabs(dx_0) <= fabs(dx_1) ? lo - 1 : lo;
                    si.r0 = r; si.r1 = r; si.w0 = 1.0; si.w1 = 0.0;
                }
            }
            a.steps_out[s] = si;
        }
    }
}

}  // namespace

cudaError_t launch_window_init(const InitArgs& a, int32_t* prog, long long n_prog, unsigned long long* ticket,
                               cudaStream_t st)
{
    long long blocks = (n_prog + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 1184) blocks = 1184;
    window_init_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, prog, n_prog, ticket);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_route_window(const WinArgs& a, int warps_per_cta, int num_sms, cudaStream_t st)
{
    const size_t smem = (size_t)warps_per_cta * a.smem_per_warp;
    void (*kern)(const WinArgs) = nullptr;
    const bool f = a.F != nullptr, w = a.Wmul != nullptr;
    if (!f) kern = route_window_kernel<false, false>;
    else if (!w) kern = route_window_kernel<true, false>;
    else kern = route_window_kernel<true, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const long long pairs = (long long)a.n_tasks * a.n_mblocks;
    long long want = (pairs + warps_per_cta - 1) / warps_per_cta;
    const unsigned grid = (unsigned)(want < num_sms ? (want < 1 ? 1 : want) : num_sms);   // one resident CTA per SM
    kern<<<grid, warps_per_cta * 32, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace txh
