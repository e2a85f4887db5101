// txh_lane.cu -- reach-parallel, time-skewed routing kernel for small ensembles (sm_100a).
//
// route_lane_kernel evaluates `nsteps` timesteps of the Muskingum update (tx_fast_hydrology/nutils.py:64-89,
// as called from muskingum.py:456 and the simulate loop muskingum.py:527-533) with LANES = REACHES.  It is the
// path for deterministic runs and small ensembles (M <= 16), where the member-per-lane window kernel
// (txh_window.cu) leaves 31 of 32 lanes idle.
//
// The network is cut into regions (LaneSchedule, txh_topology.cpp); a CTA claims a region and keeps its
// state in shared memory for every step of the launch.  Row j of a region evaluates timestep s in iteration
// k = s + off[j]; off decreases by one along every edge, so what a row needs from its upstream rows in
// iteration k was written in iteration k - 1: one __syncthreads per iteration, every row busy on its own
// timestep (time skewing: the 1,000-level dependency chain of a step costs nothing once the pipeline is
// full).  Per row and step:
//     inflow = sum of the upstream outflows               (shared-memory gathers, fixed child order)
//     o' = alpha*inflow + (p + gamma*q),   p' = beta*inflow + chi*o'      (p = beta*i + chi*o of the old state)
// Outflows crossing regions travel through streams ring[slot][member][step] in global memory: an unwritten
// cell holds EMPTY (all bits set); the consumer mirrors a stream as a virtual row, prefetches it 16 steps
// ahead with cp.async, polls a cell that is still EMPTY, and puts EMPTY back.  Regions are claimed in a
// topological order, so a region only waits for regions that are running or done (every CTA is resident).
// The forcing is interpolated per step exactly as nutils.py:21-34 does; its bracket rows live in shared
// memory and the next row is prefetched (cp.async) one bracket ahead.
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
__device__ __forceinline__ double ld_relaxed_f64(const double* p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void cp_async8(void* s, const void* g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async16(void* s, const void* g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// wait until at most `pending` of this thread's most recent groups are still in flight
__device__ __forceinline__ void cp_async_wait_pending(int pending)
{
    switch (pending) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
        case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    }
}
__device__ __forceinline__ bool is_empty(double v) { return __double_as_longlong(v) == -1ll; }
__device__ __forceinline__ double empty_cell() { return __longlong_as_double(-1ll); }

constexpr int kBatch = 16;                 // steps of a stream fetched per cp.async batch (128 bytes)
constexpr int kExtCells = 2 * kBatch;      // shared-memory window of a stream: this batch and the next

// Auxiliary word pair kept in the eighth double of a row's coefficient record.
struct RowAux { int32_t rec; int32_t issue; };

// A new forcing bracket for a row: (pr0, pr1) -> (r0, r1).  The rows of the old bracket and the prefetched
// row `pn` are in the record; anything else (irregular tables, nearest-row method jumps) is read in place.
__device__ __forceinline__ void rotate_bracket(double* cf, const double* __restrict__ Fcol, int64_t n, int R, int pr0, int pr1,
                                               int r0, int r1, int& groups_issued)
{
    RowAux aux = *reinterpret_cast<RowAux*>(cf + 7);
    // the prefetch this row issued at its previous rotation: groups committed since then may stay in flight
    cp_async_wait_pending(min(7, groups_issued - aux.issue - 1 < 0 ? 0 : groups_issued - aux.issue - 1));
    const double f0 = cf[4], f1 = cf[5], fn = cf[6];
    const int pn = min(max(pr0, pr1) + 1, R - 1);
    const double n0 = r0 == pr0 ? f0 : r0 == pr1 ? f1 : r0 == pn ? fn : __ldg(Fcol + (size_t)r0 * n);
    const double n1 = r1 == pr0 ? f0 : r1 == pr1 ? f1 : r1 == pn ? fn : __ldg(Fcol + (size_t)r1 * n);
    cf[4] = n0; cf[5] = n1;
    const int nn = min(max(r0, r1) + 1, R - 1);
    if (nn == r0) cf[6] = n0;
    else if (nn == r1) cf[6] = n1;
    else cp_async8(cf + 6, Fcol + (size_t)nn * n);
    cp_async_commit();
    aux.issue = groups_issued++;
    *reinterpret_cast<RowAux*>(cf + 7) = aux;
}

template <int MT, bool HAS_F, bool HAS_W>
__global__ void __launch_bounds__(1024, 1)
route_lane_kernel(const LaneArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_all[];
    __shared__ int sRegion;
    constexpr int MC = MT < 4 ? MT : 4;                         // members evaluated together (register tile)
    int4* sMeta = reinterpret_cast<int4*>(smem_all);
    double* sCoef = reinterpret_cast<double*>(smem_all + a.off_coef);     // [rr][8]
    double* sP = reinterpret_cast<double*>(smem_all + a.off_p);           // [MT][rr]
    double* sOb = reinterpret_cast<double*>(smem_all + a.off_obuf);       // [2][MT][rv]
    double* sExt = reinterpret_cast<double*>(smem_all + a.off_ext);       // [virt][MT][32]
    uint16_t* sChild = reinterpret_cast<uint16_t*>(smem_all + a.off_child);
    const int tid = threadIdx.x;
    const int TR = a.TR, TV = (int)blockDim.x - a.TR;
    const int rr = a.rr_stride, rv = a.rv_stride;
    const int nsteps = a.nsteps, M = a.M, ld = a.ld;
    const size_t splp = (size_t)a.splp;
    bool abandon = false;                                      // watchdog / poisoned handle: decided by a barrier vote

    for (;;) {
        if (tid == 0) {
            const long long t = (long long)atomicAdd(a.ticket, 1ull);
            sRegion = (t < a.n_regions && ld_relaxed_s32(a.status) == 0) ? (int)t : -1;
        }
        __syncthreads();
        const int reg = sRegion;
        if (reg < 0) break;
        const LaneRegionDesc rd = a.regions[reg];
        const int nr = rd.n_real, nv = rd.n_virt;
        // ---- load the region: row records, children, coefficients, p = beta*i + chi*o, forcing bracket of step 0 ----
        {
            const int4* gm = a.meta + rd.row_off;
            for (int i = tid; i < nr + nv + 1; i += blockDim.x) sMeta[i] = gm[i];
            const uint16_t* gc = a.child + rd.child_off;
            for (int i = tid; i < rd.n_child; i += blockDim.x) sChild[i] = gc[i];
        }
        int r0s = 0, r1s = 0, rns = 0;
        if (HAS_F) { r0s = __ldg(&a.steps[0].r0); r1s = __ldg(&a.steps[0].r1); rns = min(max(r0s, r1s) + 1, a.R - 1); }
        if (tid < TR) {
            for (int r = tid; r < nr; r += TR) {
                const int pos = a.meta[rd.row_off + r].x;
                const double2 ab = *reinterpret_cast<const double2*>(a.coef + 4 * (size_t)pos);
                const double2 cg = *reinterpret_cast<const double2*>(a.coef + 4 * (size_t)pos + 2);
                double* cf = sCoef + 8 * r;
                cf[0] = ab.x; cf[1] = ab.y; cf[2] = cg.x; cf[3] = cg.y;
                if (HAS_F) {
                    cf[4] = __ldg(a.F + (size_t)r0s * a.n + pos);
                    cf[5] = __ldg(a.F + (size_t)r1s * a.n + pos);
                    cf[6] = __ldg(a.F + (size_t)rns * a.n + pos);
                } else { cf[4] = cf[5] = cf[6] = 0.0; }
                RowAux aux; aux.rec = a.rec_slot ? a.rec_slot[pos] : -1; aux.issue = -1;
                *reinterpret_cast<RowAux*>(cf + 7) = aux;
                const double* og = a.O + (size_t)pos * ld;
                const double* ig = a.I + (size_t)pos * ld;
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    double p = 0.0;
                    if (m < M) p = ab.y * __ldcg(ig + m) + cg.x * __ldcg(og + m);
                    sP[m * rr + r] = p;
                }
            }
        } else {
            // the first batch of every incoming stream
            for (int v = tid - TR; v < nv; v += TV) {
                const int slot = a.meta[rd.row_off + nr + v].w;
                for (int m = 0; m < M; ++m) {
                    double* ex = sExt + ((size_t)(v * MT + m) << 5);
                    const double* g = a.ring + ((size_t)slot * M + m) * splp;
#pragma unroll
                    for (int q = 0; q < kBatch / 2; ++q) cp_async16(ex + 2 * q, g + 2 * q);
                }
            }
            cp_async_commit();
        }
        __syncthreads();

        const int niter = nsteps + rd.n_extra - 1;
        int groups_issued = 0;
        int dead = 0;
        for (int k = 0; k < niter; ++k) {
            const double* obp = sOb + (size_t)((k & 1) ^ 1) * MT * rv;
            double* obc = sOb + (size_t)(k & 1) * MT * rv;
            if (tid < TR) {
                for (int r = tid; r < nr; r += TR) {
                    const int4 mt = sMeta[r];
                    const int s = k - mt.y;
                    if ((unsigned)s >= (unsigned)nsteps) continue;
                    const int cb = mt.z, ce = sMeta[r + 1].z;
                    double* cf = sCoef + 8 * r;
                    double w0 = 0.0, w1 = 0.0;
                    int fr0 = 0, fr1 = 0;
                    if (HAS_F) {
                        const StepInterp si = a.steps[s];
                        fr0 = si.r0; fr1 = si.r1; w0 = si.w0; w1 = si.w1;
                        if (s > 0) {
                            const int pr0 = __ldg(&a.steps[s - 1].r0), pr1 = __ldg(&a.steps[s - 1].r1);
                            if (pr0 != fr0 || pr1 != fr1)
                                rotate_bracket(cf, a.F + mt.x, a.n, a.R, pr0, pr1, fr0, fr1, groups_issued);
                        }
                    }
                    const double2 ab = *reinterpret_cast<const double2*>(cf);
                    const double2 cg = *reinterpret_cast<const double2*>(cf + 2);
                    double q = 0.0, f0 = 0.0, f1 = 0.0;
                    if (HAS_F) {
                        const double2 ff = *reinterpret_cast<const double2*>(cf + 4);
                        f0 = ff.x; f1 = ff.y;
                        if (!HAS_W) q = cg.y * (w0 * f0 + w1 * f1);
                    }
                    const bool last = s + 1 == nsteps;
#pragma unroll
                    for (int m0 = 0; m0 < MT; m0 += MC) {
                        double infl[MC];
#pragma unroll
                        for (int u = 0; u < MC; ++u) infl[u] = 0.0;
                        for (int c = cb; c < ce; ++c) {
                            const int ci = sChild[c];
#pragma unroll
                            for (int u = 0; u < MC; ++u) infl[u] += obp[(m0 + u) * rv + ci];
                        }
#pragma unroll
                        for (int u = 0; u < MC; ++u) {
                            const int m = m0 + u;
                            double qm = q;
                            if (HAS_W) {
                                const int mc = min(m, a.wm_ld - 1);
                                qm = cg.y * (w0 * __ldg(a.Wmul + (size_t)fr0 * a.wm_ld + mc) * f0 +
                                             w1 * __ldg(a.Wmul + (size_t)fr1 * a.wm_ld + mc) * f1);
                            }
                            const double o = ab.x * infl[u] + (sP[m * rr + r] + qm);
                            sP[m * rr + r] = ab.y * infl[u] + cg.x * o;
                            obc[m * rv + r] = o;
                            if (m < M) {
                                if (mt.w >= 0) __stcg(a.ring + ((size_t)mt.w * M + m) * splp + s, o);
                                if (last) {
                                    __stcg(a.O + (size_t)mt.x * ld + m, o);
                                    __stcg(a.I + (size_t)mt.x * ld + m, infl[u]);
                                }
                            }
                        }
                    }
                    if (a.rec_slot) {
                        const int rec = reinterpret_cast<const RowAux*>(cf + 7)->rec;
                        const long long gs = a.rec_step_base + s + 1;
                        if (rec >= 0 && gs % a.rec_every == 0) {
                            double* out = a.rec_out + ((size_t)(gs / a.rec_every - 1) * a.rec_count + rec) * M;
                            for (int m = 0; m < M; ++m) out[m] = obc[m * rv + r];
                        }
                    }
                }
            } else {
                for (int v = tid - TR; v < nv; v += TV) {
                    const int4 mt = sMeta[nr + v];
                    const int s = k - mt.y;
                    if ((unsigned)s >= (unsigned)nsteps) continue;
                    if ((s & (kBatch - 1)) == 0) {
                        cp_async_wait_all();                                 // this batch (issued 16 iterations ago)
                        if (s + kBatch < nsteps) {
                            for (int m = 0; m < M; ++m) {
                                double* ex = sExt + ((size_t)(v * MT + m) << 5) + ((s + kBatch) & (kExtCells - 1));
                                const double* g = a.ring + ((size_t)mt.w * M + m) * splp + s + kBatch;
#pragma unroll
                                for (int q = 0; q < kBatch / 2; ++q) cp_async16(ex + 2 * q, g + 2 * q);
                            }
                        }
                        cp_async_commit();
                    }
                    for (int m = 0; m < M; ++m) {
                        double* cell = a.ring + ((size_t)mt.w * M + m) * splp + s;
                        double val = sExt[((size_t)(v * MT + m) << 5) + (s & (kExtCells - 1))];
                        if (is_empty(val)) {
                            // the prefetch came too early: poll the cell itself
                            unsigned spins = 0, nap = 32;
                            unsigned long long t0 = 0;
                            for (;;) {
                                val = ld_relaxed_f64(cell);
                                if (!is_empty(val)) break;
                                if ((++spins & 15u) == 0) {
                                    if (dead || ld_relaxed_s32(a.status) != 0) { dead = 1; val = 0.0; break; }
                                    const unsigned long long now = globaltimer_ns();
                                    if (t0 == 0) t0 = now;
                                    else if (now - t0 > a.watchdog_ns) { atomicExch(a.status, 1); dead = 1; val = 0.0; break; }
                                }
                                __nanosleep(nap);
                                if (nap < 256u) nap <<= 1;
                            }
                        }
                        __stcg(cell, empty_cell());
                        obc[m * rv + nr + v] = val;
                    }
                }
            }
            if (__syncthreads_or(dead)) { abandon = true; break; }   // also the barrier between two iterations
        }
        cp_async_wait_all();
        if (abandon) break;
    }
    // the last CTA to leave re-arms the ticket for the next launch on this handle
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(a.done, 1ull) + 1ull == (unsigned long long)gridDim.x) {
            *a.done = 0ull;
            *a.ticket = 0ull;
        }
    }
}

template <int MT>
cudaError_t launch_mt(const LaneArgs& a, int threads, size_t smem, int grid, cudaStream_t st)
{
    void (*kern)(const LaneArgs) = nullptr;
    const bool f = a.F != nullptr, w = a.Wmul != nullptr;
    if (!f) kern = route_lane_kernel<MT, false, false>;
    else if (!w) kern = route_lane_kernel<MT, true, false>;
    else kern = route_lane_kernel<MT, true, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, threads, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_route_lane(const LaneArgs& a, int mt, int threads, size_t smem, int grid, cudaStream_t st)
{
    switch (mt) {
        case 1: return launch_mt<1>(a, threads, smem, grid, st);
        case 2: return launch_mt<2>(a, threads, smem, grid, st);
        case 4: return launch_mt<4>(a, threads, smem, grid, st);
        case 8: return launch_mt<8>(a, threads, smem, grid, st);
        case 16: return launch_mt<16>(a, threads, smem, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace txh
