#!/usr/bin/env python
"""bench.py -- headline benchmark of the routing + assimilation hot path (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[2] ("C3"): the synthetic Texas-scale network (100,000 reaches,
1,000 topological levels, seed 2), a 64-member ensemble PER GPU, a 7-day run at a 5-minute step
(2,016 routing steps) with an EnKF assimilation of 500 synthetic gauges every hour (168 updates).
One bench "step" is one whole 7-day run.  The metric is reach*timestep*member updates per second.

  value : device-resident run (inputs already in HBM), CUDA events, max over ranks
  e2e   : the same run through the public Python API from PINNED HOST buffers: initial ensemble,
          forcing table, member multipliers and observations copied host->device and the final
          ensemble copied device->host inside the timed region
  roofline : route_window_kernel (the window-resident routing kernel, one launch per hourly window),
          algorithmic bytes per launch / its mean CUDA-event duration / measured HBM copy bandwidth
  cpu_baseline : the CPU oracle (C/OpenMP restatement of the reference kernels, members over host
          threads + numpy EnKF) on a bounded sample of the same workload, rank 0 at N=1 only

--impl reference times the oracle port alone (the reference is numba/Python and cannot travel to
the GPU box; its kernels are restated in oracle/txh_oracle.c and pinned against it by the goldens).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "reach*timestep*member updates/sec"
UNIT = "updates/s"
DT_S = 300.0
T0 = "2024-01-01T00:00:00Z"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # workload knobs (defaults = the configuration the metric is quoted on)
    ap.add_argument("--reaches", type=int, default=100000)
    ap.add_argument("--members", type=int, default=64, help="ensemble members per GPU")
    ap.add_argument("--days", type=float, default=7.0)
    ap.add_argument("--gauges", type=int, default=500)
    ap.add_argument("--assim-every", type=int, default=12, help="routing steps between EnKF updates")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--cpu-windows", type=int, default=3, help="hourly windows in the CPU sample")
    ap.add_argument("--sharding", default="basins", choices=["basins", "members"],
                    help="N > 1: one independent basin (network + ensemble + gauges) per GPU, no collective; or the "
                         "members of ONE network's ensemble over the GPUs, EnKF statistics combined with NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def algorithmic_bytes_per_update(M):
    """read o_prev, i_prev, write o_next, i_next (32 B) + per reach*step terms
    shared by the members: 4 coefficients (32 B), forcing (8 B), one index (4 B)."""
    return 32.0 + 44.0 / M


# ------------------------------------------------------------------------------------------------
# workload (identical numbers for the GPU arm and the CPU arm)
# ------------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, a, rank=0, world=1):
        from tx_fast_hydrology_b200 import synthetic as S
        self.a = a
        self.n = n = a.reaches
        self.M = M = a.members
        self.Mtot = M * world
        self.rank, self.world = rank, world
        self.by_basin = a.sharding == "basins" and world > 1
        if self.by_basin:
            # basin sharding: every rank owns an independent basin of the same size with its own ensemble
            # and gauges (the reference's app runs one Kalman filter per sub-basin model, app.py:130-141)
            self.Mtot = M
        seed = a.seed + (1000 * rank if self.by_basin else 0)
        self.seed = seed
        self.nsteps = int(round(a.days * 86400.0 / DT_S))
        self.every = a.assim_every
        self.nwin = self.nsteps // self.every
        self.net = S.make_network(n, seed)
        self.params = S.make_params(n, seed)
        import pandas as pd
        self.t0_ns = int(pd.Timestamp(T0).value)
        self.times, self.table = S.make_forcing(n, self.nsteps, DT_S, seed, t0_ns=self.t0_ns, rows_every=12)
        mul_all = S.make_member_multipliers(self.times.size, self.Mtot, seed)
        c0 = 0 if self.by_basin else rank * M
        self.mul = np.ascontiguousarray(mul_all[:, c0:c0 + M])
        rng = np.random.default_rng(seed + 7)
        spread = rng.uniform(0.5, 1.5, size=(n, self.Mtot))
        self.o0 = np.ascontiguousarray(self.params["o_t"][:, None] * spread[:, c0:c0 + M])
        self.gauges = S.make_gauges(self.net["endnodes"], a.gauges, seed=4)
        self.m = self.gauges.size
        self.R = 1e-2 * np.eye(self.m)             # app.py:137-138 values
        self.Q = 2.0
        self.model_dict = S.model_dict(self.net, self.params, dt_s=DT_S, t0=T0)
        self.Zp = None

    def make_observations(self, truth):
        """truth [nwin][m] -> per-member perturbed observations [nwin][m][Mtot] (all ranks identical)."""
        rng = np.random.default_rng(self.seed + 11)
        obs = truth + 0.1 * rng.standard_normal(truth.shape)                 # N(0, R), R = 1e-2 I
        noise = 0.1 * rng.standard_normal((truth.shape[0], self.m, self.Mtot))
        self.obs = obs
        self.Zp = np.ascontiguousarray(obs[:, :, None] + noise)
        return self.Zp

    def updates_per_run(self):
        """whole job: all ranks"""
        return float(self.n) * self.M * self.world * self.nsteps


# ------------------------------------------------------------------------------------------------
# clocks (pynvml sampling thread; nvidia-smi is the fallback)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, device_index):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self._nv = pynvml
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self._h = h
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for k, bit in {**self.BAD, **self.NOTE}.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self._h is not None:
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._h is not None and self._t.is_alive():
            self._t.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
class CpuSample:
    """`windows` hourly windows of the workload on the host: members routed over OpenMP threads
    (oracle.run_members, nutils.py:64-89 per member), then the ensemble update (oracle.enkf_update,
    da.py:112-126 on the sample covariance) in numpy/LAPACK."""

    def __init__(self, wl, windows):
        from oracle import oracle as O
        O.build()
        self.O, self.wl, self.windows = O, wl, windows
        net = wl.net
        ind = O.compute_indegree(net["startnodes"], net["endnodes"])
        al, be, ch, ga = O.compute_coeffs(wl.params["K"], wl.params["X"], DT_S)
        self.net = {"startnodes": net["startnodes"], "endnodes": net["endnodes"], "indegree": ind,
                    "alpha": al, "beta": be, "chi": ch, "gamma": ga}
        self.threads = O.max_threads()
        self.o_init = np.ascontiguousarray(wl.o0.T)                      # [M][n]
        self.i_init = np.stack([O.init_states(net["startnodes"], net["endnodes"], o) for o in self.o_init])
        self.q_diag = np.full(wl.n, wl.Q)
        if wl.Zp is None:
            # the CPU arm alone: observations around the initial state (timing does not depend on values)
            rng = np.random.default_rng(5)
            self.Zp = wl.params["o_t"][wl.gauges][None, :, None] + 0.1 * rng.standard_normal(
                (windows, wl.m, wl.M))
        else:
            self.Zp = wl.Zp[:windows, :, :wl.M]
        self.updates = float(wl.n) * wl.M * wl.every * windows

    def run(self):
        O, wl = self.O, self.wl
        o = self.o_init.copy(); i = self.i_init.copy()
        xp = wl.times.astype(np.float64)
        step_ns = DT_S * 1e9
        t = float(wl.t0_ns)
        t0 = time.perf_counter()
        for k in range(self.windows):
            O.run_members(self.net, o, i, wl.every, xp, wl.table, t, step_ns, wmul=wl.mul, threads=self.threads)
            t += wl.every * step_ns
            Op, Ip, _ = O.enkf_update(self.net, o.T, i.T, wl.gauges, self.Zp[k], self.q_diag, wl.R)
            o = np.ascontiguousarray(Op.T); i = np.ascontiguousarray(Ip.T)
        return time.perf_counter() - t0, o

    def describe(self):
        wl = self.wl
        return (f"{self.windows} hourly windows of the workload: {wl.every * self.windows} routing steps x "
                f"{wl.M} members x {wl.n} reaches + {self.windows} EnKF updates of {wl.m} gauges; "
                f"oracle/txh_oracle.c over {self.threads} OpenMP threads + numpy")


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = Workload(a)
    cs = CpuSample(wl, a.cpu_windows)
    for _ in range(max(1, a.warmup)):
        cs.run()
    tot = 0.0
    for _ in range(a.steps):
        dt, _ = cs.run()
        tot += dt
    val = cs.updates * a.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * tot / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(wl, a, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cs.threads, "kind": "port", "sample": cs.describe()},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(wl, a, world):
    return {
        "workload": "C3: synthetic Texas-scale network, 64-member ensemble per GPU, 7-day run at a 5-min step, "
                    "hourly EnKF of 500 gauges (BASELINE.json configs[2])",
        "reaches": wl.n * (world if wl.by_basin else 1), "reaches_per_gpu": wl.n, "levels": 1000,
        "members_per_gpu": wl.M, "members_per_ensemble": wl.Mtot,
        "routing_steps": wl.nsteps, "dt_s": DT_S, "gauges": wl.m, "enkf_updates": wl.nwin,
        "assimilate_every_steps": wl.every, "seed": a.seed,
        "sharding": ("single GPU" if world == 1 else
                     "independent basins, one per GPU, each with its own ensemble, gauges and EnKF; no collective"
                     if wl.by_basin else
                     "ensemble members of one network over ranks (network replicated); EnKF statistics combined with NCCL"),
        "l2": "256 MiB buffer written between bench steps (inside the timed region); per-run inputs "
              "(forcing 135 MB + observations + 102 MB state) exceed the 126 MB L2",
    }


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def gpu_arm(a):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tx_fast_hydrology_b200 import build, _lib
    build.build_lib()
    lib = _lib.load()
    import pandas as pd
    from tx_fast_hydrology_b200.muskingum import Muskingum
    from tx_fast_hydrology_b200.da import EnsembleKalmanFilter

    wl = Workload(a, rank, world)
    n, M, m, nsteps, every, nwin = wl.n, wl.M, wl.m, wl.nsteps, wl.every, wl.nwin
    dev = torch.device("cuda", local)

    # ---- truth run (setup, untimed): one deterministic member, gauge outflows every hour -------
    truth_mdl = Muskingum(dict(wl.model_dict), members=1)
    f1 = truth_mdl.make_forcing(times_ns=wl.times, table=wl.table)
    rec = truth_mdl.run(f1, nwin * every, record_reaches=wl.gauges, record_every=every)
    truth_mdl.network.check()
    truth = rec.cpu().numpy()[:, :, 0]
    f1.close()
    del truth_mdl
    Zp_host = wl.make_observations(truth)

    # ---- the ensemble model + filter ---------------------------------------------------------------
    d = dict(wl.model_dict)
    d["o_t"] = wl.o0
    mdl = Muskingum(d, members=M)
    idx = pd.DatetimeIndex(pd.to_datetime(wl.t0_ns + (np.arange(nwin, dtype=np.int64) + 1) * int(every * DT_S * 1e9),
                                          unit="ns", utc=True)).as_unit("ns")
    meas = pd.DataFrame(wl.obs, index=idx, columns=[d["reach_ids"][j] for j in wl.gauges])
    enkf = EnsembleKalmanFilter(mdl, meas, wl.Q, wl.R, every=pd.Timedelta(seconds=every * DT_S),
                                group=False if (wl.by_basin or world == 1) else None)
    assert enkf.Mtot == wl.Mtot

    # pinned host inputs / outputs of the end-to-end arm
    def pinned(x):
        t = torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        return t
    o0_pin = pinned(wl.o0)
    table_pin = pinned(wl.table)
    mul_pin = pinned(wl.mul)
    Zp_pin = pinned(Zp_host)
    out_pin = torch.empty((n, M), dtype=torch.float64).pin_memory()
    Zp_dev = torch.empty((nwin, m, wl.Mtot), dtype=torch.float64, device=dev)

    # device-resident inputs of the `value` arm
    forcing = mdl.make_forcing(times_ns=wl.times, table=table_pin, member_mul=mul_pin)
    mdl.upload_state(o0_pin)
    O, I = mdl.device_state
    O0, I0 = O.clone(), I.clone()
    Zp_dev.copy_(Zp_pin)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    t_start = pd.Timestamp(T0)

    def resident_run(timers=None):
        flush.fill_(1)                                   # L2 flush between bench steps
        O.copy_(O0); I.copy_(I0)
        mdl._datetime = t_start
        mdl.run_assimilating(forcing, nsteps, enkf, every, Zp_dev, timers=timers)

    e2e_parts = {}
    copy_stream = torch.cuda.Stream()
    obs_ready = torch.cuda.Event()

    def e2e_run():
        def lap(key, t0, sync=True):
            if sync:
                torch.cuda.synchronize()
            t1 = time.perf_counter()
            e2e_parts[key] = e2e_parts.get(key, 0.0) + (t1 - t0)
            return t1
        t = time.perf_counter()
        mdl.upload_state(o0_pin)                                                     # H2D
        t = lap("upload_state_ms", t)
        # H2D into the resident table, in row chunks on the forcing's copy stream: the routing launches wait only for
        # the hours they read, so the upload overlaps the first windows (its time shows up inside run_ms)
        forcing.update(wl.times, table_pin, mul_pin, overlap=True)
        f = forcing
        t = lap("forcing_enqueue_ms", t, sync=False)
        with torch.cuda.stream(copy_stream):                                         # H2D beside the first window
            Zp_dev.copy_(Zp_pin, non_blocking=True)
            obs_ready.record(copy_stream)
        t = lap("observations_enqueue_ms", t, sync=False)
        mdl._datetime = t_start
        mdl.run_assimilating(f, nsteps, enkf, every, Zp_dev, observations_ready=obs_ready)
        t = lap("run_ms", t)
        mdl.download_state(out_o=out_pin)                                            # D2H (synchronises)
        lap("download_ms", t)

    h2d = o0_pin.numel() * 8 + table_pin.numel() * 8 + mul_pin.numel() * 8 + Zp_pin.numel() * 8
    d2h = out_pin.numel() * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up ---------------------------------------------------------------------------------------
    for _ in range(a.warmup):
        resident_run()
    mdl.network.check()

    # ---- timed: K device-resident runs ---------------------------------------------------------------
    clocks = ClockSampler(local)
    timers = []
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    launches0 = int(lib.txh_launch_count())
    barrier()
    clocks.start()
    wall0 = time.perf_counter()
    ev0.record()
    for _ in range(a.steps):
        resident_run(timers)
    ev1.record()
    barrier()
    wall = time.perf_counter() - wall0
    clk = clocks.stop()
    launches = int(lib.txh_launch_count()) - launches0
    mdl.network.check()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / a.steps
    value = wl.updates_per_run() / (ms_step * 1e-3)
    final_o = mdl.download_state()[0]
    if not np.isfinite(final_o).all():
        raise SystemExit("bench.py: non-finite ensemble state after the run")

    # routing kernel roofline: mean CUDA-event duration of one launch (one hourly window)
    # (an unsharded ensemble runs inside the library, which keeps the event pairs; the member-sharded loop hands
    # back torch events)
    k_list = [e0.elapsed_time(e1) for e0, e1 in timers] if timers else list(mdl.network.route_timings())
    k_ms = float(np.mean(k_list)) if k_list else float("nan")
    ab = algorithmic_bytes_per_update(M)
    bytes_per_launch = float(n) * M * every * ab
    achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fpk:
            peak = float(json.load(fpk)["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as ft:
            traffic = json.load(ft).get("route_window_kernel_dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "route_window_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms_per_launch": k_ms,
                "launches_timed": len(k_list), "bytes_per_update": ab,
                "routing_share_of_step": k_ms * nwin / ms_step if ms_step > 0 else None}

    # ---- end to end through the public API, host buffers in and out --------------------------------
    e2e = None
    if not a.no_e2e:
        e2e_run()                                         # warm-up (allocations)
        e2e_parts.clear()
        barrier()
        t0 = time.perf_counter()
        reps = max(1, min(a.steps, 3))
        for _ in range(reps):
            e2e_run()
        barrier()
        s_e2e = max_over_ranks((time.perf_counter() - t0) / reps)
        e2e = {"value": wl.updates_per_run() / s_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * s_e2e, "runs": reps,
               "breakdown_ms": {k: round(1e3 * v / reps, 3) for k, v in e2e_parts.items()}}

    # ---- CPU baseline (rank 0, single GPU only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cs = CpuSample(wl, a.cpu_windows)
        cs.run()                                          # warm-up (page faults, OpenMP pool)
        best = min(cs.run()[0] for _ in range(2))
        cpu = {"value": cs.updates / best, "unit": UNIT, "cores": cs.threads, "kind": "port",
               "sample": cs.describe()}
        # the same on ONE core, one window (SURVEY.md section 8d asks for both)
        c1 = CpuSample(wl, 1)
        c1.threads = 1
        cpu["value_1core"] = c1.updates / c1.run()[0]

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(wl, a, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clk, "wall_s_timed_region": wall,
            "roofline_whole_step": {"achieved": value * ab / 1e9, "unit": "GB/s", "frac": value * ab / 1e9 / peak / world},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return reference_arm(a)
    return gpu_arm(a)


if __name__ == "__main__":
    sys.exit(main())
