"""Warm per-phase device times of one assimilation window (routing launch, stats, solve, apply) with CUDA
events, in the order the benchmark runs them (development aid)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork, Forcing
    n, M, m, seed, every, reps = 100000, 64, 500, 2, 12, 40
    net_d = S.make_network(n, seed); prm = S.make_params(n, seed)
    net = RiverNetwork(net_d["endnodes"]); net.compute_coeffs(prm["K"], prm["X"], 300.0)
    t0 = 1_700_000_000 * 10**9
    times, table = S.make_forcing(n, every, 300.0, seed, t0_ns=t0)
    mul = S.make_member_multipliers(times.size, M, seed)
    f = Forcing(net, times, table, mul)
    rng = np.random.default_rng(0)
    o0 = prm["o_t"][:, None] * rng.uniform(0.5, 1.5, size=(n, M))
    O = net.alloc_state(M); I = net.alloc_state(M); G = net.alloc_state(M)
    O0 = net.alloc_state(M); net.pack_host(o0, M, O0); I0 = net.alloc_state(M); net.init_inflows(O0, I0, M)
    gidx = S.make_gauges(net_d["endnodes"], m, seed=4)
    f64 = dict(dtype=torch.float64, device="cuda")
    rowsum = torch.empty(n, **f64); HX = torch.empty((m, M), **f64)
    work = torch.empty(net.enkf_work_size(m, M), **f64); W = torch.empty((m, M), **f64); T = torch.empty((M, M), **f64)
    qs = torch.full((m,), 2.0, **f64); R = torch.eye(m, **f64) * 1e-2; Dinv = torch.full((m,), 1.0 / 2.01, **f64)
    Zp = torch.as_tensor(prm["o_t"][gidx][:, None] + 0.1 * rng.standard_normal((m, M)), device="cuda")
    net.set_stats_output(rowsum, 1.0 / M)
    names = ["route_window", "stats(gather)", "solve", "apply", "whole"]
    acc = np.zeros(len(names))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    for it in range(reps + 3):
        O.copy_(O0); I.copy_(I0)
        ev[0].record()
        net.route_run(O, I, M, f, t0, int(300e9), every)
        ev[1].record()
        net.enkf_stats(O, M, gidx, None, HX)
        ev[2].record()
        net.enkf_solve(m, M, HX, Zp, rowsum, gidx, qs, R, work, W, T, Dinv, 1)
        ev[3].record()
        net.enkf_apply(O, I, M, None, 0, M, 0, rowsum, T, gidx, qs, W, G)
        ev[4].record()
        torch.cuda.synchronize()
        if it >= 3:
            for k in range(4):
                acc[k] += ev[k].elapsed_time(ev[k + 1])
            acc[4] += ev[0].elapsed_time(ev[4])
    net.check()
    print(json.dumps({k: round(1e3 * v / reps, 1) for k, v in zip(names, acc)}), "(us)")


if __name__ == "__main__":
    main()
