"""Quick device timing of txh_route_run (development aid, not the bench contract)."""
import argparse
import json
import sys
import os
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--M", type=int, default=64)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--sched", type=str, default="")
    ap.add_argument("--longchain", action="store_true")
    a = ap.parse_args()
    import torch
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork, Forcing
    net_d = S.make_longchain_network() if a.longchain else S.make_network(a.n, a.seed)
    n = net_d["endnodes"].size
    prm = S.make_params(n, a.seed)
    sp = [int(x) for x in a.sched.split(",")] if a.sched else None
    t = time.time()
    net = RiverNetwork(net_d["endnodes"], sp)
    tb = time.time() - t
    net.compute_coeffs(prm["K"], prm["X"], 300.0)
    info = net.schedule_info()
    t0 = 1_700_000_000 * 10**9
    times, table = S.make_forcing(n, a.steps, 300.0, a.seed, t0_ns=t0)
    mul = S.make_member_multipliers(times.size, a.M, a.seed) if a.M > 1 else None
    f = Forcing(net, times, table, mul)
    rng = np.random.default_rng(0)
    O = net.alloc_state(a.M); I = net.alloc_state(a.M)
    o0 = prm["o_t"][:, None] * rng.uniform(0.5, 1.5, size=(n, a.M))
    net.pack_host(o0, a.M, O)
    net.init_inflows(O, I, a.M)
    for _ in range(2):
        net.route_run(O, I, a.M, f, t0, int(300e9), a.steps)
    net.check()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    best = 1e30; tot = 0.0
    for _ in range(a.reps):
        ev0.record()
        net.route_run(O, I, a.M, f, t0, int(300e9), a.steps)
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1); best = min(best, ms); tot += ms
    net.check()
    upd = n * a.M * a.steps
    ab = 32 + 44.0 / a.M + (0 if mul is None else 0)
    out = {"n": n, "M": a.M, "steps": a.steps, "sched": info, "topo_build_s": round(tb, 4),
           "ms_best": round(best, 4), "ms_mean": round(tot / a.reps, 4),
           "us_per_step": round(1e3 * best / a.steps, 2),
           "updates_per_s": upd / (best * 1e-3), "GBps_algorithmic": upd * ab / (best * 1e-3) / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
