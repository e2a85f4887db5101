"""Sweep the schedule parameters of the dataflow routing kernel on one network (development aid).
Prints one line per setting: params, tasks, ms per window."""
import argparse
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--M", type=int, default=64)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--long", type=str, default="32")
    ap.add_argument("--spine", type=str, default="32")
    ap.add_argument("--pocket", type=str, default="48")
    ap.add_argument("--slots", type=str, default="12")
    ap.add_argument("--link", type=str, default="8")
    a = ap.parse_args()
    import torch
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork, Forcing
    net_d = S.make_network(a.n, a.seed)
    n = net_d["endnodes"].size
    prm = S.make_params(n, a.seed)
    t0 = 1_700_000_000 * 10**9
    times, table = S.make_forcing(n, a.steps, 300.0, a.seed, t0_ns=t0)
    mul = S.make_member_multipliers(times.size, a.M, a.seed) if a.M > 1 else None
    rng = np.random.default_rng(0)
    o0 = prm["o_t"][:, None] * rng.uniform(0.5, 1.5, size=(n, a.M))
    grid = [[int(x) for x in v.split(",")] for v in (a.long, a.spine, a.pocket, a.slots, a.link)]
    for sp in itertools.product(*grid):
        try:
            net = RiverNetwork(net_d["endnodes"], list(sp))
        except Exception as e:
            print(json.dumps({"sched": sp, "error": str(e)[:80]})); continue
        net.compute_coeffs(prm["K"], prm["X"], 300.0)
        info = net.schedule_info()
        f = Forcing(net, times, table, mul)
        O = net.alloc_state(a.M); I = net.alloc_state(a.M)
        net.pack_host(o0, a.M, O)
        net.init_inflows(O, I, a.M)
        for _ in range(2):
            net.route_run(O, I, a.M, f, t0, int(300e9), a.steps)
        net.check()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(a.reps):
            ev0.record()
            net.route_run(O, I, a.M, f, t0, int(300e9), a.steps)
            ev1.record(); torch.cuda.synchronize()
            best = min(best, ev0.elapsed_time(ev1))
        net.check()
        print(json.dumps({"sched": sp, "ms": round(best, 4), "tasks": info["n_tasks"], "spine": info["n_spine"],
                          "pocket": info["n_pocket"], "cp_tasks": info["cp_tasks"], "cp_cost": info["cp_cost"]}), flush=True)
        f.close(); net.close()
        del O, I


if __name__ == "__main__":
    main()
