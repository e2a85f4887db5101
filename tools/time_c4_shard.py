"""Device time of a 24-hour deterministic run on one shard of the C4 forest (development aid): how the lane kernel's
launch time depends on the shard size (strong scaling of bench.py's c4_basins arm)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork, Forcing
    nsteps, seed = 288, 3
    for n, nb in ((2_700_000, 64), (1_350_000, 32), (675_000, 16), (337_500, 8)):
        net_d = S.make_network(n, seed, n_basins=nb)
        prm = S.make_params(n, seed)
        net = RiverNetwork(net_d["endnodes"])
        net.compute_coeffs(prm["K"], prm["X"], 300.0)
        t0 = 1_700_000_000 * 10**9
        times, table = S.make_forcing(n, nsteps, 300.0, seed, t0_ns=t0, rows_every=12)
        f = Forcing(net, times, table, None)
        O = net.alloc_state(1); I = net.alloc_state(1)
        net.pack_host(prm["o_t"][:, None].copy(), 1, O)
        net.init_inflows(O, I, 1)
        O0, I0 = O.clone(), I.clone()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ls = net.lane_schedule(1)
        reg = ls["regions"]
        res = {}
        for cold in (False, True):
            best, tot, reps = 1e30, 0.0, 5
            for it in range(reps + 2):
                if cold:
                    flush.fill_(1)
                O.copy_(O0); I.copy_(I0)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                net.route_run(O, I, 1, f, t0, int(300e9), nsteps)
                e1.record(); torch.cuda.synchronize()
                if it >= 2:
                    ms = e0.elapsed_time(e1); best = min(best, ms); tot += ms
            res["cold" if cold else "warm"] = (round(best, 3), round(tot / reps, 3))
        net.check()
        print(json.dumps({"reaches": n, "basins": nb, "regions": int(ls["n_regions"]), "max_real": int(ls["max_real"]),
                          "max_virt": int(ls["max_virt"]), "n_slots": int(ls["n_slots"]),
                          "max_height": int(reg[:, :].max(axis=0)[3]) if reg.shape[1] > 3 else None,
                          "ms_best_mean": res}))
        del net, f, O, I, O0, I0, flush
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
