"""A split network with one dense KalmanFilter per sub-model (app/app.py:130-141) under AsyncSimulation: the filters of
the sub-models that are ready together batched into ONE chain of launches (txh_kfb_filter) vs one chain per sub-model.
Prints one JSON line per mode (wall time of the collection run, kernel launches, agreement of the two modes)."""
import asyncio
import json
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def build(n, seed, ncuts, T):
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.muskingum import Muskingum
    from tx_fast_hydrology_b200.da import KalmanFilter
    net = S.make_network(n, seed)
    prm = S.make_params(n, seed, well_posed=True)
    d = S.model_dict(net, prm, dt_s=3600.0)
    mdl = Muskingum(d)
    t0 = mdl.datetime.value
    rng = np.random.default_rng(seed)
    times = t0 + (np.arange(T, dtype=np.int64) + 1) * int(3600e9)
    idx = pd.DatetimeIndex(pd.to_datetime(times, unit="ns", utc=True)).as_unit("ns")
    df = pd.DataFrame(rng.gamma(0.5, 2.0, size=(T, n)), index=idx, columns=d["reach_ids"])
    # balanced sub-basins (HUC-like): walking upstream -> downstream, cut wherever the not-yet-cut subtree reaches `target`
    lev, _ = mdl.network.levels()
    end = net["endnodes"]
    open_sz = np.ones(n, dtype=np.int64)
    cuts = []
    target = max(8, n // ncuts)
    for j in np.argsort(lev, kind="stable"):
        e = end[j]
        if e == j:
            continue
        if open_sz[j] >= target:
            cuts.append(int(j))
        else:
            open_sz[e] += open_sz[j]
    mc = mdl.split(cuts)
    mt = t0 + np.arange(0, T + 1, dtype=np.int64) * int(3600e9)
    midx = pd.DatetimeIndex(pd.to_datetime(mt, unit="ns", utc=True)).as_unit("ns")
    for k, s in mc.models.items():
        if s.n < 8:
            continue
        m = min(6, max(1, s.n // 40))
        gl = np.sort(rng.choice(s.n, size=m, replace=False))
        mdf = pd.DataFrame(rng.uniform(0.5, 8.0, size=(mt.size, m)), index=midx, columns=[s.reach_ids[j] for j in gl])
        s.bind_callback(KalmanFilter(s, mdf, 2.0 * np.eye(s.n), 1e-2 * np.eye(m), 2.0 * np.eye(s.n)), key="kf")
    return mc, df


def main():
    import torch
    from tx_fast_hydrology_b200._lib import load
    from tx_fast_hydrology_b200.simulation import AsyncSimulation
    n, seed, ncuts, T = 6000, 17, 40, 12
    res = {}
    for batched in (True, False, True, False):          # the first pair warms up (module load, allocator); the second is reported
        mc, df = build(n, seed, ncuts, T)
        sim = AsyncSimulation(mc, df)
        sim.batch_filters = batched
        torch.cuda.synchronize()
        l0 = load().txh_launch_count(); t0 = time.perf_counter()
        out = asyncio.run(sim.simulate())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        sizes = sorted(m.n for m in mc.models.values())
        warm = batched not in res
        if warm:
            res[batched] = None
            continue
        res[batched] = out
        print(json.dumps({"batched": batched, "reaches": n, "sub_models": len(mc.models), "with_filter": sum(1 for m in mc.models.values() if m.callbacks),
                          "sub_model_sizes": [sizes[0], sizes[len(sizes) // 2], sizes[-1]], "steps": T, "wall_s": round(dt, 3),
                          "kernel_launches": int(load().txh_launch_count() - l0)}), flush=True)
    err = max(float(np.abs(res[True][k].values - res[False][k].values).max() / max(1e-300, np.abs(res[False][k].values).max())) for k in res[True])
    print(json.dumps({"max_rel_diff_batched_vs_per_model": err}))


if __name__ == "__main__":
    main()
