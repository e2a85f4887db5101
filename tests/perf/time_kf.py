"""Per-step time of the dense per-sub-basin Kalman filter (the reference's operational path, app/app.py:130-141:
one `KalmanFilter` per sub-model, `filter()` after every step) -- the CUDA drop-in against the CPU oracle
(oracle/txh_oracle.c `_aqat_par` on all host threads + numpy/LAPACK), same inputs, parity printed.

    python tools/time_kf.py [n ...]
"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def frame(times_ns, table, cols):
    idx = pd.DatetimeIndex(pd.to_datetime(times_ns, unit="ns", utc=True)).as_unit("ns")
    return pd.DataFrame(table, index=idx, columns=cols)


def one(n, steps=24, seed=11, cpu=True):
    import torch
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.muskingum import Muskingum
    from tx_fast_hydrology_b200.da import KalmanFilter
    from oracle import oracle
    m = max(4, n // 50)
    net_d = S.make_network(n, seed)
    prm = S.make_params(n, seed, well_posed=True)
    d = S.model_dict(net_d, prm, dt_s=300.0)
    rng = np.random.default_rng(seed)
    t0 = pd.Timestamp(d["datetime"]).value
    times, table = S.make_forcing(n, steps, 300.0, seed, t0_ns=t0)
    gidx = np.sort(S.make_gauges(net_d["endnodes"], m, seed=4))
    mt = t0 + np.arange(steps // 12 + 2, dtype=np.int64) * int(3600e9)
    meas = prm["o_t"][gidx][None, :] * rng.uniform(0.8, 1.2, size=(mt.size, m))
    Q = 2.0 * np.eye(n); R = 1e-2 * np.eye(m); P0 = 2.0 * np.eye(n)       # app/app.py:137-139
    out = {"n": n, "m": m, "steps": steps}
    # --- CUDA drop-in
    mdl = Muskingum(d)
    kf = KalmanFilter(mdl, frame(mt, meas, [d["reach_ids"][j] for j in gidx]), Q, R, P0)
    mdl.bind_callback(kf, key="kf")
    df = frame(times, table, d["reach_ids"])
    end = pd.Timestamp(t0 + steps * int(300e9), tz="UTC")
    it = mdl.simulate(df, end_time=end)
    next(it); next(it)                                                    # warm-up: start filter + two steps
    torch.cuda.synchronize(); t = time.perf_counter()
    k = 0
    for _ in it:
        k += 1
    torch.cuda.synchronize()
    out["gpu_ms_per_step"] = 1e3 * (time.perf_counter() - t) / k
    o_gpu = mdl.o_t_next.copy(); P_gpu = kf.P_t_next
    # --- CPU oracle
    if cpu:
        om = oracle.OracleModel(net_d["startnodes"], net_d["endnodes"], prm["K"], prm["X"], prm["o_t"], 300.0, t0_ns=t0)
        okf = oracle.OracleKalmanFilter(om, mt, meas, gidx, Q, R, P0.copy())
        om.callbacks["kf"] = okf
        it = om.simulate(times, table, end_ns=t0 + steps * int(300e9))
        next(it); next(it)
        t = time.perf_counter(); k = 0
        for _ in it:
            k += 1
        out["cpu_ms_per_step"] = 1e3 * (time.perf_counter() - t) / k
        out["cpu_threads"] = oracle.max_threads()
        out["relerr_o"] = float(np.abs(o_gpu - om.o_t_next).max() / np.abs(om.o_t_next).max())
        out["relerr_P"] = float(np.abs(P_gpu - okf.P_t_next).max() / np.abs(okf.P_t_next).max())
        out["speedup"] = out["cpu_ms_per_step"] / out["gpu_ms_per_step"]
    return out


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    sizes = [int(a) for a in args] or [500, 1000, 2000, 4000]
    for n in sizes:
        print(json.dumps(one(n, cpu="--gpu-only" not in sys.argv)), flush=True)
