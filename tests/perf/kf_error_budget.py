"""Measured error of the dense Kalman filter / smoother against the reference goldens, next to the bound their
conditioning allows (development aid; prints one JSON line).  Run on the GPU box."""
import json
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity import relerr, normerr                         # noqa: E402
import test_gpu_model as T                                  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
out = {}
from tx_fast_hydrology_b200.da import KalmanFilter, KalmanSmoother   # noqa: E402
g = np.load(os.path.join(G, "kalman_n120.npz"))
mdl, d = T.model_from(g)
df = T.frame(g["times"], g["table"], d["reach_ids"])
mdf = T.frame(g["meas_times"], g["meas"], [d["reach_ids"][j] for j in g["gauge_cols"]])
kf = KalmanFilter(mdl, mdf, g["Q"], g["R"], g["P0"])
mdl.bind_callback(kf, key="kf")
O, gains, Pd = [], [], []
for state in mdl.simulate(df):
    O.append(state.o_t_next.copy()); gains.append(kf.gain.copy()); Pd.append(np.diag(kf.P_t_next).copy())
out["kf"] = {"o_elem": relerr(np.stack(O), g["O"]), "o_norm": normerr(np.stack(O), g["O"]),
             "gains_norm": normerr(np.stack(gains), g["gains"]),
             "gains_norm_per_step": [normerr(a, b) for a, b in zip(gains, g["gains"])],
             "P_diag_norm": normerr(np.stack(Pd), g["P_diag"]), "P_norm": normerr(kf.P_t_next, g["P_final"]),
             "P_elem": relerr(kf.P_t_next, g["P_final"]), "K_norm": normerr(kf.K, g["K_final"]),
             "cond_P_final": float(np.linalg.cond(g["P_final"]))}
g = np.load(os.path.join(G, "smoother_n60.npz"))
mdl, d = T.model_from(g)
df = T.frame(g["times"], g["table"], d["reach_ids"])
mdf = T.frame(g["meas_times"], g["meas"], [d["reach_ids"][j] for j in g["gauge_idx"]])
ks = KalmanSmoother(mdl, mdf, g["Q"], g["R"], g["P0"])
mdl.bind_callback(ks, key="ks")
for _ in mdl.simulate(df):
    pass
ts = sorted(ks.o_hat_s.index)
conds = [float(np.linalg.cond(ks.P_p[t].cpu().numpy())) for t in ks.datetimes[1:]]
out["smoother"] = {"o_hat_s_norm": normerr(ks.o_hat_s.loc[ts].values, g["o_hat_s"]),
                   "i_hat_s_norm": normerr(ks.i_hat_s.loc[ts].values, g["i_hat_s"]),
                   "o_hat_s_elem": relerr(ks.o_hat_s.loc[ts].values, g["o_hat_s"]),
                   "P_s_first_norm": normerr(ks.P_s[ts[0]].cpu().numpy(), g["P_s_first"]),
                   "P_f_last_norm": normerr(ks.P_f[ts[-1]].cpu().numpy(), g["P_f_last"]),
                   "cond_P_p_max": max(conds), "cond_P_p": conds}
print(json.dumps(out))
