"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference (imported from /root/reference) on seeded synthetic inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Outputs (committed): kernels_n60.npz, kernels_n160.npz, model_c1.npz,
kalman_n120.npz, checkpoint_n80.npz, split_n300.npz, smoother_n60.npz, quirks_n90.npz,
checkpoint_kf_n70.npz.  Every file stores its inputs next to the
reference's outputs, so replaying it needs neither the reference nor the
generator.
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from tx_fast_hydrology.muskingum import Muskingum            # noqa: E402
from tx_fast_hydrology.nutils import (_ax_bu, _ax, _apply_gain, _ap_par, _aqat_par,   # noqa: E402
                                      interpolate_sample)
from tx_fast_hydrology.da import KalmanFilter                 # noqa: E402
from tx_fast_hydrology.simulation import CheckPoint           # noqa: E402
from tx_fast_hydrology_b200 import synthetic as S             # noqa: E402

T0 = "2024-01-01T00:00:00Z"


def frame(times_ns, table, cols):
    idx = pd.DatetimeIndex(pd.to_datetime(times_ns, unit="ns", utc=True)).as_unit("ns")
    return pd.DataFrame(table, index=idx, columns=cols)


def kernels(n, seed, fname):
    net = S.make_network(n, seed)
    prm = S.make_params(n, seed)
    d = S.model_dict(net, prm, dt_s=300.0, t0=T0)
    mdl = Muskingum(d)
    rng = np.random.default_rng(seed + 77)
    heads = mdl.startnodes[mdl.indegree == 0]
    q = rng.gamma(0.5, 2.0, size=n)
    i_prev = mdl.i_t_next.copy(); o_prev = mdl.o_t_next.copy()
    i1, o1 = _ax_bu(heads, mdl.endnodes, mdl.alpha, mdl.beta, mdl.chi, mdl.gamma,
                    i_prev, o_prev, q, mdl.indegree)
    i2, o2 = _ax(heads, mdl.endnodes, mdl.alpha, mdl.beta, mdl.chi, i_prev, o_prev,
                 mdl.indegree)
    gain = rng.standard_normal(n)
    ig, og = _apply_gain(heads, mdl.endnodes, gain, mdl.indegree)
    Pm = rng.standard_normal((n, n))
    P = Pm @ Pm.T / n + np.eye(n)
    ap = _ap_par(P, np.empty((n, n)), heads, mdl.endnodes, mdl.alpha, mdl.beta, mdl.chi,
                 mdl.indegree).copy()
    aqat_sym = np.ascontiguousarray(_aqat_par(P, np.empty((n, n)), heads, mdl.endnodes,
                                              mdl.alpha, mdl.beta, mdl.chi, mdl.indegree))
    aqat_gen = np.ascontiguousarray(_aqat_par(Pm, np.empty((n, n)), heads, mdl.endnodes,
                                              mdl.alpha, mdl.beta, mdl.chi, mdl.indegree))
    # interpolation: inside, on a knot, before, after; linear + nearest
    xp = (1.7e18 + np.arange(6) * 3.6e12)
    fp = rng.standard_normal((6, 7))
    xs = np.array([xp[0] - 5.0e11, xp[0], xp[2] + 9.0e11, xp[3], xp[4] + 1.8e12,
                   xp[5], xp[5] + 1e12])
    lin = np.stack([interpolate_sample(float(x), xp, fp, 1) for x in xs])
    near = np.stack([interpolate_sample(float(x), xp, fp, 0) for x in xs])
    np.savez_compressed(
        os.path.join(HERE, fname), endnodes=mdl.endnodes, K=mdl.K, X=mdl.X, dt=300.0,
        indegree=mdl.indegree, alpha=mdl.alpha, beta=mdl.beta, chi=mdl.chi,
        gamma=mdl.gamma, o_init=prm["o_t"], i_init=i_prev, q=q,
        axbu_i=i1, axbu_o=o1, ax_i=i2, ax_o=o2, gain=gain, gain_i=ig, gain_o=og,
        P_sym=P, P_gen=Pm, ap=ap, aqat_sym=aqat_sym, aqat_gen=aqat_gen,
        xp=xp, fp=fp, xs=xs, interp_lin=lin, interp_near=near)
    print(fname, "ok")


def model_c1():
    """BASELINE.json configs[0]: 1,000 reaches, 288 five-minute steps, Muskingum.simulate."""
    n, T, seed = 1000, 288, 1
    net = S.make_network(n, seed)
    prm = S.make_params(n, seed)
    d = S.model_dict(net, prm, dt_s=300.0, t0=T0)
    mdl = Muskingum(d)
    t0_ns = mdl.datetime.value
    times, table = S.make_forcing(n, T, 300.0, seed, t0_ns=t0_ns)
    df = frame(times, table, d["reach_ids"])
    keep = list(range(0, T, 12)) + [T - 1]
    O = []; I = []
    total = np.zeros(n)
    for k, state in enumerate(mdl.simulate(df)):
        total += state.o_t_next
        if k in keep:
            O.append(state.o_t_next.copy()); I.append(state.i_t_next.copy())
    assert k == T - 1
    np.savez_compressed(
        os.path.join(HERE, "model_c1.npz"), endnodes=mdl.endnodes, K=mdl.K, X=mdl.X,
        o_init=prm["o_t"], dt=300.0, t0_ns=t0_ns, times=times, table=table,
        keep=np.asarray(keep), O=np.stack(O), I=np.stack(I), o_sum=total,
        final_time_ns=mdl.datetime.value)
    print("model_c1.npz ok")


def kalman():
    n, m, T, seed = 120, 8, 24, 11
    net = S.make_network(n, seed)
    prm = S.make_params(n, seed, well_posed=True)
    d = S.model_dict(net, prm, dt_s=300.0, t0=T0)
    mdl = Muskingum(d)
    t0_ns = mdl.datetime.value
    times, table = S.make_forcing(n, T, 300.0, seed, t0_ns=t0_ns, rows_every=6)
    df = frame(times, table, d["reach_ids"])
    gidx = S.make_gauges(net["endnodes"], m, seed=seed)
    rng = np.random.default_rng(seed + 5)
    gidx_cols = rng.permutation(gidx)                 # unsorted columns: exercises da.py:36-44
    mt = t0_ns + np.arange(0, T + 1, 3, dtype=np.int64) * int(300e9)
    mt = mt[mt <= t0_ns + 18 * int(300e9)]            # measurements end before the run does
    meas = rng.uniform(0.5, 8.0, size=(mt.size, m))
    mdf = frame(mt, meas, [d["reach_ids"][j] for j in gidx_cols])
    Rm = rng.standard_normal((m, m)); R = 1e-2 * np.eye(m) + 1e-3 * (Rm @ Rm.T)
    Q = 2.0 * np.eye(n); P0 = Q.copy()
    kf = KalmanFilter(mdl, mdf, Q, R, P0)
    mdl.bind_callback(kf, key="kf")
    O = []; I = []; Pd = []; gains = []
    for state in mdl.simulate(df):
        O.append(state.o_t_next.copy()); I.append(state.i_t_next.copy())
        Pd.append(np.diag(kf.P_t_next).copy()); gains.append(kf.gain.copy())
    np.savez_compressed(
        os.path.join(HERE, "kalman_n120.npz"), endnodes=mdl.endnodes, K=mdl.K, X=mdl.X,
        o_init=prm["o_t"], dt=300.0, t0_ns=t0_ns, times=times, table=table,
        gauge_cols=gidx_cols, meas_times=mt, meas=meas, R=R, Q=Q, P0=P0,
        O=np.stack(O), I=np.stack(I), P_diag=np.stack(Pd), gains=np.stack(gains),
        P_final=kf.P_t_next, K_final=kf.K, sorted_idx=kf.reach_indices)
    print("kalman_n120.npz ok")


def checkpoint():
    """CheckPoint + save_state/load_state rewind (simulation.py:169-211, muskingum.py:573-588)."""
    n, T, seed = 80, 36, 21
    net = S.make_network(n, seed)
    prm = S.make_params(n, seed)
    d = S.model_dict(net, prm, dt_s=300.0, t0=T0)
    mdl = Muskingum(d)
    t0_ns = mdl.datetime.value
    times, table = S.make_forcing(n, T, 300.0, seed, t0_ns=t0_ns)
    df = frame(times, table, d["reach_ids"])
    cp = CheckPoint(mdl, timedelta=3600)
    mdl.bind_callback(cp, key="checkpoint")
    for state in mdl.simulate(df):
        pass
    o_end = mdl.o_t_next.copy(); t_end = mdl.datetime.value
    saved_o = mdl.saved_states["o_t_next"].copy(); saved_i = mdl.saved_states["i_t_next"].copy()
    saved_t = mdl.saved_states["datetime"].value
    mdl.load_state()
    o_loaded = mdl.o_t_next.copy(); t_loaded = mdl.datetime.value
    for state in mdl.simulate(df):       # second cycle from the checkpoint
        pass
    np.savez_compressed(
        os.path.join(HERE, "checkpoint_n80.npz"), endnodes=mdl.endnodes, K=mdl.K, X=mdl.X,
        o_init=prm["o_t"], dt=300.0, t0_ns=t0_ns, times=times, table=table,
        o_end=o_end, t_end=t_end, saved_o=saved_o, saved_i=saved_i, saved_t=saved_t,
        o_loaded=o_loaded, t_loaded=t_loaded, o_end2=mdl.o_t_next.copy(),
        t_end2=mdl.datetime.value, saved_t2=mdl.saved_states["datetime"].value)
    print("checkpoint_n80.npz ok")


def split_collection():
    """Muskingum.split + AsyncSimulation (muskingum.py:607-714, simulation.py:94-166): a network cut at
    three reaches, hourly step with forcing rows aligned 1:1 with the steps, the sub-models'
    hydrographs, and the un-split run they must reproduce."""
    import asyncio
    from tx_fast_hydrology.simulation import AsyncSimulation
    n, T, seed = 300, 24, 31
    net = S.make_network(n, seed)
    prm = S.make_params(n, seed)
    d = S.model_dict(net, prm, dt_s=3600.0, t0=T0)
    d["dx"] = np.ones(n)
    d["paths"] = [[0.0] for _ in range(n)]
    mdl = Muskingum(d)
    t0_ns = mdl.datetime.value
    rng = np.random.default_rng(seed + 3)
    times = t0_ns + (np.arange(T, dtype=np.int64) + 1) * int(3600e9)
    table = rng.gamma(0.5, 2.0, size=(T, n))
    df = frame(times, table, d["reach_ids"])
    # cut at three non-outlet reaches with a decent subtree above them
    order = np.argsort(-np.bincount(net["endnodes"], minlength=n) - rng.uniform(0, 0.5, n))
    cuts = [int(j) for j in order if net["endnodes"][j] != j][:3]
    mc = mdl.split(cuts, create_state_space=False)
    names = list(mc.models.keys())
    sub_reach = {k: np.asarray([int(r) for r in mc.models[k].reach_ids]) for k in names}
    sub_i0 = {k: mc.models[k].i_t_next.copy() for k in names}
    sub_indeg = {k: mc.models[k].indegree.copy() for k in names}
    conns = []
    for k in names:
        for c in mc.models[k].sinks:
            conns.append([int(c.upstream_model.name), int(c.downstream_model.name), int(c.upstream_index),
                          int(c.downstream_index)])
    sim = AsyncSimulation(mc, df)
    outputs = asyncio.run(sim.simulate())
    # un-split run
    ref = Muskingum(d)
    whole = [ref.o_t_next.copy()]
    for state in ref.simulate(df):
        whole.append(state.o_t_next.copy())
    out = dict(endnodes=mdl.endnodes, K=mdl.K, X=mdl.X, o_init=prm["o_t"], dt=3600.0, t0_ns=t0_ns, times=times,
               table=table, cuts=np.asarray(cuts), n_models=len(names), connections=np.asarray(sorted(conns)),
               whole=np.stack(whole))
    for k in names:
        out[f"reach_{k}"] = sub_reach[k]; out[f"i0_{k}"] = sub_i0[k]; out[f"indegree_{k}"] = sub_indeg[k]
        out[f"out_{k}"] = outputs[k].values
        out[f"out_times_{k}"] = outputs[k].index.astype("int64").values
    np.savez_compressed(os.path.join(HERE, "split_n300.npz"), **out)
    print("split_n300.npz ok", len(names), "models", conns)


def smoother():
    """KalmanSmoother (da.py:139-264): forward filter with stored covariances + RTS backward pass."""
    from tx_fast_hydrology.da import KalmanSmoother
    n, m, T, seed = 60, 6, 10, 41
    net = S.make_network(n, seed)
    prm = S.make_params(n, seed, well_posed=True)
    d = S.model_dict(net, prm, dt_s=300.0, t0=T0)
    mdl = Muskingum(d)
    t0_ns = mdl.datetime.value
    times, table = S.make_forcing(n, T, 300.0, seed, t0_ns=t0_ns, rows_every=5)
    df = frame(times, table, d["reach_ids"])
    gidx = np.sort(S.make_gauges(net["endnodes"], m, seed=seed))
    rng = np.random.default_rng(seed + 5)
    mt = t0_ns + np.arange(0, T + 1, dtype=np.int64) * int(300e9)        # a measurement at every model time
    meas = rng.uniform(0.5, 8.0, size=(mt.size, m))
    mdf = frame(mt, meas, [d["reach_ids"][j] for j in gidx])
    Rm = rng.standard_normal((m, m)); R = 1e-2 * np.eye(m) + 1e-3 * (Rm @ Rm.T)
    Q = 0.5 * np.eye(n); P0 = Q.copy()
    ks = KalmanSmoother(mdl, mdf, Q, R, P0)
    mdl.bind_callback(ks, key="ks")
    for state in mdl.simulate(df):
        pass
    ts = sorted(ks.o_hat_s.index)
    np.savez_compressed(
        os.path.join(HERE, "smoother_n60.npz"), endnodes=mdl.endnodes, K=mdl.K, X=mdl.X, o_init=prm["o_t"], dt=300.0,
        t0_ns=t0_ns, times=times, table=table, gauge_idx=gidx, meas_times=mt, meas=meas, R=R, Q=Q, P0=P0,
        o_final=mdl.o_t_next, i_final=mdl.i_t_next,
        o_hat_s=ks.o_hat_s.loc[ts].values, i_hat_s=ks.i_hat_s.loc[ts].values,
        smooth_times=np.asarray([pd.Timestamp(t).value for t in ts]),
        P_s_first=ks.P_s[ts[0]], P_f_last=ks.P_f[ts[-1]], n_times=len(ks.datetimes))
    print("smoother_n60.npz ok", len(ks.datetimes))


def quirks():
    """User-mutated coefficients and the variable-timestep quirk, through the reference's step():
    set_transmissive_boundary (muskingum.py:567-571) in the middle of a run, and step(p, timedelta=dt2)
    followed by default steps (muskingum.py:447-449: the second call sees dt == self.dt and keeps the
    coefficients of the 600 s call)."""
    n, seed = 90, 23
    net = S.make_network(n, seed)
    prm = S.make_params(n, seed)
    d = S.model_dict(net, prm, dt_s=300.0, t0=T0)
    rng = np.random.default_rng(seed)
    q = rng.gamma(0.5, 2.0, size=(16, n))
    # (a) transmissive boundary installed after 5 steps
    mdl = Muskingum(d)
    for s in range(5):
        mdl.step(q[s])
    o_a5, i_a5 = mdl.o_t_next.copy(), mdl.i_t_next.copy()
    tb = np.sort(rng.choice(n, size=6, replace=False))
    mdl.set_transmissive_boundary(tb)
    for s in range(5, 10):
        mdl.step(q[s])
    o_a10, i_a10 = mdl.o_t_next.copy(), mdl.i_t_next.copy()
    coef_a = np.stack([mdl.alpha, mdl.beta, mdl.chi, mdl.gamma])
    # (b) variable timestep
    mdl = Muskingum(dict(d))
    mdl.step(q[0])
    mdl.step(q[1], timedelta=pd.to_timedelta(600, unit="s"))
    o_b2, t_b2 = mdl.o_t_next.copy(), mdl.datetime.value
    coef_b2 = np.stack([mdl.alpha, mdl.beta, mdl.chi, mdl.gamma])
    mdl.step(q[2])                                   # default timedelta: runs on the 600 s coefficients
    mdl.step(q[3])
    o_b4, i_b4, t_b4 = mdl.o_t_next.copy(), mdl.i_t_next.copy(), mdl.datetime.value
    coef_b4 = np.stack([mdl.alpha, mdl.beta, mdl.chi, mdl.gamma])
    np.savez_compressed(
        os.path.join(HERE, "quirks_n90.npz"), endnodes=mdl.endnodes, K=mdl.K, X=mdl.X, o_init=prm["o_t"],
        dt=300.0, t0_ns=pd.Timestamp(T0).value, q=q, tb=tb, o_a5=o_a5, i_a5=i_a5, o_a10=o_a10, i_a10=i_a10,
        coef_a=coef_a, o_b2=o_b2, t_b2=t_b2, coef_b2=coef_b2, o_b4=o_b4, i_b4=i_b4, t_b4=t_b4, coef_b4=coef_b4)
    print("quirks_n90.npz ok")


def checkpoint_kf():
    """The operational cycle of app.py:56-80,125-141: 'checkpoint' bound before 'kf' (SURVEY.md A.7), a run,
    load_state (model state, clock and the filter's covariance rewind: simulation.py:196-206, da.py:83-89),
    the measurement table REASSIGNED (app.py:75-80), and a second run from the checkpoint."""
    n, m, T, seed = 70, 6, 30, 29
    net = S.make_network(n, seed)
    prm = S.make_params(n, seed, well_posed=True)
    d = S.model_dict(net, prm, dt_s=300.0, t0=T0)
    mdl = Muskingum(d)
    t0_ns = mdl.datetime.value
    times, table = S.make_forcing(n, 2 * T, 300.0, seed, t0_ns=t0_ns, rows_every=6)
    cols = d["reach_ids"]
    df1 = frame(times[times <= t0_ns + T * int(300e9)], table[times <= t0_ns + T * int(300e9)], cols)
    gidx = S.make_gauges(net["endnodes"], m, seed=seed)
    rng = np.random.default_rng(seed + 1)
    gcols = rng.permutation(gidx)
    mt1 = t0_ns + np.arange(0, T + 1, 3, dtype=np.int64) * int(300e9)
    meas1 = rng.uniform(0.5, 8.0, size=(mt1.size, m))
    mdf1 = frame(mt1, meas1, [cols[j] for j in gcols])
    R = 1e-2 * np.eye(m); Q = 2.0 * np.eye(n); P0 = Q.copy()
    cp = CheckPoint(mdl, timedelta=3600)
    mdl.bind_callback(cp, key="checkpoint")
    kf = KalmanFilter(mdl, mdf1, Q, R, P0)
    mdl.bind_callback(kf, key="kf")
    for state in mdl.simulate(df1):
        pass
    o_end, t_end = mdl.o_t_next.copy(), mdl.datetime.value
    P_end = kf.P_t_next.copy()
    saved_o, saved_i = mdl.saved_states["o_t_next"].copy(), mdl.saved_states["i_t_next"].copy()
    saved_t = mdl.saved_states["datetime"].value
    saved_P = kf.saved_states["P_t_next"].copy()
    mdl.load_state()
    o_loaded, t_loaded, P_loaded = mdl.o_t_next.copy(), mdl.datetime.value, kf.P_t_next.copy()
    # a new forecast cycle: new forcing frame, new measurement frame (different rows, same gauges)
    t1 = mdl.datetime.value
    sel = (times >= t1) & (times <= t1 + T * int(300e9))
    df2 = frame(times[sel], table[sel], cols)
    mt2 = t1 + np.arange(0, T + 1, 2, dtype=np.int64) * int(300e9)
    meas2 = rng.uniform(0.5, 8.0, size=(mt2.size, m))
    kf.measurements = frame(mt2, meas2, [cols[j] for j in gcols]).iloc[:, np.argsort(gcols)]
    O2 = []
    for state in mdl.simulate(df2):
        O2.append(state.o_t_next.copy())
    np.savez_compressed(
        os.path.join(HERE, "checkpoint_kf_n70.npz"), endnodes=mdl.endnodes, K=mdl.K, X=mdl.X, o_init=prm["o_t"],
        dt=300.0, t0_ns=t0_ns, times=times, table=table, T=T, gauge_cols=gcols, meas_times1=mt1, meas1=meas1,
        meas_times2=mt2, meas2=meas2, R=R, Q=Q, P0=P0, o_end=o_end, t_end=t_end, P_end=P_end, saved_o=saved_o,
        saved_i=saved_i, saved_t=saved_t, saved_P=saved_P, o_loaded=o_loaded, t_loaded=t_loaded, P_loaded=P_loaded,
        O2=np.stack(O2), o_end2=mdl.o_t_next.copy(), i_end2=mdl.i_t_next.copy(), t_end2=mdl.datetime.value,
        P_end2=kf.P_t_next.copy(), saved_t2=mdl.saved_states["datetime"].value)
    print("checkpoint_kf_n70.npz ok")


def split_kf():
    """The product's operating mode (app/app.py:121-166): Muskingum.split, ONE dense KalmanFilter bound to every
    sub-model that has gauges, AsyncSimulation over the collection.  Same network, cuts and forcing as
    split_n300.npz; gauges sampled per sub-model, measurements hourly."""
    import asyncio
    from tx_fast_hydrology.simulation import AsyncSimulation
    n, T, seed = 300, 24, 31
    net = S.make_network(n, seed)
    prm = S.make_params(n, seed, well_posed=True)
    d = S.model_dict(net, prm, dt_s=3600.0, t0=T0)
    d["dx"] = np.ones(n)
    d["paths"] = [[0.0] for _ in range(n)]
    mdl = Muskingum(d)
    t0_ns = mdl.datetime.value
    rng = np.random.default_rng(seed + 3)
    times = t0_ns + (np.arange(T, dtype=np.int64) + 1) * int(3600e9)
    table = rng.gamma(0.5, 2.0, size=(T, n))
    df = frame(times, table, d["reach_ids"])
    order = np.argsort(-np.bincount(net["endnodes"], minlength=n) - rng.uniform(0, 0.5, n))
    cuts = [int(j) for j in order if net["endnodes"][j] != j][:3]
    mc = mdl.split(cuts, create_state_space=False)
    names = list(mc.models.keys())
    out = dict(endnodes=mdl.endnodes, K=mdl.K, X=mdl.X, o_init=prm["o_t"], dt=3600.0, t0_ns=t0_ns, times=times,
               table=table, cuts=np.asarray(cuts), n_models=len(names))
    mt = t0_ns + np.arange(0, T + 1, dtype=np.int64) * int(3600e9)
    mt = mt[mt <= t0_ns + 18 * int(3600e9)]
    out["meas_times"] = mt
    kfs = {}
    for k in names:
        sub = mc.models[k]
        m = min(4, max(1, sub.n // 12))
        if sub.n < 6:
            out[f"gauges_{k}"] = np.zeros(0, dtype=np.int64)
            continue
        gl = np.sort(rng.choice(sub.n, size=m, replace=False))           # local gauge indices
        meas = rng.uniform(0.5, 8.0, size=(mt.size, m))
        mdf = frame(mt, meas, [sub.reach_ids[j] for j in gl])
        R = 1e-2 * np.eye(m); Q = 2.0 * np.eye(sub.n); P0 = Q.copy()
        kf = KalmanFilter(sub, mdf, Q, R, P0)
        sub.bind_callback(kf, key="kf")
        kfs[k] = kf
        out[f"gauges_{k}"] = gl; out[f"meas_{k}"] = meas
    sim = AsyncSimulation(mc, df)
    outputs = asyncio.run(sim.simulate())
    for k in names:
        out[f"reach_{k}"] = np.asarray([int(r) for r in mc.models[k].reach_ids])
        out[f"out_{k}"] = outputs[k].values
        if k in kfs:
            out[f"P_{k}"] = kfs[k].P_t_next
            out[f"i_end_{k}"] = mc.models[k].i_t_next
    np.savez_compressed(os.path.join(HERE, "split_kf_n300.npz"), **out)
    print("split_kf_n300.npz ok", len(names), "models,", len(kfs), "filters")


if __name__ == "__main__":
    only = set(sys.argv[1:])
    if only:                                  # e.g. `make_golden.py quirks checkpoint_kf`: just these fixtures
        for name in only:
            globals()[name]()
        sys.exit(0)
    smoother()
    split_collection()
    kernels(60, 7, "kernels_n60.npz")
    kernels(160, 8, "kernels_n160.npz")
    model_c1()
    kalman()
    checkpoint()
