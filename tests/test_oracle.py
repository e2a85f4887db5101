"""The CPU oracle replayed against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  Integer artefacts and -- because the C restatement keeps the
reference's operation order with FMA contraction off -- the floating-point kernels are bit-exact."""
import os

import numpy as np
import pytest


@pytest.mark.parametrize("fname", ["kernels_n60.npz", "kernels_n160.npz"])
def test_kernels_bit_exact(oracle, golden_dir, fname):
    O = oracle
    g = np.load(os.path.join(golden_dir, fname))
    en = g["endnodes"]; n = en.size; sn = np.arange(n)
    ind = O.compute_indegree(sn, en)
    assert (ind == g["indegree"]).all()
    a, b, c, ga = O.compute_coeffs(g["K"], g["X"], float(g["dt"]))
    for x, k in zip((a, b, c, ga), ("alpha", "beta", "chi", "gamma")):
        assert (x == g[k]).all()
    heads = sn[ind == 0]
    assert (O.init_states(sn, en, g["o_init"]) == g["i_init"]).all()
    i1, o1 = O._ax_bu(heads, en, a, b, c, ga, g["i_init"], g["o_init"], g["q"], ind)
    assert (i1 == g["axbu_i"]).all() and (o1 == g["axbu_o"]).all()
    i2, o2 = O._ax(heads, en, a, b, c, g["i_init"], g["o_init"], ind)
    assert (i2 == g["ax_i"]).all() and (o2 == g["ax_o"]).all()
    ig, og = O._apply_gain(heads, en, g["gain"], ind)
    assert (ig == g["gain_i"]).all() and (og == g["gain_o"]).all()
    assert (O._ap_par(g["P_sym"], np.empty((n, n)), heads, en, a, b, c, ind) == g["ap"]).all()
    for key in ("sym", "gen"):
        assert (O._aqat_par(g["P_" + key], np.empty((n, n)), heads, en, a, b, c, ind) == g["aqat_" + key]).all()
    lin = np.stack([O.interpolate_sample(float(x), g["xp"], g["fp"], 1) for x in g["xs"]])
    near = np.stack([O.interpolate_sample(float(x), g["xp"], g["fp"], 0) for x in g["xs"]])
    assert (lin == g["interp_lin"]).all() and (near == g["interp_near"]).all()
    # the interpreted restatement agrees with the C one
    ip, op = O.py_ax_bu(heads, en, a, b, c, ga, g["i_init"], g["o_init"], g["q"], ind)
    assert (op == o1).all() and (ip == i1).all()


def test_model_c1_bit_exact(oracle, golden_dir):
    """BASELINE.json configs[0] through OracleModel.simulate (muskingum.py:499-536)."""
    O = oracle
    g = np.load(os.path.join(golden_dir, "model_c1.npz"))
    n = g["endnodes"].size
    mdl = O.OracleModel(np.arange(n), g["endnodes"], g["K"], g["X"], g["o_init"], float(g["dt"]), int(g["t0_ns"]))
    keep = set(int(k) for k in g["keep"])
    Os, Is = [], []
    total = np.zeros(n)
    for k, st in enumerate(mdl.simulate(g["times"], g["table"])):
        total += st.o_t_next
        if k in keep:
            Os.append(st.o_t_next.copy()); Is.append(st.i_t_next.copy())
    assert k == 287 and mdl.time_ns == int(g["final_time_ns"])
    assert (np.stack(Os) == g["O"]).all() and (np.stack(Is) == g["I"]).all() and (total == g["o_sum"]).all()


def test_kalman_filter(oracle, golden_dir):
    """OracleKalmanFilter == KalmanFilter (da.py:14-136) to LAPACK rounding."""
    O = oracle
    g = np.load(os.path.join(golden_dir, "kalman_n120.npz"))
    n = g["endnodes"].size
    mdl = O.OracleModel(np.arange(n), g["endnodes"], g["K"], g["X"], g["o_init"], float(g["dt"]), int(g["t0_ns"]))
    kf = O.OracleKalmanFilter(mdl, g["meas_times"], g["meas"], g["gauge_cols"], g["Q"], g["R"], g["P0"])
    assert (kf.reach_indices == g["sorted_idx"]).all()
    mdl.callbacks["kf"] = kf
    Os, Pd = [], []
    for st in mdl.simulate(g["times"], g["table"]):
        Os.append(st.o_t_next.copy()); Pd.append(np.diag(kf.P_t_next).copy())
    assert np.abs(np.stack(Os) - g["O"]).max() <= 1e-11 * np.abs(g["O"]).max()
    assert np.abs(np.stack(Pd) - g["P_diag"]).max() <= 1e-11 * np.abs(g["P_diag"]).max()
    assert np.abs(kf.K - g["K_final"]).max() <= 1e-11 * np.abs(g["K_final"]).max()


def test_run_members_equals_per_member_loop(oracle):
    """The threaded member-batched baseline == stepping _ax_bu per member."""
    O = oracle
    from tx_fast_hydrology_b200 import synthetic as S
    n, M, T, seed = 800, 5, 17, 4
    net = S.make_network(n, seed); prm = S.make_params(n, seed)
    sn, en = net["startnodes"], net["endnodes"]
    ind = O.compute_indegree(sn, en)
    a, b, c, g = O.compute_coeffs(prm["K"], prm["X"], 300.0)
    t0 = 1_700_000_000 * 10**9
    times, table = S.make_forcing(n, T, 300.0, seed, t0_ns=t0, rows_every=5)
    mul = S.make_member_multipliers(times.size, M, seed)
    rng = np.random.default_rng(0)
    o0 = rng.uniform(0.1, 5.0, size=(M, n)); i0 = np.stack([O.init_states(sn, en, o0[k]) for k in range(M)])
    ref = {"startnodes": sn, "endnodes": en, "indegree": ind, "alpha": a, "beta": b, "chi": c, "gamma": g}
    o1 = o0.copy(); i1 = i0.copy()
    O.run_members(ref, o1, i1, T, times.astype(np.float64), table, float(t0), 300e9, wmul=mul, threads=3)
    xp = times.astype(np.float64)
    for k in range(M):
        o, i = o0[k].copy(), i0[k].copy()
        for s in range(T):
            x = float(t0) + (s + 1) * 300e9
            ix = int(np.searchsorted(xp, x))
            if ix == 0:
                q = mul[0, k] * table[0]
            elif ix >= xp.size:
                q = mul[-1, k] * table[-1]
            else:
                frac = (x - xp[ix - 1]) / ((x - xp[ix - 1]) + (xp[ix] - x))
                q = ((1 - frac) * mul[ix - 1, k]) * table[ix - 1] + (frac * mul[ix, k]) * table[ix]
            i, o = O._ax_bu(sn[ind == 0], en, a, b, c, g, i, o, q, ind)
        assert (o == o1[k]).all() and (i == i1[k]).all()
