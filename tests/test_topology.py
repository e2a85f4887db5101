"""Topology pass and dataflow schedule of libtxh (host code, exact integers; no GPU needed)."""
import numpy as np
import pytest

import sched_sim
from tx_fast_hydrology_b200 import synthetic as S


def _net(libtxh, endnodes, sp=None):
    from tx_fast_hydrology_b200.network import RiverNetwork
    return RiverNetwork(endnodes, sp)


@pytest.mark.parametrize("n,seed,basins", [(1, 0, 1), (2, 1, 1), (7, 2, 1), (300, 3, 1), (5000, 4, 3), (40000, 5, 7)])
def test_topology_exact(libtxh, oracle, n, seed, basins):
    net = S.make_network(n, seed, n_basins=basins)
    en, sn = net["endnodes"], net["startnodes"]
    rn = _net(libtxh, en)
    ind = oracle.compute_indegree(sn, en)
    assert (rn.indegree() == ind).all()                                  # muskingum.py:322-330
    assert (rn.headwaters() == sn[ind == 0]).all()                       # muskingum.py:444
    assert (rn.visit_order() == oracle.visit_order(sn, en, ind)).all()   # nutils.py:72-88
    lev, nl = rn.levels()
    assert (lev == oracle.levels(en)).all() and nl == lev.max() + 1
    order, off = rn.level_order()
    assert sorted(order.tolist()) == list(range(n))
    for l in range(nl):
        blk = order[off[l]:off[l + 1]]
        assert (lev[blk] == l).all() and (np.diff(blk) > 0).all()
    # chains: v continues the chain of its only upstream reach iff indegree[v] == 1
    cid, cpos, clen = rn.chains()
    for j in range(n):
        if ind[j] == 1:
            u = int(np.flatnonzero((en == j) & (sn != j))[0]) if n <= 5000 else None
            if u is not None:
                assert cid[j] == cid[u] and cpos[j] == cpos[u] + 1
        else:
            assert cpos[j] == 0
    assert (np.bincount(cid, minlength=clen.size) == clen).all()
    # paths: follow the upstream reach of highest level; path length == level of its last reach + 1
    pid, ppos = rn.paths()
    assert (ppos <= lev).all() and (ppos[ind == 0] == 0).all()


def test_levels_against_networkx(libtxh):
    nx = pytest.importorskip("networkx")
    net = S.make_network(600, 9, n_basins=2)
    en = net["endnodes"]
    G = nx.DiGraph()
    G.add_nodes_from(range(en.size))
    G.add_edges_from((j, int(e)) for j, e in enumerate(en) if j != e)
    rn = _net(libtxh, en)
    lev, nl = rn.levels()
    assert nl == nx.dag_longest_path_length(G) + 1
    for gen, nodes in enumerate(nx.topological_generations(G)):
        assert (lev[list(nodes)] == gen).all()


def test_rejects_cycles_and_bad_indices(libtxh):
    from tx_fast_hydrology_b200._lib import TxhError
    with pytest.raises(TxhError):
        _net(libtxh, np.array([1, 2, 0], dtype=np.int64))
    with pytest.raises(TxhError):
        _net(libtxh, np.array([1, 5], dtype=np.int64))


@pytest.mark.parametrize("n,seed,sp,order", [
    (1000, 1, None, "random"), (4000, 3, None, "lifo"), (3000, 4, (8, 4, 6, 3, 2), "random"),
    (200, 5, (4, 8, 8, 1, 1), "fifo"), (2000, 7, (2, 3, 2, 0, 3), "random"), (1500, 8, (16, 32, 48, 12, 64), "random"),
    (12, 11, None, "random"), (1, 12, None, "fifo")])
def test_schedule_protocol(libtxh, oracle, n, seed, sp, order):
    """The descriptors the kernel consumes, executed on the CPU under the kernel's dataflow protocol in
    a seeded random order, reproduce _ax_bu over several steps; counters never underflow, every task
    runs every step, no event arrives before its task re-armed."""
    O = oracle
    net = S.make_network(n, seed, n_basins=2 if n > 500 else 1)
    prm = S.make_params(n, seed)
    rn = _net(libtxh, net["endnodes"], sp)
    en, sn = net["endnodes"], net["startnodes"]
    ind = O.compute_indegree(sn, en)
    info = rn.schedule_info()
    sch = rn.schedule()
    sch["n_side"] = n
    a, b, c, g = rn.compute_coeffs(prm["K"], prm["X"], 300.0)
    for x, y in zip((a, b, c, g), O.compute_coeffs(prm["K"], prm["X"], 300.0)):
        assert (x == y).all()
    pos = sch["pos_of_reach"]
    assert sorted(pos.tolist()) == list(range(n))
    rop = np.empty(n, dtype=np.int64); rop[pos] = np.arange(n)
    coef = np.stack([a, b, c, g], 1)[rop]
    o0 = prm["o_t"]; i0 = O.init_states(sn, en, o0)
    rng = np.random.default_rng(0)
    T = 4
    qs = [rng.gamma(0.5, 2.0, n) for _ in range(T)]
    Os = o0[rop].copy(); Is = i0[rop].copy()
    ex = sched_sim.simulate(sch, coef, Os, Is, lambda s: qs[s][rop], T, seed=seed, order=order)
    assert ex == T * info["n_tasks"]
    Or, Ir = o0.copy(), i0.copy()
    for s in range(T):
        Ir, Or = O._ax_bu(sn[ind == 0], en, a, b, c, g, Ir, Or, qs[s], ind)
    assert np.abs(Os[pos] - Or).max() <= 1e-12 * np.abs(Or).max()
    assert np.abs(Is[pos] - Ir).max() <= 1e-12 * np.abs(Ir).max()
    # the window-mode descriptors over the same row layout (route_window_kernel)
    win = rn.window_schedule()
    assert win["tasks"][:, 1].sum() == n and (np.diff(np.sort(win["tasks"][:, 0])) > 0).all()
    Ow = o0[rop].copy(); Iw = i0[rop].copy()
    exw = sched_sim.simulate_window(win, coef, Ow, Iw, lambda s: qs[s][rop], T)
    assert exw == T * win["n_tasks"]
    assert np.abs(Ow[pos] - Or).max() <= 1e-12 * np.abs(Or).max()
    assert np.abs(Iw[pos] - Ir).max() <= 1e-12 * np.abs(Ir).max()


def test_texas_scale_schedule(libtxh):
    """BASELINE.json configs[1] network: ~100k reaches, ~1k levels; schedule statistics are sane."""
    net = S.make_network(100_000, 2)
    rn = _net(libtxh, net["endnodes"])
    lev, nl = rn.levels()
    assert 900 <= nl <= 1100
    info = rn.schedule_info()
    assert info["row_fallbacks"] == 0 and info["slots_used"] <= 12
    assert info["cp_tasks"] < 64 and 1000 < info["n_tasks"] < 10000
    t = rn.schedule()["tasks"]
    assert t[:, 1].sum() - t[t[:, 8] == 2, 1].sum() == 100_000 + t[t[:, 8] == 3, 1].sum()   # PRE and FIX share rows


def test_longchain_network(libtxh):
    net = S.make_longchain_network()
    rn = _net(libtxh, net["endnodes"])
    _, nl = rn.levels()
    assert nl == 10000
    assert rn.schedule_info()["cp_tasks"] < 200
