"""The reach-parallel (lane) schedule of csrc/txh_topology.cpp, checked on the CPU: structural invariants and a
protocol emulation of route_lane_kernel against the oracle's restatement of nutils.py:64-89."""
import numpy as np
import pytest

import lane_sim


@pytest.mark.parametrize("n,seed,cap,basins", [(60, 3, 16, 1), (400, 5, 32, 2), (1000, 1, 0, 1), (3000, 8, 97, 3),
                                               (3000, 8, 3000, 1)])
def test_lane_schedule_emulated_against_oracle(oracle, n, seed, cap, basins):
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork
    netd = S.make_network(n, seed, n_basins=basins)
    prm = S.make_params(n, seed)
    end = netd["endnodes"]
    net = RiverNetwork(end)
    sched = net.lane_schedule(1, cap)
    lane_sim.check_invariants(sched, end)
    al, be, ch, ga = oracle.compute_coeffs(prm["K"], prm["X"], 300.0)
    ind = oracle.compute_indegree(netd["startnodes"], end)
    rng = np.random.default_rng(seed)
    nsteps = 7
    q = rng.gamma(0.5, 2.0, size=(nsteps, n))
    o0 = prm["o_t"].copy()
    i0 = oracle.init_states(netd["startnodes"], end, o0)
    o, i = o0.copy(), i0.copy()
    ref = []
    heads = netd["startnodes"][ind == 0]                            # muskingum.py:444
    for s in range(nsteps):
        i, o = oracle._ax_bu(heads, end, al, be, ch, ga, i, o, q[s], ind)
        ref.append(o.copy())
    go, gi, traj = lane_sim.run(sched, None, al, be, ch, ga, o0, i0, q)
    scale = np.abs(np.array(ref)).max()
    assert np.abs(traj - np.array(ref)).max() <= 1e-12 * scale
    assert np.abs(go - o).max() <= 1e-12 * scale and np.abs(gi - i).max() <= 1e-12 * np.abs(i).max()


def test_lane_schedule_long_chain_and_member_tiles(oracle):
    """BASELINE configs[4]-like shape (unbranched stem + tributaries) and the larger member tiles."""
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork
    netd = S.make_longchain_network(stem=1500, tribs=1500, seed=5)
    end = netd["endnodes"]
    net = RiverNetwork(end)
    for M, cap in ((1, 256), (4, 128), (16, 64)):
        sched = net.lane_schedule(M, cap)
        lane_sim.check_invariants(sched, end)
        assert sched["max_real"] <= cap
        assert sched["member_tile"] >= M


def test_lane_schedule_large_network_scales():
    """100k reaches (BASELINE configs[1]): one region per SM-sized share, every reach exactly once."""
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork
    netd = S.make_network(100000, 2)
    net = RiverNetwork(netd["endnodes"])
    sched = net.lane_schedule(1)
    rows = sched["rows"]
    real = rows[rows[:, 0] >= 0, 0]
    assert real.size == 100000 and np.unique(real).size == 100000
    assert sched["n_regions"] >= 100 and sched["max_real"] <= 2048
