"""BASELINE.json configs[1], [3] and [4] at their full network sizes, through the C ABI, against the CPU oracle
(oracle/txh_oracle.c: nutils.py:64-89 looped per member).  Step counts are cut to what the oracle finishes in
seconds; the full-length runs are covered by size-independent properties (linearity, launch splitting,
idempotence of the recorded trajectory) and by bench.py's in-run parity number.
Tolerance: element-wise relative error <= 1e-9 with a floor of 1e-12 of the largest element (tests/parity.py)."""
import numpy as np
import pytest

from parity import relerr

pytestmark = pytest.mark.gpu
RTOL = 1e-9
T0 = 1_700_000_000 * 10**9
DT = int(300e9)


@pytest.fixture(scope="module")
def torch_cuda(libtxh):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no CPU fallback exists)"
    return torch


def _oracle_net(oracle, net_d, coef):
    al, be, ch, ga = coef
    return {"startnodes": net_d["startnodes"], "endnodes": net_d["endnodes"],
            "indegree": oracle.compute_indegree(net_d["startnodes"], net_d["endnodes"]),
            "alpha": al, "beta": be, "chi": ch, "gamma": ga}


def _run_case(torch, oracle, net_d, seed, M, T, rows_every=12, members_checked=None, chunks=None):
    """GPU run of T steps (optionally in several calls) vs the oracle on the same inputs."""
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork, Forcing
    n = net_d["endnodes"].size
    prm = S.make_params(n, seed)
    net = RiverNetwork(net_d["endnodes"])
    coef = net.compute_coeffs(prm["K"], prm["X"], 300.0)
    times, table = S.make_forcing(n, T, 300.0, seed, t0_ns=T0, rows_every=rows_every)
    mul = S.make_member_multipliers(times.size, M, seed) if M > 1 else None
    rng = np.random.default_rng(seed)
    o0 = prm["o_t"][:, None] * (rng.uniform(0.5, 1.5, size=(n, M)) if M > 1 else np.ones((n, 1)))
    f = Forcing(net, times, table, mul)
    O = net.alloc_state(M); I = net.alloc_state(M)
    net.pack_host(o0, M, O)
    net.init_inflows(O, I, M)
    t = T0
    for ns in (chunks or [T]):
        net.route_run(O, I, M, f, t, DT, ns)
        t += ns * DT
    net.check()
    mc = M if members_checked is None else members_checked
    o_gpu = net.unpack_host(O, M)[:, :mc]; i_gpu = net.unpack_host(I, M)[:, :mc]
    o_ref = np.ascontiguousarray(o0[:, :mc].T)
    i_ref = np.stack([oracle.init_states(net_d["startnodes"], net_d["endnodes"], x) for x in o_ref])
    oracle.run_members(_oracle_net(oracle, net_d, coef), o_ref, i_ref, T, times.astype(np.float64), table,
                       float(T0), float(DT), wmul=None if mul is None else np.ascontiguousarray(mul[:, :mc]))
    eo, ei = relerr(o_gpu, o_ref.T), relerr(i_gpu, i_ref.T)
    f.close(); net.close()
    return eo, ei


def test_c2_texas_scale_deterministic_336_steps(torch_cuda, oracle):
    """configs[1]: ~100k reaches, ~1k levels, single deterministic member, 28 hours at a 5-min step -- 336 steps,
    28 forcing brackets -- in one call (the lane kernel's whole-run launch)."""
    from tx_fast_hydrology_b200 import synthetic as S
    eo, ei = _run_case(torch_cuda, oracle, S.make_network(100_000, 2), 2, 1, 336)
    assert eo < RTOL and ei < RTOL, (eo, ei)


def test_c2_run_split_over_calls_is_the_same_run(torch_cuda, oracle):
    """The same network, 100 steps as 1 + 37 + 50 + 12: state handed over in HBM between launches, forcing
    brackets entered in the middle of a call."""
    from tx_fast_hydrology_b200 import synthetic as S
    eo, ei = _run_case(torch_cuda, oracle, S.make_network(100_000, 2), 2, 1, 100, chunks=[1, 37, 50, 12])
    assert eo < RTOL and ei < RTOL, (eo, ei)


def test_c4_conus_scale_24_steps(torch_cuda, oracle):
    """configs[3]: ~2.7M reaches in 64 independent basins (seed 3), deterministic, 24 steps on ONE GPU: ten times
    more regions than SMs, so regions are claimed in waves and streams wait for regions of earlier waves."""
    from tx_fast_hydrology_b200 import synthetic as S
    eo, ei = _run_case(torch_cuda, oracle, S.make_network(2_700_000, 3, n_basins=64), 3, 1, 24)
    assert eo < RTOL and ei < RTOL, (eo, ei)


def test_c5_long_chain_1024_members(torch_cuda, oracle):
    """configs[4]: 10k-reach unbranched main stem + 10k tributary reaches, 1024 members (every member checked),
    24 steps with per-member forcing multipliers: the affine hop along the stem's segments (window kernel)."""
    from tx_fast_hydrology_b200 import synthetic as S
    eo, ei = _run_case(torch_cuda, oracle, S.make_longchain_network(10000, 10000, seed=5), 5, 1024, 24)
    assert eo < RTOL and ei < RTOL, (eo, ei)


@pytest.mark.parametrize("M", [1, 2, 3, 8, 16])
def test_c5_long_chain_small_ensembles(torch_cuda, oracle, monkeypatch, M):
    """The same long-chain network on the reach-parallel kernel (forced up to 16 members): a 10k-deep skew,
    member tiles 1..16 with padding members (M = 3)."""
    from tx_fast_hydrology_b200 import synthetic as S
    monkeypatch.setenv("TXH_ROUTE_KERNEL", "lane")
    eo, ei = _run_case(torch_cuda, oracle, S.make_longchain_network(10000, 10000, seed=5), 5, M, 30, rows_every=7)
    assert eo < RTOL and ei < RTOL, (eo, ei)


def test_lane_kernel_recording_and_launch_splitting(torch_cuda, oracle, monkeypatch):
    """Recorded hydrographs (every 3rd step) from the lane kernel with a launch that is split (TXH_LANE_SPL) at
    a multiple of the recording cadence: identical to the single-launch run, and equal to the oracle's
    trajectory at the recorded steps."""
    torch = torch_cuda
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork, Forcing
    n, seed, T, every = 5000, 41, 90, 3
    net_d = S.make_network(n, seed, n_basins=2)
    prm = S.make_params(n, seed)
    times, table = S.make_forcing(n, T, 300.0, seed, t0_ns=T0, rows_every=5)
    rec_reach = np.sort(np.random.default_rng(seed).choice(n, size=40, replace=False))
    out = {}
    for spl in ("0", "31"):
        monkeypatch.setenv("TXH_LANE_SPL", spl)
        net = RiverNetwork(net_d["endnodes"])
        coef = net.compute_coeffs(prm["K"], prm["X"], 300.0)
        f = Forcing(net, times, table)
        O = net.alloc_state(1); I = net.alloc_state(1)
        net.pack_host(prm["o_t"][:, None], 1, O); net.init_inflows(O, I, 1)
        rec = torch.zeros((T // every, rec_reach.size, 1), dtype=torch.float64, device="cuda")
        net.route_run(O, I, 1, f, T0, DT, T, rec_reach=rec_reach, rec_every=every, rec_out=rec)
        net.check()
        out[spl] = (rec.cpu().numpy()[:, :, 0], net.unpack_host(O, 1)[:, 0])
        f.close(); net.close()
    assert (out["0"][0] == out["31"][0]).all() and (out["0"][1] == out["31"][1]).all()
    ref = _oracle_net(oracle, net_d, coef)
    o = prm["o_t"].copy()[None, :]
    i = oracle.init_states(net_d["startnodes"], net_d["endnodes"], o[0])[None, :]
    o = np.ascontiguousarray(o); i = np.ascontiguousarray(i)
    traj = []
    for k in range(T // every):
        oracle.run_members(ref, o, i, every, times.astype(np.float64), table, float(T0 + k * every * DT), float(DT))
        traj.append(o[0, rec_reach].copy())
    assert relerr(out["0"][0], np.stack(traj)) < RTOL
