"""Host-side logic of the multi-GPU paths on CPU: member-sharded EnKF statistics combined with
torch.distributed (gloo, world_size 2) equal the single-process statistics; basin sharding packs
whole basins."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tx_fast_hydrology_b200.sharding import combine_statistics
    rng = np.random.default_rng(7)
    n, m, M = 50, 6, 4
    X = rng.standard_normal((n, world * M))
    obs = np.sort(rng.choice(n, m, replace=False))
    loc = X[:, rank * M:(rank + 1) * M]
    rowsum = torch.from_numpy(loc.sum(axis=1).copy())
    HX = torch.from_numpy(np.ascontiguousarray(loc[obs]))
    mean, HXall = combine_statistics(rowsum, HX, world * M, group=None)
    ok = np.allclose(mean.numpy(), X.mean(axis=1), rtol=0, atol=1e-15) and np.array_equal(HXall.numpy(), X[obs])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_member_sharded_statistics_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_basin_sharding_balances_whole_basins():
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.sharding import shard_basins, extract_shard
    net = S.make_network(30000, 3, n_basins=24)
    parts = shard_basins(net["basin"], 4)
    sizes = [int(np.isin(net["basin"], p).sum()) for p in parts]
    assert sum(sizes) == 30000 and max(sizes) <= 1.35 * (30000 / 4)
    assert sorted(b for p in parts for b in p) == list(range(24))
    sub, idx = extract_shard(net["endnodes"], net["basin"], parts[1])
    assert sub.size == sizes[1] and (sub < sub.size).all()
    # an extracted shard is closed under "downstream": its endnodes map back to the same reaches
    assert (idx[sub] == net["endnodes"][idx]).all()


def test_coefficient_arrays_are_read_only_and_tracked():
    """`model.alpha` ... hand out read-only arrays; assignment (or set_transmissive_boundary, muskingum.py:567-571)
    installs a copy and marks the device copy stale -- no O(n) comparison before a launch (host logic only)."""
    import numpy as np
    import pytest
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.muskingum import Muskingum
    n = 60
    mdl = Muskingum(S.model_dict(S.make_network(n, 5), S.make_params(n, 5), dt_s=300.0))
    assert not mdl._coef_dirty and not mdl.alpha.flags.writeable
    with pytest.raises(ValueError):
        mdl.alpha[0] = 1.0
    a0 = mdl.alpha.copy()
    mdl.set_transmissive_boundary(np.array([2, 7]))
    assert mdl._coef_dirty
    assert mdl.alpha[2] == 1.0 and mdl.beta[7] == 0.0 and mdl.chi[2] == 0.0 and mdl.gamma[7] == 0.0
    assert np.array_equal(np.delete(mdl.alpha, [2, 7]), np.delete(a0, [2, 7]))
    with pytest.raises(ValueError):
        mdl.beta = np.zeros(n + 1)
    c = mdl.copy()
    assert np.array_equal(c.alpha, mdl.alpha) and c.alpha is not mdl.alpha
