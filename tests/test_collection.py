"""`Muskingum.split` / `ModelCollection` / `AsyncSimulation` against the golden
vectors of the unmodified reference (tests/golden/split_n300.npz, made by make_golden.split_collection)."""
import asyncio
import os

import numpy as np
import pandas as pd
import pytest

RTOL = 1e-9


from parity import relerr, normerr        # element-wise with a floor / max-norm (dense matrices)


def _model(g):
    from tx_fast_hydrology_b200.muskingum import Muskingum
    n = g["endnodes"].size
    d = {"name": "golden", "datetime": pd.Timestamp(int(g["t0_ns"]), tz="UTC"),
         "timedelta": pd.to_timedelta(float(g["dt"]), unit="s"), "reach_ids": [str(i) for i in range(n)],
         "startnodes": np.arange(n, dtype=np.int64), "endnodes": g["endnodes"].astype(np.int64),
         "K": g["K"].astype(np.float64), "X": g["X"].astype(np.float64), "o_t": g["o_init"].astype(np.float64),
         "dx": np.ones(n), "paths": [[0.0] for _ in range(n)]}
    return Muskingum(d), d


def _frame(g, cols):
    idx = pd.DatetimeIndex(pd.to_datetime(g["times"], unit="ns", utc=True)).as_unit("ns")
    return pd.DataFrame(g["table"], index=idx, columns=cols)


def test_split_structure(libtxh, golden_dir, tmp_path):
    """Components, their numbering, reach membership, inflow correction and connections (host only)."""
    from tx_fast_hydrology_b200.muskingum import ModelCollection
    g = np.load(os.path.join(golden_dir, "split_n300.npz"))
    mdl, d = _model(g)
    mc = mdl.split([int(c) for c in g["cuts"]])
    assert list(mc.models.keys()) == [str(k) for k in range(int(g["n_models"]))]
    conns = []
    for k, sub in mc.models.items():
        assert ([int(r) for r in sub.reach_ids] == g[f"reach_{k}"]).all()
        assert (sub.indegree == g[f"indegree_{k}"]).all()                   # exact integers
        assert relerr(sub.i_t_next, g[f"i0_{k}"]) < 1e-15
        assert (sub.startnodes == np.arange(sub.n)).all()
        for c in sub.sinks:
            conns.append([int(c.upstream_model.name), int(c.downstream_model.name), int(c.upstream_index),
                          int(c.downstream_index)])
    assert (np.asarray(sorted(conns)) == g["connections"]).all()
    assert mc.datetime == mdl.datetime and mc.timedelta == mdl.timedelta
    # JSON round trip of the collection keeps models and wiring
    path = str(tmp_path / "collection.json")
    mc.dump_model_collection(path)
    mc2 = ModelCollection.from_file(path)
    assert list(mc2.models.keys()) == list(mc.models.keys())
    for k in mc.models:
        assert (mc2.models[k].K == mc.models[k].K).all() and mc2.models[k].reach_ids == mc.models[k].reach_ids
        assert len(mc2.models[k].sinks) == len(mc.models[k].sinks)
        assert len(mc2.models[k].sources) == len(mc.models[k].sources)


@pytest.mark.gpu
@pytest.mark.parametrize("with_callback", [False, True])
def test_async_simulation_matches_reference(libtxh, golden_dir, with_callback):
    """Sub-model hydrographs == the reference's AsyncSimulation, and together == the un-split run; once through
    the device-resident fast path and once through the per-step path (a callback bound to every sub-model)."""
    from tx_fast_hydrology_b200.simulation import AsyncSimulation, CheckPoint
    g = np.load(os.path.join(golden_dir, "split_n300.npz"))
    mdl, d = _model(g)
    df = _frame(g, d["reach_ids"])
    mc = mdl.split([int(c) for c in g["cuts"]])
    if with_callback:
        for sub in mc.models.values():
            sub.bind_callback(CheckPoint(sub, timedelta=7200), key="checkpoint")
    sim = AsyncSimulation(mc, df)
    outputs = asyncio.run(sim.simulate())
    whole = np.empty_like(g["whole"])
    for k, sub in mc.models.items():
        out = outputs[k]
        # the reference's frame carries a microsecond index (pandas 3 default unit); compare instants
        assert (out.index.as_unit("ns").astype("int64").values == g[f"out_times_{k}"] * 1000).all()
        assert relerr(out.values, g[f"out_{k}"]) < RTOL
        whole[:, g[f"reach_{k}"]] = out.values
    assert relerr(whole, g["whole"]) < RTOL
    assert mc.datetime.value == int(g["times"][-1])


def test_load_nhd_geojson(libtxh, tmp_path):
    """muskingum.py:877-917: COMID/toCOMID -> indices, missing downstream id -> self-loop outlet, defaults."""
    import json
    from tx_fast_hydrology_b200.muskingum import Muskingum, load_nhd_geojson
    comid = [501, 77, 9001, 12, 345, 60]
    to = [77, 12, 12, 999999, 501, 345]                  # 999999 is not a feature: reach 12 is the outlet
    feats = [{"attributes": {"COMID": c, "toCOMID": t, "Shape_Length": 1.5 + k},
              "geometry": {"paths": [[[0.0, 1.0 * k], [1.0, 1.0 * k]]]}} for k, (c, t) in enumerate(zip(comid, to))]
    path = str(tmp_path / "nhd.json")
    with open(path, "w") as f:
        json.dump({"features": feats}, f)
    obj = load_nhd_geojson(path)
    assert obj["reach_ids"] == [str(c) for c in comid]
    assert (obj["startnodes"] == np.arange(6)).all()
    assert (obj["endnodes"] == np.array([1, 3, 3, 3, 0, 4])).all()
    assert (obj["K"] == 3600.0).all() and (obj["X"] == 0.29).all() and (obj["o_t"] == 1e-3).all()
    assert (obj["dx"] == 1.5 + np.arange(6)).all() and len(obj["paths"]) == 6
    mdl = Muskingum(obj)
    assert (mdl.indegree == np.array([1, 1, 0, 2, 1, 0])).all()
