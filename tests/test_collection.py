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


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [True, False])
def test_collection_with_a_kalman_filter_per_sub_model(libtxh, golden_dir, batched):
    """The product's operating mode (app/app.py:121-166) against the unmodified reference: the network split into
    sub-models, ONE dense KalmanFilter bound to each, AsyncSimulation over the collection -- sub-model hydrographs,
    final inflows and final covariances.  `batched`: the filters of the sub-models that are ready together run as
    ONE chain of launches on the union of their networks (txh_kfb_filter), or one chain per sub-model."""
    from tx_fast_hydrology_b200.da import KalmanFilter
    from tx_fast_hydrology_b200.simulation import AsyncSimulation
    g = np.load(os.path.join(golden_dir, "split_kf_n300.npz"))
    mdl, d = _model(g)
    df = _frame(g, d["reach_ids"])
    mc = mdl.split([int(c) for c in g["cuts"]])
    assert len(mc.models) == int(g["n_models"])
    idx = pd.DatetimeIndex(pd.to_datetime(g["meas_times"], unit="ns", utc=True)).as_unit("ns")
    kfs = {}
    for k, sub in mc.models.items():
        gl = g[f"gauges_{k}"]
        if gl.size == 0:
            continue
        mdf = pd.DataFrame(g[f"meas_{k}"], index=idx, columns=[sub.reach_ids[j] for j in gl])
        kfs[k] = KalmanFilter(sub, mdf, 2.0 * np.eye(sub.n), 1e-2 * np.eye(gl.size), 2.0 * np.eye(sub.n))
        sub.bind_callback(kfs[k], key="kf")
    sim = AsyncSimulation(mc, df)
    sim.batch_filters = batched
    from tx_fast_hydrology_b200._lib import load
    launches = load().txh_launch_count()
    outputs = asyncio.run(sim.simulate())
    launches = load().txh_launch_count() - launches
    print(f"batched={batched}: {launches} kernel launches for the collection run")
    for k, sub in mc.models.items():
        assert ([int(r) for r in sub.reach_ids] == g[f"reach_{k}"]).all()
        assert relerr(outputs[k].values, g[f"out_{k}"]) < RTOL
        if k in kfs:
            assert relerr(sub.i_t_next, g[f"i_end_{k}"]) < RTOL
            assert normerr(kfs[k].P_t_next, g[f"P_{k}"]) < RTOL         # covariance: max-norm (tests/parity.py)
            assert kfs[k].K.shape == (sub.n, g[f"gauges_{k}"].size) and kfs[k].gain.shape == (sub.n,)
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


def test_native_geojson_scanner_equals_the_reference_loader(libtxh, tmp_path):
    """txh_scan_nhd_geojson (csrc/txh_geojson.cpp) == the reference's per-feature loops (muskingum.py:877-917,
    restated here with json + a dict) on 40,000 features: shuffled ids, keys in varying order, null / missing /
    unknown toCOMID, geometry with nested arrays and braces inside strings."""
    import json
    import time
    from tx_fast_hydrology_b200.muskingum import load_nhd_geojson
    rng = np.random.default_rng(5)
    n = 40_000
    ids = rng.permutation(np.arange(10_000_000, 10_000_000 + 3 * n))[:n]
    down = rng.integers(0, n, size=n)
    feats = []
    for k in range(n):
        r = rng.random()
        to = None if r < 0.02 else (int(ids[down[k]]) if r < 0.95 else 777)      # null / a feature / an unknown id
        attrs = {"COMID": int(ids[k]), "toCOMID": to, "Shape_Length": float(rng.uniform(0.1, 9.0)),
                 "GNIS_NAME": 'a } [ \\" b'}
        if k % 3 == 0:
            attrs = dict(reversed(list(attrs.items())))
        if k % 7 == 0:
            attrs.pop("toCOMID")
        geom = {"paths": [[[float(k), 0.5], [float(k) + 1, 1.5e-3]]], "note": "{[}"}
        feats.append({"geometry": geom, "attributes": attrs} if k % 2 else {"attributes": attrs, "geometry": geom})
    path = str(tmp_path / "big.json")
    with open(path, "w") as f:
        json.dump({"type": "x", "features": feats, "tail": [1, {"features": []}]}, f)
    t0 = time.perf_counter()
    obj = load_nhd_geojson(path, load_paths=False)
    t_native = time.perf_counter() - t0
    # the reference's logic
    index_of = {}
    for k in range(n):
        index_of.setdefault(int(ids[k]), k)
    end = np.array([index_of.get(ft["attributes"].get("toCOMID"), k) for k, ft in enumerate(feats)])
    assert obj["reach_ids"] == [str(int(x)) for x in ids]
    assert (obj["endnodes"] == end).all() and (obj["startnodes"] == np.arange(n)).all()
    assert (obj["dx"] == np.array([ft["attributes"]["Shape_Length"] for ft in feats])).all()
    assert obj["paths"] == [] and t_native < 5.0
    with open(str(tmp_path / "bad.json"), "w") as f:
        f.write('{"features": [{"attributes": {"toCOMID": 3, "Shape_Length": 1.0}}]}')
    from tx_fast_hydrology_b200._lib import TxhError
    with pytest.raises(TxhError):
        load_nhd_geojson(str(tmp_path / "bad.json"))
