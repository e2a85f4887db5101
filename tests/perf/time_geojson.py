"""load_nhd_geojson at scale: the native single-pass scanner (csrc/txh_geojson.cpp) against the reference's way
(json.load + per-feature Python loops + a pandas reindex, muskingum.py:877-917 restated).  CPU only."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tx_fast_hydrology_b200.muskingum import load_nhd_geojson     # noqa: E402


def main(n):
    rng = np.random.default_rng(1)
    ids = rng.permutation(np.arange(1_000_000, 1_000_000 + 2 * n))[:n]
    down = rng.integers(0, n, size=n)
    path = os.path.join(tempfile.gettempdir(), f"nhd_{n}.json")
    with open(path, "w") as f:
        f.write('{"features": [')
        for k in range(n):
            pts = ", ".join(f"[{-97.0 + 1e-4 * j + 1e-6 * k:.6f}, {30.0 + 2e-4 * j:.6f}]" for j in range(12))
            f.write(("," if k else "") + '{"attributes": {"COMID": %d, "toCOMID": %d, "Shape_Length": %.6f, "StreamOrde": 2}, '
                    '"geometry": {"paths": [[%s]]}}' % (ids[k], ids[down[k]] if k % 50 else 0, 0.5 + 1e-6 * k, pts))
        f.write("]}")
    size_mb = os.path.getsize(path) / 1e6
    t0 = time.perf_counter(); obj = load_nhd_geojson(path, load_paths=False); t_native = time.perf_counter() - t0
    t0 = time.perf_counter()
    with open(path) as f:
        d = json.load(f)
    node_ids = [i["attributes"]["COMID"] for i in d["features"]]
    target = [i["attributes"]["toCOMID"] for i in d["features"]]
    dx = np.asarray([i["attributes"]["Shape_Length"] for i in d["features"]])
    paths = [np.asarray(i["geometry"]["paths"]) for i in d["features"]]
    m = pd.Series(np.arange(len(node_ids)), index=node_ids)
    start = m.reindex(node_ids, fill_value=-1).values
    end = m.reindex(target, fill_value=-1).values.copy()
    for i in range(len(start)):
        if end[i] == -1:
            end[i] = start[i]
    t_ref = time.perf_counter() - t0
    assert (end == obj["endnodes"]).all() and (dx == obj["dx"]).all()
    os.remove(path)
    print(json.dumps({"features": n, "file_MB": round(size_mb, 1), "native_scan_s": round(t_native, 3),
                      "reference_style_s": round(t_ref, 3), "speedup": round(t_ref / t_native, 1),
                      "native_MB_per_s": round(size_mb / t_native, 1), "cpu_count": os.cpu_count()}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300_000)
