"""Region timeline of one route_lane_kernel launch (TXH_LANE_TRACE=<file>, development aid).

    TXH_LANE_TRACE=gpurun_out/lane.bin python tests/perf/run_configs.py c2 ; python tools/lane_trace.py gpurun_out/lane.bin
"""
import sys

import numpy as np


def main(path):
    raw = open(path, "rb").read()
    hd = np.frombuffer(raw[:32], dtype=np.int64)
    nreg, nsteps, threads, grid = (int(x) for x in hd)
    regs = np.frombuffer(raw[32:32 + 32 * nreg], dtype=np.int32).reshape(nreg, 8)
    tr = np.frombuffer(raw[32 + 32 * nreg:], dtype=np.uint64).reshape(nreg, 8).astype(np.float64)
    t0 = tr[:, 0][tr[:, 0] > 0].min()
    claim, loaded, lag, end = ((tr[:, k] - t0) / 1e3 for k in range(4))
    lag = np.where(tr[:, 2] > 0, lag, loaded)
    iters = tr[:, 4]
    loop = end - lag
    per_iter = loop / np.maximum(iters, 1)
    print(f"regions {nreg}  steps {nsteps}  threads {threads}  grid {grid}  launch span {end.max():.1f} us")
    print(f"per-iteration after the last lag wait: median {np.median(per_iter) * 1e3:.0f} ns  p10 {np.percentile(per_iter, 10) * 1e3:.0f}  "
          f"p90 {np.percentile(per_iter, 90) * 1e3:.0f}  max {per_iter.max() * 1e3:.0f}")
    print(f"load phase: median {np.median(loaded - claim):.1f} us   lag wait: median {np.median(lag - loaded):.1f} us  max {np.max(lag - loaded):.1f} us")
    print(f"slow-path polls per region: median {np.median(tr[:, 6]):.0f}  max {tr[:, 6].max():.0f}  total {tr[:, 6].sum():.0f}")
    order = np.argsort(end)
    print("last regions to finish:  reg  height rows virt extra | claim  loaded  lag-over  end (us) | ns/iter  slow polls  sm")
    for g in order[-12:]:
        print(f"   {g:5d} {regs[g, 6]:5d} {regs[g, 1]:5d} {regs[g, 2]:4d} {regs[g, 5]:5d} | {claim[g]:8.1f} {loaded[g]:8.1f} {lag[g]:8.1f} {end[g]:8.1f} | "
              f"{per_iter[g] * 1e3:7.0f} {tr[g, 6]:8.0f} {tr[g, 5]:4.0f}")
    by_h = {}
    for g in range(nreg):
        by_h.setdefault(int(regs[g, 6]), []).append(g)
    print("by height: height  regions  median lag-over  median end  median ns/iter")
    for h in sorted(by_h):
        idx = by_h[h]
        print(f"   {h:4d} {len(idx):6d} {np.median(lag[idx]):10.1f} {np.median(end[idx]):10.1f} {np.median(per_iter[idx]) * 1e3:8.0f}")


if __name__ == "__main__":
    main(sys.argv[1])
