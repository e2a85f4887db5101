"""Summarise a TXH_TRACE_FILE timeline dump of the dataflow routing kernel (development aid)."""
import sys

import numpy as np


def main(path):
    raw = np.fromfile(path, dtype=np.uint64)
    pairs, ns, ntasks, nmb = (int(x) for x in raw[:4].view(np.int64))
    tr = raw[4:].reshape(ns, pairs, 4)
    t_pop = tr[..., 0].astype(np.int64); t_beg = tr[..., 1].astype(np.int64); t_cmp = tr[..., 2].astype(np.int64)
    t_end = (tr[..., 3] >> np.uint64(16)).astype(np.int64)
    kind = (tr[..., 3] & np.uint64(0xff)).astype(np.int64)
    smid = ((tr[..., 3] >> np.uint64(8)) & np.uint64(0xff)).astype(np.int64)
    # globaltimer low bits were shifted out of 64 bits: compare on the truncated (48-bit) scale
    m48 = (1 << 48) - 1
    t_pop &= m48; t_beg &= m48; t_cmp &= m48
    t0 = t_pop[t_pop > 0].min()
    span = (t_end.max() - t0) / 1e3
    print(f"pairs={pairs} steps={ns} tasks={ntasks} nmb={nmb}  span={span:.1f} us  ({span / ns:.1f} us/step)")
    names = {0: "POCKET", 1: "PRE", 2: "LINK", 3: "FIX"}
    for k in (0, 1, 2, 3):
        m = kind == k
        if not m.any():
            continue
        wait = (t_beg - t_pop)[m] / 1e3; run = (t_cmp - t_beg)[m] / 1e3; nfy = (t_end - t_cmp)[m] / 1e3
        print(f"{names[k]:7s} n={m.sum():7d}  pop+desc {np.median(wait):7.2f}/{wait.mean():7.2f}  "
              f"body {np.median(run):7.2f}/{run.mean():7.2f}/{run.max():7.2f}  notify {np.median(nfy):6.2f}/{nfy.mean():6.2f}  (median/mean[/max] us)")
    busy = (t_end - t_beg).sum() / 1e3
    nsm = len(np.unique(smid))
    print(f"warp-busy time {busy:.0f} us over {nsm} SMs -> avg busy warps {busy / span:.1f}")
    for s in range(min(ns, 12)):
        print(f" step {s:3d}: first start {(t_beg[s].min() - t0) / 1e3:9.1f}  last end {(t_end[s].max() - t0) / 1e3:9.1f} us")
    # chain timeline in step 0: sort CHAIN tasks by end time, show the tail
    s = 0
    m = np.flatnonzero(kind[s] == 2)
    order = m[np.argsort(t_end[s][m])]
    print(" last LINK tasks of step 0 (pair, begin, body_end, end) us:")
    for p in order[-12:]:
        print(f"   {p:6d} {(t_beg[s][p] - t0) / 1e3:9.2f} {(t_cmp[s][p] - t0) / 1e3:9.2f} {(t_end[s][p] - t0) / 1e3:9.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
