"""Summarise a TXH_TRACE_FILE timeline dump of the window routing kernel (development aid)."""
import sys

import numpy as np


def main(path):
    raw = np.fromfile(path, dtype=np.uint64)
    pairs, ns, ntasks, nmb = (int(x) for x in raw[:4].view(np.int64))
    assert nmb < 0, "not a window-kernel trace"
    tr = raw[4:].reshape(pairs, 4 + ns).astype(np.int64)
    claim, loaded, end = tr[:, 0], tr[:, 1], tr[:, 2]
    kind = tr[:, 3] & 0xff
    smid = (tr[:, 3] >> 8) & 0xffff
    pub = tr[:, 4:]
    t0 = claim.min()
    span = (end.max() - t0) / 1e3
    print(f"pairs={pairs} steps={ns} span={span:.1f} us  SMs={len(np.unique(smid))}")
    for k, name in ((0, "POCKET"), (1, "SEG")):
        m = kind == k
        if not m.any():
            continue
        ld = (loaded - claim)[m] / 1e3; run = (end - loaded)[m] / 1e3
        st = (claim[m] - t0) / 1e3; en = (end[m] - t0) / 1e3
        per_step = np.diff(pub[m], axis=1) / 1e3 if ns > 1 else np.zeros((m.sum(), 1))
        print(f"{name:6s} n={m.sum():6d} load {np.median(ld):6.2f}/{ld.max():6.2f}  run {np.median(run):7.2f}/{run.mean():7.2f}/{run.max():7.2f}"
              f"  step {np.median(per_step):6.2f}/{per_step.mean():6.2f}/{per_step.max():6.2f}  first claim {st.min():7.1f} last claim {st.max():7.1f}"
              f"  last end {en.max():7.1f}  (us)")
    busy = (end - claim).sum() / 1e3
    print(f"warp-busy {busy:.0f} us -> avg busy warps {busy / span:.1f}")
    # the last tasks to finish: their publish times per step
    order = np.argsort(end)[-6:]
    for p in order:
        print(f" pair {p:6d} kind {kind[p]} claim {(claim[p]-t0)/1e3:7.1f} loaded {(loaded[p]-t0)/1e3:7.1f} pub "
              + " ".join(f"{(x-t0)/1e3:6.1f}" for x in pub[p]) + f" end {(end[p]-t0)/1e3:7.1f}")
    # claims over time
    edges = np.linspace(0, span, 11)
    cl = np.histogram((claim - t0) / 1e3, bins=edges)[0]
    en = np.histogram((end - t0) / 1e3, bins=edges)[0]
    print(" time bins (us):", " ".join(f"{e:6.0f}" for e in edges[1:]))
    print(" claims        :", " ".join(f"{c:6d}" for c in cl))
    print(" ends          :", " ".join(f"{c:6d}" for c in en))


if __name__ == "__main__":
    main(sys.argv[1])
