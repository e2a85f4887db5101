"""Critical path of a route_window_kernel launch from its TXH_TRACE_FILE timeline (development aid).

    python tools/trace_critical.py trace.bin [--n 100000 --seed 2]

Walks back from the last publish of the launch: the predecessor of (task, step) is whichever finished last of its own
previous step, its producers' same step, and its load; prints how the span divides over those edge kinds."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("trace")
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--slow", action="store_true", help="also list the tasks that pace the critical path (its 'own' edges)")
    a = ap.parse_args()
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork
    net = RiverNetwork(S.make_network(a.n, a.seed)["endnodes"])
    ws = net.window_schedule()
    tasks, prod = ws["tasks"], ws["prod"]
    raw = np.fromfile(a.trace, dtype=np.uint64)
    pairs, ns, ntasks, nmb = (int(x) for x in raw[:4].view(np.int64))
    assert nmb == -1 and ntasks == tasks.shape[0], (nmb, ntasks, tasks.shape)
    tr = raw[4:].reshape(pairs, 4 + ns).astype(np.int64)
    t0 = tr[:, 0].min()
    claim, loaded, end = tr[:, 0] - t0, tr[:, 1] - t0, tr[:, 2] - t0
    kind = tr[:, 3] & 0xff
    smid = (tr[:, 3] >> 8) & 0xffff
    pub = tr[:, 4:] - t0
    print(f"tasks {ntasks} steps {ns} span {end.max() / 1e3:.1f} us; segments {(kind == 1).sum()} pockets {(kind == 0).sum()}")
    T = int(np.argmax(pub[:, ns - 1])); s = ns - 1
    acc = {"own": [0, 0.0], "hop": [0, 0.0], "hop_same_sm": [0, 0.0], "load": [0, 0.0], "claim": [0, 0.0]}
    path = []
    while True:
        t_now = pub[T, s]
        cands = []
        if s > 0:
            cands.append((pub[T, s - 1], "own", T, s - 1))
        else:
            cands.append((loaded[T], "load", T, -1))
        po, npr = tasks[T, 5], tasks[T, 6]
        for p in prod[po:po + npr]:
            cands.append((pub[p, s], "hop_same_sm" if smid[p] == smid[T] else "hop", int(p), s))
        tp, k, Tn, sn = max(cands)
        acc[k][0] += 1; acc[k][1] += (t_now - tp) / 1e3
        path.append((T, s, k, t_now / 1e3))
        if k == "load":
            acc["claim"][0] += 1; acc["claim"][1] += (loaded[T] - claim[T]) / 1e3
            print(f"path starts at task {T} (kind {kind[T]}), claimed at {claim[T] / 1e3:.1f} us, loaded at {loaded[T] / 1e3:.1f} us")
            break
        T, s = Tn, sn
    for k, (c, us) in acc.items():
        print(f"  {k:12s} edges {c:5d}  total {us:8.1f} us  mean {us / max(c, 1):6.2f} us")
    # how the wave structure looks: claims over time
    edges = np.linspace(0, end.max() / 1e3, 11)
    print(" bins (us):", " ".join(f"{e:6.0f}" for e in edges[1:]))
    print(" claims   :", " ".join(f"{c:6d}" for c in np.histogram(claim / 1e3, bins=edges)[0]))
    print(" ends     :", " ".join(f"{c:6d}" for c in np.histogram(end / 1e3, bins=edges)[0]))
    # per-step publish time of the final task and the ripple of step 0 down the longest producer chain
    print(" last task publishes (us):", " ".join(f"{x / 1e3:6.1f}" for x in pub[path[0][0]]))
    print(" path head (task, step, edge, t):", path[:6], "... tail:", path[-6:])
    if a.slow:
        slow_stages(a.trace, a.n, a.seed)



def slow_stages(trace, n=100000, seed=2):
    """Which tasks own the 'own' edges of the critical path, and what do they look like."""
    from tx_fast_hydrology_b200 import synthetic as S
    from tx_fast_hydrology_b200.network import RiverNetwork
    net = RiverNetwork(S.make_network(n, seed)["endnodes"])
    ws = net.window_schedule()
    tasks, prod = ws["tasks"], ws["prod"]
    raw = np.fromfile(trace, dtype=np.uint64)
    pairs, ns, ntasks, nmb = (int(x) for x in raw[:4].view(np.int64))
    tr = raw[4:].reshape(pairs, 4 + ns).astype(np.int64)
    t0 = tr[:, 0].min()
    loaded = tr[:, 1] - t0
    pub = tr[:, 4:] - t0
    kind = tr[:, 3] & 0xff
    T = int(np.argmax(pub[:, ns - 1])); s = ns - 1
    own = {}
    while True:
        cands = [(pub[T, s - 1], "own", T, s - 1)] if s > 0 else [(loaded[T], "load", T, -1)]
        po, npr = tasks[T, 5], tasks[T, 6]
        for p in prod[po:po + npr]:
            cands.append((pub[p, s], "hop", int(p), s))
        tp, k, Tn, sn = max(cands)
        if k == "own":
            own.setdefault(T, []).append((s, (pub[T, s] - tp) / 1e3))
        if k == "load":
            break
        T, s = Tn, sn
    for T, lst in own.items():
        d = tasks[T]
        steps = np.diff(pub[T]) / 1e3
        print(f"task {T} kind {kind[T]} len {d[1]} n_words {d[4]} n_prod {d[6]} n_in {d[8]} n_out {d[9]}: own edges {[(s, round(x, 2)) for s, x in lst]}; "
              f"its step times {np.round(steps, 2)}")
        # how long after its last input did each step publish?
        po, npr = tasks[T, 5], tasks[T, 6]
        if npr:
            last_in = np.max(pub[prod[po:po + npr]], axis=0)
            print("   publish - last input (us):", np.round((pub[T] - last_in) / 1e3, 2))



if __name__ == "__main__":
    main()
