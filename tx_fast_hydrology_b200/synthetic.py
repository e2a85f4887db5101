"""Deterministic synthetic river networks, forcings and gauges.

The reference ships no generator, fixtures or data; the
benchmark configurations of BASELINE.json are therefore synthesised here,
following SURVEY.md section 8(d): a recursive Hack's-law tree (main-stem
length ~ size**0.6), tributaries joining the stem no closer to its head than
their own depth (so the stem is the longest path and the number of
topological levels equals the stem length), indegree <= 4, reach ids randomly
permuted so index order is not topological order, outlets encoded as
self-loops (`muskingum.py:897-902`), `startnodes == arange(n)`.

Everything is numpy + a fixed `np.random.default_rng(seed)`; no I/O.
"""
import numpy as np

HACK_EXPONENT = 0.6
MAX_TRIBS_PER_REACH = 3          # + the stem's own upstream reach -> indegree <= 4


def _split_sizes(rng, total, k):
    """k positive integers summing to `total`, heavy-tailed (Pareto weights)."""
    if k == 1:
        return np.array([total], dtype=np.int64)
    w = rng.pareto(1.2, size=k) + 0.05
    extra = total - k
    sizes = np.floor(extra * w / w.sum()).astype(np.int64)
    short = extra - int(sizes.sum())
    if short > 0:
        idx = rng.choice(k, size=short, replace=True)
        np.add.at(sizes, idx, 1)
    return sizes + 1


def _grow_tree(rng, size, down, next_id, stem_scale):
    """Append one tree of `size` reaches to `down` (list of downstream ids, -1 =
    outlet placeholder); returns the id of its outlet reach.  Iterative."""
    # stack entries: (subtree size, id of the reach it drains into or -1)
    root_out = None
    stack = [(int(size), -1)]
    while stack:
        sz, target = stack.pop()
        ls = int(round(stem_scale * sz ** HACK_EXPONENT))
        ls = max(1, min(ls, sz))
        if sz <= 2:
            ls = sz
        first = next_id[0]
        next_id[0] += ls
        # stem reaches first..first+ls-1, headwater first; reach p drains to p+1
        down.extend(range(first + 1, first + ls))
        down.append(target)
        outlet = first + ls - 1
        if target == -1:
            root_out = outlet
        rest = sz - ls
        if rest <= 0:
            continue
        cap = MAX_TRIBS_PER_REACH * (ls - 1) if ls > 1 else 0
        if cap == 0:
            # cannot branch (stem of one reach): extend as an upstream chain
            # hanging off the single reach -- keeps sizes exact
            stack.append((rest, first))
            continue
        k = int(min(rest, max(1, min(cap, round(1.6 * (ls - 1) * rng.uniform(0.6, 1.0))))))
        sizes = _split_sizes(rng, rest, k)
        used = np.zeros(ls, dtype=np.int64)
        order = np.argsort(-sizes, kind="stable")
        for t in order:
            s_t = int(sizes[t])
            d_t = int(round(stem_scale * s_t ** HACK_EXPONENT))
            d_t = max(1, min(d_t, s_t))
            lo = min(max(1, d_t), ls - 1)
            p = int(rng.integers(lo, ls))
            # probe for a stem reach with free capacity, downstream first
            q = p
            while q < ls and used[q] >= MAX_TRIBS_PER_REACH:
                q += 1
            if q >= ls:
                q = p
                while q >= 1 and used[q] >= MAX_TRIBS_PER_REACH:
                    q -= 1
                if q < 1:
                    q = p          # over capacity: accept indegree > 4 (rare)
            used[q] += 1
            stack.append((s_t, first + q))
    return root_out


def make_network(n, seed, n_basins=1, stem_scale=1.0, permute=True, basin_sizes=None):
    """Returns dict(startnodes, endnodes int64[n], basin int64[n] = basin id of
    each reach).  `n_basins` independent trees (log-uniform sizes unless
    `basin_sizes` is given)."""
    rng = np.random.default_rng(seed)
    n = int(n)
    if basin_sizes is None:
        if n_basins == 1:
            basin_sizes = np.array([n], dtype=np.int64)
        else:
            w = np.exp(rng.uniform(np.log(1.0), np.log(200.0), size=n_basins))
            basin_sizes = np.maximum(1, np.floor(n * w / w.sum())).astype(np.int64)
            basin_sizes[np.argmax(basin_sizes)] += n - int(basin_sizes.sum())
    basin_sizes = np.asarray(basin_sizes, dtype=np.int64)
    assert int(basin_sizes.sum()) == n and (basin_sizes > 0).all()
    down = []
    next_id = [0]
    basin = np.empty(n, dtype=np.int64)
    for b, sz in enumerate(basin_sizes):
        lo = next_id[0]
        _grow_tree(rng, int(sz), down, next_id, stem_scale)
        basin[lo:next_id[0]] = b
    end = np.asarray(down, dtype=np.int64)
    assert end.size == n
    ids = np.arange(n, dtype=np.int64)
    end = np.where(end < 0, ids, end)                  # outlets -> self-loops
    if permute:
        perm = rng.permutation(n).astype(np.int64)     # old id -> new id
        new_end = np.empty(n, dtype=np.int64)
        new_end[perm] = perm[end]
        new_basin = np.empty(n, dtype=np.int64)
        new_basin[perm] = basin
        end, basin = new_end, new_basin
    return {"startnodes": ids, "endnodes": end, "basin": basin}


def make_longchain_network(stem=10000, tribs=10000, seed=5, max_trib_depth=20, permute=True):
    """Config 5: an unbranched `stem`-reach main stem plus ~`tribs` tributary
    reaches in short chains (depth <= max_trib_depth) joining at random stem
    positions."""
    rng = np.random.default_rng(seed)
    down = list(range(1, stem)) + [-1]
    used = np.zeros(stem, dtype=np.int64)
    nid = stem
    left = tribs
    while left > 0:
        d = int(min(left, rng.integers(1, max_trib_depth + 1)))
        p = int(rng.integers(max(1, d), stem))
        if used[p] >= MAX_TRIBS_PER_REACH:
            continue
        used[p] += 1
        down.extend(range(nid + 1, nid + d))
        down.append(p)
        nid += d
        left -= d
    n = nid
    end = np.asarray(down, dtype=np.int64)
    ids = np.arange(n, dtype=np.int64)
    end = np.where(end < 0, ids, end)
    if permute:
        perm = rng.permutation(n).astype(np.int64)
        new_end = np.empty(n, dtype=np.int64)
        new_end[perm] = perm[end]
        end = new_end
    return {"startnodes": ids, "endnodes": end, "basin": np.zeros(n, dtype=np.int64)}


def make_params(n, seed, well_posed=False):
    """K ~ U[300,7200] s (or U[150,600] for the alpha>0 'well-posed' set),
    X ~ U[0.05,0.45], o_t ~ U[0.1,10]."""
    rng = np.random.default_rng(seed + 1000003)
    if well_posed:
        K = rng.uniform(150.0, 600.0, size=n)
    else:
        K = rng.uniform(300.0, 7200.0, size=n)
    X = rng.uniform(0.05, 0.45, size=n)
    o_t = rng.uniform(0.1, 10.0, size=n)
    return {"K": K, "X": X, "o_t": o_t}


def make_forcing(n, nsteps, dt_s, seed, t0_ns=0, rows_every=12):
    """Hourly-style forcing table: rows every `rows_every` model steps, starting
    at the model start time, `nsteps/rows_every + 1` rows; q ~ Gamma(0.5, 2.0)
    times a smooth storm envelope.  Returns (times_ns int64[R], table f64[R][n])."""
    rng = np.random.default_rng(seed + 2000003)
    R = nsteps // rows_every + 1
    step_ns = int(round(dt_s * 1e9))
    times = t0_ns + np.arange(R, dtype=np.int64) * (rows_every * step_ns)
    base = rng.gamma(0.5, 2.0, size=n)
    x = np.linspace(0.0, 1.0, R)
    env = 0.15 + np.exp(-0.5 * ((x - 0.35) / 0.12) ** 2) + 0.5 * np.exp(-0.5 * ((x - 0.75) / 0.08) ** 2)
    jitter = rng.uniform(0.8, 1.2, size=(R, 1))
    table = (env[:, None] * jitter) * base[None, :]
    return times, np.ascontiguousarray(table)


def make_member_multipliers(R, M, seed, sigma=0.2):
    """Multiplicative lognormal forcing perturbations, one per (table row, member)."""
    rng = np.random.default_rng(seed + 3000003)
    return np.ascontiguousarray(np.exp(sigma * rng.standard_normal((R, M)) - 0.5 * sigma * sigma))


def make_gauges(endnodes, m, seed=4):
    """m gauged reaches sampled without replacement, biased to high-order
    (large contributing area) reaches.  Returns ascending reach indices."""
    rng = np.random.default_rng(seed + 4000003)
    end = np.asarray(endnodes, dtype=np.int64)
    n = end.size
    # contributing count via one topological sweep
    indeg = np.bincount(end, minlength=n)
    indeg[end == np.arange(n)] -= 1
    area = np.ones(n, dtype=np.float64)
    pending = indeg.copy()
    frontier = list(np.flatnonzero(pending == 0))
    while frontier:
        nxt = []
        for j in frontier:
            e = end[j]
            if e != j:
                area[e] += area[j]
                pending[e] -= 1
                if pending[e] == 0:
                    nxt.append(e)
        frontier = nxt
    w = area ** 0.5
    idx = rng.choice(n, size=min(m, n), replace=False, p=w / w.sum())
    return np.sort(idx.astype(np.int64))


def model_dict(net, params, dt_s=300.0, name="synthetic", t0="2024-01-01T00:00:00Z"):
    """Assemble the JSON-like dict `Muskingum.__init__` takes (muskingum.py:139)."""
    import pandas as pd
    n = net["startnodes"].size
    return {
        "name": name,
        "datetime": pd.Timestamp(t0),
        "timedelta": pd.to_timedelta(dt_s, unit="s"),
        "reach_ids": [str(i) for i in range(n)],
        "startnodes": net["startnodes"].copy(),
        "endnodes": net["endnodes"].copy(),
        "K": params["K"].copy(),
        "X": params["X"].copy(),
        "o_t": params["o_t"].copy(),
    }
