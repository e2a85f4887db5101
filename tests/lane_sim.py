"""CPU emulation of route_lane_kernel's protocol (csrc/txh_lane.cu) on the schedule the library builds.

Test infrastructure: it executes exactly what the kernel's threads do -- regions in ticket order, rows
evaluated time-skewed (row j does step s in iteration s + off[j]), upstream values read from the buffer the
previous iteration wrote, outflows crossing regions through per-slot streams -- with numpy in place of the
CUDA threads.  It checks the schedule (skews, children, streams, ticket order), not the arithmetic of the GPU.
"""
import numpy as np


def _children(R, kids, i, n_real, n_virt):
    """Region-local children of row i: the two inline ones (skipping the zero row) + the further ones."""
    zero = n_real + n_virt
    c01 = int(R[i, 2])
    cs = [c for c in (c01 & 0xffff, (c01 >> 16) & 0xffff) if c != zero]
    cs += [int(c) for c in kids[R[i, 4]:R[i, 4] + R[i, 3]]]
    return cs


def check_invariants(sched, endnodes):
    """Structural properties the kernel relies on."""
    n = endnodes.size
    regions, rows, child = sched["regions"], sched["rows"], sched["child"]
    seen = np.zeros(n, dtype=np.int64)
    producer_region = {}
    consumer_count = {}
    for g, (row_off, n_real, n_virt, child_off, n_child, n_extra, height, _) in enumerate(regions):
        R = rows[row_off:row_off + n_real + n_virt]
        assert (R[:n_real, 0] >= 0).all() and (R[n_real:, 0] == -1).all()
        seen[R[:n_real, 0]] += 1
        assert R[:, 1].min() >= 0 and R[:, 1].max() == n_extra - 1 and n_extra <= 65535
        assert int(R[:n_real, 3].sum()) == n_child and (R[n_real:, 3] == 0).all()
        # rows with further children come last; before them rows are grouped by skew phase
        has_x = R[:n_real, 3] > 0
        assert (np.diff(has_x.astype(int)) >= 0).all()
        local = {int(j): i for i, j in enumerate(R[:n_real, 0])}
        kids = child[child_off:child_off + n_child]
        virt_used = np.zeros(n_virt, dtype=np.int64)
        for i in range(n_real):
            j = int(R[i, 0])
            cs = _children(R, kids, i, n_real, n_virt)
            if n <= 5000:
                ups = np.flatnonzero((endnodes == j) & (np.arange(n) != j))
                assert len(cs) == len(ups)
                assert sorted(int(R[c, 0]) for c in cs if c < n_real) == sorted(int(u) for u in ups if int(u) in local)
            for c in cs:
                assert R[c, 1] == R[i, 1] - 1                          # one barrier between producer and consumer
                if c >= n_real:
                    virt_used[c - n_real] += 1
            if R[i, 5] >= 0:
                assert int(R[i, 5]) not in producer_region
                producer_region[int(R[i, 5])] = g
                d = int(endnodes[j])
                assert d != j and d not in local                       # published rows drain into another region
            else:
                d = int(endnodes[j])
                assert d == j or d in local
        assert (virt_used == 1).all()
        for v in range(n_virt):
            slot = int(R[n_real + v, 5])
            consumer_count[slot] = consumer_count.get(slot, 0) + 1
            assert producer_region[slot] < g                           # producers hold smaller tickets
    assert (seen == 1).all()
    assert sorted(producer_region) == list(range(sched["n_slots"]))
    assert all(consumer_count.get(s, 0) == 1 for s in range(sched["n_slots"]))


def run(sched, pos_like, alpha, beta, chi, gamma, o, i, q_steps):
    """`nsteps` routing steps; q_steps [nsteps][n] lateral inflow per step (reach order), o/i [n] reach order.
    Returns (o, i) after the last step and the full outflow trajectory [nsteps][n]."""
    del pos_like
    nsteps, n = q_steps.shape
    regions, rows, child = sched["regions"], sched["rows"], sched["child"]
    ring = np.full((max(1, sched["n_slots"]), nsteps), np.nan)
    o_out = np.array(o, dtype=np.float64); i_out = np.array(i, dtype=np.float64)
    traj = np.zeros((nsteps, n))
    for (row_off, n_real, n_virt, child_off, n_child, n_extra, height, _) in regions:
        R = rows[row_off:row_off + n_real + n_virt]
        kids = child[child_off:child_off + n_child]
        reach = R[:n_real, 0]
        clist = [_children(R, kids, r, n_real, n_virt) for r in range(n_real)]
        p = beta[reach] * i[reach] + chi[reach] * o[reach]
        ob = np.full((2, n_real + n_virt), np.nan)
        for k in range(nsteps + n_extra - 1):
            cur, prev = ob[k & 1], ob[(k & 1) ^ 1]
            for r in range(n_real):
                s = k - R[r, 1]
                if not 0 <= s < nsteps:
                    continue
                j = reach[r]
                inflow = 0.0
                for c in clist[r]:
                    v = prev[c]
                    assert not np.isnan(v), "row read before its producer wrote"
                    inflow += v
                on = alpha[j] * inflow + (p[r] + gamma[j] * q_steps[s, j])
                p[r] = beta[j] * inflow + chi[j] * on
                cur[r] = on
                traj[s, j] = on
                if R[r, 5] >= 0:
                    ring[R[r, 5], s] = on
                if s == nsteps - 1:
                    o_out[j] = on; i_out[j] = inflow
            for v in range(n_virt):
                s = k - R[n_real + v, 1]
                if not 0 <= s < nsteps:
                    continue
                val = ring[R[n_real + v, 5], s]
                assert not np.isnan(val), "stream consumed before it was produced (ticket order broken)"
                ring[R[n_real + v, 5], s] = np.nan
                cur[n_real + v] = val
    assert np.isnan(ring).all()                                        # every cell consumed and put back
    return o_out, i_out, traj
