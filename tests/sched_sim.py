"""Test-only CPU interpreter of the device schedule (NOT a product path).

Executes the very descriptor arrays the routing kernel consumes (`txh_get_schedule`)
under the kernel's dataflow protocol -- pending counters, re-arm on completion,
same-step / next-step notifications, ready set -- with tasks picked in a seeded
random order, so that both the descriptors and the protocol's invariants are
checked without a GPU.  Arithmetic follows the kernel's (POCKET / PRE / LINK / FIX).
"""
import numpy as np

ROW = 0x80000000
POCKET, PRE, LINK, FIX = 0, 1, 2, 3


def simulate(sched, coef, O, I, q_of_step, nsteps, seed=0, order="random"):
    """coef [n][4], O/I [n] (schedule order, float64; updated in place),
    q_of_step(s) -> [n] schedule-order forcing.  Returns #tasks executed."""
    tasks = sched["tasks"]; hdr = sched["hdr"]; inw = sched["inw"]; nfy = sched["notify"]
    nt = tasks.shape[0]
    rng = np.random.default_rng(seed)
    pending = tasks[:, 6].astype(np.int64).copy()           # need0
    stepno = np.zeros(nt, dtype=np.int64)
    ready = [k for k in range(nt) if pending[k] == 0]
    running_guard = np.zeros(nt, dtype=bool)
    scratch = np.full(32, np.nan)
    side_buf = np.full(max(1, int(sched.get('n_side', O.size))), np.nan)
    # prefix product of alpha along each segment (rows whose header carries the continue bit)
    cumA = np.empty(O.size)
    for p in range(O.size):
        cumA[p] = cumA[p - 1] * coef[p][0] if (p > 0 and (int(hdr[p]) & 1)) else coef[p][0]
    executed = 0
    while ready:
        if order == "random":
            k = ready.pop(int(rng.integers(len(ready))))
        elif order == "lifo":
            k = ready.pop()
        else:
            k = ready.pop(0)
        begin, ln, in_off, nfy_off, n_same, n_next, need0, need, kind, nwords = (int(x) for x in tasks[k, :10])
        s = int(stepno[k])
        assert s < nsteps and pending[k] == 0 and not running_guard[k]
        running_guard[k] = True
        q = q_of_step(s)
        w = in_off
        if kind == POCKET:
            acc = 0.0
            for p in range(begin, begin + ln):
                h = int(hdr[p]); infl = acc if (h & 1) else 0.0
                for _ in range((h >> 6) & 0x1ffffff):
                    x = int(inw[w]); w += 1
                    if x & ROW:
                        infl += O[x & 0x7fffffff]
                    else:
                        assert not np.isnan(scratch[x]); infl += scratch[x]; scratch[x] = np.nan
                a, b, c, g = coef[p]
                on = a * infl + (b * I[p] + c * O[p] + g * q[p])
                I[p] = infl; O[p] = on
                if h >> 31:                            # pocket root: push to the spine's side slab
                    side_buf[int(inw[w])] = on; w += 1
                sl = (h >> 1) & 31
                if sl:
                    scratch[sl - 1] = on
                acc = on
        elif kind == PRE:
            B = 0.0
            cur = int(tasks[k, 10])                    # side_off
            for p in range(begin, begin + ln):
                h = int(hdr[p]); side = 0.0
                for _ in range((h >> 6) & 0x1fff):
                    assert not np.isnan(side_buf[cur]); side += side_buf[cur]; side_buf[cur] = np.nan; cur += 1
                a, b, c, g = coef[p]
                infl = side + (B if (h & 1) else 0.0)
                B = a * infl + (b * I[p] + c * O[p] + g * q[p])
                I[p] = side; O[p] = B
        elif kind == LINK:
            o = 0.0
            for _ in range(ln):                      # one record per segment of the path
                last = int(inw[w]) & 0x7fffffff; nl = int(inw[w + 1]); w += 2
                oin = o
                for _ in range(nl):
                    x = int(inw[w]); w += 1
                    assert x & ROW
                    oin += O[x & 0x7fffffff]
                o = cumA[last] * oin + O[last]
                O[last] = o
        else:
            oin = 0.0
            for _ in range(nwords):
                x = int(inw[w]); w += 1
                assert x & ROW
                oin += O[x & 0x7fffffff]
            op = oin
            for p in range(begin, begin + ln):
                on = O[p]
                if p + 1 < begin + ln:
                    on = cumA[p] * oin + O[p]
                    O[p] = on
                I[p] = op + I[p]
                op = on
        executed += 1
        # completion protocol (route_dataflow_kernel)
        assert pending[k] == 0, "an event for the next step arrived before the task re-armed"
        stepno[k] = s + 1
        more = s + 1 < nsteps
        if more:
            pending[k] = need
        running_guard[k] = False
        targets = list(nfy[nfy_off:nfy_off + n_same]) + (list(nfy[nfy_off + n_same:nfy_off + n_same + n_next]) if more else [])
        for t in targets:
            t = int(t)
            pending[t] -= 1
            assert pending[t] >= 0, "dependency counter underflow"
            if pending[t] == 0:
                assert stepno[t] < nsteps
                ready.append(t)
    assert (stepno == nsteps).all(), "not every task ran every step"
    return executed


WIN_SLOT, WIN_OWN = 0x80000000, 0x40000000
WPOCKET, WSEG = 0, 1


def simulate_window(w, coef, O, I, q_of_step, nsteps):
    """Test-only CPU interpreter of the WINDOW-mode descriptors (`txh_get_window_schedule`): every task
    keeps its rows locally for all `nsteps` steps and exchanges rows through the slot ring, exactly as
    route_window_kernel does; tasks run one after another in descriptor order (producers first).
    coef [n][4], O/I [n] schedule order (updated in place).  Returns #task-steps executed."""
    tasks, hdr, inw, prod = w["tasks"], w["hdr"], w["inw"], w["prod"]
    nt = tasks.shape[0]
    ring = np.full((nsteps, max(1, w["n_slots"])), np.nan)
    done = np.zeros(nt, dtype=bool)
    cumA = np.empty(O.size)
    for p in range(O.size):
        cumA[p] = cumA[p - 1] * coef[p][0] if (p > 0 and (int(hdr[p]) & 1)) else coef[p][0]
    qs = [q_of_step(s) for s in range(nsteps)]
    executed = 0
    for k in range(nt):
        begin, ln, kind, in_off, nwords, prod_off, nprod, out_slot, n_in = (int(x) for x in tasks[k, :9])
        for pr in prod[prod_off:prod_off + nprod]:
            assert pr < k and done[pr], "producer not ordered before its consumer"
        Ol = O[begin:begin + ln].copy(); Il = I[begin:begin + ln].copy()
        for s in range(nsteps):
            q = qs[s]
            wv = in_off
            if kind == WPOCKET:
                scratch = np.full(32, np.nan)
                acc = 0.0
                for r in range(ln):
                    p = begin + r
                    h = int(hdr[p]); infl = acc if (h & 1) else 0.0
                    for _ in range((h >> 6) & 0x1ffffff):
                        x = int(inw[wv]); wv += 1
                        if x & WIN_SLOT:
                            v = ring[s, x & 0x3fffffff]; assert not np.isnan(v); infl += v; ring[s, x & 0x3fffffff] = np.nan
                        elif x & WIN_OWN:
                            assert (x & 0x3fffffff) < r; infl += Ol[x & 0x3fffffff]
                        else:
                            assert not np.isnan(scratch[x]); infl += scratch[x]; scratch[x] = np.nan
                    a, b, c, g = coef[p]
                    on = a * infl + (b * Il[r] + c * Ol[r] + g * q[p])
                    Il[r] = infl; Ol[r] = on
                    if h >> 31:
                        ring[s, int(inw[wv])] = on; wv += 1
                    sl = (h >> 1) & 31
                    if sl:
                        scratch[sl - 1] = on
                    acc = on
            else:
                ent = [int(inw[wv + t]) for t in range(n_in)]; wv += n_in
                B = 0.0
                for r in range(ln):                                   # PRE
                    p = begin + r
                    h = int(hdr[p]); side = 0.0
                    for _ in range((h >> 6) & 0x1fff):
                        x = int(inw[wv]); wv += 1
                        assert x & WIN_SLOT
                        v = ring[s, x & 0x3fffffff]; assert not np.isnan(v); side += v; ring[s, x & 0x3fffffff] = np.nan
                    a, b, c, g = coef[p]
                    infl = side + (B if (h & 1) else 0.0)
                    B = a * infl + (b * Il[r] + c * Ol[r] + g * q[p])
                    Il[r] = side; Ol[r] = B
                oin = 0.0                                             # hop
                for x in ent:
                    v = ring[s, x & 0x3fffffff]; assert not np.isnan(v); oin += v; ring[s, x & 0x3fffffff] = np.nan
                out = cumA[begin + ln - 1] * oin + Ol[ln - 1]
                if out_slot >= 0:
                    ring[s, out_slot] = out
                op = oin                                              # FIX
                for r in range(ln):
                    on = out if r == ln - 1 else cumA[begin + r] * oin + Ol[r]
                    Ol[r] = on
                    Il[r] = op + Il[r]
                    op = on
            assert wv == in_off + nwords, "input stream not consumed exactly"
            executed += 1
        O[begin:begin + ln] = Ol; I[begin:begin + ln] = Il
        done[k] = True
    assert np.isnan(ring).all(), "a published row was never consumed (the ring would not be clean for the next launch)"
    return executed
