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
