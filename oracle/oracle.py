"""CPU parity oracle for the routing + assimilation hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package
(`tx_fast_hydrology_b200/`).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module.

Parity status: PINNED against the reference executed live in the build
container (`tests/golden/make_golden.py` -> `tests/golden/*.npz`, replayed by
`tests/test_oracle.py`).  The reference has no golden vectors of its own
(SURVEY.md section 4 / 8c).

Two layers:
  * thin ctypes wrappers over `oracle/txh_oracle.c` (the C restatement of the
    numba kernels) carrying the reference's own names and signatures
    (`nutils.py:64-214`), and
  * numpy restatements of the host-side logic: `Muskingum.step_iter` /
    `simulate_iter` (`muskingum.py:435-536`), `KalmanFilter.filter`
    (`da.py:91-136`) and the ensemble (sample-covariance) form of the same
    algebra, which the reference does not have.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    so = os.path.join(_HERE, "libtxh_oracle.so")
    src = os.path.join(_HERE, "txh_oracle.c")
    if force or not os.path.exists(so) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.run(["make", "-s", "-C", _HERE, "-B"], check=True)
    return so


def build_ref():
    """Copy the unmodified reference package into oracle/_ref/ (only where /root/reference exists: the build
    container).  Returns the directory, or None when there is no copy."""
    subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=False, stdout=subprocess.DEVNULL)
    return ref_dir()


def ref_dir():
    d = os.path.join(_HERE, "_ref")
    return d if os.path.exists(os.path.join(d, "tx_fast_hydrology", "nutils.py")) else None


def reference_nutils():
    """The reference's own `tx_fast_hydrology.nutils` (numba kernels) from oracle/_ref, or None."""
    import importlib
    import sys
    d = ref_dir()
    if d is None:
        return None
    try:
        if d not in sys.path:
            sys.path.insert(0, d)
        return importlib.import_module("tx_fast_hydrology.nutils")
    except Exception:
        return None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.txo_visit_order.restype = ctypes.c_int64
        _LIB.txo_run_members.restype = ctypes.c_int64
        _LIB.txo_max_threads.restype = ctypes.c_int
    return _LIB


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_i64p)


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_f64p)


def max_threads():
    """Host threads this process may use: its CPU affinity (launchers such as torchrun export OMP_NUM_THREADS=1,
    which would otherwise pin the oracle to one core; every parallel region here names its thread count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, int(lib().txo_max_threads()))


# --------------------------------------------------------------------------
# nutils.py kernels (same names, same argument order, same return values)
# --------------------------------------------------------------------------
def _ax_bu(startnodes, endnodes, alpha, beta, chi, gamma, i_t_prev, o_t_prev,
           q_t_next, indegree):
    """nutils.py:64-89."""
    n = endnodes.size
    sn, snp = _i(startnodes); en, enp = _i(endnodes); ind, indp = _i(indegree)
    a, ap = _f(alpha); b, bp = _f(beta); c, cp = _f(chi); g, gp = _f(gamma)
    ip, ipp = _f(i_t_prev); op, opp = _f(o_t_prev); q, qp = _f(q_t_next)
    i_n = np.empty(n); o_n = np.empty(n); work = np.empty(n, dtype=np.int64)
    lib().txo_ax_bu(ctypes.c_int64(n), ctypes.c_int64(sn.size), snp, enp, ap, bp, cp,
                    gp, ipp, opp, qp, indp, i_n.ctypes.data_as(_f64p),
                    o_n.ctypes.data_as(_f64p), work.ctypes.data_as(_i64p))
    return i_n, o_n


def _ax(startnodes, endnodes, alpha, beta, chi, i_t_prev, o_t_prev, indegree):
    """nutils.py:91-114."""
    n = endnodes.size
    sn, snp = _i(startnodes); en, enp = _i(endnodes); ind, indp = _i(indegree)
    a, ap = _f(alpha); b, bp = _f(beta); c, cp = _f(chi)
    ip, ipp = _f(i_t_prev); op, opp = _f(o_t_prev)
    i_n = np.empty(n); o_n = np.empty(n); work = np.empty(n, dtype=np.int64)
    lib().txo_ax(ctypes.c_int64(n), ctypes.c_int64(sn.size), snp, enp, ap, bp, cp,
                 ipp, opp, indp, i_n.ctypes.data_as(_f64p),
                 o_n.ctypes.data_as(_f64p), work.ctypes.data_as(_i64p))
    return i_n, o_n


def _apply_gain(startnodes, endnodes, gain, indegree):
    """nutils.py:116-134."""
    n = endnodes.size
    sn, snp = _i(startnodes); en, enp = _i(endnodes); ind, indp = _i(indegree)
    g, gp = _f(gain)
    i_n = np.empty(n); o_n = np.empty(n); work = np.empty(n, dtype=np.int64)
    lib().txo_apply_gain(ctypes.c_int64(n), ctypes.c_int64(sn.size), snp, enp, gp,
                         indp, i_n.ctypes.data_as(_f64p), o_n.ctypes.data_as(_f64p),
                         work.ctypes.data_as(_i64p))
    return i_n, o_n


def numba_init_inflows(a, indices, b):
    """nutils.py:136-141 (in place on `a`; no self-loop guard)."""
    assert a.dtype == np.float64 and a.flags.c_contiguous
    idx, idxp = _i(indices); bb, bp = _f(b)
    lib().txo_init_inflows(ctypes.c_int64(idx.size), a.ctypes.data_as(_f64p), idxp, bp)


def _ap_par(P, out, startnodes, endnodes, alpha, beta, chi, indegree, threads=0):
    """nutils.py:157-169.  Writes into the caller's `out` and returns it."""
    m, n = P.shape
    assert m == n
    assert out.flags.c_contiguous and out.dtype == np.float64
    PP, Pp = _f(P)
    sn, snp = _i(startnodes); en, enp = _i(endnodes); ind, indp = _i(indegree)
    a, ap = _f(alpha); b, bp = _f(beta); c, cp = _f(chi)
    lib().txo_ap(ctypes.c_int64(n), ctypes.c_int64(sn.size), Pp,
                 out.ctypes.data_as(_f64p), snp, enp, ap, bp, cp, indp,
                 ctypes.c_int(threads or max_threads()))
    return out


def _aqat_par(P, out, startnodes, endnodes, alpha, beta, chi, indegree, threads=0):
    """nutils.py:194-214.  Returns the matrix the reference's caller sees (the
    `out.T` view), as a fresh C-order array; `out` is used as scratch."""
    m, n = P.shape
    assert m == n
    PP, Pp = _f(P)
    sn, snp = _i(startnodes); en, enp = _i(endnodes); ind, indp = _i(indegree)
    a, ap = _f(alpha); b, bp = _f(beta); c, cp = _f(chi)
    res = np.empty((n, n))
    tmp = out if (out.flags.c_contiguous and out.dtype == np.float64) else np.empty((n, n))
    lib().txo_aqat(ctypes.c_int64(n), ctypes.c_int64(sn.size), Pp,
                   res.ctypes.data_as(_f64p), tmp.ctypes.data_as(_f64p), snp, enp,
                   ap, bp, cp, indp, ctypes.c_int(threads or max_threads()))
    return res


def interpolate_sample(x, xp, fp, method=1):
    """nutils.py:5-39."""
    xpp, xpc = _f(xp); fpp, fpc = _f(fp)
    T, m = fpp.shape
    res = np.empty(m)
    lib().txo_interpolate_sample(ctypes.c_double(x), ctypes.c_int64(T),
                                 ctypes.c_int64(m), xpc, fpc, ctypes.c_int(method),
                                 res.ctypes.data_as(_f64p))
    return res


def interpolate_samples(xs, xp, fp, method=1):
    """nutils.py:41-50."""
    return np.stack([interpolate_sample(float(x), xp, fp, method) for x in xs])


def compute_indegree(startnodes, endnodes):
    """muskingum.py:322-330."""
    sn, snp = _i(startnodes); en, enp = _i(endnodes)
    out = np.empty(sn.size, dtype=np.int64)
    lib().txo_indegree(ctypes.c_int64(sn.size), snp, enp, out.ctypes.data_as(_i64p))
    return out


def compute_coeffs(K, X, dt):
    """muskingum.py:332-360 -> (alpha, beta, chi, gamma)."""
    k, kp = _f(K); x, xp = _f(X)
    n = k.size
    out = [np.empty(n) for _ in range(4)]
    lib().txo_coeffs(ctypes.c_int64(n), kp, xp, ctypes.c_double(dt),
                     *[o.ctypes.data_as(_f64p) for o in out])
    return tuple(out)


def init_states(startnodes, endnodes, o_t_next):
    """muskingum.py:410-419 -> i_t_next (self-loop inflow included)."""
    sn, snp = _i(startnodes); en, enp = _i(endnodes); o, op = _f(o_t_next)
    i = np.empty(sn.size)
    lib().txo_init_states(ctypes.c_int64(sn.size), snp, enp, op, i.ctypes.data_as(_f64p))
    return i


def visit_order(startnodes, endnodes, indegree):
    """Order in which nutils.py:72-88 evaluates reaches (ascending headwaters)."""
    sn, snp = _i(startnodes); en, enp = _i(endnodes); ind, indp = _i(indegree)
    heads = sn[ind == 0]
    hd, hdp = _i(heads)
    n = en.size
    order = np.empty(n, dtype=np.int64); work = np.empty(n, dtype=np.int64)
    c = lib().txo_visit_order(ctypes.c_int64(n), ctypes.c_int64(hd.size), hdp, enp, indp,
                              order.ctypes.data_as(_i64p), work.ctypes.data_as(_i64p))
    assert c == n, "not a forest of in-trees"
    return order


def run_members(net, o_state, i_state, nsteps, xp, fp, t0, dt_ns, wmul=None, threads=0):
    """Member-batched CPU run (muskingum.py:499-533 per member, members over threads).
    `o_state`, `i_state` are [M][n] and are advanced in place.  Returns #updates."""
    sn, snp = _i(net["startnodes"]); en, enp = _i(net["endnodes"])
    ind, indp = _i(net["indegree"])
    heads, hp = _i(sn[ind == 0])
    a, ap = _f(net["alpha"]); b, bp = _f(net["beta"]); c, cp = _f(net["chi"])
    g, gp = _f(net["gamma"])
    x, xpp = _f(xp); f, fpp = _f(fp)
    assert o_state.flags.c_contiguous and i_state.flags.c_contiguous
    M, n = o_state.shape
    if wmul is not None:
        w, wp = _f(wmul)
        assert w.shape == (f.shape[0], M)
    else:
        wp = _f64p()
    return int(lib().txo_run_members(
        ctypes.c_int64(n), ctypes.c_int64(heads.size), ctypes.c_int64(M),
        ctypes.c_int64(nsteps), hp, enp, ap, bp, cp, gp, indp,
        ctypes.c_int64(f.shape[0]), xpp, fpp, wp, ctypes.c_double(t0),
        ctypes.c_double(dt_ns), o_state.ctypes.data_as(_f64p),
        i_state.ctypes.data_as(_f64p), ctypes.c_int(threads or max_threads())))


# --------------------------------------------------------------------------
# Pure-Python restatement (small cases; cross-checks the C file)
# --------------------------------------------------------------------------
def py_ax_bu(startnodes, endnodes, alpha, beta, chi, gamma, i_t_prev, o_t_prev,
             q_t_next, indegree):
    """nutils.py:64-89, line by line, in interpreted Python."""
    n = endnodes.size
    i_t_next = np.zeros(n); o_t_next = np.zeros(n)
    indegree_t = indegree.copy()
    for k in range(startnodes.size):
        s = int(startnodes[k]); e = int(endnodes[s])
        while indegree_t[s] == 0:
            o_t_next[s] += (alpha[s] * i_t_next[s] + beta[s] * i_t_prev[s]
                            + chi[s] * o_t_prev[s] + gamma[s] * q_t_next[s])
            if s != e:
                i_t_next[e] += o_t_next[s]
            indegree_t[e] -= 1
            s = e; e = int(endnodes[s])
    return i_t_next, o_t_next


# --------------------------------------------------------------------------
# Topology definitions (last row)
# --------------------------------------------------------------------------
def levels(endnodes):
    """level[j] = 0 for headwaters else 1 + max(level[u] : u -> j), u != j."""
    en = np.asarray(endnodes, dtype=np.int64)
    n = en.size
    sn = np.arange(n, dtype=np.int64)
    ind = compute_indegree(sn, en)
    order = visit_order(sn, en, ind)
    lev = np.zeros(n, dtype=np.int64)
    for j in order:
        e = en[j]
        if e != j and lev[e] < lev[j] + 1:
            lev[e] = lev[j] + 1
    return lev


# --------------------------------------------------------------------------
# Host-side model logic
# --------------------------------------------------------------------------
class OracleModel:
    """State + step/simulate semantics of `Muskingum` (muskingum.py:139-177,
    435-536), with the callback firing order of the reference."""

    def __init__(self, startnodes, endnodes, K, X, o_t, dt, t0_ns=0):
        self.startnodes = np.asarray(startnodes, dtype=np.int64)
        self.endnodes = np.asarray(endnodes, dtype=np.int64)
        self.n = self.startnodes.size
        self.K = np.asarray(K, dtype=np.float64)
        self.X = np.asarray(X, dtype=np.float64)
        self.dt = float(dt)
        self.time_ns = int(t0_ns)
        self.indegree = compute_indegree(self.startnodes, self.endnodes)
        self.o_t_next = np.array(o_t, dtype=np.float64)
        # muskingum.py:159-161: init_states then prev[:] = next[:]
        self.i_t_next = init_states(self.startnodes, self.endnodes, self.o_t_next)
        self.o_t_prev = self.o_t_next.copy()
        self.i_t_prev = self.i_t_next.copy()
        self.alpha, self.beta, self.chi, self.gamma = compute_coeffs(self.K, self.X, self.dt)
        self.callbacks = {}

    @property
    def heads(self):
        return self.startnodes[self.indegree == 0]      # muskingum.py:444

    def step(self, p_t_next):
        """muskingum.py:435-465."""
        o_prev, i_prev = self.o_t_next, self.i_t_next
        for cb in self.callbacks.values():
            cb.on_step_start()
        i_n, o_n = _ax_bu(self.heads, self.endnodes, self.alpha, self.beta, self.chi,
                          self.gamma, i_prev, o_prev, p_t_next, self.indegree)
        self.o_t_next, self.o_t_prev = o_n, o_prev
        self.i_t_next, self.i_t_prev = i_n, i_prev
        self.time_ns += int(round(self.dt * 1e9))
        for cb in self.callbacks.values():
            cb.on_step_end()

    def simulate(self, times_ns, table, end_ns=None):
        """muskingum.py:499-536: generator over steps; forcing sampled at t+dt."""
        xp = np.asarray(times_ns, dtype=np.int64).astype(np.float64)
        if end_ns is None:
            end_ns = int(times_ns[-1])
        for cb in self.callbacks.values():
            cb.on_simulation_start()
        step_ns = int(round(self.dt * 1e9))
        while self.time_ns < end_ns:
            p = interpolate_sample(float(self.time_ns + step_ns), xp, table)
            self.step(p)
            yield self
        for cb in self.callbacks.values():
            cb.on_simulation_end()


class OracleKalmanFilter:
    """`KalmanFilter` (da.py:14-136) restated on an OracleModel.  `reach_indices`
    are the gauged reach indices in measurement-column order; columns/R are
    permuted to ascending index order as da.py:36-44 does."""

    def __init__(self, model, meas_times_ns, meas, reach_indices, Q_cov, R_cov, P_t_init,
                 every_ns=None):
        self.model = model
        reach_indices = np.asarray(reach_indices, dtype=np.int64)
        perm = np.argsort(reach_indices)
        self.reach_indices = reach_indices[perm]
        self.meas_times = np.asarray(meas_times_ns, dtype=np.int64).astype(np.float64)
        self.latest_ns = int(meas_times_ns[-1])
        self.meas = np.ascontiguousarray(np.asarray(meas, dtype=np.float64)[:, perm])
        self.R_cov = np.asarray(R_cov, dtype=np.float64)[perm, :][:, perm]
        self.Q_cov = Q_cov
        self.P_t_next = P_t_init
        s = np.zeros(model.n, dtype=bool)
        s[self.reach_indices] = True
        self.s = s
        self.every_ns = every_ns          # None = every step (the reference)

    def _due(self):
        if self.model.time_ns > self.latest_ns:       # da.py:51,58
            return False
        if self.every_ns is not None and (self.model.time_ns % self.every_ns) != 0:
            return False
        return True

    def on_simulation_start(self):
        if self._due():
            self.filter()

    def on_step_start(self):
        pass

    def on_step_end(self):
        if self._due():
            self.filter()

    def on_simulation_end(self):
        pass

    def filter(self):
        """da.py:91-136."""
        mdl = self.model
        P_prev = self.P_t_next
        s = self.s
        Z = interpolate_sample(float(mdl.time_ns), self.meas_times, self.meas)
        dz = Z - mdl.o_t_next[s]
        out = np.empty(P_prev.shape)
        P = _aqat_par(P_prev, out, mdl.heads, mdl.endnodes, mdl.alpha, mdl.beta,
                      mdl.chi, mdl.indegree)
        P += self.Q_cov
        K = P[:, s] @ np.linalg.inv(P[s][:, s] + self.R_cov)
        gain = K @ dz
        P = P - K @ P[s]
        i_g, o_g = _apply_gain(mdl.heads, mdl.endnodes, gain, mdl.indegree)
        mdl.i_t_next += i_g
        mdl.o_t_next += o_g
        self.P_t_next, self.P_t_prev = P, P_prev
        self.K, self.dz, self.gain = K, dz, gain


def enkf_update(net, O, I, obs_idx, Zp, Q_diag, R_cov):
    """Ensemble form of da.py:112-126 (the reference has no ensemble filter;
    SURVEY.md section 8c row 'New capability').

    O, I      : [n][M] forecast outflows / inflows, one column per member
    obs_idx   : gauged reach indices, ascending (da.py:33-44 ordering)
    Zp        : [m][M] per-member (perturbed) observations
    Q_diag    : [n] diagonal of the model-noise covariance Q
    R_cov     : [m][m]
    P := sample covariance of the forecast ensemble + Q  replaces A P A^T + Q
    (da.py:115-117); gain / innovation / gain application follow da.py:112,
    119-126 per member.  Returns (O_post, I_post, K).
    """
    O = np.asarray(O, dtype=np.float64); I = np.asarray(I, dtype=np.float64)
    n, M = O.shape
    s = np.asarray(obs_idx, dtype=np.int64)
    mean = O.sum(axis=1) / M
    A = O - mean[:, None]                                    # anomalies
    HA = A[s]
    dz = Zp - O[s]                                           # da.py:112 per member
    P_xs = A @ HA.T / (M - 1)                                # P[:, s]
    P_xs[s, np.arange(s.size)] += Q_diag[s]                  # + Q[:, s]  (Q diagonal)
    S = P_xs[s] + R_cov                                      # P[s][:, s] + R
    K = P_xs @ np.linalg.inv(S)                              # da.py:119
    gain = K @ dz                                            # da.py:121, per member
    heads = net["startnodes"][net["indegree"] == 0]
    O_post = O.copy(); I_post = I.copy()
    for k in range(M):
        i_g, o_g = _apply_gain(heads, net["endnodes"], np.ascontiguousarray(gain[:, k]),
                               net["indegree"])              # da.py:124
        I_post[:, k] += i_g                                  # da.py:125
        O_post[:, k] += o_g                                  # da.py:126
    return O_post, I_post, K
