"""Free functions of the reference's kernel module (tx_fast_hydrology/nutils.py) under their
own names and signatures, so code that imports them by name keeps working:

    _ax_bu, _ax, _apply_gain, _ap_par, _aqat_par           -> device launches through libtxh
    interpolate_sample, interpolate_samples                 -> host (scalar bookkeeping)

The routing entry points take and return host numpy arrays exactly as the numba kernels do
(fresh output arrays, inputs never mutated; nutils.py:68-70) and run on the GPU; they exist for
interface fidelity -- hot loops should use `Muskingum.run` / the C ABI with device-resident state.
"""
import numpy as np

from .network import RiverNetwork

_NETS = {}


def _net_for(endnodes):
    end = np.ascontiguousarray(endnodes, dtype=np.int64)
    key = (end.size, hash(end.tobytes()))
    net = _NETS.get(key)
    if net is None:
        if len(_NETS) > 8:
            _NETS.clear()
        net = _NETS[key] = RiverNetwork(end)
    return net


def interpolate_sample(x, xp, fp, method=1):
    """nutils.py:5-39: row interpolation of the (T x m) table `fp` at scalar `x`."""
    xp = np.asarray(xp)
    n = xp.shape[0]
    ix = int(np.searchsorted(xp, x))
    if ix == 0:
        return np.array(fp[0], dtype=np.float64)
    if ix >= n:
        return np.array(fp[n - 1], dtype=np.float64)
    dx_0 = x - xp[ix - 1]
    dx_1 = xp[ix] - x
    if method == 1:
        frac = dx_0 / (dx_0 + dx_1)
        return (1 - frac) * fp[ix - 1] + (frac) * fp[ix]
    return np.array(fp[ix - 1] if abs(dx_0) <= abs(dx_1) else fp[ix], dtype=np.float64)


def interpolate_samples(xs, xp, fp, method=1):
    """nutils.py:41-50."""
    out = np.zeros((len(xs), fp.shape[1]), dtype=np.float64)
    for i, x in enumerate(xs):
        out[i, :] = interpolate_sample(x, xp, fp, method=method)
    return out


def _step(endnodes, alpha, beta, chi, gamma, i_t_prev, o_t_prev, q):
    import torch
    net = _net_for(endnodes)
    g = np.zeros_like(alpha) if gamma is None else gamma
    net.set_coeffs(alpha, beta, chi, g)
    O = net.alloc_state(1); I = net.alloc_state(1)
    net.pack_host(np.ascontiguousarray(o_t_prev, dtype=np.float64).reshape(-1, 1), 1, O)
    net.pack_host(np.ascontiguousarray(i_t_prev, dtype=np.float64).reshape(-1, 1), 1, I)
    qd = None if q is None else torch.as_tensor(np.ascontiguousarray(q, dtype=np.float64), device='cuda')
    net.route_step(O, I, 1, qd)
    return net.unpack_host(I, 1)[:, 0], net.unpack_host(O, 1)[:, 0]


def _ax_bu(startnodes, endnodes, alpha, beta, chi, gamma, i_t_prev, o_t_prev, q_t_next, indegree):
    """nutils.py:64-89."""
    return _step(endnodes, alpha, beta, chi, gamma, i_t_prev, o_t_prev, q_t_next)


def _ax(startnodes, endnodes, alpha, beta, chi, i_t_prev, o_t_prev, indegree):
    """nutils.py:91-114."""
    return _step(endnodes, alpha, beta, chi, None, i_t_prev, o_t_prev, None)


def _apply_gain(startnodes, endnodes, gain, indegree):
    """nutils.py:116-134: o = gain, i[j] = sum of the gains of the reaches draining into j."""
    net = _net_for(endnodes)
    n = net.n
    G = net.alloc_state(1); O = net.alloc_state(1); I = net.alloc_state(1)
    net.pack_host(np.ascontiguousarray(gain, dtype=np.float64).reshape(-1, 1), 1, G)
    net.apply_gain(G, O, I, 1)
    return net.unpack_host(I, 1)[:, 0], net.unpack_host(O, 1)[:, 0]


def _ap_par(P, out, startnodes, endnodes, alpha, beta, chi, indegree):
    """nutils.py:157-169: out[:, c] = A . P[:, c]; writes into and returns the caller's `out`."""
    m, n = P.shape
    assert (m == n)
    net = _net_for(endnodes)
    net.set_coeffs(alpha, beta, chi, np.zeros_like(alpha))
    X = net.alloc_state(n); scr = net.alloc_state(n)
    net.pack_host(np.ascontiguousarray(P, dtype=np.float64), n, X)
    net.route_apply(X, scr, n)
    out[:, :] = net.unpack_host(X, n)
    return out


_ap = _ap_par


def _aqat_par(P, out, startnodes, endnodes, alpha, beta, chi, indegree):
    """nutils.py:194-214: two passes with `out = out.T` between; returns the transposed view of `out`."""
    m, n = P.shape
    assert (m == n)
    _ap_par(P, out, startnodes, endnodes, alpha, beta, chi, indegree)
    view = out.T
    res = np.empty((n, n))
    _ap_par(np.ascontiguousarray(view), res, startnodes, endnodes, alpha, beta, chi, indegree)
    view[:, :] = res
    return view


_aqat = _aqat_par
