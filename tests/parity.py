"""Error metrics of the parity tests (test infrastructure).

relerr  : ELEMENT-WISE relative error with an absolute floor,
              max_k |a_k - b_k| / max(|b_k|, FLOOR * max|b|),   FLOOR = 1e-12,
          used for every state comparison (outflows, inflows, hydrographs).  The bar is BASELINE.json's
          north_star: <= 1e-9 in FP64.  The floor only protects exact zeros and denormal-scale entries: an
          element 1e-12 of the largest one still has to agree to 1e-9 of ITS OWN magnitude.  After an ensemble
          update the element is compared to 1e-9 of max(|itself|, |its forecast|): see relerr(scale=...).
normerr : max-norm relative error, max|a - b| / max|b|, for dense matrices whose small entries are
          differences of O(1) numbers (covariances, Kalman gains): an element-wise bound is not attainable
          there by any two correct FP64 implementations (LAPACK vs. the device solve), see the comments at
          the call sites.
"""
import numpy as np

FLOOR = 1e-12


def relerr(a, b, floor=FLOOR, scale=None):
    """`scale` (optional, same shape as b): magnitudes of the TERMS the element was summed from.  An ensemble
    update forms o + gain, and where the gain cancels the forecast (observed: |o + gain| down to 5e-11 of |o|) no
    FP64 implementation, the reference's included, can hold the result to 1e-9 of its own size -- only to 1e-9 of
    the forecast it was computed from.  The denominator is then max(|b|, scale, floor * max|b|)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    top = float(np.abs(b).max())
    if top == 0.0:
        return float(np.abs(a).max())
    den = np.maximum(np.abs(b), floor * top)
    if scale is not None:
        den = np.maximum(den, np.abs(scale))
    return float((np.abs(a - b) / den).max())


def normerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()) / max(1e-300, float(np.abs(b).max()))


def enkf_tolerance(O_forecast, gauges, q_diag, R_cov, base=1e-9):
    """Element-wise tolerance for a state that went through an ensemble update.  Both sides solve with the
    innovation covariance S = HA HA^T/(M-1) + Q[s,s] + R -- the oracle by an explicit inverse exactly as the
    reference does (da.py:119), the device by a Cholesky / Woodbury solve -- so each carries a forward error of
    order eps * cond(S) RELATIVE TO THE GAIN ROW, whatever the size of the element.  The bound is
    max(1e-9, 64 eps cond(S)); O_forecast is [n][M]."""
    O = np.asarray(O_forecast, dtype=np.float64)
    M = O.shape[1]
    HA = O[gauges] - O[gauges].mean(axis=1, keepdims=True)
    S = HA @ HA.T / (M - 1) + np.diag(np.broadcast_to(q_diag, (O.shape[0],))[gauges]) + np.asarray(R_cov)
    return max(base, 64.0 * np.finfo(np.float64).eps * float(np.linalg.cond(S)))
