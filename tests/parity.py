"""Error metrics of the parity tests (test infrastructure).

relerr  : ELEMENT-WISE relative error with an absolute floor,
              max_k |a_k - b_k| / max(|b_k|, FLOOR * max|b|),   FLOOR = 1e-12,
          used for every state comparison (outflows, inflows, hydrographs).  The bar is BASELINE.json's
          north_star: <= 1e-9 in FP64.  The floor only protects exact zeros and denormal-scale entries: an
          element 1e-12 of the largest one still has to agree to 1e-9 of ITS OWN magnitude.
normerr : max-norm relative error, max|a - b| / max|b|, for dense matrices whose small entries are
          differences of O(1) numbers (covariances, Kalman gains): an element-wise bound is not attainable
          there by any two correct FP64 implementations (LAPACK vs. the device solve), see the comments at
          the call sites.
"""
import numpy as np

FLOOR = 1e-12


def relerr(a, b, floor=FLOOR):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    scale = float(np.abs(b).max())
    if scale == 0.0:
        return float(np.abs(a).max())
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor * scale)).max())


def normerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()) / max(1e-300, float(np.abs(b).max()))
