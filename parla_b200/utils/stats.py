"""Convergence-rate fits used by the reference's tests and notebooks (parla/utils/stats.py:6-55).  Host-side
numpy on the logged error histories (``log.errors``)."""
import warnings

import numpy as np


def _fit(design, logy):
    coef = np.linalg.lstsq(design, logy, rcond=None)[0]
    ss_tot = np.sum((logy - np.mean(logy)) ** 2)
    ss_res = np.sum((logy - design @ coef) ** 2)
    return coef, 1 - ss_res / ss_tot


def _positive(x, y):
    x, y = np.asarray(x, dtype=float).ravel(), np.asarray(y, dtype=float).ravel()
    assert x.size == y.size
    if np.any(y <= 0):
        warnings.warn('Dropping samples "i" where y[i] <= 0.')
        x, y = x[y > 0], y[y > 0]
    return x, y


def loglinear_fit(x, y):
    """Least-squares fit log(y) ~ a + b x; returns ([a, b], R^2)  (stats.py:6-28)."""
    x, y = _positive(x, y)
    return _fit(np.column_stack([np.ones(x.size), x]), np.log(y))


def loglog_fit(x, y):
    """Least-squares fit log(y) ~ a + b log(x); returns ([a, b], R^2)  (stats.py:31-55)."""
    x, y = _positive(x, y)
    if np.any(x <= 0):
        raise ValueError('Input x must be positive.')
    return _fit(np.column_stack([np.ones(x.size), np.log(x)]), np.log(y))
