"""Rangefinders.  Mirrors parla/comps/rangefinders.py: interface (:78-123) and ``RF1`` (:126-188)."""
import warnings

import numpy as np

from .. import distla
from .sketchers.aware import RowSketcher


class RangeFinder:

    def __call__(self, A, k, tol, rng):
        raise NotImplementedError()

    exec = __call__


class RF1(RangeFinder):

    def __init__(self, rso: RowSketcher):
        self.rso = rso

    def __call__(self, A, k, tol, rng):
        assert k > 0                                               # rangefinders.py:176-177
        assert k <= min(A.shape)
        if not np.isnan(tol):
            msg = """
            This RangeFinder implementation cannot directly control
            approximation error. Parameter "tol" is being ignored.
            """
            warnings.warn(msg)
        rng = np.random.default_rng(rng)
        S = self.rso(A, k, rng)
        Y = distla.mm(A, S)                                        # :186
        return distla.orth(Y)                                      # :187

    exec = __call__
