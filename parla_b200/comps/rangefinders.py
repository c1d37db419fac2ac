"""Rangefinders.  Mirrors parla/comps/rangefinders.py: interface (:78-123) and ``RF1`` (:126-188)."""
import warnings

import numpy as np

from .. import distla
from .sketchers.aware import RowSketcher


def rf1(A, k, num_pass, rng):
    """rangefinders.py:22-71: RS1 (Gaussian, num_pass - 1 power-iteration passes) -> RF1."""
    from .sketchers import oblivious
    from .sketchers.aware import RS1
    from ..utils import linalg_wrappers as ulaw
    rng = np.random.default_rng(rng)
    rso_ = RS1(sketch_op_gen=oblivious.SkOpGA(), num_pass=num_pass - 1, stabilizer=ulaw.orth, passes_per_stab=1)
    return RF1(rso_)(A, k, 0.0, rng)      # (:70 passes tol = 0.0, which RF1 reports as ignored)


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
