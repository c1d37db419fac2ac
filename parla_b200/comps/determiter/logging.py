"""Runtime / error log of a sketch-and-precondition solve.

Field-for-field mirror of parla/comps/determiter/logging.py:4-94 (tests read ``log.errors``); the
extra fields (``iters``, ``istop``, ``passes_over_A``) are additions.
"""
import numpy as np


class SketchAndPrecondLog:

    def __init__(self):
        self.time_sketch = 0.0
        self.time_factor = 0.0
        self.time_presolve = 0.0
        self.time_convert = 0.0
        self._time_setup = 0.0
        self.time_iterate = 0.0
        self.times = None
        self.errors = None
        self.error_desc = """Fill in."""
        self.iters = 0
        self.istop = 0
        self.passes_over_A = 0

    @property
    def time_setup(self):
        # logging.py:63-70
        self._time_setup = self.time_sketch + self.time_factor + self.time_convert
        return self._time_setup

    def wrap_up(self, iter_errors, init_error):
        # logging.py:72-94
        iter_errors = np.atleast_1d(np.asarray(iter_errors, dtype=float))
        setup = self.time_setup
        ramp = np.linspace(0, self.time_iterate, iter_errors.size, endpoint=True)
        self.times = np.concatenate(([setup], setup + self.time_presolve + ramp))
        self.errors = np.concatenate(([init_error], iter_errors))
