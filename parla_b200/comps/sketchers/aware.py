"""Data-aware row sketching by power iteration.

Mirrors parla/comps/sketchers/aware.py: ``RowSketcher`` interface (:33-73) and ``RS1`` (:76-184).
All products with A are FP64 DMMA GEMMs (pla_gemm_f64); the stabiliser is normally
``parla_b200.utils.linalg_wrappers.orth`` (Householder QR on the device).
"""
import numpy as np

from ... import distla
from ...parallel import RowSharded
from ...utils.sketching import as_device_operator, shard_context


def _dense(S):
    return as_device_operator(S).to_dense()


def _tall_test_matrix(gen, A, k, rng):
    """gen(m, k, rng) restricted to this rank's rows of A (row-sharded A) or in full."""
    if not isinstance(A, RowSharded):
        return _dense(gen(A.shape[0], k, rng))
    S = as_device_operator(gen(A.shape[0], k, rng))
    if hasattr(S, "seed"):                               # virtual Gaussian: generate only the local rows
        from ... import kernels as K
        loc = K.philox_normal_fill(A.local.shape[0], k, S.seed, S.scale, row_offset=A.row_offset, device=A.device)
    else:
        loc = S.to_dense()[A.row_offset:A.row_offset + A.local.shape[0]].contiguous()
    return RowSharded(loc, A.row_offset, A.m_global, A.group)


def rs1(A, k, num_pass, rng, stabilizer=None, passes_per_stab=1, sketch_op_gen=None):
    """Procedural wrapper (aware.py:9-30): Gaussian start, ``num_pass`` power-iteration passes, QR-stabilised
    after every pass.  (The extra keyword arguments are conveniences of this package.)"""
    from . import oblivious
    from ...utils import linalg_wrappers as ulaw
    assert num_pass >= 0
    assert k >= 1
    assert k <= min(A.shape)
    gen = oblivious.SkOpGA() if sketch_op_gen is None else sketch_op_gen
    return RS1(gen, num_pass, ulaw.orth if stabilizer is None else stabilizer, passes_per_stab)(A, k, rng)


class RowSketcher:

    def __call__(self, A, k, rng):
        raise NotImplementedError()

    exec = __call__


class RS1(RowSketcher):

    def __init__(self, sketch_op_gen, num_pass, stabilizer, passes_per_stab):
        self.sketch_op_gen = sketch_op_gen
        self.num_pass = num_pass
        self.stabilizer = stabilizer
        self.passes_per_stab = passes_per_stab

    def __call__(self, A, k, rng):
        assert self.num_pass >= 0                                   # aware.py:160
        rng = np.random.default_rng(rng)
        passes_done = 0
        if self.num_pass % 2 == 0:                                  # :163-164
            S = _dense(self.sketch_op_gen(A.shape[1], k, rng))
        else:                                                       # :165-169
            S = distla.mm_t(A, _tall_test_matrix(self.sketch_op_gen, A, k, rng))
            passes_done += 1
            if self.passes_per_stab == 1:
                S = self.stabilizer(S)
        q = (self.num_pass - passes_done) // 2
        while q > 0:                                                # :174-183
            S = distla.mm(A, S)
            passes_done += 1
            if passes_done % self.passes_per_stab == 0:
                S = self.stabilizer(S)
            S = distla.mm_t(A, S)
            passes_done += 1
            if passes_done % self.passes_per_stab == 0:
                S = self.stabilizer(S)
            q -= 1
        return S

    exec = __call__
