"""Data-oblivious sketching-operator generators (OO layer over utils.sketching).

Mirrors parla/comps/sketchers/oblivious.py:25-55: ``SketchOpGen.__call__(n_rows, n_cols, rng)``.
"""
from ...utils import sketching as usk


class SketchOpGen:

    def __call__(self, n_rows, n_cols, rng):
        raise NotImplementedError()

    exec = __call__


class SkOpGA(SketchOpGen):
    """Gaussian operator generator (oblivious.py:38-45)."""

    def __init__(self, normalize=True):
        self.normalize = normalize

    def __call__(self, n_rows, n_cols, rng):
        return usk.gaussian_operator(n_rows, n_cols, rng, self.normalize)

    exec = __call__


class SkOpSJ(SketchOpGen):
    """SJLT generator (oblivious.py:48-55)."""

    def __init__(self, vec_nnz=8):
        self.vec_nnz = vec_nnz

    def __call__(self, n_rows, n_cols, rng):
        return usk.sjlt_operator(n_rows, n_cols, rng, self.vec_nnz)

    exec = __call__


class SkOpTC(SketchOpGen):
    """SRCT (subsampled randomized cosine transform) generator (oblivious.py:68-72)."""

    def __call__(self, n_rows, n_cols, rng):
        return usk.srct_operator(n_rows, n_cols, rng)

    exec = __call__


class SkOpON(SketchOpGen):
    """Orthonormal operator generator (oblivious.py:31-35)."""

    def __call__(self, n_rows, n_cols, rng):
        return usk.orthonormal_operator(n_rows, n_cols, rng)

    exec = __call__


class SkOpSS(SketchOpGen):
    """Sparse-sign operator generator (oblivious.py:58-65; the reference's __call__ forgets to return)."""

    def __init__(self, density=0.05):
        self.density = density

    def __call__(self, n_rows, n_cols, rng):
        return usk.sparse_sign_operator(n_rows, n_cols, rng, self.density)

    exec = __call__


class SkOpIN(SketchOpGen):
    """Index (row / column sampling) operator generator (oblivious.py:75-82)."""

    def __init__(self, indices=None):
        self.indices = indices

    def __call__(self, n_rows, n_cols, rng):
        return usk.sampling_operator(n_rows, n_cols, rng, self.indices)

    exec = __call__
