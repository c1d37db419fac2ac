"""parla/utils/linalg_wrappers.py:6-7 on the device: ``orth`` = Q factor of an economic Householder QR
(a Householder TSQR when the argument is ``RowSharded``)."""
from .. import distla


def orth(S):
    return distla.orth(S)
