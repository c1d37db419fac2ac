"""parla/utils/linalg_wrappers.py:6-7 on the device: ``orth`` = Q factor of an economic Householder QR."""
from .. import kernels as K


def orth(S):
    return K.qr_economic(S)[0]
