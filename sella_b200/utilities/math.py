"""CUDA mirror of sella/utilities/math.pyx (the part on the hot path).

``modified_gram_schmidt(Xin, Yin=None, eps1, eps2, maxiter)`` has the reference's
signature, return value (an (n, nx_kept) array) and error behaviour
(``RuntimeError("MGS failed.")``), sella/utilities/math.pyx:143-159; the work is done by
``sb_mgs`` (csrc/subspace.cu)."""
import numpy as np

from .. import kernels as K
from .._host import up_cols, down_cols


def modified_gram_schmidt(Xin, Yin=None, eps1=1.e-15, eps2=1.e-6, maxiter=100):
    Xin = np.asarray(Xin, dtype=np.float64)
    if Xin.shape[1] == 0:
        return Xin
    X = up_cols(Xin)
    Y = None
    if Yin is not None and np.asarray(Yin).shape[1] > 0:
        if np.asarray(Yin).shape[0] != Xin.shape[0]:
            raise RuntimeError("MGS failed.")          # mgs returns -1 on a shape mismatch
        Y = up_cols(Yin)
    nkept, status = K.mgs(X, Y, eps1=eps1, eps2=eps2, maxiter=maxiter)
    nx = int(nkept[0])
    if nx < 0:
        raise RuntimeError("MGS failed.")
    return down_cols(X, nx)
