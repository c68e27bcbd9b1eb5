"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the reference's iterated modified Gram-Schmidt,
``sella/utilities/math.pyx:74-140`` (``cdef mgs``) and its Python wrapper
``modified_gram_schmidt`` (``math.pyx:143-159``).

Pinned against the compiled reference itself (``oracle/ref_loader.py``) through
the fixtures in ``tests/golden/`` (``make_golden.py``).
"""
import numpy as np

MGS_SHAPE_MISMATCH = -1   # math.pyx:89-90
MGS_MAXITER = -2          # math.pyx:132-133


def _sweep(x, basis, ncols, normtot, eps2):
    """One pass of 'project out column j, renormalise' over ``ncols`` columns of
    ``basis`` (math.pyx:107-115 for Y, :118-126 for accepted X columns).

    Returns (normtot, dropped).  The running product of the post-projection
    norms is what the reference uses both as rank test (< eps2 -> drop) and as
    convergence test (1 - normtot <= eps1 -> accept).
    """
    for j in range(ncols):
        b = basis[:, j]
        x -= (b @ x) * b
        nrm = np.linalg.norm(x)
        normtot *= nrm
        if normtot < eps2:
            return normtot, True
        x /= nrm
    return normtot, False


def mgs(X, Y=None, eps1=1e-15, eps2=1e-6, maxiter=100):
    """In-place orthonormalisation of the columns of X against Y and against
    each other.  Returns the number of columns kept (>=0) or a negative code.

    Y is assumed orthonormal already (the wrapper below takes care of it).
    """
    n, nx = X.shape
    ny = 0
    if Y is not None:
        if Y.shape[0] != n:
            return MGS_SHAPE_MISMATCH
        ny = Y.shape[1]

    kept = 0
    for i in range(nx):
        if i != kept:
            X[:, kept] = X[:, i]
        x = X[:, kept]
        x /= np.linalg.norm(x)
        for _ in range(maxiter):
            normtot = 1.0
            if ny:
                normtot, _ = _sweep(x, Y, ny, normtot, eps2)
            # the rank test is applied after each of the two sweeps even when a
            # sweep was empty (math.pyx:116-117,127-128)
            if normtot >= eps2:
                normtot, _ = _sweep(x, X, kept, normtot, eps2)
            if normtot < eps2:
                break                      # slot `kept` is reused by the next column
            if 0.0 <= 1.0 - normtot <= eps1:
                kept += 1                  # a whole sweep changed nothing: accept
                break
        else:
            return MGS_MAXITER
    X[:, kept:] = 0.0                      # math.pyx:136-138
    return kept


def modified_gram_schmidt(Xin, Yin=None, eps1=1e-15, eps2=1e-6, maxiter=100):
    """math.pyx:143-159: copies, orthonormalises Y first, raises on failure."""
    if Xin.shape[1] == 0:
        return Xin
    Yout = None
    if Yin is not None:
        Yout = np.array(Yin, dtype=np.float64, order="C", copy=True)
        ny = mgs(Yout, None, eps1=eps1, eps2=eps2, maxiter=maxiter)
        Yout = Yout[:, :ny]
    Xout = np.array(Xin, dtype=np.float64, order="C", copy=True)
    nx = mgs(Xout, Yout, eps1=eps1, eps2=eps2, maxiter=maxiter)
    if nx < 0:
        raise RuntimeError("MGS failed.")
    return Xout[:, :nx]
