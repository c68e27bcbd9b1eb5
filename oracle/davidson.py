"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the reference's Jacobi-Davidson / Rayleigh-Ritz partial
diagonaliser, ``sella/eigensolvers.py``:

  exact          eigensolvers.py:9-28
  rayleigh_ritz  eigensolvers.py:31-112
  expand         eigensolvers.py:115-153   (lanczos, gd, jd0, jd0_alt, mjd0, mjd0_alt)

The metric matrix ``B`` of the reference (always the identity at every call site
in the package: peswrapper.py:537-539) is kept as an argument so the generalised
small eigenproblem ``eigh(V^T Y~, V^T B V)`` is reproduced literally.

Pinned against the reference through ``tests/golden`` (make_golden.py).
"""
import numpy as np
from scipy.linalg import eigh, solve

from .orth import modified_gram_schmidt
from .hessian import symmetrize_Y


def exact(A, gamma=None, P=None):
    """Dense fall-back (eigensolvers.py:9-28).  For an operator A the matrix is
    rebuilt by probing with the *rows* of P's eigenvector matrix (:23-25; any
    orthonormal basis gives the same sum) and symmetrised."""
    if isinstance(A, np.ndarray):
        lams, vecs = eigh(A)
    else:
        n = A.shape[0]
        probes = np.eye(n) if P is None else exact(P)[1]
        dense = np.zeros((n, n))
        for row in probes:
            dense += np.outer(row, A.dot(row))
        lams, vecs = eigh(0.5 * (dense + dense.T))
    return lams, vecs, lams[None, :] * vecs


def expand(V, Y, P, B, lams, vecs, shift, method="jd0", seeking=0):
    """Correction vector for the Ritz pair ``seeking`` (eigensolvers.py:115-153)."""
    d, k = V.shape
    Vr = V @ vecs
    R = Y @ vecs - (B @ Vr) * lams[None, :]
    r = R[:, seeking]
    M = P - shift * B
    if method == "lanczos":
        return r
    if method == "gd":
        return np.linalg.solve(M, r)
    if method == "jd0":
        # bordered system [[M, v], [v^T, 0]] [t; eps] = -[r; 0]   (:133-139)
        v = Vr[:, seeking]
        K = np.block([[M, v[:, None]], [v[None, :], np.zeros((1, 1))]])
        rhs = np.zeros(d + 1)
        rhs[:d] = -r
        return solve(K, rhs)[:d]
    if method == "jd0_alt":
        v = Vr[:, seeking]
        Mr = solve(M, r)
        Mv = solve(M, v)
        den = v @ Mv
        if abs(den) < 1e-12:
            return Mr
        return Mv * (v @ Mr / den) - Mr
    if method == "mjd0":
        K = np.block([[M, Vr], [Vr.T, np.zeros((k, k))]])
        rhs = np.zeros(d + k)
        rhs[:d] = -r
        return solve(K, rhs)[:d]
    if method == "mjd0_alt":
        Mr = solve(M, r)
        MV = solve(M, Vr)
        coef = solve(Vr.T @ MV, Vr.T @ Mr)
        return solve(M, Vr @ coef - r)
    raise ValueError("Unknown diagonalization method {}".format(method))


def rayleigh_ritz(A, gamma, P, B=None, v0=None, vref=None, vreftol=0.99,
                  method="jd0", maxiter=None, trace=None):
    """eigensolvers.py:31-112.

    ``trace`` (oracle-only, optional list) receives one dict per outer iteration
    with the rotated Ritz values / residual norms / chosen target, so the GPU
    iterates can be compared step by step.
    """
    n = A.shape[0]
    if B is None:
        B = np.eye(n)
    if maxiter is None:
        maxiter = 2 * n + 1
    if gamma <= 0:
        return exact(A, gamma, P)

    if v0 is not None:
        V = modified_gram_schmidt(v0.reshape((-1, 1)))
    else:
        # negative-curvature eigenvectors of the preconditioner (at least one)
        P_lams, P_vecs, _ = exact(P, 0)
        nneg = max(1, int(np.sum(P_lams < 0)))
        V = modified_gram_schmidt(P_vecs[:, :nneg])
    AV = A.dot(V)

    symm = 2
    while True:
        k = V.shape[1]
        Asub = V.T @ symmetrize_Y(V, AV, symm=symm)
        lams, rot = eigh(Asub, V.T @ B @ V)          # lower triangles only
        nneg = max(1, int(np.sum(lams < 0)))
        AV = AV @ rot                                # rotate the basis to Ritz vectors
        V = V @ rot
        if k >= min(n, maxiter):
            return lams, V, AV

        Yt = symmetrize_Y(V, AV, symm=symm)
        R = Yt[:, :nneg] - (B @ V[:, :nneg]) * lams[None, :nneg]
        Rnorm = np.linalg.norm(R, axis=0)

        if vref is not None and abs(V[:, 0] @ vref) > vreftol:
            return lams, V, AV

        target = None
        for idx in range(nneg):
            if k == 1 or Rnorm[idx] >= gamma * abs(lams[idx]):
                target = idx
                break
        if trace is not None:
            trace.append(dict(k=k, lams=lams.copy(), rnorm=Rnorm.copy(), target=target))
        if target is None:
            return lams, V, AV
        ri = R[:, target]

        t = expand(V, Yt, P, B, lams, np.eye(k), lams[target], method, target)
        t = t / np.linalg.norm(t)
        if np.linalg.norm(t - V @ (V.T @ t)) < 1e-2:
            t = ri / np.linalg.norm(ri)              # Lanczos step instead (:92-94)

        t = modified_gram_schmidt(t[:, None], V)
        if t.shape[1] == 0:                          # :99-109
            for rj in R.T:
                t = modified_gram_schmidt(rj[:, None], V)
                if t.shape[1] == 1:
                    break
            else:
                t = modified_gram_schmidt(np.random.normal(size=(n, 1)), V)
                if t.shape[1] == 0:
                    return lams, V, AV

        V = np.hstack([V, t])
        AV = np.hstack([AV, A.dot(t)])
