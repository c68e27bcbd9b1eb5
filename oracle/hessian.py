"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the reference's secant-pair symmetrisation and symmetric
multi-secant quasi-Newton updates, ``sella/hessian_update.py``:

  symmetrize_Y2   hessian_update.py:12-24
  symmetrize_Y    hessian_update.py:27-37
  update_H        hessian_update.py:40-111   (dispatcher, first-update scaling,
                                              final (B+B^T)/2)
  TS-BFGS / PSB / Greenstadt / BFGS / DFP / SR1 corrections   :114-152

LAPACK (scipy ``eigh``/``lstsq``/``solve``) is the third-party arithmetic the
reference itself calls; it is available here, so the same drivers are used.

Pinned against the reference through ``tests/golden`` (make_golden.py).
"""
import numpy as np
from scipy.linalg import eigh, lstsq, solve


def _sequential_symmetrizer(S, Y):
    """Correction dY such that S^T (Y + dY) is symmetric, built one column at a
    time so column i only borrows from S[:, :i]  (hessian_update.py:12-24)."""
    k = S.shape[1]
    dY = np.zeros_like(Y)
    YtS = Y.T @ S
    StS = S.T @ S
    dYtS = np.zeros_like(YtS)
    for i in range(1, k):
        mismatch = YtS[i, :i].T - YtS[:i, i] - dYtS[:i, i]
        coef = np.linalg.lstsq(StS[:i, :i], mismatch, rcond=None)[0]
        dY[:, i] = -S[:, :i] @ coef
        dYtS[i, :] = -StS[:, :i] @ coef
    return dY


def symmetrize_Y(S, Y, symm):
    """hessian_update.py:27-37."""
    if symm is None or S.shape[1] == 1:
        return Y
    if symm == 2:
        return Y + _sequential_symmetrizer(S, Y)
    skew_lower = np.tril(S.T @ Y - Y.T @ S, -1).T
    if symm == 0:
        return Y + S @ lstsq(S.T @ S, skew_lower)[0]
    if symm == 1:
        return Y + Y @ lstsq(S.T @ Y, skew_lower)[0]
    raise ValueError("Unknown symmetrization method {}".format(symm))


def _rank2k(U, J, S):
    """Common shape of the Powell-symmetric family, hessian_update.py:124-125:
    U J^T + J U^T - U (J^T S) U^T."""
    UJt = U @ J.T
    return (UJt + UJt.T) - U @ (J.T @ S) @ U.T


def _delta_ts_bfgs(B, S, Y, lams, vecs):
    """hessian_update.py:118-125.  |B| S is formed from the eigenpairs of B."""
    J = Y - B @ S
    absBS = vecs @ (np.abs(lams)[:, None] * (vecs.T @ S))
    X = S.T @ Y @ Y.T + S.T @ absBS @ absBS.T
    U = lstsq(X @ S, X)[0].T
    return _rank2k(U, J, S)


def _delta_psb(B, S, Y):
    """hessian_update.py:128-132."""
    J = Y - B @ S
    U = solve(S.T @ S, S.T).T
    return _rank2k(U, J, S)


def _delta_greenstadt(B, S, Y):
    """hessian_update.py:147-152."""
    BS = B @ S
    J = Y - BS
    U = solve(S.T @ BS, BS.T).T
    return _rank2k(U, J, S)


def _delta_dfp(B, S, Y):
    """hessian_update.py:135-139."""
    J = Y - B @ S
    U = solve(S.T @ Y, Y.T).T
    return _rank2k(U, J, S)


def _delta_bfgs(B, S, Y):
    """hessian_update.py:114-115."""
    BS = B @ S
    return Y @ solve(Y.T @ S, Y.T) - BS @ solve(S.T @ BS, BS.T)


def _delta_sr1(B, S, Y):
    """hessian_update.py:142-144."""
    R = Y - B @ S
    return R @ solve(R.T @ S, R.T)


def update_H(B, S, Y, method="TS-BFGS", symm=2, lams=None, vecs=None):
    """hessian_update.py:40-111 (CPU branch; the reference's optional torch
    branch :69-75,160-203 computes the same TS-BFGS formula on device)."""
    if S.ndim == 1:
        if np.linalg.norm(S) < 1e-8:
            return B                                   # :49-52 no-op
        S = S[:, None]
    if Y.ndim == 1:
        Y = Y[:, None]

    Yt = symmetrize_Y(S, Y, symm)

    if B is None:
        # :58-67 scaled identity, scale = geometric mean of |Ritz values of S^T Y~|
        theta = np.abs(eigh(S.T @ Yt)[0])
        theta = np.maximum(theta, 1e-12)
        B = np.exp(np.mean(np.log(theta))) * np.eye(S.shape[0])

    if lams is None or vecs is None:
        lams, vecs = eigh(B)

    if method == "BFGS_auto":                          # :80-87
        method = "TS-BFGS"
        if np.all(lams > 0):
            if np.all(eigh(S.T @ Yt, S.T @ S)[0] > 0):
                method = "BFGS"

    if method == "TS-BFGS":
        delta = _delta_ts_bfgs(B, S, Yt, lams, vecs)
    elif method == "PSB":
        delta = _delta_psb(B, S, Yt)
    elif method == "Greenstadt":
        delta = _delta_greenstadt(B, S, Yt)
    elif method == "BFGS":
        delta = _delta_bfgs(B, S, Yt)
    elif method == "DFP":
        delta = _delta_dfp(B, S, Yt)
    elif method == "SR1":
        delta = _delta_sr1(B, S, Yt)
    else:
        raise ValueError("Unknown update method {}".format(method))

    Bplus = delta + B
    return 0.5 * (Bplus + Bplus.T)                     # :104-109
