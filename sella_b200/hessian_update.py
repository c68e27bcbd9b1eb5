"""CUDA mirror of sella/hessian_update.py: ``update_H`` and ``symmetrize_Y`` with the
reference's signatures (numpy in / numpy out, one system).

All of the reference's methods run on the device: TS-BFGS, PSB, Greenstadt, DFP, BFGS,
SR1 and BFGS_auto, with symm in {None, 0, 1, 2}; there is no CPU path."""
import numpy as np
import torch

from . import kernels as K
from ._host import up_mat, up_cols, down_cols, zeros, raise_status
from ._lib import I, _p, _stream, call

_METHODS = {"TS-BFGS": 0, "PSB": 1, "Greenstadt": 2, "DFP": 3, "BFGS": 4, "SR1": 5, "BFGS_auto": 6}


def _prep(S, Y, first, n, symm=2, k=None):
    kcap = S.shape[1]
    k = kcap if k is None else k
    if symm is None or k == 1:
        symm = 2                                   # identity for one pair (hessian_update.py:28-29)
    if symm not in (0, 1, 2):
        raise ValueError("Unknown symmetrization method {}".format(symm))
    Ytil = torch.zeros_like(S)
    lam0, skip, status = zeros(1), zeros(1, dtype=torch.int32), zeros(1, dtype=torch.int32)
    kvec = torch.full((1,), k, dtype=torch.int32, device=S.device)
    call("sb_update_prep", _p(S), _p(Y), _p(Ytil), I(kcap), _p(kvec), I(n), I(n), I(int(first)), I(int(symm)),
         _p(lam0), _p(skip), _p(status), _p(None), I(1), _stream())
    return Ytil, lam0, skip, status, kvec


def symmetrize_Y(S, Y, symm):
    """sella/hessian_update.py:27-37."""
    S = np.asarray(S, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if symm is None or S.shape[1] == 1:
        return Y
    Sd, Yd = up_cols(S), up_cols(Y)
    Ytil, _, _, status, _ = _prep(Sd, Yd, False, S.shape[0], symm)
    raise_status(status, "symmetrize_Y")
    return down_cols(Ytil)


def update_H(B, S, Y, method="TS-BFGS", symm=2, lams=None, vecs=None):
    """sella/hessian_update.py:40-111.  Returns the updated Hessian (new array); a
    single step shorter than 1e-8 returns B itself, as the reference does."""
    S = np.asarray(S, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if S.ndim == 1:
        if np.linalg.norm(S) < 1e-8:
            return B
        S = S[:, None]
    if Y.ndim == 1:
        Y = Y[:, None]
    if method not in _METHODS:
        raise ValueError("Unknown update method {}".format(method))
    n, k = S.shape
    m = _METHODS[method]
    kcap = 2 * k if m in (4, 6) else k             # BFGS writes 2k (U, J) pairs
    if kcap > 32:
        raise NotImplementedError("update_H: too many secant pairs for the device kernels")
    pad = np.zeros((n, kcap - k))
    Sd, Yd = up_cols(np.hstack([S, pad])), up_cols(np.hstack([Y, pad]))
    first = B is None
    Ytil, lam0, skip, status, kvec = _prep(Sd, Yd, first, n, symm, k)
    Bd = zeros(1, n, n) if first else up_mat(B)
    evals, Vt = zeros(1, n), zeros(1, n, n)
    need_spectrum = m in (0, 6)
    if first:
        call("sb_fill_scaled_identity", _p(Bd), _p(evals), _p(Vt), _p(lam0), I(n), I(n), _p(skip), I(1), _stream())
    elif need_spectrum:
        if lams is not None and vecs is not None:
            evals = up_mat(np.asarray(lams)[None, :])[0].contiguous()
            Vt = up_mat(np.asarray(vecs).T)
        else:
            K.eigh(Bd, evals=evals, Vt=Vt, status=status)
    BS, VtS, aC, aBS = (torch.zeros_like(Sd) for _ in range(4))
    K.hv_ld(Bd, Sd, BS, k)
    if need_spectrum:
        K.hv_ld(Vt, Sd, VtS, k)
        call("sb_abs_scale", _p(VtS), _p(evals), _p(aC), I(kcap), I(n), _p(skip), I(1), _stream())
        K.hv_ld(Vt, aC, aBS, k, transposed=True)
    U, J, W, Xw = (torch.zeros_like(Sd) for _ in range(4))
    kout = torch.zeros(1, dtype=torch.int32, device=Sd.device)
    call("sb_update_mid", _p(Sd), _p(Ytil), _p(BS), _p(aBS if need_spectrum else None), _p(U), _p(J), _p(W), _p(Xw),
         I(kcap), _p(kvec), I(n), I(m), _p(skip), _p(status), _p(None), _p(evals.view(1, n) if need_spectrum else None),
         _p(kout), I(1), _stream())
    call("sb_update_apply", _p(Bd), _p(U), _p(J), _p(W), I(kcap), _p(kout), I(n), _p(skip), I(1), _stream())
    raise_status(status, "update_H")
    return Bd[0].cpu().numpy()
