"""CUDA mirror of sella/hessian_update.py: ``update_H`` and ``symmetrize_Y`` with the
reference's signatures (numpy in / numpy out, one system).

Supported on the device path: methods TS-BFGS, PSB, Greenstadt with symm=2 (Sella's
defaults, sella/linalg.py:149-150); other methods / symm values raise
NotImplementedError rather than silently running on the CPU."""
import numpy as np
import torch

from . import kernels as K
from ._host import up_mat, up_cols, down_cols, zeros, raise_status
from ._lib import I, _p, _stream, call

_METHODS = {"TS-BFGS": 0, "PSB": 1, "Greenstadt": 2}


def _prep(S, Y, first, n):
    k = S.shape[1]
    Ytil = torch.zeros_like(S)
    lam0, skip, status = zeros(1), zeros(1, dtype=torch.int32), zeros(1, dtype=torch.int32)
    kvec = torch.full((1,), k, dtype=torch.int32, device=S.device)
    call("sb_update_prep", _p(S), _p(Y), _p(Ytil), I(k), _p(kvec), I(n), I(n), I(int(first)), _p(lam0), _p(skip),
         _p(status), _p(None), I(1), _stream())
    return Ytil, lam0, skip, status, kvec


def symmetrize_Y(S, Y, symm):
    """sella/hessian_update.py:27-37 (symm=2 path on the device)."""
    S = np.asarray(S, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    if symm is None or S.shape[1] == 1:
        return Y
    if symm != 2:
        raise NotImplementedError("symmetrize_Y: only symm=2 is on the CUDA path")
    Sd, Yd = up_cols(S), up_cols(Y)
    Ytil, _, _, status, _ = _prep(Sd, Yd, False, S.shape[0])
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
    if method not in _METHODS or (symm != 2 and S.shape[1] > 1):
        raise NotImplementedError("update_H(method=%r, symm=%r) is not on the CUDA path" % (method, symm))
    n, k = S.shape
    if k > 16:
        raise NotImplementedError("update_H: more than 16 secant pairs")
    Sd, Yd = up_cols(S), up_cols(Y)
    first = B is None
    Ytil, lam0, skip, status, kvec = _prep(Sd, Yd, first, n)
    Bd = zeros(1, n, n) if first else up_mat(B)
    evals, Vt = zeros(1, n), zeros(1, n, n)
    if first:
        call("sb_fill_scaled_identity", _p(Bd), _p(evals), _p(Vt), _p(lam0), I(n), I(n), _p(skip), I(1), _stream())
    elif _METHODS[method] == 0:
        if lams is not None and vecs is not None:
            evals = up_mat(np.asarray(lams)[None, :])[0].contiguous()
            Vt = up_mat(np.asarray(vecs).T)
        else:
            K.eigh(Bd, evals=evals, Vt=Vt, status=status)
    BS, VtS, aC, aBS = (torch.zeros_like(Sd) for _ in range(4))
    K.hv_ld(Bd, Sd, BS, k)
    m = _METHODS[method]
    if m == 0:
        K.hv_ld(Vt, Sd, VtS, k)
        call("sb_abs_scale", _p(VtS), _p(evals), _p(aC), I(k), I(n), _p(skip), I(1), _stream())
        K.hv_ld(Vt, aC, aBS, k, transposed=True)
    U, J, W, Xw = (torch.zeros_like(Sd) for _ in range(4))
    call("sb_update_mid", _p(Sd), _p(Ytil), _p(BS), _p(aBS if m == 0 else None), _p(U), _p(J), _p(W), _p(Xw), I(k),
         _p(kvec), I(n), I(m), _p(skip), _p(status), _p(None), I(1), _stream())
    call("sb_update_apply", _p(Bd), _p(U), _p(J), _p(W), I(k), _p(kvec), I(n), _p(skip), I(1), _stream())
    raise_status(status, "update_H")
    return Bd[0].cpu().numpy()
