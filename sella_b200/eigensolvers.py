"""CUDA mirror of sella/eigensolvers.py: ``rayleigh_ritz`` and ``exact`` with the
reference's signatures (numpy in / numpy out, one system).

``A`` may be a dense symmetric ndarray (applied on the device with the batched H.V
kernel) or any operator with ``.shape`` and ``.dot`` (applied on the host, one vector at
a time, exactly as the reference calls ``A.dot``; this is how a finite-difference
``NumericalHessian`` plugs in).  ``P`` is the dense preconditioner.  Expansion methods on
the device: 'jd0', 'jd0_alt' (same correction equation, solved in P's eigenbasis), 'gd',
'lanczos', 'mjd0', 'mjd0_alt' (again one equation, bordered vs block-eliminated form).  The metric B of the reference is
the identity at every call site (peswrapper.py:537-539) and is not supported otherwise.
"""
import numpy as np
import torch

from . import kernels as K
from ._host import up, up_mat, down_cols, zeros, raise_status
from ._lib import I, D, _p, _stream, call

_METHOD = {"jd0": 0, "jd0_alt": 0, "gd": 1, "lanczos": 2, "mjd0": 3, "mjd0_alt": 3}
KCAP = 32


def exact(A, gamma=None, P=None):
    """sella/eigensolvers.py:9-28 (dense A on the device; operators are densified by
    probing, then diagonalised on the device)."""
    if not isinstance(A, np.ndarray):
        n = A.shape[0]
        probes = np.eye(n) if P is None else exact(np.asarray(P))[1]
        dense = np.zeros((n, n))
        for row in probes:
            dense += np.outer(row, A.dot(row))
        A = 0.5 * (dense + dense.T)
    w, Vt, status = K.eigh(up_mat(A))
    raise_status(status, "eigh")
    lams = w[0].cpu().numpy()
    vecs = Vt[0].T.contiguous().cpu().numpy()
    return lams, vecs, lams[None, :] * vecs


def rayleigh_ritz(A, gamma, P, B=None, v0=None, vref=None, vreftol=0.99, method="jd0", maxiter=None):
    """sella/eigensolvers.py:31-112."""
    n = A.shape[0]
    if B is not None and not np.array_equal(B, np.eye(n)):
        raise NotImplementedError("rayleigh_ritz: only the identity metric is on the CUDA path")
    if vref is not None:
        raise NotImplementedError("rayleigh_ritz: vref (optbench hook) is not on the CUDA path")
    if method not in _METHOD:
        raise ValueError("Unknown diagonalization method {}".format(method))
    if gamma <= 0:
        return exact(A, gamma, P)
    if maxiter is None:
        maxiter = 2 * n + 1
    kcap = min(KCAP, max(2, min(n, maxiter)))
    dense = isinstance(A, np.ndarray)
    Ad = up_mat(A) if dense else None
    meth = _METHOD[method]

    def apply(vec_dev):                       # vec_dev: [1, n] device -> [1, n] device
        if dense:
            return K.hv(Ad, vec_dev.view(1, 1, n)).view(1, n)
        return up(np.asarray(A.dot(vec_dev[0].cpu().numpy())).ravel()).view(1, n)

    i32 = dict(dtype=torch.int32)
    V, AV, Yw = zeros(1, kcap, n), zeros(1, kcap, n), zeros(1, kcap, n)
    ksz, ninit, nhist, state, status = (zeros(1, **i32) for _ in range(5))
    lams, rv, rvhat = zeros(1, kcap), zeros(1, 2, n), zeros(1, 2, n)
    theta, that, t, vnew = zeros(1), zeros(1, n), zeros(1, n), zeros(1, n)

    Pd = up_mat(P)
    pl = Pvt = None
    need_P_spectrum = (v0 is None) or meth in (0, 1, 3)
    Vhat = zeros(1, kcap, n) if meth == 3 else None
    p_identity = bool(np.array_equal(P, np.eye(n)))
    if need_P_spectrum and not (p_identity and v0 is not None and meth == 0):
        pl, Pvt, st = K.eigh(Pd)
        raise_status(st, "eigh(P)")
    if v0 is not None:
        call("sb_davidson_init", _p(up(np.asarray(v0).ravel()).view(1, n)), _p(None), _p(None), I(0), _p(V),
             I(kcap), I(n), _p(ksz), _p(ninit), _p(nhist), _p(state), _p(status), _p(None), I(1), _stream())
    else:
        call("sb_davidson_init", _p(None), _p(pl), _p(Pvt), I(1), _p(V), I(kcap), I(n), _p(ksz), _p(ninit),
             _p(nhist), _p(state), _p(status), _p(None), I(1), _stream())
    k = 0
    for j in range(int(ninit[0])):
        AV[0, j] = apply(V[:, j].contiguous())[0]
        k += 1
    ksz.fill_(k)
    maxiter_eff = min(n, maxiter)
    while True:
        call("sb_davidson_rr", _p(V), _p(AV), I(kcap), _p(ksz), I(n), D(float(gamma)), I(maxiter_eff), _p(lams),
             _p(rv), _p(theta), _p(state), _p(status), I(1), _stream())
        if int(state[0]) != 0:
            break
        if meth == 2:
            tin = None
        elif p_identity and v0 is not None and meth == 0:
            tin = None
        elif meth == 3:
            K.hv_ld(Pvt, rv, rvhat, 1)
            K.hv_ld(Pvt, V, Vhat, k)
            call("sb_davidson_mjd_coeff", _p(Vhat), I(kcap), _p(ksz), _p(rvhat), _p(pl), _p(theta), _p(that), I(n),
                 _p(state), _p(status), I(1), _stream())
            K.hv_ld(Pvt, that.view(1, 1, n), t.view(1, 1, n), 1, transposed=True)
            tin = t
        else:
            K.hv_ld(Pvt, rv, rvhat, 2)
            call("sb_davidson_jd_coeff", _p(rvhat), _p(pl), _p(theta), _p(that), I(n), I(meth), _p(state), I(1),
                 _stream())
            K.hv_ld(Pvt, that.view(1, 1, n), t.view(1, 1, n), 1, transposed=True)
            tin = t
        use_identity = int(tin is None and meth == 0)
        call("sb_davidson_expand", _p(tin), _p(rv), _p(theta), _p(V), _p(Yw), I(kcap), _p(ksz), I(n),
             I(use_identity), I(int(meth == 2)), _p(vnew), _p(state), _p(status), I(1), _stream())
        if int(state[0]) != 0:
            break
        AV[0, k] = apply(vnew)[0]
        k += 1
        ksz.fill_(k)
    st = int(status[0])
    if st & 1:
        raise RuntimeError("MGS failed.")
    return lams[0, :k].cpu().numpy(), down_cols(V, k), down_cols(AV, k)
