"""Batched device operators (torch CUDA tensors in/out) over the C ABI.

Shapes: matrices [b, n, n]; vector blocks [b, nvec, n] ("vector-major": each
vector contiguous); per-system scalars [b].  All float64, contiguous, on the
current CUDA device.  `active` is an optional int32 [b] mask.
"""
import torch

from . import _lib
from ._lib import I, D, _p, _stream, call, check_f64, require_cuda


def _mask(active):
    if active is None:
        return None
    if active.dtype != torch.int32:
        active = active.to(torch.int32)
    return active.contiguous()


def hv(A, X, transposed=False, active=None, out=None):
    """Y[b,v] = A[b] @ X[b,v]  (or A[b].T @ X[b,v])."""
    require_cuda()
    squeeze = X.dim() == 2
    if squeeze:
        X = X.unsqueeze(1)
    X = X.contiguous()
    check_f64(A, X)
    b, n, _ = A.shape
    nvec = X.shape[1]
    Y = torch.empty_like(X) if out is None else out
    if active is not None and out is None:
        Y.zero_()
    active = _mask(active)
    call("sb_hv", _p(A), _p(X), _p(Y), _p(active), I(b), I(n), I(nvec), I(int(transposed)), _stream())
    return Y.squeeze(1) if squeeze else Y


def quadratic_pes(A, xstar, x, active=None, f=None, g=None):
    """f, g of the fixed quadratic surfaces 1/2 (x-x*)^T A (x-x*)."""
    require_cuda()
    check_f64(A, xstar, x)
    b, n, _ = A.shape
    f = torch.empty(b, dtype=torch.float64, device=A.device) if f is None else f
    g = torch.empty_like(x) if g is None else g
    work = torch.empty_like(x)
    active = _mask(active)
    call("sb_quadratic_pes", _p(A), _p(xstar), _p(x), _p(f), _p(g), _p(work), _p(active), I(b), I(n),
         _stream())
    return f, g


class EighWorkspace:
    def __init__(self, batch, n, device):
        self.work = torch.empty((batch, n, n), dtype=torch.float64, device=device)
        self.small = torch.empty((3, batch, n), dtype=torch.float64, device=device)
        self.work2 = None           # allocated on first use by eigh(): GEMM-based accumulation of Q^T

    def blocked(self):
        if self.work2 is None:
            b, n, _ = self.work.shape
            self.work2 = torch.empty(b * (n * n + 32 * n + 256 * ((n + 15) // 16)), dtype=torch.float64,
                                     device=self.work.device)
        return self.work2


def eigvalsh(A, active=None, evals=None, ws=None, status=None):
    """Ascending eigenvalues [b,n] only (no eigenvector work)."""
    return eigh(A, active=active, evals=evals, ws=ws, status=status, vectors=False)[0]


def eigh(A, active=None, evals=None, Vt=None, ws=None, status=None, vectors=True):
    """Ascending eigenvalues [b,n] and eigenvectors as ROWS of Vt [b,n,n] (vectors=False: Vt is None)."""
    require_cuda()
    check_f64(A)
    b, n, _ = A.shape
    dev = A.device
    evals = torch.empty((b, n), dtype=torch.float64, device=dev) if evals is None else evals
    if vectors:
        Vt = torch.empty((b, n, n), dtype=torch.float64, device=dev) if Vt is None else Vt
    else:
        Vt = None
    ws = EighWorkspace(b, n, dev) if ws is None else ws
    status = torch.zeros(b, dtype=torch.int32, device=dev) if status is None else status
    active = _mask(active)
    if vectors and n >= 64 and b * n * n >= (1 << 22):
        call("sb_eigh_blocked", _p(A), _p(evals), _p(Vt), _p(ws.work), _p(ws.small), _p(ws.blocked()), _p(status),
             _p(active), I(b), I(n), _stream())
    else:
        call("sb_eigh", _p(A), _p(evals), _p(Vt), _p(ws.work), _p(ws.small), _p(status), _p(active), I(b),
             I(n), _stream())
    return evals, Vt, status


# --------------------------------------------------------------------------
# thin wrappers used by the batched engine (all arguments preallocated tensors)
# --------------------------------------------------------------------------
from ._lib import LL  # noqa: E402


def hv_ld(A, X, Y, nvec, transposed=False, active=None):
    """In-place variant on [b, ldv, n] blocks: first `nvec` slots only."""
    b, n, _ = A.shape
    call("sb_hv_ld", _p(A), _p(X), _p(Y), _p(active), I(b), I(n), I(nvec), I(X.shape[1]),
         I(int(transposed)), _stream())
    return Y


def hv_rect(A, m, X, Y, nvec, transposed=False, active=None):
    """hv_ld on the first m rows of every A[b] (A: [b, rows, n] contiguous): transposed=False gives
    Y[b,v,:m] = A[b,:m] @ X[b,v]; transposed=True gives Y[b,v,:] = A[b,:m].T @ X[b,v,:m]."""
    b, rows, n = A.shape
    call("sb_hv_rect", _p(A), LL(rows * n), I(int(m)), _p(X), _p(Y), _p(active), I(b), I(n), I(nvec), I(X.shape[1]),
         I(int(transposed)), _stream())
    return Y


def mgs(X, Y=None, eps1=1e-15, eps2=1e-6, maxiter=100, active=None):
    """Batched modified_gram_schmidt: X [b,nx,n] (modified in place), Y [b,ny,n]."""
    require_cuda()
    check_f64(X, Y)
    b, nx, n = X.shape
    ny = 0 if Y is None else Y.shape[1]
    Ywork = None if Y is None else torch.empty_like(Y)
    nkept = torch.zeros(b, dtype=torch.int32, device=X.device)
    status = torch.zeros(b, dtype=torch.int32, device=X.device)
    call("sb_mgs", _p(X), I(nx), _p(Y), _p(Ywork), I(ny), I(n), D(eps1), D(eps2), I(maxiter),
         _p(nkept), _p(status), _p(_mask(active)), I(b), _stream())
    return nkept, status


def eigh_update(evals, Vt, U, J, C, active=None):
    """(evals, Vt) of B  ->  (evals, Vt) of B + U J^T + J U^T - U sym(C) U^T, in place.

    U, J: [b, k, n] (k <= 16), C: [b, k, k].  Secular-equation update (secular.cu)."""
    require_cuda()
    check_f64(evals, Vt, U, J)
    b, k, n = U.shape
    dev = U.device
    Cmat = torch.zeros((b, 32, 33), dtype=torch.float64, device=dev)
    Cmat[:, :k, :k] = C
    Cmat = Cmat.reshape(b, 32 * 33).contiguous()
    P = torch.zeros((b, 2 * k, n), dtype=torch.float64, device=dev)
    Z = torch.zeros_like(P)
    sig = torch.zeros((b, 2 * k), dtype=torch.float64, device=dev)
    nterm = torch.zeros(b, dtype=torch.int32, device=dev)
    skip = torch.zeros(b, dtype=torch.int32, device=dev)
    if active is not None:
        skip = (1 - _mask(active)).to(torch.int32)
    status = torch.zeros(b, dtype=torch.int32, device=dev)
    work = torch.empty((b, n, n), dtype=torch.float64, device=dev)
    qwork = torch.empty((b, n, n), dtype=torch.float64, device=dev)
    kvec = torch.full((b,), k, dtype=torch.int32, device=dev)
    call("sb_lowrank_factor", _p(U.contiguous()), _p(J.contiguous()), _p(Cmat), I(k), _p(kvec), I(n), _p(P),
         _p(sig), _p(nterm), _p(skip), I(b), _stream())
    hv_ld(Vt, P, Z, 2 * k)
    call("sb_secular_update", _p(evals), _p(Vt), _p(Z), I(2 * k), _p(sig), _p(nterm), I(n), _p(work),
         _p(qwork), _p(status), _p(skip), I(b), _stream())
    return evals, Vt, status


# --------------------------------------------------------------------------
# dense algebra of the internal-coordinate path (csrc/dense.cu)
# --------------------------------------------------------------------------
def gemm(A, B, transA=False, transB=False, alpha=1.0, beta=0.0, out=None, active=None):
    """C[b] = alpha op(A[b]) op(B[b]) + beta C[b] on fp64 tensor-core tiles.  A, B: [b, r, c]
    row-major (a 2-D operand is shared by the batch)."""
    require_cuda()
    check_f64(A, B)
    sA = 0 if A.dim() == 2 else A.shape[-2] * A.shape[-1]
    sB = 0 if B.dim() == 2 else B.shape[-2] * B.shape[-1]
    batch = A.shape[0] if A.dim() == 3 else (B.shape[0] if B.dim() == 3 else 1)
    M, K = (A.shape[-1], A.shape[-2]) if transA else (A.shape[-2], A.shape[-1])
    K2, N = (B.shape[-1], B.shape[-2]) if transB else (B.shape[-2], B.shape[-1])
    if K != K2:
        raise ValueError("gemm: inner dimensions differ (%d, %d)" % (K, K2))
    if out is None:
        out = torch.empty((batch, M, N), dtype=torch.float64, device=A.device)
        beta = 0.0
    check_f64(out)
    call("sb_gemm", I(int(transA)), I(int(transB)), I(M), I(N), I(K), D(float(alpha)), _p(A), I(A.shape[-1]), LL(sA),
         _p(B), I(B.shape[-1]), LL(sB), D(float(beta)), _p(out), I(N), LL(M * N), _p(_mask(active)), I(batch), _stream())
    return out


def qr(A, active=None, want_q=True):
    """Economy QR of A [b, m, n], m >= n: Q [b, m, n], R [b, n, n] (LAPACK sign convention);
    want_q=False: (None, R), half the work."""
    require_cuda()
    check_f64(A)
    b, m, n = A.shape
    fac = A.clone()
    Q = torch.empty((b, m, n), dtype=torch.float64, device=A.device) if want_q else None
    R = torch.empty((b, n, n), dtype=torch.float64, device=A.device)
    work = torch.empty(b * (m * n + 32 * n + 256 * ((n + 15) // 16)), dtype=torch.float64, device=A.device)
    call("sb_qr", _p(fac), I(m), I(n), _p(Q), _p(R), _p(work), _p(_mask(active)), I(b), _stream())
    return Q, R


def potrf(A, active=None, status=None):
    """Upper Cholesky factor R (A = R^T R) of symmetric positive definite A [b, n, n]; A is not modified."""
    require_cuda()
    check_f64(A)
    b, n, _ = A.shape
    R = A.clone()
    status = torch.zeros(b, dtype=torch.int32, device=A.device) if status is None else status
    call("sb_potrf", _p(R), I(n), _p(status), _p(_mask(active)), I(b), _stream())
    return R, status


def trtri(R, active=None, status=None):
    """Inverse of the upper-triangular R [b, n, n]."""
    require_cuda()
    check_f64(R)
    b, n, _ = R.shape
    X = torch.empty_like(R)
    work = torch.empty_like(R)
    status = torch.zeros(b, dtype=torch.int32, device=R.device) if status is None else status
    call("sb_trtri", _p(R), _p(X), _p(work), I(n), _p(status), _p(_mask(active)), I(b), _stream())
    return X, status
