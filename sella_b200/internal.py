"""Batched internal coordinates on the device: values, Wilson B-matrix and second-
derivative contractions for an explicit coordinate list (translations, bonds, angles,
dihedrals), the linear-algebra half of sella/internal.py (BaseInternals.calc :1735-1778,
.jacobian :1780-1902, .hessian/.hessian_rdot :2189-2575 with SparseInternalHessians
ldot/rdot, sella/linalg.py:601-646).

The topology search of the reference's ``Internals`` (find_all_bonds/angles/dihedrals,
internal.py:3366-3671) is host-side set-up and out of scope: pass the index arrays.
"""
import numpy as np
import torch

from ._host import dev
from ._lib import I, _p, _stream, call, check_f64


def _iarr(a, width):
    a = np.zeros((0, width), dtype=np.int32) if a is None else np.asarray(a, dtype=np.int32).reshape(-1, width)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def _tvec(t, n, k):
    if t is None:
        return None
    t = np.asarray(t, dtype=np.float64).reshape(n, k, 3)
    return torch.from_numpy(np.ascontiguousarray(t)).to(dev())


class BatchedInternals:
    """translations: [(atom, dim), ...]; bonds [(i, j)]; angles [(i, j, k)]; dihedrals
    [(i, j, k, l)]; tvec_*: PBC shift vectors (ncvec @ cell) per coordinate, or None."""

    def __init__(self, natoms, translations=None, bonds=None, angles=None, dihedrals=None,
                 tvec_bonds=None, tvec_angles=None, tvec_dihedrals=None, rotation_ref=None):
        self.natoms, self.n = int(natoms), 3 * int(natoms)
        self.trans, self.bonds = _iarr(translations, 2), _iarr(bonds, 2)
        self.angles, self.diheds = _iarr(angles, 3), _iarr(dihedrals, 4)
        self.ntrans, self.nbonds = self.trans.shape[0], self.bonds.shape[0]
        self.nangles, self.ndihedrals = self.angles.shape[0], self.diheds.shape[0]
        self.nother = 0
        # the three rotation coordinates of the whole configuration relative to `rotation_ref`
        # ([natoms, 3] shared, or [b, natoms, 3]); they come last (the reference's _names order)
        self.nrotations = 0
        self.rot_ref = None
        if rotation_ref is not None:
            r = np.asarray(rotation_ref, dtype=np.float64)
            r = r.reshape((-1, self.natoms, 3))
            r = r - r.mean(axis=1, keepdims=True)          # Rotation.__init__ centres it (internal.py:1041)
            self.rot_ref = torch.from_numpy(np.ascontiguousarray(r)).to(dev())
            self.rot_refstride = 0 if r.shape[0] == 1 else self.n
            self.nrotations = 3
            self.qprev = None
            self._rot_work = None
        self.nstd = self.ntrans + self.nbonds + self.nangles + self.ndihedrals
        self.nint = self.nstd + self.nrotations
        self.tb = _tvec(tvec_bonds, self.nbonds, 1)
        self.ta = _tvec(tvec_angles, self.nangles, 2)
        self.td = _tvec(tvec_dihedrals, self.ndihedrals, 3)

    def _topo(self):
        return (_p(self.trans), I(self.ntrans), _p(self.bonds), I(self.nbonds), _p(self.angles), I(self.nangles),
                _p(self.diheds), I(self.ndihedrals), _p(self.tb), _p(self.ta), _p(self.td))

    def calc(self, x, jacobian=False, active=None):
        """q [b, nint] (and B [b, nint, n])."""
        check_f64(x)
        b = x.shape[0]
        q = torch.zeros((b, self.nint), dtype=torch.float64, device=x.device)
        B = torch.zeros((b, self.nint, self.n), dtype=torch.float64, device=x.device) if jacobian else None
        if self.nstd:
            # the kernel writes coordinate c at q + b*nstd + c (and rows likewise): run it on its own block
            qs = q if not self.nrotations else torch.zeros((b, self.nstd), dtype=torch.float64, device=x.device)
            Bs = B if (B is None or not self.nrotations) else torch.zeros((b, self.nstd, self.n), dtype=torch.float64,
                                                                           device=x.device)
            call("sb_internals_qB", *self._topo(), _p(x), I(self.n), _p(qs), _p(Bs), _p(active), I(b), _stream())
            if self.nrotations:
                q[:, :self.nstd] = qs
                if B is not None:
                    B[:, :self.nstd] = Bs
        if self.nrotations:
            self._rotation(x, q, B, None, None, active)
        return (q, B) if jacobian else q

    def _rotation(self, x, q, B, v, D, active):
        from ._lib import LL
        b = x.shape[0]
        if self.qprev is None or self.qprev.shape[0] != b:
            self.qprev = torch.zeros((b, 4), dtype=torch.float64, device=x.device)
            self.qprev[:, 0] = 1.0
            self._rot_work = torch.empty((b, 14 * self.n), dtype=torch.float64, device=x.device)
        nint, n, o = self.nint, self.n, self.nstd
        call("sb_rotation", _p(x), I(self.natoms), _p(self.rot_ref), LL(self.rot_refstride), _p(self.qprev),
             _p(None if q is None else q[:, o:]), LL(nint), _p(None if B is None else B[:, o:]), LL(nint * n),
             _p(None if v is None else v[:, o:]), LL(nint), _p(D), _p(self._rot_work), _p(active), I(b), _stream())

    def jacobian(self, x, active=None):
        return self.calc(x, jacobian=True, active=active)[1]

    def ldot(self, x, v, active=None):
        """D [b, n, n] = sum_c v[b,c] d2q_c/dx2."""
        check_f64(x, v)
        b = x.shape[0]
        D = torch.zeros((b, self.n, self.n), dtype=torch.float64, device=x.device)
        if self.nstd:
            vs = v if not self.nrotations else v[:, :self.nstd].contiguous()
            call("sb_internals_hess", *self._topo(), _p(x), I(self.n), _p(vs), _p(D), _p(None), _p(None), _p(active),
                 I(b), _stream())
        if self.nrotations:
            self._rotation(x, None, None, v, D, active)
        return D

    def rdot(self, x, w, active=None):
        """R [b, nint, n]: row c = (d2q_c/dx2) w[b]."""
        if self.nrotations:
            raise NotImplementedError("rdot with rotation coordinates is not on the CUDA path yet")
        check_f64(x, w)
        b = x.shape[0]
        R = torch.zeros((b, self.nint, self.n), dtype=torch.float64, device=x.device)
        call("sb_internals_hess", *self._topo(), _p(x), I(self.n), _p(None), _p(None), _p(w), _p(R), _p(active),
             I(b), _stream())
        return R


class WilsonAlgebra:
    """The linear algebra InternalPES builds on the Wilson matrix Bm [b, nint, ncart]
    (sella/peswrapper.py:674-736, 1011-1082, 1124-1127, 1176-1183), batched on the device:

        Q, R            economy QR, Unred = Q                       _get_jacobian_qr  :674-709
        Binv            R^-1 Q^T  [b, ncart, nint]                  _get_Binv         :711-736
        gradient(g)     g_cart @ Binv                               eval              :1124-1127
        drdx(J)         J R^-1 (reduced) and J R^-1 Q^T (internal)  _compute_basis_int:1050-1082
        Hc(Dc, Dq)      Binv^T (D_cons - D_int) Binv                _compute_Hc_int   :1011-1031
        reduce(H)       Q^T H Q  and  df_pred                       get_df_pred       :1176-1183
        H0(h0)          P diag(h0) P,  P = Q Q^T                    _range_space_projector :72-82

    A rank-deficient Jacobian (|R_ii| < 1e-6 max|R_ii|; the reference then switches to an SVD,
    :691-704) is reported through ``rank_deficient`` and not handled on the device."""

    def __init__(self, Bm):
        from . import kernels as K
        check_f64(Bm)
        self.b, self.nint, self.ncart = Bm.shape
        if self.nint < self.ncart:
            raise ValueError("the Wilson matrix needs nint >= ncart")
        self.Q, self.R = K.qr(Bm)
        rd = torch.diagonal(self.R, dim1=1, dim2=2).abs()
        self.rank_deficient = rd.min(dim=1).values < 1e-6 * rd.max(dim=1).values
        self.Rinv, self.status = K.trtri(self.R)
        self.Binv = K.gemm(self.Rinv, self.Q, transB=True)

    def gradient(self, g_cart):
        from . import kernels as K
        return K.gemm(g_cart.view(self.b, 1, self.ncart).contiguous(), self.Binv).view(self.b, self.nint)

    def drdx(self, J):
        """J: constraint Jacobian wrt Cartesians, [b, nc, ncart] or [nc, ncart] shared."""
        from . import kernels as K
        red = K.gemm(J, self.Rinv)                              # [b, nc, ncart]: in the basis Q
        return red, K.gemm(red, self.Q, transB=True)            # [b, nc, nint]

    def Hc(self, D_cons, D_int):
        from . import kernels as K
        T = K.gemm((D_cons - D_int).contiguous(), self.Binv)    # [b, ncart, nint]
        return K.gemm(self.Binv, T, transA=True)                # [b, nint, nint]

    def reduce(self, H):
        from . import kernels as K
        return K.gemm(self.Q, K.gemm(H, self.Q), transA=True)   # [b, ncart, ncart]

    def df_pred(self, dx, g, H):
        from . import kernels as K
        dxr = K.gemm(dx.view(self.b, 1, self.nint).contiguous(), self.Q)          # [b, 1, ncart]
        gr = K.gemm(g.view(self.b, 1, self.nint).contiguous(), self.Q)
        Hr = self.reduce(H)
        Hd = K.gemm(dxr, Hr)
        return (gr * dxr).sum(dim=(1, 2)) + 0.5 * (Hd * dxr).sum(dim=(1, 2))

    def H0(self, h0):
        """h0 [nint] or [b, nint]: diagonal guess projected onto range(Bm)."""
        from . import kernels as K
        h = h0 if h0.dim() == 2 else h0.unsqueeze(0).expand(self.b, self.nint)
        DQ = (h.unsqueeze(2) * self.Q).contiguous()                               # diag(h0) Q
        core = K.gemm(self.Q, DQ, transA=True)                                    # Q^T diag(h0) Q
        return K.gemm(self.Q, K.gemm(core, self.Q, transB=True))                  # Q core Q^T
