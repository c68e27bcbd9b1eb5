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
                 tvec_bonds=None, tvec_angles=None, tvec_dihedrals=None):
        self.natoms, self.n = int(natoms), 3 * int(natoms)
        self.trans, self.bonds = _iarr(translations, 2), _iarr(bonds, 2)
        self.angles, self.diheds = _iarr(angles, 3), _iarr(dihedrals, 4)
        self.ntrans, self.nbonds = self.trans.shape[0], self.bonds.shape[0]
        self.nangles, self.ndihedrals = self.angles.shape[0], self.diheds.shape[0]
        self.nother = self.nrotations = 0
        self.nint = self.ntrans + self.nbonds + self.nangles + self.ndihedrals
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
        call("sb_internals_qB", *self._topo(), _p(x), I(self.n), _p(q), _p(B), _p(active), I(b), _stream())
        return (q, B) if jacobian else q

    def jacobian(self, x, active=None):
        return self.calc(x, jacobian=True, active=active)[1]

    def ldot(self, x, v, active=None):
        """D [b, n, n] = sum_c v[b,c] d2q_c/dx2."""
        check_f64(x, v)
        b = x.shape[0]
        D = torch.zeros((b, self.n, self.n), dtype=torch.float64, device=x.device)
        call("sb_internals_hess", *self._topo(), _p(x), I(self.n), _p(v), _p(D), _p(None), _p(None), _p(active),
             I(b), _stream())
        return D

    def rdot(self, x, w, active=None):
        """R [b, nint, n]: row c = (d2q_c/dx2) w[b]."""
        check_f64(x, w)
        b = x.shape[0]
        R = torch.zeros((b, self.nint, self.n), dtype=torch.float64, device=x.device)
        call("sb_internals_hess", *self._topo(), _p(x), I(self.n), _p(None), _p(None), _p(w), _p(R), _p(active),
             I(b), _stream())
        return R
