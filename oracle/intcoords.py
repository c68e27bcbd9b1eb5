"""ORACLE (test infrastructure, never shipped on the product path).

A set of internal coordinates with values, Wilson B-matrix and EXACT second derivatives,
the part of the reference's ``BaseInternals`` that ``InternalPES`` consumes:

  calc            sella/internal.py:1735-1778
  jacobian        sella/internal.py:1780-1902
  hessian().ldot  sella/internal.py:2189-2305 + sella/linalg.py:601-620   (sum_c v_c d2q_c/dx2)
  hessian_rdot    sella/internal.py:2307-2575                             (row c = (d2q_c/dx2) w)
  wrap            sella/internal.py:2577-2587                             (dihedral differences)
  check_for_bad_internals  sella/internal.py:3704-3736                    (angles within atol of 0 / pi)
  guess_hessian   sella/internal.py:3738-3830                             (Lindh-type diagonal model)

The primal formulas are the reference's (internal.py:58-80, the same as oracle/internals.py);
derivatives come from second-order forward-mode AD ("hyper-dual" numbers) vectorised over the
coordinates of one kind in numpy, where the reference uses JAX.  Coordinate order is the
reference's ``_names`` order: translations, bonds, angles, dihedrals.

PARITY UNPINNED against the reference itself (JAX is not installed here); pinned against
oracle/internals.py (dual-number gradients + central differences) in tests/test_internal_pes.py and,
for the assembly (ldot/rdot), against the reference's SparseInternalHessians (tests/golden).
"""
import numpy as np

# covalent radii (Angstrom) of Cordero et al., Dalton Trans. 2008, 2832 -- the table ASE ships as
# ase.data.covalent_radii and the reference reads (internal.py:3371, 3744); index = atomic number
COVALENT_RADII = np.array([
    0.20, 0.31, 0.28, 1.28, 0.96, 0.84, 0.76, 0.71, 0.66, 0.57, 0.58, 1.66, 1.41, 1.21, 1.11, 1.07, 1.05, 1.02,
    1.06, 2.03, 1.76, 1.70, 1.60, 1.53, 1.39, 1.39, 1.32, 1.26, 1.24, 1.32, 1.22, 1.22, 1.20, 1.19, 1.20, 1.20,
    1.16, 2.20, 1.95, 1.90, 1.75, 1.64, 1.54, 1.47, 1.46, 1.42, 1.39, 1.45, 1.44, 1.42, 1.39, 1.39, 1.38, 1.39,
    1.40, 2.44, 2.15, 2.07, 2.04, 2.03, 2.01, 1.99, 1.98, 1.98, 1.96, 1.94, 1.92, 1.92, 1.89, 1.90, 1.87, 1.87,
    1.75, 1.70, 1.62, 1.51, 1.44, 1.41, 1.36, 1.36, 1.32, 1.45, 1.46, 1.48, 1.40, 1.50, 1.50])
BOHR = 0.5291772105638411          # ase.units.Bohr  (CODATA 2014, ASE's default)
HARTREE = 27.211386024367243       # ase.units.Hartree


class HD:
    """Second-order forward-mode AD over k variables, vectorised over m coordinates:
    v [m], g [m, k], h [m, k, k] (symmetric)."""

    def __init__(self, v, g, h):
        self.v, self.g, self.h = v, g, h

    @staticmethod
    def var(values, index, k):
        m = len(values)
        g = np.zeros((m, k)); g[:, index] = 1.0
        return HD(np.asarray(values, float), g, np.zeros((m, k, k)))

    def _lift(self, o):
        if isinstance(o, HD):
            return o
        return HD(np.broadcast_to(np.asarray(o, float), self.v.shape).copy(), np.zeros_like(self.g), np.zeros_like(self.h))

    def __add__(self, o):
        o = self._lift(o); return HD(self.v + o.v, self.g + o.g, self.h + o.h)
    __radd__ = __add__

    def __sub__(self, o):
        o = self._lift(o); return HD(self.v - o.v, self.g - o.g, self.h - o.h)

    def __neg__(self):
        return HD(-self.v, -self.g, -self.h)

    def __mul__(self, o):
        o = self._lift(o)
        cross = self.g[:, :, None] * o.g[:, None, :]
        return HD(self.v * o.v, self.g * o.v[:, None] + o.g * self.v[:, None],
                  self.h * o.v[:, None, None] + o.h * self.v[:, None, None] + cross + cross.transpose(0, 2, 1))
    __rmul__ = __mul__

    def chain(self, f, f1, f2):
        """f(self) given f, f', f'' evaluated at self.v."""
        return HD(f, f1[:, None] * self.g,
                  f1[:, None, None] * self.h + f2[:, None, None] * self.g[:, :, None] * self.g[:, None, :])

    def recip(self):
        r = 1.0 / self.v
        return self.chain(r, -r * r, 2.0 * r ** 3)

    def __truediv__(self, o):
        return self * self._lift(o).recip()

    def sqrt(self):
        s = np.sqrt(self.v)
        return self.chain(s, 0.5 / s, -0.25 / s ** 3)

    def acos(self):
        u = 1.0 - self.v ** 2
        return self.chain(np.arccos(np.clip(self.v, -1.0, 1.0)), -1.0 / np.sqrt(u), -self.v / u ** 1.5)


def hd_atan2(y, x):
    r2 = x.v ** 2 + y.v ** 2
    gy, gx = x.v / r2, -y.v / r2                      # d atan2 / dy, / dx
    hyy, hxx, hxy = -2 * x.v * y.v / r2 ** 2, 2 * x.v * y.v / r2 ** 2, (y.v ** 2 - x.v ** 2) / r2 ** 2
    g = gy[:, None] * y.g + gx[:, None] * x.g
    yx = y.g[:, :, None] * x.g[:, None, :]
    h = (gy[:, None, None] * y.h + gx[:, None, None] * x.h
         + hyy[:, None, None] * y.g[:, :, None] * y.g[:, None, :] + hxx[:, None, None] * x.g[:, :, None] * x.g[:, None, :]
         + hxy[:, None, None] * (yx + yx.transpose(0, 2, 1)))
    return HD(np.arctan2(y.v, x.v), g, h)


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _kind(pos, idx, tvec):
    """pos (N,3); idx (m, na) atoms of m coordinates of one kind; tvec (m, na-1, 3).
    Returns value [m], gradient [m, 3 na], Hessian [m, 3 na, 3 na] in the coordinates' own atoms."""
    m, na = idx.shape
    k = 3 * na
    p = [[HD.var(pos[idx[:, a], d], 3 * a + d, k) for d in range(3)] for a in range(na)]
    d_ = [[p[a + 1][d] - p[a][d] + tvec[:, a, d] for d in range(3)] for a in range(na - 1)]
    if na == 2:                                       # internal.py:58-60
        q = _dot(d_[0], d_[0]).sqrt()
    elif na == 3:                                     # internal.py:63-68
        d1 = [-c for c in d_[0]]
        d2 = d_[1]
        q = (_dot(d1, d2) / (_dot(d1, d1).sqrt() * _dot(d2, d2).sqrt())).acos()
    else:                                             # internal.py:71-80
        d1, d2, d3 = d_
        c12, c23 = _cross(d1, d2), _cross(d2, d3)
        q = hd_atan2(_dot(d2, _cross(c12, c23)), _dot(d2, d2).sqrt() * _dot(c12, c23))
    return q.v, q.g, q.h


class CoordinateSet:
    """translations [(atom, dim)], bonds [(i, j)], angles [(i, j, k)], dihedrals [(i, j, k, l)];
    tvecs: dict kind -> (m, na-1, 3) periodic shift vectors (ncvecs @ cell) or None."""

    def __init__(self, natoms, translations=(), bonds=(), angles=(), dihedrals=(), tvecs=None, numbers=None,
                 atol=15.0):
        self.natoms, self.ndof = int(natoms), 3 * int(natoms)
        self.trans = np.asarray(list(translations), dtype=int).reshape(-1, 2)
        self.lists = dict(bonds=np.asarray(list(bonds), dtype=int).reshape(-1, 2),
                          angles=np.asarray(list(angles), dtype=int).reshape(-1, 3),
                          dihedrals=np.asarray(list(dihedrals), dtype=int).reshape(-1, 4))
        tvecs = tvecs or {}
        self.tvecs = {}
        for kind, arr in self.lists.items():
            t = tvecs.get(kind)
            self.tvecs[kind] = np.zeros((len(arr), arr.shape[1] - 1, 3)) if t is None else \
                np.asarray(t, float).reshape(len(arr), arr.shape[1] - 1, 3)
        self.ntrans, self.nbonds = len(self.trans), len(self.lists["bonds"])
        self.nangles, self.ndihedrals = len(self.lists["angles"]), len(self.lists["dihedrals"])
        self.nother = self.nrotations = 0
        self.nint = self.ntrans + self.nbonds + self.nangles + self.ndihedrals
        self.numbers = None if numbers is None else np.asarray(numbers, dtype=int)
        self.atol = atol * np.pi / 180.0

    # -- values and derivatives ------------------------------------------------------------
    def _blocks(self, pos):
        pos = np.asarray(pos, float).reshape(-1, 3)
        out = []
        for kind in ("bonds", "angles", "dihedrals"):
            idx = self.lists[kind]
            if len(idx):
                out.append((idx,) + _kind(pos, idx, self.tvecs[kind]))
        return pos, out

    def calc(self, pos):
        pos, blocks = self._blocks(pos)
        q = [pos[self.trans[:, 0], self.trans[:, 1]]] + [b[1] for b in blocks]
        return np.concatenate(q) if q else np.zeros(0)

    def jacobian(self, pos):
        pos, blocks = self._blocks(pos)
        B = np.zeros((self.nint, self.ndof))
        B[np.arange(self.ntrans), 3 * self.trans[:, 0] + self.trans[:, 1]] = 1.0
        row = self.ntrans
        for idx, _, g, _ in blocks:
            m, na = idx.shape
            cols = (3 * idx[:, :, None] + np.arange(3)[None, None, :]).reshape(m, 3 * na)
            np.add.at(B, (np.arange(row, row + m)[:, None], cols), g)
            row += m
        return B

    def ldot(self, pos, v):
        """sum_c v_c d2q_c/dx2, dense (ndof, ndof)."""
        pos, blocks = self._blocks(pos)
        D = np.zeros((self.ndof, self.ndof))
        row = self.ntrans
        for idx, _, _, h in blocks:
            m, na = idx.shape
            cols = (3 * idx[:, :, None] + np.arange(3)[None, None, :]).reshape(m, 3 * na)
            np.add.at(D, (cols[:, :, None], cols[:, None, :]), v[row:row + m, None, None] * h)
            row += m
        return D

    def rdot(self, pos, w):
        """(nint, ndof): row c = (d2q_c/dx2) w."""
        pos, blocks = self._blocks(pos)
        R = np.zeros((self.nint, self.ndof))
        row = self.ntrans
        for idx, _, _, h in blocks:
            m, na = idx.shape
            cols = (3 * idx[:, :, None] + np.arange(3)[None, None, :]).reshape(m, 3 * na)
            hw = np.einsum("mij,mj->mi", h, w[cols])
            np.add.at(R, (np.arange(row, row + m)[:, None], cols), hw)
            row += m
        return R

    # -- bookkeeping -----------------------------------------------------------------------
    def wrap(self, vec):
        lo = self.ntrans + self.nbonds + self.nangles
        hi = lo + self.ndihedrals
        vec[lo:hi] = (vec[lo:hi] + np.pi) % (2 * np.pi) - np.pi
        return vec

    def bad_angles(self, pos):
        """Indices (within the angle list) of angles within atol of 0 or pi, or None."""
        if not self.nangles:
            return None
        pos = np.asarray(pos, float).reshape(-1, 3)
        val = _kind(pos, self.lists["angles"], self.tvecs["angles"])[0]
        bad = np.nonzero(~((self.atol < val) & (val < np.pi - self.atol)))[0]
        return bad if len(bad) else None

    def guess_hessian(self, pos, h0cart=70.0):
        """Diagonal of the model Hessian (internal.py:3738-3830), eV / Angstrom^2 (or rad^2)."""
        if self.numbers is None:
            raise ValueError("guess_hessian needs atomic numbers")
        pos = np.asarray(pos, float).reshape(-1, 3)
        rc = COVALENT_RADII[self.numbers]
        h0 = [np.full(self.ntrans, h0cart)]
        b = self.lists["bonds"]
        nb = np.zeros(self.natoms, dtype=int)
        np.add.at(nb, b.ravel(), 1)

        def blen(pairs, tv):
            return np.linalg.norm(pos[pairs[:, 1]] - pos[pairs[:, 0]] + tv, axis=1)
        if len(b):
            r = blen(b, self.tvecs["bonds"][:, 0])
            h0.append(0.3601 * np.exp(-1.944 * (r - rc[b].sum(1)) / BOHR) * HARTREE / BOHR ** 2)
        a = self.lists["angles"]
        if len(a):
            rab, rbc = blen(a[:, :2], self.tvecs["angles"][:, 0]), blen(a[:, 1:], self.tvecs["angles"][:, 1])
            cab, cbc = rc[a[:, :2]].sum(1), rc[a[:, 1:]].sum(1)
            h0.append((0.089 + 0.11 * np.exp(-0.44 * (rab + rbc - cab - cbc) / BOHR)
                       / (cab * cbc / BOHR ** 2) ** -0.42) * HARTREE)
        d = self.lists["dihedrals"]
        if len(d):
            bc = d[:, 1:3]
            rbc = blen(bc, self.tvecs["dihedrals"][:, 1])
            cbc = rc[bc].sum(1)
            L = nb[bc].sum(1) - 2
            h0.append((0.0015 + 14.0 * L ** 0.57 * np.exp(-2.85 * (rbc - cbc) / BOHR)
                       / (rbc * cbc / BOHR ** 2) ** 4.0) * HARTREE)
        return np.abs(np.concatenate(h0))
