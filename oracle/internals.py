"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the reference's internal-coordinate primal formulas
(sella/internal.py:58-80 bond / angle / dihedral, :466-470 translation) with first
derivatives by forward-mode dual numbers in numpy and second derivatives by central
differences of those analytic gradients.

PARITY UNPINNED: the reference differentiates the same formulas with JAX, which is not
installed in this image, so the reference itself cannot be run; the oracle is anchored
the way the reference's own test anchors its derivatives (finite differences,
tests/internal/test_get_internal.py:25-55, rtol = atol = 1e-7).
"""
import numpy as np


class Dual:
    """value + gradient wrt k independent variables (forward mode)."""

    def __init__(self, v, g):
        self.v, self.g = v, g

    @staticmethod
    def const(c, k):
        return Dual(float(c), np.zeros(k))

    def _lift(self, o):
        return o if isinstance(o, Dual) else Dual.const(o, len(self.g))

    def __add__(self, o):
        o = self._lift(o); return Dual(self.v + o.v, self.g + o.g)
    __radd__ = __add__

    def __sub__(self, o):
        o = self._lift(o); return Dual(self.v - o.v, self.g - o.g)

    def __rsub__(self, o):
        return self._lift(o) - self

    def __neg__(self):
        return Dual(-self.v, -self.g)

    def __mul__(self, o):
        o = self._lift(o); return Dual(self.v * o.v, self.g * o.v + self.v * o.g)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._lift(o); return Dual(self.v / o.v, (self.g * o.v - self.v * o.g) / o.v ** 2)


def dsqrt(a):
    s = np.sqrt(a.v); return Dual(s, a.g * 0.5 / s)


def dacos(a):
    if a.v >= 1.0:
        return Dual.const(0.0, len(a.g))
    if a.v <= -1.0:
        return Dual.const(np.pi, len(a.g))
    return Dual(np.arccos(a.v), -a.g / np.sqrt(1.0 - a.v ** 2))


def datan2(y, x):
    r2 = x.v ** 2 + y.v ** 2
    return Dual(np.arctan2(y.v, x.v), (x.v * y.g - y.v * x.g) / r2)


def _cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _sub(a, b, t):
    return [a[i] - b[i] + t[i] for i in range(3)]


def bond(p, t):
    d = _sub(p[1], p[0], t[0]); return dsqrt(_dot(d, d))


def angle(p, t):
    e = _sub(p[1], p[0], t[0]); d1 = [-c for c in e]
    d2 = _sub(p[2], p[1], t[1])
    return dacos(_dot(d1, d2) / (dsqrt(_dot(d1, d1)) * dsqrt(_dot(d2, d2))))


def dihedral(p, t):
    d1 = _sub(p[1], p[0], t[0]); d2 = _sub(p[2], p[1], t[1]); d3 = _sub(p[3], p[2], t[2])
    c12, c23 = _cross(d1, d2), _cross(d2, d3)
    numer = _dot(d2, _cross(c12, c23))
    denom = dsqrt(_dot(d2, d2)) * _dot(c12, c23)
    return datan2(numer, denom)


_FUNCS = {2: bond, 3: angle, 4: dihedral}


def coord_value_grad(pos, atoms, tvec):
    """pos (N,3); atoms tuple (m); tvec (m-1,3) -> value, gradient (m,3)."""
    m = len(atoms)
    k = 3 * m
    p = [[Dual(pos[a, d], np.eye(k)[3 * ia + d]) for d in range(3)] for ia, a in enumerate(atoms)]
    t = [[Dual.const(tvec[i, d], k) for d in range(3)] for i in range(m - 1)]
    q = _FUNCS[m](p, t)
    return q.v, q.g.reshape(m, 3)


def evaluate(pos, translations=(), bonds=(), angles=(), dihedrals=(), tvecs=None, hess_h=1e-5):
    """Returns q (nint,), B (nint, 3N), and the list of dense per-coordinate Hessians
    (3N x 3N each; zeros for translations).  tvecs: dict kind -> array or None."""
    pos = np.asarray(pos, dtype=float).reshape(-1, 3)
    N = len(pos)
    tvecs = tvecs or {}
    q, rows, hessians = [], [], []
    for (a, dim) in translations:
        q.append(pos[a, dim]); r = np.zeros(3 * N); r[3 * a + dim] = 1.0
        rows.append(r); hessians.append(np.zeros((3 * N, 3 * N)))
    for kind, lst in (("bonds", bonds), ("angles", angles), ("dihedrals", dihedrals)):
        for i, atoms in enumerate(lst):
            atoms = tuple(int(a) for a in atoms)
            m = len(atoms)
            tv = np.zeros((m - 1, 3)) if tvecs.get(kind) is None else np.asarray(tvecs[kind])[i].reshape(m - 1, 3)
            v, g = coord_value_grad(pos, atoms, tv)
            q.append(v)
            r = np.zeros(3 * N)
            for ia, a in enumerate(atoms):
                r[3 * a:3 * a + 3] = g[ia]
            rows.append(r)
            H = np.zeros((3 * N, 3 * N))
            for ia, a in enumerate(atoms):
                for d in range(3):
                    pp = pos.copy(); pp[a, d] += hess_h
                    pm = pos.copy(); pm[a, d] -= hess_h
                    gp = coord_value_grad(pp, atoms, tv)[1]
                    gm = coord_value_grad(pm, atoms, tv)[1]
                    dg = (gp - gm) / (2 * hess_h)
                    for ja, a2 in enumerate(atoms):
                        H[3 * a + d, 3 * a2:3 * a2 + 3] = dg[ja]
            hessians.append(0.5 * (H + H.T))
    return np.array(q), np.array(rows), hessians
