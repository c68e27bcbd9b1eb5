"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the reference's state/glue objects for the Cartesian
saddle-search loop:

  FiniteDifferenceHessian  <- NumericalHessian           sella/linalg.py:14-101
  OperatorSum              <- MatrixSum                  sella/linalg.py:104-140
  ApproxHessian            <- ApproximateHessian         sella/linalg.py:143-353 (CPU branch)
  split_constraints        <- _split_cons_subspace       sella/peswrapper.py:51-69
  CartesianPES             <- PES                        sella/peswrapper.py:214-606
                              (get_HL_projected :363-386, _calc_basis :395-407,
                               get_scons :429-438, _update_basis :467-481,
                               diag :508-556, converged :558-568, kick :578-602)

``CartesianPES`` is driven by a plain callable ``x -> (f, g)`` and a *linear*
constraint set ``C x = c`` (what ``Constraints.fix_translation`` produces); the
constraint Hessian ``Hc`` of a linear constraint is identically zero, which is
also what the reference evaluates in that case.

Pinned: the first four through ``tests/golden`` directly; ``CartesianPES`` and
``oracle/driver.py`` by running the reference's own ``PES`` / ``Sella`` classes
with duck-typed atoms (``tests/golden/make_golden.py``, ``ref_harness.py``).
"""
import numpy as np
from scipy.linalg import eigh, qr

from .davidson import rayleigh_ritz
from .hessian import symmetrize_Y, update_H


class FiniteDifferenceHessian:
    """Matrix-free H v ~ |v| (g(x0 + eta v^) - g(x0)) / eta with a canonical sign
    for the displacement and a record of every (v, Hv) pair (linalg.py:39-95)."""
    dtype = np.dtype("float64")

    def __init__(self, func, x0, g0, eta, threepoint=False, Uproj=None):
        self.func, self.eta, self.threepoint, self.Uproj = func, eta, threepoint, Uproj
        self.x0, self.g0 = x0.copy(), g0.copy()
        self.ntrue = len(x0)
        n = self.ntrue if Uproj is None else Uproj.shape[1]
        self.shape = (n, n)
        self.calls = 0
        self.Vs = np.empty((self.ntrue, 0))
        self.AVs = np.empty((self.ntrue, 0))

    @staticmethod
    def canonical_sign(v, g0, x0):
        """linalg.py:59-73."""
        vg, vx = v @ g0, v @ x0
        if abs(vg) > 1e-4:
            return 1.0 if vg < 0 else -1.0
        if abs(vx) > 1e-4:
            return 1.0 if vx < 0 else -1.0
        for vi in v:
            if vi > 1e-4:
                return 1.0
            if vi < -1e-4:
                return -1.0
        return 1.0

    def matvec(self, v):
        self.calls += 1
        v = np.asarray(v, dtype=float).ravel()
        if self.Uproj is not None:
            v = self.Uproj @ v
        sign = self.canonical_sign(v, self.g0, self.x0)
        vnorm = np.linalg.norm(v)
        if vnorm < 1e-12:
            return np.zeros(self.shape[0])
        vnorm *= sign
        _, gp = self.func(self.x0 + self.eta * v / vnorm)
        if self.threepoint:
            _, gm = self.func(self.x0 - self.eta * v / vnorm)
            Av = vnorm * (gp - gm) / (2 * self.eta)
        else:
            Av = vnorm * (gp - self.g0) / self.eta
        self.Vs = np.hstack((self.Vs, v[:, None]))
        self.AVs = np.hstack((self.AVs, Av[:, None]))
        return Av if self.Uproj is None else self.Uproj.T @ Av

    def dot(self, X):
        X = np.asarray(X)
        if X.ndim == 1:
            return self.matvec(X)
        return np.column_stack([self.matvec(X[:, j]) for j in range(X.shape[1])])

    def __add__(self, other):
        return OperatorSum(self, other)

    def __sub__(self, other):
        return OperatorSum(self, -other)


class OperatorSum:
    """linalg.py:104-140: dense terms are pre-summed, operators applied in turn."""

    def __init__(self, *terms):
        self.shape = terms[0].shape
        dense = None
        self.terms = []
        for t in terms:
            if isinstance(t, np.ndarray):
                dense = t.copy() if dense is None else dense + t
            else:
                self.terms.append(t)
        if dense is not None:
            self.terms.append(dense)

    def dot(self, X):
        out = np.zeros_like(np.asarray(X, dtype=float))
        for t in self.terms:
            out = out + t.dot(X)
        return out


class ApproxHessian:
    """Dense approximate Hessian B (None == identity, 'uninitialised') with a
    lazily cached eigendecomposition (linalg.py:143-353)."""

    def __init__(self, dim, ncart, B0=None, update_method="TS-BFGS", symm=2,
                 initialized=False):
        self.dim, self.ncart = dim, ncart
        self.shape = (dim, dim)
        self.update_method, self.symm = update_method, symm
        self.initialized = initialized
        self.set_B(B0)

    def set_B(self, target):
        self._evals = self._evecs = None
        if target is None:
            self.B = None
            self.initialized = False
            return
        if np.isscalar(target):
            target = target * np.eye(self.dim)
        else:
            self.initialized = True
        self.B = target

    def _spectrum(self):
        if self._evals is None and self.B is not None:
            self._evals, self._evecs = eigh(self.B)

    @property
    def evals(self):
        self._spectrum()
        return self._evals

    @evals.setter
    def evals(self, v):
        self._evals = v

    @property
    def evecs(self):
        self._spectrum()
        return self._evecs

    @evecs.setter
    def evecs(self, v):
        self._evecs = v

    def update(self, dx, dg):
        """linalg.py:274-304: the very first update only sets the Cartesian block
        (scaled identity + one update); later ones pass B's eigenpairs along."""
        B = np.zeros(self.shape) if self.B is None else self.B.copy()
        if not self.initialized:
            self.initialized = True
            nc = self.ncart
            B[:nc, :nc] = update_H(None, dx[:nc], dg[:nc], method=self.update_method,
                                   symm=self.symm)
            self.set_B(B)
            return
        self.set_B(update_H(B, dx, dg, method=self.update_method, symm=self.symm,
                            lams=self.evals, vecs=self.evecs))

    def project(self, U):
        Bp = None if self.B is None else U.T @ self.B @ U
        return ApproxHessian(U.shape[1], 0, Bp, self.update_method, self.symm)

    def asarray(self):
        return self.B if self.B is not None else np.eye(self.dim)

    def __matmul__(self, v):
        return v if self.B is None else self.B @ v

    dot = __matmul__


def split_constraints(drdx, tol_factor=1e-6):
    """Orthonormal bases of row-space(drdx) and its complement by a pivoted full
    QR of drdx^T (peswrapper.py:51-69)."""
    Q, R, _ = qr(drdx.T, mode="full", pivoting=True, check_finite=False)
    d = np.abs(np.diag(R))
    ncons = int(np.sum(d > tol_factor * d[0])) if (d.size and d[0] > 0) else 0
    return Q[:, :ncons], Q[:, ncons:]


class CartesianPES:
    """Reference ``PES`` semantics over ``func: x -> (f, g)`` and linear
    constraints ``C x = c`` (C: ncons x n; may be empty)."""
    int = None
    n_cell_dof = 0

    def __init__(self, func, x0, C=None, c=None, eta=1e-4, v0=None,
                 eigensolver="jd0", H0=None, hessian_function=None):
        self.func = func
        self.x = np.array(x0, dtype=float)
        self.dim = self.ncart = len(self.x)
        self.C = np.zeros((0, self.dim)) if C is None else np.asarray(C, float)
        self.c = np.zeros(self.C.shape[0]) if c is None else np.asarray(c, float)
        self.eta, self.v0, self.eigensolver = eta, v0, eigensolver
        self.hessian_function = hessian_function          # x -> (n, n); peswrapper.py:228,290
        self.H = ApproxHessian(self.dim, self.ncart, H0, initialized=H0 is not None)
        self.neval = 0
        self.first_diag = True
        self.curr = dict(x=None, f=None, g=None)
        self.last = self.curr.copy()
        self._basis = None

    # -- geometry ---------------------------------------------------------
    def get_x(self):
        return self.x.copy()

    def set_x(self, target):
        diff = target - self.x
        self.x = np.array(target, dtype=float)
        return diff, diff, self.curr.get("g", np.zeros_like(diff))

    def eval(self):
        self.neval += 1
        return self.func(self.x)

    def _calc_eg(self, x):
        self.neval += 1
        return self.func(x)

    # -- constraints --------------------------------------------------------
    def get_drdx(self):
        return self.C

    def get_res(self):
        return self.C @ self.x - self.c

    def _calc_basis(self):
        if self._basis is None:        # linear constraints: geometry independent
            Ucons, Ufree = split_constraints(self.C)
            self._basis = (self.C, Ucons, np.eye(self.dim), Ufree)
        return self._basis

    def _update_basis(self, basis=None):
        drdx, Ucons, Unred, Ufree = basis if basis is not None else self._calc_basis()
        self.curr.update(drdx=drdx, Ucons=Ucons, Unred=Unred, Ufree=Ufree)
        g = self.curr["g"]
        self.curr["L"] = None if g is None else np.linalg.lstsq(drdx.T, g, rcond=None)[0]

    def _update(self, feval=True):
        """peswrapper.py:440-465: evaluate at most once per distinct geometry."""
        key = self.x.tobytes()
        new_point = True
        if self.curr["x"] is not None and key == self.curr.get("key"):
            if feval and self.curr["f"] is None:
                new_point = False
            else:
                return False
        f, g = self.eval() if feval else (None, None)
        if new_point:
            self.last = self.curr.copy()
        self.curr.update(x=self.x.copy(), key=key, f=f, g=g)
        self._update_basis()
        return True

    def get_f(self):
        self._update()
        return self.curr["f"]

    def get_g(self):
        self._update()
        return self.curr["g"].copy()

    def get_Unred(self):
        self._update(False)
        return self.curr["Unred"]

    def get_Ufree(self):
        self._update(False)
        return self.curr["Ufree"]

    def get_Ucons(self):
        self._update(False)
        return self.curr["Ucons"]

    def get_scons(self):
        Ucons = self.get_Ucons()
        return -Ucons @ np.linalg.lstsq(self.get_drdx() @ Ucons, self.get_res(),
                                        rcond=None)[0]

    # -- Hessians -----------------------------------------------------------
    def get_H(self):
        return self.H

    def get_Hc(self):
        return np.zeros((self.dim, self.dim))          # linear constraints

    def get_HL_projected(self, U):
        B = self.H.B
        Bp = None
        if B is not None:
            Bp = U.T @ B @ U
            L = self.curr.get("L")
            if L is not None and L.size > 0:
                Bp = Bp - U.T @ self.get_Hc() @ U
        return ApproxHessian(U.shape[1], 0, Bp, self.H.update_method, self.H.symm)

    # -- partial diagonalisation ----------------------------------------------
    def diag(self, gamma=0.1, threepoint=False, maxiter=None):
        if self.curr["f"] is None:
            self._update(True)
        Ufree = self.get_Ufree()
        nfree = Ufree.shape[1]
        if nfree == 0:
            return
        P = self.get_HL_projected(Ufree)
        no_model = P.B is None
        if no_model or self.first_diag:
            v0 = self.v0 if self.v0 is not None else self.get_g() @ Ufree
            if v0 is not None and np.linalg.norm(v0) < 1e-12:
                v0 = None
        else:
            v0 = None
        P = np.eye(nfree) if no_model else P.asarray()

        Hop = FiniteDifferenceHessian(self._calc_eg, self.get_x(), self.get_g(),
                                      self.eta, threepoint, Ufree)
        Hc = self.get_Hc()
        self.last_rr = rayleigh_ritz(Hop - Ufree.T @ Hc @ Ufree, gamma, P, v0=v0,
                                     method=self.eigensolver, maxiter=maxiter)
        Vs, AVs = Hop.Vs, Hop.AVs
        Asub = Vs.T @ symmetrize_Y(Vs, AVs, symm=2) - Vs.T @ Hc @ Vs
        _, X = eigh(Asub)
        self.H.update(Vs @ X, AVs @ X)
        self.first_diag = False

    # -- convergence / stepping ----------------------------------------------
    def converged(self, fmax, cmax=1e-5):
        Ufree = self.get_Ufree()
        fproj = -(Ufree @ (Ufree.T @ self.get_g())).reshape((-1, 3))
        f1 = np.linalg.norm(fproj, axis=1).max()
        c1 = np.linalg.norm(self.get_res())
        return (f1 < fmax) and (c1 < cmax), f1, c1

    def kick(self, dx, diag=False, **diag_kwargs):
        x0, f0, g0 = self.get_x(), self.get_f(), self.get_g()
        B0 = self.H.asarray()
        dx_i, dx_f, g_par = self.set_x(x0 + dx)
        df_pred = g0 @ dx_i + (dx_i @ B0 @ dx_i) / 2.0
        dg = self.get_g() - g_par
        df = self.get_f() - f0
        ratio = None if abs(df_pred) < 1e-14 else df / df_pred
        if self.last["x"] is not None and self.last["g"] is not None:
            self.H.update(dx_f, dg)
        if diag:
            if self.hessian_function is not None:         # peswrapper.py:596-600
                self.calculate_hessian()
            else:
                self.diag(**diag_kwargs)
        return ratio

    def calculate_hessian(self):
        """peswrapper.py:604-606."""
        self.H.set_B(np.array(self.hessian_function(self.get_x()), dtype=float))


class NonlinearPES(CartesianPES):
    """Cartesian PES with position-dependent constraints: internal coordinates (bonds, angles,
    dihedrals, single-atom translations; oracle/internals.py) held at target values, optionally
    next to linear rows.  Follows the reference's generic PES: get_drdx / get_res evaluated at the
    current geometry (peswrapper.py:388-393), basis per geometry (:395-407), multipliers
    L = lstsq(drdx^T, g) (:467-481) and Hc = sum_i L_i d2r_i/dx2 (:343-352)."""

    def __init__(self, func, x0, coords, targets=None, C=None, c=None, **kw):
        CartesianPES.__init__(self, func, x0, C, c, **kw)
        self.coords = {k: [tuple(int(a) for a in t) for t in coords.get(k, ())]
                       for k in ("translations", "bonds", "angles", "dihedrals")}
        self.ndih = len(self.coords["dihedrals"])
        # "rotation_ref": reference geometry of the three whole-system rotation coordinates
        # (Constraints.fix_rotation, internal.py:2825-2859); they come last
        self.rot_ref = coords.get("rotation_ref")
        self.q_prev = None
        q0 = self._q(self.x)[0]
        self.targets = q0.copy() if targets is None else np.asarray(targets, float)

    def _q(self, x, L=None):
        from . import internals as oi
        from . import rotation as orot
        c = self.coords
        q, B, H = oi.evaluate(x.reshape(-1, 3), c["translations"], c["bonds"], c["angles"], c["dihedrals"])
        if self.rot_ref is not None:
            unit = [np.eye(3)[k] for k in range(3)]
            out = orot.rotation(x, self.rot_ref, self.q_prev, np.zeros(3) if L is None else L)
            vals, J, self.q_prev = out[0], out[1], out[2]
            q = np.concatenate([q, vals])
            B = np.vstack([B.reshape(-1, x.size), J])
            # per-coordinate Hessians are only ever used contracted with multipliers (get_Hc)
            H = list(H) + [("rot", k) for k in range(3)]
        return q, B, H

    def get_drdx(self):
        return np.vstack([self.C, self._q(self.x)[1]])

    def get_res(self):
        r = self._q(self.x)[0] - self.targets
        if self.ndih:
            hi = len(r) - (3 if self.rot_ref is not None else 0)
            r[hi - self.ndih:hi] = (r[hi - self.ndih:hi] + np.pi) % (2 * np.pi) - np.pi
        return np.concatenate([self.C @ self.x - self.c, r])

    def _calc_basis(self):
        drdx = self.get_drdx()
        Ucons, Ufree = split_constraints(drdx)
        return drdx, Ucons, np.eye(self.dim), Ufree

    def get_Hc(self):
        from . import rotation as orot
        L = self.curr["L"]
        H = self._q(self.x)[2]
        nlin = self.C.shape[0]
        out = np.zeros((self.dim, self.dim))
        for Li, Hi in zip(L[nlin:], H):
            if not isinstance(Hi, tuple):
                out += Li * Hi
        if self.rot_ref is not None:
            out += orot.rotation(self.x, self.rot_ref, self.q_prev, L[-3:])[3]
        return out
