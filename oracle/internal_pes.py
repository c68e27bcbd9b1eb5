"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the reference's ``InternalPES`` (sella/peswrapper.py:609-1288) over
``oracle.intcoords.CoordinateSet`` -- a search in redundant internal coordinates with a geodesic
back-transformation:

  __init__ (H0 = P diag(h0) P)          peswrapper.py:609-661, _range_space_projector :72-82
  _get_jacobian_qr / _get_Binv          :674-736   (economy QR; SVD when |R_ii| < 1e-6 max|R_ii|)
  _set_x_iterative                      :749-839   (Newton on the internal residual)
  _set_x_ode / _q_ode                   :841-880, 1200-1221   (geodesic ODE, scipy LSODA, atol 1e-6)
  set_x / _add_proj_delta               :883-927
  _project_to_constraints               :928-994   (Newton projection in the constraint subspace)
  get_x (dihedral unwrapping)           :996-1008
  _compute_Hc_int                       :1011-1031
  get_drdx / _compute_basis_int         :1046-1082
  eval (g_int = g_cart Binv)            :1124-1127
  get_df_pred                           :1176-1183
  get_projected_forces                  :1185-1194
  wrap_dx                               :1196-1197

and of the base-class methods it inherits unchanged (``oracle.pes.CartesianPES``).  Not restated:
dummy atoms, rotation coordinates, cell degrees of freedom, ``update_internals`` (re-detection
of the coordinate set, optimize.py:382-410 -- the caller rebuilds the object instead).

``integrator``: "lsoda" is the reference's integrator.  "rk" is NOT the reference's: it is the
Dormand-Prince 5(4) scheme with per-system step control that the CUDA engine runs
(sella_b200/batched_internal.py), restated here so that engine and oracle can be compared step by
step; tests/test_internal_pes.py checks that both integrators lead to the same converged geometry.

PINNED against the reference's own ``InternalPES`` / ``MaxInternalStep`` / ``Sella.step`` code as far as this
image allows: ``tests/golden/internal_loop.npz`` holds eight trajectories produced by the UNMODIFIED reference
files driven through ``oracle/ref_internal_harness.py`` (slabs with held atoms: prfo, qn, frozen B+, Newton
stepper; a cluster with bonds + angles + dihedrals; free clusters through the SVD branch, saddle search and
minimisation), and ``tests/test_oracle_golden.py::test_internal_pes_oracle_matches_reference_internal_pes``
reproduces every step (positions 2e-8 A, trust radius 1e-8, rho 1e-6, final Hessian 1e-6).  What stays
unpinned is only what JAX computes in the reference: the coordinate derivatives themselves (both sides of that
comparison take them from ``oracle/intcoords.py``; they are checked by finite differences and, for the assembly,
against the reference's ``SparseInternalHessians``).
"""
import numpy as np
from scipy.integrate import LSODA
from scipy.linalg import qr, solve_triangular

from .pes import ApproxHessian, CartesianPES, split_constraints

# Dormand-Prince 5(4) tableau (Dormand & Prince 1980)
DP_C = np.array([0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0])
DP_A = [
    [],
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
DP_B5 = np.array([35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0])
DP_B4 = np.array([5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40])
RK_ATOL, RK_RTOL, RK_MAXSTEPS = 1e-7, 1e-4, 64


def range_space_projector(B):
    """peswrapper.py:72-82."""
    Q, R, _ = qr(B, mode="full", pivoting=True, check_finite=False)
    rdiag = np.abs(np.diag(R))
    rcond = max(B.shape) * np.finfo(float).eps
    nkeep = int(np.sum(rdiag > rcond * rdiag[0])) if (rdiag.size and rdiag[0] > 0) else 0
    return Q[:, :nkeep] @ Q[:, :nkeep].T


class InternalPES(CartesianPES):
    n_cell_dof = 0

    def __init__(self, func, pos0, ints, cons=None, targets=None, eta=1e-4, v0=None, eigensolver="jd0",
                 H0=None, iterative_stepper=0, exact_geodesic=True, integrator="lsoda"):
        self.func = func
        self.pos = np.array(pos0, dtype=float).ravel()
        self.int = ints
        self.cons = cons
        self.ncons = 0 if cons is None else cons.nint
        if self.ncons:
            self.targets = cons.calc(self.pos) if targets is None else np.asarray(targets, float)
        self.dim, self.ncart = ints.nint, ints.ndof
        self.eta, self.v0, self.eigensolver = eta, v0, eigensolver
        self.hessian_function = None
        self.neval = 0
        self.first_diag = True
        self.curr = dict(x=None, f=None, g=None)
        self.last = self.curr.copy()
        self.iterative_stepper, self.exact_geodesic, self.integrator = iterative_stepper, exact_geodesic, integrator
        self._cache = {}
        self.ode_nfev = 0
        if H0 is None:
            P = range_space_projector(ints.jacobian(self.pos))
            H0 = P @ np.diag(ints.guess_hessian(self.pos)) @ P
        # set_H(H0, initialized=False) still ends up initialised: set_B marks any array as such
        # (linalg.py:232-247), so the first secant update is an ordinary one
        self.H = ApproxHessian(self.dim, self.ncart, H0, initialized=True)

    # -- per-geometry factorisations (cached like the reference's _LRU2 caches) -------------------
    def _geom(self):
        key = self.pos.tobytes()
        c = self._cache
        if c.get("key") != key:
            c.clear()
            c["key"] = key
        return c

    def _jacobian_qr(self):
        c = self._geom()
        if "QR" not in c:
            B = self.int.jacobian(self.pos)
            Q, R = np.linalg.qr(B, mode="reduced")
            rd = np.abs(np.diag(R))
            if len(rd) and rd.min() < 1e-6 * rd.max():
                U, S, VT = np.linalg.svd(B, full_matrices=False)
                nn = int(np.sum(S > 1e-6))
                Q, R = U[:, :nn], np.diag(S[:nn]) @ VT[:nn]
                c["Binv"] = VT[:nn].T @ np.diag(1.0 / S[:nn]) @ U[:, :nn].T
            c["QR"] = (Q, R)
        return c["QR"]

    def _Binv(self):
        c = self._geom()
        if "Binv" not in c:
            Q, R = self._jacobian_qr()
            if "Binv" not in c:                       # the SVD branch has stored it already
                c["Binv"] = solve_triangular(R, Q.T, check_finite=False)
        return c["Binv"]

    # -- constraints ---------------------------------------------------------------------------
    def _cons_jac(self):
        return np.zeros((0, self.ncart)) if not self.ncons else self.cons.jacobian(self.pos)

    def get_res(self):
        if not self.ncons:
            return np.zeros(0)
        return self.cons.wrap(self.cons.calc(self.pos) - self.targets)

    def get_drdx(self):
        return self._cons_jac() @ self._Binv()

    def _calc_basis(self):
        c = self._geom()
        if "basis" not in c:
            Q, R = self._jacobian_qr()
            nint = Q.shape[0]
            J = self._cons_jac()
            if J.shape[0] == 0:
                c["basis"] = (np.zeros((0, nint)), np.zeros((nint, 0)), Q, Q)
            else:
                if R.shape[0] == R.shape[1]:
                    red = solve_triangular(R.T, J.T, lower=True, check_finite=False).T
                else:
                    red = J @ (self._Binv() @ Q)
                Vcons, Vfree = split_constraints(red)
                c["basis"] = (red @ Q.T, Q @ Vcons, Q, Q @ Vfree)
        return c["basis"]

    def get_Hc(self):
        c = self._geom()
        key = ("Hc", None if self.curr["L"] is None else self.curr["L"].tobytes())
        if key not in c:
            L = self.curr["L"]
            Binv = self._Binv()
            if L is None or L.size == 0:
                c[key] = np.zeros((self.dim, self.dim))
            else:
                Dc = self.cons.ldot(self.pos, L)
                Lint = L @ self._cons_jac() @ Binv
                Dq = self.int.ldot(self.pos, Lint)
                c[key] = Binv.T @ (Dc - Dq) @ Binv
        return c[key]

    # -- geometry ------------------------------------------------------------------------------
    def get_x(self):
        x = self.int.calc(self.pos)
        if self.curr["x"] is not None and self.int.ndihedrals:
            lo = self.int.ntrans + self.int.nbonds + self.int.nangles
            hi = lo + self.int.ndihedrals
            d = x[lo:hi] - self.curr["x"][lo:hi]
            x[lo:hi] = self.curr["x"][lo:hi] + (d + np.pi) % (2 * np.pi) - np.pi
        return x

    def wrap_dx(self, dx):
        return self.int.wrap(dx)

    def _q_ode(self, t, y):
        self.ode_nfev += 1
        x, dxdt, g = y.reshape((3, self.ncart))
        self.pos = x.copy()
        D = self.int.rdot(self.pos, dxdt)
        Binv = self._Binv() if self.exact_geodesic else self._ode_Binv
        out = -Binv @ (D @ np.column_stack((dxdt, g)))
        return np.concatenate([dxdt, out[:, 0], out[:, 1]])

    def _set_x_ode(self, target):
        dx = self.wrap_dx(target - self.get_x())
        Binv = self._Binv()
        self._ode_Binv = Binv
        g0 = self.curr.get("g")
        y0 = np.concatenate([self.pos, Binv @ dx, Binv @ (np.zeros_like(dx) if g0 is None else g0)])
        if self.integrator == "lsoda":
            ode = LSODA(self._q_ode, 0.0, y0, t_bound=1.0, atol=1e-6)
            y, t0 = y0, 0.0
            while ode.status == "running":
                ode.step()
                y, t0 = ode.y, ode.t
                self.pos = y[:self.ncart].copy()
                if self.int.bad_angles(self.pos) is not None:
                    break
                if ode.nfev > 1000:
                    raise RuntimeError("Geometry update ODE is taking too long to converge!")
            if ode.status == "failed":
                raise RuntimeError("Geometry update ODE failed to converge!")
        else:
            y, t0 = self._rk(y0)
        y = y.reshape((3, self.ncart))
        self.pos = y[0].copy()
        B = self.int.jacobian(self.pos)
        return t0 * dx, t0 * B @ y[1], B @ y[2]

    def _rk(self, y0):
        """Dormand-Prince 5(4) from t = 0 to 1; step control on the mixed error norm
        max_i |e_i| / (atol + rtol max(|y_i|, |ynew_i|)), h <- h min(5, max(0.2, 0.9 err^-1/5)).
        Stops early (like the reference's loop) when an angle comes within atol of 0 or pi."""
        t, h, y = 0.0, 1.0, y0.copy()
        k1 = self._q_ode(t, y)
        for _ in range(RK_MAXSTEPS):
            h = min(h, 1.0 - t)
            ks = [k1]
            for s in range(1, 7):
                ys = y + h * sum(a * k for a, k in zip(DP_A[s], ks))
                ks.append(self._q_ode(t + DP_C[s] * h, ys))
            ynew = y + h * sum(b * k for b, k in zip(DP_B5, ks))
            e = h * sum((b5 - b4) * k for b5, b4, k in zip(DP_B5, DP_B4, ks))
            err = np.max(np.abs(e) / (RK_ATOL + RK_RTOL * np.maximum(np.abs(y), np.abs(ynew))))
            fac = 5.0 if err == 0.0 else min(5.0, max(0.2, 0.9 * err ** -0.2))
            if err <= 1.0:
                t, y, k1 = t + h, ynew, ks[6]           # FSAL: k7 is the slope at the new point
                self.pos = y[:self.ncart].copy()
                if t >= 1.0 - 1e-14 or self.int.bad_angles(self.pos) is not None:
                    return y, min(t, 1.0)
            h *= fac
        raise RuntimeError("Geometry update ODE is taking too long to converge!")

    def _set_x_iterative(self, target, max_iter=20):
        pos0 = self.pos.copy()
        x0 = self.get_x()
        dx_initial = target - x0
        g0 = self._Binv() @ self.curr.get("g", np.zeros_like(dx_initial))
        rms_prev, initial_rms, stagn = np.inf, None, 0
        for it in range(max_iter):
            res = self.wrap_dx(target - self.get_x())
            rms = np.linalg.norm(res) / np.sqrt(len(res))
            if initial_rms is None:
                initial_rms = rms
            if rms < 1e-8:
                break
            if rms > initial_rms * 2.0:
                self.pos = pos0
                return None
            if it > 3:
                if rms > rms_prev * 0.95:
                    stagn += 1
                    if stagn >= 3:
                        if rms > initial_rms * 0.5:
                            self.pos = pos0
                            return None
                        break
                else:
                    stagn = 0
            rms_prev = rms
            self.pos = self.pos + np.linalg.lstsq(self.int.jacobian(self.pos), res, rcond=None)[0]
            if self.int.bad_angles(self.pos) is not None:
                self.pos = pos0
                return None
        fres = self.wrap_dx(target - self.get_x())
        if np.linalg.norm(fres) / np.sqrt(len(dx_initial)) > 1e-6:
            self.pos = pos0
            return None
        return dx_initial, self.get_x() - x0, self.int.jacobian(self.pos) @ g0

    def _project_to_constraints(self, target_tol=1e-7, max_iter=8, safety_limit=0.05):
        if not self.ncons:
            return False
        moved = False
        for _ in range(max_iter):
            r = self.get_res()
            if np.linalg.norm(r, ord=np.inf) < target_tol:
                return moved
            drdx, Ucons, _, _ = self._calc_basis()
            if Ucons.shape[1] == 0:
                return moved
            s = np.linalg.lstsq(drdx @ Ucons, -r, rcond=None)[0]
            dx = self._Binv() @ (Ucons @ s)
            if np.linalg.norm(dx, ord=np.inf) > safety_limit:
                return moved
            self.pos = self.pos + dx
            moved = True
        return moved

    def set_x(self, target):
        res = self._set_x_iterative(target) if self.iterative_stepper else None
        if res is None:
            res = self._set_x_ode(target)
        q_after = self.int.calc(self.pos).copy()
        moved = self._project_to_constraints()
        dx_initial, dx_final, g_final = res
        if moved:
            dx_final = dx_final + self.int.wrap(self.int.calc(self.pos) - q_after)
        return dx_initial, dx_final, g_final

    def save(self):
        self._saved = self.pos.copy()

    def restore(self):
        self.pos = self._saved.copy()

    # -- evaluation ------------------------------------------------------------------------------
    def eval(self):
        self.neval += 1
        f, g_cart = self.func(self.pos)
        return f, g_cart @ self._Binv()

    def _calc_eg(self, x):
        self.save()
        self.set_x(x)
        out = self.eval()
        self.restore()
        return out

    def _update(self, feval=True):
        """PES._update (peswrapper.py:440-465) keyed on the Cartesian geometry."""
        key = self.pos.tobytes()
        new_point = True
        if self.curr["x"] is not None and key == self.curr.get("key"):
            if feval and self.curr["f"] is None:
                new_point = False
            else:
                return False
        x = self.get_x()
        basis = self._calc_basis()
        f, g = self.eval() if feval else (None, None)
        if new_point:
            self.last = self.curr.copy()
        self.curr.update(x=x, key=key, f=f, g=g)
        self._update_basis(basis)
        return True

    def get_HL_projected(self, U):
        Bp = U.T @ self.H.B @ U
        L = self.curr.get("L")
        if L is not None and L.size > 0:
            Bp = Bp - U.T @ self.get_Hc() @ U
        return ApproxHessian(U.shape[1], 0, Bp, self.H.update_method, self.H.symm)

    def get_df_pred(self, dx, g, H):
        U = self.get_Unred()
        dxr, gr = dx @ U, g @ U
        return gr @ dxr + (dxr @ (U.T @ H @ U) @ dxr) / 2.0

    def converged(self, fmax, cmax=1e-5):
        g, Ufree = self.get_g(), self.get_Ufree()
        fproj = -((Ufree @ (Ufree.T @ g)) @ self.int.jacobian(self.pos)).reshape((-1, 3))
        f1 = np.linalg.norm(fproj, axis=1).max()
        c1 = np.linalg.norm(self.get_res())
        return (f1 < fmax) and (c1 < cmax), f1, c1

    def kick(self, dx, diag=False, **diag_kwargs):
        x0, f0, g0 = self.get_x(), self.get_f(), self.get_g()
        B0 = self.H.asarray()
        dx_i, dx_f, g_par = self.set_x(x0 + dx)
        df_pred = self.get_df_pred(dx_i, g0, B0)
        dg = self.get_g() - g_par
        df = self.get_f() - f0
        ratio = None if abs(df_pred) < 1e-14 else df / df_pred
        if self.last["x"] is not None and self.last["g"] is not None:
            self.H.update(dx_f, dg)
        if diag:
            self.diag(**diag_kwargs)
        return ratio
