"""Numpy blueprint of the engine's COMPACT representation of the approximate Hessian
(test infrastructure; the CUDA path in sella_b200/csrc/compact.cu follows it operation by operation):

    B = lam0 * I + VR^T diag(theta - lam0) VR,      VR [m, n] orthonormal rows, theta ascending

i.e. m explicit eigenpairs (theta_i, row i of VR) and the eigenvalue lam0 on the whole orthogonal
complement (multiplicity n - m), whose eigenvectors are never stored.  A quasi-Newton Hessian that
started as lam0*I (hessian_update.py:58-67) and received low-rank secant updates has exactly this form,
with m = number of rank-one terms applied so far; every pass over the eigenvectors costs O(n m)
instead of O(n^2), and once m reaches n the representation IS the dense eigendecomposition.

Everything the optimiser needs from eigh(B) (linalg.py:174-195; stepper.py:75-185;
eigensolvers.py:115-139; hessian_update.py:118-125) follows from
    f(B) x = f(lam0) x + VR^T [ (f(theta) - f(lam0)) * (VR x) ].
"""
import numpy as np


class CompactSpectrum:
    def __init__(self, n, lam0):
        self.n, self.lam0 = n, float(lam0)
        self.theta = np.zeros(0)
        self.VR = np.zeros((0, n))

    @property
    def m(self):
        return len(self.theta)

    def dense(self):
        return self.lam0 * np.eye(self.n) + self.VR.T @ ((self.theta - self.lam0)[:, None] * self.VR)

    def apply(self, f, x):
        c = self.VR @ x
        return f(self.lam0) * x + self.VR.T @ ((f(self.theta) - f(self.lam0)) * c)

    # ------------------------------------------------------------------ eigen-update
    def update(self, P, sig, drop_tol=4e-13):
        """B <- B + sum_t sig_t p_t p_t^T with orthonormal rows p_t of P [T, n].
        kernels: hv (Z = VR P^T), hvt (W = VR^T Z), append_a (p_perp, MGS among the candidates), hv + hvt
        again on the unit candidates, append_b (final MGS, rows appended with eigenvalue lam0, their
        coefficients), secular update on the m + T' explicit rows."""
        m, n = self.m, self.n
        T = len(sig)
        Z1 = P @ self.VR.T                            # hv : z_t = VR p_t            [T, m]
        if m < n:
            Pp = P - Z1 @ self.VR                     # hvt: components outside span(VR)
            # append_a: orthonormalise the T candidates among themselves (MGS, two sweeps), drop
            # negligible ones -> unit vectors whose VR-components are only relatively small
            cand = []
            for t in range(T):
                v = Pp[t].copy()
                for _ in range(2):
                    for q in cand:
                        v -= (q @ v) * q
                nrm = np.linalg.norm(v)
                if nrm <= drop_tol or len(cand) >= n - m:
                    continue
                cand.append(v / nrm)
            Q = np.array(cand).reshape(len(cand), n)
            # second projection of the UNIT candidates (hv + hvt): orthogonal to VR to rounding
            Q = Q - (Q @ self.VR.T) @ self.VR
            # append_b: re-orthonormalise among themselves, coefficients against the original p_t
            new = []
            for j in range(len(Q)):
                v = Q[j].copy()
                for _ in range(2):
                    for q in new:
                        v -= (q @ v) * q
                new.append(v / np.linalg.norm(v))
            Q = np.array(new).reshape(len(new), n)
            Z = np.hstack([Z1, P @ Q.T])
            self.VR = np.vstack([self.VR, Q])
            self.theta = np.concatenate([self.theta, np.full(len(new), self.lam0)])
        else:
            Z = Z1
        # the secular-equation update of the explicit rows (here: a small dense eigh of the same matrix)
        K = np.diag(self.theta) + (Z.T * sig[None, :]) @ Z
        w, E = np.linalg.eigh(K)
        self.theta = w
        self.VR = E.T @ self.VR

    # ------------------------------------------------------------------ step model input
    def poles(self, g, width=None):
        """Merged ascending pole list for the step kernels: the explicit eigenvalues with coefficients
        VR g, the cluster pole lam0 with coefficient |g_perp| at its sorted position (only if m < n),
        zero-weight copies of lam0 as padding up to `width`.  Returns (ev, vg, rowmap, gperp) with
        rowmap[i] = explicit row, -1 = cluster, -2 = padding."""
        m, n = self.m, self.n
        c = self.VR @ g
        gperp = g - self.VR.T @ c
        gam = np.linalg.norm(gperp)
        pos = int(np.sum(self.theta < self.lam0))
        width = (m + (1 if m < n else 0)) if width is None else width
        npad = width - m - (1 if m < n else 0)
        assert npad >= 0
        ev = np.concatenate([self.theta[:pos], [self.lam0] * ((1 if m < n else 0) + npad), self.theta[pos:]])
        vg = np.concatenate([c[:pos], ([gam] if m < n else []) + [0.0] * npad, c[pos:]])
        rowmap = np.concatenate([np.arange(pos), ([-1] if m < n else []) + [-2] * npad, np.arange(pos, m)]).astype(int)
        return ev, vg, rowmap, gperp

    def lift(self, coef, rowmap, gperp):
        """s = sum_i coef_i * (eigenvector of pole i): explicit rows, and g_perp/|g_perp| for the cluster."""
        s = np.zeros(self.n)
        gam = np.linalg.norm(gperp)
        for ci, r in zip(coef, rowmap):
            if r >= 0:
                s += ci * self.VR[r]
            elif r == -1 and gam > 0:
                s += ci * gperp / gam
        return s

    # ------------------------------------------------------------------ Davidson correction (jd0)
    def jd0(self, r, v, theta_ritz):
        """t = -a + eps b, a = (B - theta)^-1 r, b = (B - theta)^-1 v, eps = (v.a)/(v.b)
        (eigensolvers.py:123-139 in P's eigenbasis)."""
        rh, vh = self.VR @ r, self.VR @ v
        d = self.theta - theta_ritz
        d0 = self.lam0 - theta_ritz
        full = self.m >= self.n
        va = np.sum(vh * rh / d) + (0.0 if full else (v @ r - vh @ rh) / d0)
        vb = np.sum(vh * vh / d) + (0.0 if full else (v @ v - vh @ vh) / d0)
        eps = va / vb
        that = (1.0 / d - (0.0 if full else 1.0 / d0)) * (eps * vh - rh)
        t = self.VR.T @ that
        if not full:
            t = t + (eps * v - r) / d0
        return t
