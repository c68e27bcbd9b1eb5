"""Batched saddle searches in redundant internal coordinates: the reference's ``InternalPES``
(sella/peswrapper.py:609-1288) + ``MaxInternalStep`` (optimize/restricted_step.py:186-243) over a
leading batch dimension, one coordinate list shared by the batch.

The optimiser's variable is q [b, nint] (translations, bonds, angles, dihedrals of
``BatchedInternals``), the approximate Hessian H is nint x nint, and every geometry carries its own
factorisation of the Wilson matrix Bw = dq/dx [nint, ncart]:

    Bw = Q R (economy QR, Unred = Q)                       _get_jacobian_qr     :674-709
    B+ = R^-1 Q^T,  g_int = g_cart B+                      _get_Binv / eval     :711-736, 1124-1127
    constraint rows J = Bw[rows]:  J R^-1 = (Vcons Rc)^T   _compute_basis_int   :1050-1082
    L = Rc^-1 Vcons^T Q^T g                                _update_basis        :467-481
    Hc = B+^T (Dc.L - Dq.L_int) B+ = Q Rinv^T D Rinv Q^T   _compute_Hc_int      :1011-1031

Everything the restricted step needs lives in the ncart-dimensional basis Q (ncart <= nint):
HL_r = Q^T H Q - Rinv^T D Rinv, the step model Bp = P_f HL_r P_f + sigma P_c (P_c = Vcons Vcons^T), one
ncart x ncart eigensolve per geometry, and the eigenvectors lifted back with ONE GEMM (W = Vt_r Q^T) so that
the max-internal-step search (sb_qn_mis / sb_rfo_mis) and the Davidson code of the Cartesian engine
(inherited: the preconditioner spectrum is (evals, W), zero rows beyond ncart) run unchanged.

Moving to a target q is the geodesic of peswrapper.py:841-880, 1200-1221, integrated on the device by
Dormand-Prince 5(4) steps with PER-SYSTEM step control (the reference calls scipy's LSODA with
atol 1e-6 and scipy's default rtol 1e-3; this scheme runs ten times tighter, atol 1e-7 / rtol 1e-4, and is restated in oracle/internal_pes.py for
step-by-step parity), followed by the Newton projection onto the constraint manifold (:928-994).

A rank-deficient Wilson matrix (a free molecule: the reference's SVD branch, :691-704) is served by the
eigendecomposition of Bw^T Bw instead of the QR (`_factor`); status bit 256 flags a system whose rank
differs from the batch's.  Not on the device (NotImplementedError): dummy atoms, the iterative stepper, the "violation alone exceeds the radius" branch of the
restricted step, re-detection of the coordinate list when an angle becomes linear (the engine
reports `bad_internals()`; the Sella shell rebuilds the object, optimize.py:382-410).
"""
import numpy as np
import torch

from . import kernels as K
from ._lib import I, D, LL, _p, _stream, call, check_f64
from .batched import BatchedSella, _Spec

SB_ST_WILSON_RANK = 256        # (1..64 are the library's SB_ST_* bits, include/sella_b200.h)
SB_ST_GEODESIC = 512

# Dormand-Prince 5(4)
_C = (0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
_A = ((),
      (1 / 5,),
      (3 / 40, 9 / 40),
      (44 / 45, -56 / 15, 32 / 9),
      (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
      (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
      (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84))
_B5 = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0)
_B4 = (5179 / 57600, 0.0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40)
RK_ATOL, RK_RTOL, RK_MAXSTEPS = 1e-7, 1e-4, 64


def _row(v):
    """[b, n] -> [b, 1, n] contiguous (a row vector per system for the thin GEMMs)."""
    return v.reshape(v.shape[0], 1, -1).contiguous()


class _SurfaceInQ:
    """What the inherited Davidson code sees as `surface`: energy and INTERNAL gradient at a target q,
    reached from the engine's current geometry along the geodesic (PES._calc_eg, peswrapper.py:420-427:
    save / set_x / eval / restore -- nothing is committed)."""

    def __init__(self, eng):
        self.eng = eng

    def evaluate(self, qt, f_out, g_out, active=None):
        eng = self.eng
        if active is not None:
            qt = torch.where(active.view(-1, 1) > 0, qt, eng.x)
        pos, geo = eng._set_x(qt)[:2]
        eng.cart.evaluate(pos, f_out, eng._gc)
        g_out.copy_(eng._to_internal(geo, eng._gc))


class BatchedInternalSella(BatchedSella):
    _internal = True

    def __init__(self, surface, pos0, ints, cons_rows=None, cons_targets=None, h0=None, H0=None,
                 weights=(1.0, 1.0, 1.0, 1.0), exact_geodesic=True, rs=None, atol=15.0, **kw):
        """surface: Cartesian evaluator (`evaluate(x[b, ncart], f, g, active)`); pos0 [b, ncart];
        ints: BatchedInternals; cons_rows: positions of the constrained coordinates within it, held at
        cons_targets [nc] / [b, nc] (None / NaN: their values at pos0); h0 [nint] or [b, nint]: diagonal model Hessian
        (Internals.guess_hessian), projected onto range(Bw) as the reference does, or H0 [b, nint, nint];
        weights: (wx, wb, wa, wd) of MaxInternalStep."""
        check_f64(pos0)
        for k in ("constraints", "hessian_function", "v0", "spectrum", "eig_mode"):
            if kw.get(k) is not None:
                raise NotImplementedError("%s is not available with internal coordinates" % k)
        if getattr(ints, "nrotations", 0):
            raise NotImplementedError("rotation coordinates are not available with internal=True")
        self.ints = ints
        self.cart = surface
        self.pos = pos0.clone()
        b, ncart = pos0.shape
        nint = ints.nint
        if nint < ncart:
            raise ValueError("the coordinate list has fewer coordinates (%d) than Cartesian degrees of freedom (%d): "
                             "its Wilson matrix cannot have full column rank" % (nint, ncart))
        self.ncart = ncart
        dev = pos0.device
        self.exact_geodesic = bool(exact_geodesic)
        self.atol = float(atol) * np.pi / 180.0
        rows = np.zeros(0, dtype=np.int64) if cons_rows is None else np.asarray(cons_rows, dtype=np.int64)
        self.nc = len(rows)
        self.nnull, self.svd_path = 0, False
        self.nfree_int = ncart - self.nc
        self.rows = torch.from_numpy(rows).to(dev)
        self._gc = torch.zeros(b, ncart, dtype=torch.float64, device=dev)
        self.geo = None
        q0 = ints.calc(self.pos)
        super().__init__(_SurfaceInQ(self), q0, rs="mis" if rs is None else rs, spectrum="dense", eig_mode="update",
                         **kw)
        if self.rs != "mis":
            raise NotImplementedError("internal coordinates run with rs='mis' (the reference's default for them)")
        n = self.n
        f64 = dict(dtype=torch.float64, device=dev)
        wx, wb, wa, wd = (float(w) for w in weights)
        self.wmis = torch.tensor([wx] * ints.ntrans + [wb] * ints.nbonds + [wa] * ints.nangles
                                 + [wd] * ints.ndihedrals, **f64)
        self.dih = (ints.ntrans + ints.nbonds + ints.nangles, ints.nstd)
        # the spectrum of H itself (|H| S of the TS-BFGS update) next to the model spectrum (evals, Vt)
        self.evalsB, self.VtB = torch.zeros(b, n, **f64), torch.zeros(b, n, n, **f64)
        tg = None
        if self.nc:
            cur = q0[:, self.rows]
            if cons_targets is None:
                tg = cur.clone()
            else:
                t = torch.from_numpy(np.array(np.broadcast_to(
                    np.asarray(cons_targets, dtype=np.float64), (b, self.nc)))).to(dev)
                tg = torch.where(torch.isnan(t), cur, t)
        self.targets = tg
        self.scons = torch.zeros(b, n, **f64)
        self.gfull = torch.zeros(b, n, **f64)
        self.g_par = torch.zeros(b, n, **f64)
        self.sr, self.dxf = torch.zeros(b, n, **f64), torch.zeros(b, n, **f64)
        self.evalsHL = torch.zeros(b, ncart, **f64)
        self.Vg_r = torch.zeros(b, ncart, **f64)
        self.first_diag = True
        self.ode_steps = 0
        # ---- H is held twice, as the compact engine with track_B does (batched.py): densely (`_B`, for
        # the projections Q^T H Q and H.v) and as explicit eigenpairs (evalsB, VtB rows) + the eigenvalue
        # lam0 = 0 on their complement -- the model Hessian P diag(h0) P has rank <= ncart and every secant
        # update adds rank 2, so the eigen-update works on a few hundred rows instead of nint
        if self.kcap > 16:
            raise NotImplementedError("kcap <= 16 with internal coordinates")
        zi = lambda *sh: torch.zeros(*sh, dtype=torch.int32, device=dev)      # noqa: E731
        self.sp = self.spB = _Spec(b, n, dev, self.evalsB, self.VtB)
        self.tracked_B = self._B
        self.lam0.zero_()
        import os
        self.split_rotation = os.environ.get("SB_SPLIT_ROTATION", "1") != "0"
        self.split_min_rows = int(os.environ.get("SB_SPLIT_MIN_ROWS", "24"))
        self.sec_aux, self.ncand = zi(b, n + 4), zi(b)
        for sec, zc in ((self.sec1, 2), (self.seck, 2 * self.kcap)):
            for k in ("W1", "Qc", "D2", "W2"):
                sec[k] = torch.zeros(b, zc, n, **f64)
        # ---- geometry at the start and the model Hessian (peswrapper.py:641-652)
        # rank of the Wilson matrix at the start decides the factorisation (one choice per batch: a free
        # molecule is rank deficient at every geometry, a slab with held atoms at none)
        sv = torch.linalg.svdvals(self.ints.jacobian(self.pos)[:1].cpu())[0]
        self.nnull = int((sv <= 1e-6).sum())
        self.svd_path = self.nnull > 0
        self.nfree_int = ncart - self.nc - self.nnull
        self.geo = self._geometry(self.pos)
        if H0 is None:
            if h0 is None:
                raise ValueError("internal coordinates need the diagonal model Hessian h0 (Internals.guess_hessian) "
                                 "or a full H0")
            h = torch.from_numpy(np.array(h0, dtype=np.float64)).to(dev).abs()
            Qm = self.geo["Q"]
            core = K.gemm(Qm, (h.reshape(-1, n, 1) * Qm).contiguous(), transA=True)    # Q^T diag(h0) Q
            th, Wc, _ = K.eigh((0.5 * (core + core.transpose(1, 2))).contiguous(), status=self.status)
            rows = K.gemm(Wc, Qm, transB=True)           # eigenvectors of P diag(h0) P = Q core Q^T, lifted
            m = ncart - self.nnull                       # (rank-deficient Wilson matrix: the null columns of Q
            self.VtB[:, :m] = rows[:, self.nnull:]       # give exact zero eigenvalues, the lowest ones -- skipped)
            self.evalsB[:, :m] = th[:, self.nnull:]
            self._B.copy_(K.gemm(rows, (th[:, :, None] * rows).contiguous(), transA=True))
        else:
            check_f64(H0)
            m = n
            self._B.copy_(0.5 * (H0 + H0.transpose(1, 2)))
            K.eigh(self._B, evals=self.evalsB, Vt=self.VtB, ws=self.eig_ws, status=self.status)
        self._B.copy_(0.5 * (self._B + self._B.transpose(1, 2)))
        self.sp.mrows.fill_(m)
        self.sp.rb = self.sp.mmin = m
        self.H_initialized = True
        self.eig_valid = True
        self.Vt.zero_()

    # ------------------------------------------------------------------ Wilson-matrix algebra per geometry
    def _wrap(self, v):
        lo, hi = self.dih
        if hi > lo:
            v[:, lo:hi] = torch.remainder(v[:, lo:hi] + np.pi, 2.0 * np.pi) - np.pi
        return v

    def _factor(self, Bw, want_q=True, idx=None):
        """(Q, Rinv) with Bw = Q Rinv^-1 on range(Bw) and B+ = Rinv Q^T (peswrapper.py:674-736).

        Full column rank (slabs / crystals with held atoms): R = chol(Bw^T Bw), Rinv = R^-1, Q = Bw Rinv (the
        factors of the economy QR up to signs).  Rank deficient (a free
        molecule: the six rigid-body directions, the reference's SVD branch :691-704): the same two factors
        from the eigendecomposition of Bw^T Bw = V S^2 V^T -- Rinv = V S^-1, Q = Bw Rinv with ZERO columns for
        the null directions (singular values <= 1e-6 as in the reference); eigenvalues come out ascending, so
        the null directions are the first `nnull` columns of every system."""
        if not self.svd_path:
            # R from the Cholesky factor of Bw^T Bw (one GEMM + n/32 small panels instead of a Householder QR),
            # Q = Bw R^-1 where it is needed: orthonormal to kappa(Bw)^2 eps ~ 1e-14 for the Wilson matrices of
            # bonded networks (kappa ~ 10), the accuracy class of the reference's own B+ = R^-1 Q^T
            G = K.gemm(Bw, Bw, transA=True)
            R, st0 = K.potrf((0.5 * (G + G.transpose(1, 2))).contiguous())
            Rinv, st = K.trtri(R)
            st = st | st0
            Q = K.gemm(Bw, Rinv) if want_q else None
            rd = torch.diagonal(R, dim1=1, dim2=2).abs()
            bad = (rd.min(dim=1).values < 1e-6 * rd.max(dim=1).values).to(torch.int32) * SB_ST_WILSON_RANK | st
            if idx is None:
                self.status |= bad
            else:
                self.status[idx] |= bad
            return Q, Rinv
        G = K.gemm(Bw, Bw, transA=True)
        st = torch.zeros(Bw.shape[0], dtype=torch.int32, device=self.dev)
        w, Vt, _ = K.eigh((0.5 * (G + G.transpose(1, 2))).contiguous(), status=st)
        keep = w > 1e-12
        st |= (keep.sum(dim=1) != self.ncart - self.nnull).to(torch.int32) * SB_ST_WILSON_RANK
        if idx is None:
            self.status |= st
        else:
            self.status[idx] |= st
        sinv = torch.where(keep, w.clamp(min=1e-300).rsqrt(), torch.zeros_like(w))
        Rinv = (Vt * sinv[:, :, None]).transpose(1, 2).contiguous()
        return (K.gemm(Bw, Rinv) if want_q else None), Rinv

    def _geometry(self, pos):
        """q, the factors of the Wilson matrix and the constraint basis at `pos` [b, ncart]."""
        q, Bw = self.ints.calc(pos, jacobian=True)
        Q, Rinv = self._factor(Bw)
        geo = dict(pos=pos, q=q, Bw=Bw, Q=Q, Rinv=Rinv)
        if self.nc:
            J = Bw[:, self.rows].contiguous()                               # cons.jacobian(): rows of Bw
            red = K.gemm(J, Rinv)                                            # drdx in the basis Q [b, nc, ncart]
            Vc, Rc = K.qr(red.transpose(1, 2).contiguous())                  # red^T = Vcons Rc
            Rcinv, st2 = K.trtri(Rc)
            rc = torch.diagonal(Rc, dim1=1, dim2=2).abs()
            self.status |= (rc.min(dim=1).values < 1e-6 * rc.max(dim=1).values).to(torch.int32) * 16
            self.status |= st2
            geo.update(J=J, red=red, Vc=Vc, Rc=Rc, Rcinv=Rcinv,
                       res=self._wrap(q[:, self.rows] - self.targets) if self.targets is not None
                       else torch.zeros_like(q[:, self.rows]))
        return geo

    def _to_internal(self, geo, gc):
        """g_int = g_cart B+ = Q (Rinv^T g_cart)."""
        t = K.gemm(_row(gc), geo["Rinv"])
        return K.gemm(t, geo["Q"], transB=True).view(self.batch, self.n)

    def _binv(self, geo, v):
        """B+ v = Rinv (Q^T v) for v [b, k, nint] -> [b, k, ncart]."""
        return K.gemm(K.gemm(v, geo["Q"]), geo["Rinv"], transB=True)

    def _free_project(self, geo, V):
        """P_f V for V [b, k, nint]: P_f = Q (I - Vcons Vcons^T) Q^T."""
        r = K.gemm(V, geo["Q"])
        if self.nc:
            r = r - K.gemm(K.gemm(r, geo["Vc"]), geo["Vc"], transB=True)
        return K.gemm(r.contiguous(), geo["Q"], transB=True)

    def _hc_apply(self, geo, V):
        """Hc V = Q (Rinv^T D Rinv) Q^T V for V [b, k, nint]."""
        return K.gemm(K.gemm(K.gemm(V, geo["Q"]), geo["HcR"]), geo["Q"], transB=True)

    def _model(self):
        """Multipliers, constraint Hessian and the spectrum of the step model at the current geometry,
        gradient and H (get_HL_projected, peswrapper.py:363-386, in the basis Q)."""
        geo = self.geo
        b, n, nc, ncart = self.batch, self.n, self.nc, self.ncart
        Q, Rinv = geo["Q"], geo["Rinv"]
        Hr = K.gemm(Q, K.gemm(self._B, Q), transA=True)                     # Q^T H Q
        hnorm = self._hnorm()
        if nc:
            gr = K.gemm(_row(self.g), Q)                                    # (Q^T g)^T
            L = K.gemm(K.gemm(gr, geo["Vc"]), geo["Rcinv"], transB=True)    # Rc^-1 Vcons^T Q^T g  [b, 1, nc]
            Lint = K.gemm(K.gemm(L, geo["red"]), Q, transB=True)            # L J B+
            vq = -Lint.view(b, n)
            vq[:, self.rows] += L.view(b, nc)
            Dm = self.ints.ldot(self.pos, vq.contiguous())                  # Dc.L - Dq.L_int
            HcR = K.gemm(Rinv, K.gemm(Dm, Rinv), transA=True)
            geo["HcR"], geo["L"] = HcR, L.view(b, nc)
            HLr = Hr - HcR
            hnorm = hnorm + HcR.flatten(1).norm(dim=1)
            Pc = K.gemm(geo["Vc"], geo["Vc"], transB=True)
            Pf = (torch.eye(ncart, dtype=torch.float64, device=self.dev) - Pc).contiguous()
            Bp = K.gemm(Pf, K.gemm(HLr.contiguous(), Pf))
            sigma = 1.0 + 8.0 * hnorm
            Bp = 0.5 * (Bp + Bp.transpose(1, 2)) + sigma[:, None, None] * Pc
        else:
            HLr = Hr
            Bp = 0.5 * (Hr + Hr.transpose(1, 2))
            sigma = 1.0 + 8.0 * hnorm
        HLr = 0.5 * (HLr + HLr.transpose(1, 2))
        if self.nnull:
            # the null directions of the Wilson matrix (zero columns of Q: exact zero rows / columns of HL_r) are not
            # part of Unred: lift them to sigma like the constraint directions
            k = self.nnull
            dz = sigma[:, None] * torch.ones(k, dtype=torch.float64, device=self.dev)[None, :]
            Bp.diagonal(dim1=1, dim2=2)[:, :k] += dz
            HLr.diagonal(dim1=1, dim2=2)[:, :k] += dz
        geo["HLr"] = HLr
        evr, Vtr, _ = K.eigh(Bp.contiguous(), status=self.status)
        geo["evr"], geo["Vtr"] = evr, Vtr
        # lifted eigenvectors: rows of W = Vt_r Q^T; the model spectrum the Davidson code and the step use
        self.Vt[:, :ncart] = K.gemm(Vtr, Q, transB=True)
        self.evals[:, :ncart] = evr
        self.evals[:, ncart:] = sigma[:, None]
        geo["model"] = True

    # ------------------------------------------------------------------ geodesic
    def _rhs(self, y, geo0, idx=None):
        """peswrapper.py:1200-1221 for y = (x, dx/dt, g) [m, 3, ncart]; idx: the systems these m rows belong to
        (None: the whole batch, in order)."""
        pos = y[:, 0].contiguous()
        # B+ w = Rinv Rinv^T Bw^T w (semi-normal equations): no Q, which is half of a QR
        if self.exact_geodesic:
            Bw = self.ints.jacobian(pos)
            Rinv = self._factor(Bw, want_q=False, idx=idx)[1]
        elif idx is None:
            Bw, Rinv = geo0["Bw"], geo0["Rinv"]
        else:
            Bw, Rinv = geo0["Bw"][idx].contiguous(), geo0["Rinv"][idx].contiguous()
        Rd = self.ints.rdot(pos, y[:, 1].contiguous())                      # [m, nint, ncart]
        X = y[:, 1:3].contiguous()
        u = K.gemm(K.gemm(K.gemm(X, Rd, transB=True), Bw), Rinv)            # (Rinv^T Bw^T (Rd X))^T
        out = K.gemm(u, Rinv, transB=True)
        return torch.cat([y[:, 1:2], -out], dim=1)

    def _integrate(self, y0, geo0):
        """Dormand-Prince 5(4) from t = 0 to 1 with a step size per system (oracle/internal_pes.py:_rk).
        Systems that have arrived leave the working set: every further iteration (rejected or shortened steps
        of the few systems with a strongly curved path) runs on the unfinished ones only."""
        b = self.batch
        f64 = dict(dtype=torch.float64, device=self.dev)
        yout = y0.clone()
        idx = None                                   # None: all systems, in order
        y, t, h = y0, torch.zeros(b, **f64), torch.ones(b, **f64)
        k1 = self._rhs(y, geo0)
        for _ in range(RK_MAXSTEPS):
            m = y.shape[0]
            h = torch.minimum(h, 1.0 - t)
            hh = h.view(m, 1, 1)
            ks = [k1]
            for s in range(1, 7):
                acc = None
                for a, k in zip(_A[s], ks):
                    if a != 0.0:
                        acc = a * k if acc is None else acc + a * k
                ks.append(self._rhs(y + hh * acc, geo0, idx))
            inc = sum(w * k for w, k in zip(_B5, ks) if w != 0.0)
            e = hh * sum((w5 - w4) * k for w5, w4, k in zip(_B5, _B4, ks))
            ynew = y + hh * inc
            scale = RK_ATOL + RK_RTOL * torch.maximum(y.abs(), ynew.abs())
            err = (e.abs() / scale).flatten(1).max(dim=1).values
            fac = torch.where(err == 0.0, torch.full_like(err, 5.0),
                              torch.clamp(0.9 * err.clamp(min=1e-300) ** -0.2, 0.2, 5.0))
            ok = err <= 1.0
            okm = ok.view(m, 1, 1)
            y = torch.where(okm, ynew, y)
            k1 = torch.where(okm, ks[6], k1)
            t = torch.where(ok, t + h, t)
            done = ok & (t >= 1.0 - 1e-14)
            h = h * fac
            self.ode_steps += 1
            ndone = int(done.sum().item())
            if ndone:
                if idx is None:
                    yout = torch.where(done.view(m, 1, 1), y, yout)
                else:
                    yout[idx[done]] = y[done]
            if ndone == m:
                break
            if ndone:
                keep = ~done
                idx = torch.nonzero(keep).flatten() if idx is None else idx[keep]
                y, k1, t, h = y[keep].contiguous(), k1[keep].contiguous(), t[keep], h[keep]
        else:
            left = torch.arange(b, device=self.dev) if idx is None else idx
            self.status[left] |= SB_ST_GEODESIC
            yout[left] = y
        return yout

    def _set_x(self, target):
        """InternalPES.set_x (peswrapper.py:883-903) from the current geometry, WITHOUT committing:
        returns (pos, geo, dx_initial, dx_final, g_par)."""
        b, n = self.batch, self.n
        geo0 = self.geo
        dx = self._wrap(target - self.x)
        g0 = self.g if self._evaluated else torch.zeros_like(dx)
        vg = self._binv(geo0, torch.stack([dx, g0], dim=1).contiguous())    # B+ dx, B+ g
        y = self._integrate(torch.cat([self.pos.view(b, 1, -1), vg], dim=1).contiguous(), geo0)
        pos = y[:, 0].contiguous()
        geo = self._geometry(pos)
        fin = K.gemm(y[:, 1:3].contiguous(), geo["Bw"], transB=True)        # B y1 (tangent displacement), B y2
        dx_final, g_par = fin[:, 0].contiguous(), fin[:, 1].contiguous()
        if self.nc:
            pos, geo, dx_final = self._project_to_constraints(pos, geo, dx_final)
        return pos, geo, dx, dx_final, g_par

    def _project_to_constraints(self, pos, geo, dx_final, target_tol=1e-7, max_iter=8, safety_limit=0.05):
        """peswrapper.py:928-994 + _add_proj_delta :905-926: Newton steps in the constraint subspace,
        dq = Ucons s with (drdx Ucons) s = -r, i.e. dx_cart = Rinv Vcons Rc^-T (-r)."""
        q_after = geo["q"]
        moved = torch.zeros(self.batch, dtype=torch.bool, device=self.dev)
        go = torch.ones(self.batch, dtype=torch.bool, device=self.dev)
        for _ in range(max_iter):
            r = geo["res"]
            go = go & (r.abs().max(dim=1).values >= target_tol)
            if not bool(go.any().item()):
                break
            s = -K.gemm(_row(r), geo["Rcinv"])                               # (-Rc^-T r)^T
            dxc = K.gemm(K.gemm(s, geo["Vc"], transB=True), geo["Rinv"], transB=True).view(self.batch, -1)
            go = go & (dxc.abs().max(dim=1).values <= safety_limit)
            if not bool(go.any().item()):
                break
            pos = torch.where(go.view(-1, 1), pos + dxc, pos)
            moved = moved | go
            geo = self._geometry(pos)
        if bool(moved.any().item()):
            dq = self._wrap(geo["q"] - q_after)
            dx_final = torch.where(moved.view(-1, 1), dx_final + dq, dx_final)
        return pos, geo, dx_final

    # ------------------------------------------------------------------ hooks of the inherited Davidson code
    def _free_dim(self):
        return self.nfree_int

    def _diag_start_vector(self, part):
        """peswrapper.py:521-528: the FIRST diagonalisation starts from Ufree^T g although a model exists."""
        if not self.first_diag:
            return None
        return self._free_project(self.geo, _row(self.g)).view(self.batch, self.n)

    def _hvp(self, vec, vstride, mask, maskval, active):
        b, n, kc = self.batch, self.n, self.kcap
        call("sb_hvp_prepare", _p(vec), LL(vstride), _p(self.x), _p(self.g), D(self.eta), _p(self.xdisp),
             _p(self.signnorm), I(n), _p(mask), I(maskval), I(b), _stream())
        self.surface.evaluate(self.xdisp, self.fplus, self.gplus, active=active)
        gbase, eta_eff = self.g, self.eta
        if self.threepoint:
            call("sb_hvp_prepare", _p(vec), LL(vstride), _p(self.x), _p(self.g), D(-self.eta), _p(self.xminus),
                 _p(self.signnorm), I(n), _p(mask), I(maskval), I(b), _stream())
            self.surface.evaluate(self.xminus, self.fminus, self.gminus, active=active)
            gbase, eta_eff = self.gminus, 2.0 * self.eta
        kslot = self.ksz.clone()
        call("sb_hvp_finish", _p(vec), LL(vstride), _p(self.gplus), _p(gbase), _p(self.signnorm), D(eta_eff),
             _p(self.AV), _p(self.Vs), _p(self.AVs), I(kc), _p(self.ksz), _p(self.nhist), I(n),
             _p(mask), I(maskval), I(b), _stream())
        # the operator of peswrapper.py:531-537 is Ufree^T (H_fd - Hc) Ufree: subtract Hc v in the slot just
        # filled, then project the block onto the free space (idempotent for the slots filled earlier)
        if self.nc:
            v = vec if vec.dim() == 2 and vec.is_contiguous() else vec.contiguous()
            hcv = self._hc_apply(self.geo, _row(v)).view(b, n)
            idx = kslot.clamp(max=kc - 1).long().view(b, 1, 1).expand(b, 1, n)
            sel = ((mask == maskval) if mask is not None else torch.ones(b, dtype=torch.bool, device=self.dev))
            cur = torch.gather(self.AV, 1, idx).view(b, n)
            self.AV.scatter_(1, idx, torch.where(sel.view(b, 1), cur - hcv, cur).view(b, 1, n))
        self.AV.copy_(self._free_project(self.geo, self.AV))

    def _flush_history(self, part, nv, nl):
        """PES.diag tail (peswrapper.py:541-551): Atilde = Vs^T sym(Vs, AVs) - Vs^T Hc Vs, Ritz rotation of the
        operator history, one block update of H."""
        hcvs = self._hc_apply(self.geo, self.Vs) if self.nc else None
        call("sb_history_ritz", _p(self.Vs), _p(self.AVs), I(self.kcap), _p(self.nhist), I(self.n), _p(self.nvec),
             _p(self.dav_state), _p(self.status), _p(hcvs), I(self.batch), _stream())
        self._update(self.Vs, self.AVs, self.upk, self.nvec, nv, part)

    def _hnorm(self):
        """[b] largest |eigenvalue| of H (explicit pairs; the complement's eigenvalue is 0)."""
        R = max(1, self.sp.rb)
        live = torch.arange(R, device=self.dev)[None, :] < self.sp.mrows[:, None]
        return (self.evalsB[:, :R].abs() * live).max(dim=1).values

    def _update(self, S, Y, bufs, kvec, nv, active, bs_ready=False, abs_ready=False):
        """ApproximateHessian.update (linalg.py:274-304): the compact eigen-update of the Cartesian engine on H's
        explicit eigenpairs (complement eigenvalue 0), the dense copy carried along by sb_update_apply."""
        self._update_compact(S, Y, bufs, kvec, nv, active)
        self._sync_rows()                 # exact bounds on the explicit-row counts (one small read)

    def _run_diag(self, part):
        if not self.geo.get("model"):
            self._model()
        self._diag(part)
        self.first_diag = False
        self.geo["model"] = False                  # H has changed

    # ------------------------------------------------------------------ public
    def ensure_evaluated(self):
        if not self._evaluated:
            self.cart.evaluate(self.pos, self.f, self._gc)
            self.g.copy_(self._to_internal(self.geo, self._gc))
            self._evaluated = True

    def step(self, active=None):
        """One Sella.step (optimize.py:317-440) for every system."""
        if active is not None:
            raise NotImplementedError("per-system masks are not available with internal coordinates")
        b, n, ncart, nc = self.batch, self.n, self.ncart, self.nc
        if not self.initialized:
            self.ensure_evaluated()
            if self.eig:
                self._run_diag(None)
                self.since_diag.fill_(-1)
            self.initialized = True
        geo = self.geo
        if not geo.get("model"):
            self._model()
        # ---- restricted step (restricted_step.py:28-62): g' = Ufree^T (g + H scons)
        sadd = None
        if nc:
            sc = -K.gemm(_row(geo["res"]), geo["Rcinv"])                     # -Rc^-T res
            self.scons.copy_(K.gemm(K.gemm(sc, geo["Vc"], transB=True), geo["Q"], transB=True).view(b, n))
            K.hv_ld(self._B, self.scons.view(b, 1, n), self.gfull.view(b, 1, n), 1)
            self.gfull.add_(self.g)
            sadd = self.scons
        else:
            self.gfull.copy_(self.g)
        gr = K.gemm(_row(self.gfull), geo["Q"])
        self.Vg_r.copy_(K.gemm(gr, geo["Vtr"], transB=True).view(b, ncart))
        if nc + self.nnull:
            # the poles at sigma are the constraint (and null) directions: Ufree^T removes them
            self.Vg_r[:, ncart - nc - self.nnull:] = 0.0
        if self.method == "qn":
            call("sb_qn_mis", _p(self.Vg_r), _p(geo["evr"]), _p(self.Vt), _p(self.delta), I(self.order), I(n),
                 _p(self.s), _p(self.smag), _p(self.alpha), _p(self.status), _p(None), _p(sadd), I(ncart),
                 LL(n * n), _p(self.wmis), I(b), _stream())
        else:
            call("sb_rfo_mis", _p(self.Vg_r), _p(geo["evr"]), _p(self.Vt), _p(self.delta), I(self.order), I(n),
                 I(1 if self.method == "prfo" else 0), _p(self.s), _p(self.smag), _p(self.alpha), _p(self.status),
                 _p(None), _p(sadd), I(ncart), LL(n * n), _p(self.wmis), I(b), _stream())
        # ---- re-diagonalise?  eigenvalues of Unred^T (H - Hc) Unred (optimize.py:362-371)
        ev_evals = geo["evr"]
        if self.eig and int((self.since_diag >= int(self._ipar[2])).any().item()):
            K.eigvalsh(geo["HLr"].contiguous(), evals=self.evalsHL, status=self.status)
            ev_evals = self.evalsHL
        call("sb_ev_decide", _p(ev_evals), I(ncart), I(1), _p(self.since_diag), _p(self.ev), self._dpar,
             self._ipar, _p(None), I(b), _stream())
        # ---- kick (peswrapper.py:578-602): geodesic to q + s, evaluate, rho, secant update
        self._kick_core(self.s)
        if int(self.ev.sum().item()) > 0:
            self._run_diag(self.ev)

    def _kick_core(self, s):
        b, n = self.batch, self.n
        target = self.x + s
        pos, geo, dx_i, dx_f, g_par = self._set_x(target)
        self.cart.evaluate(pos, self.fnew, self._gc)
        self.gnew.copy_(self._to_internal(geo, self._gc))
        # get_x: dihedrals continue from the previous values (peswrapper.py:996-1008)
        self.xnew.copy_(self.x + self._wrap(geo["q"] - self.x))
        # get_df_pred (:1176-1183) with Unred of the NEW geometry: dx_r = Q^T dx_initial
        self.sr.copy_(K.gemm(K.gemm(_row(dx_i), geo["Q"]), geo["Q"], transB=True).view(b, n))
        K.hv_ld(self._B, self.sr.view(b, 1, n), self.up1["BS"], 1)
        self.g_par.copy_(g_par)
        self.dxf.copy_(dx_f)
        call("sb_kick_finish", _p(self.x), _p(self.f), _p(self.g), _p(self.xnew), _p(self.fnew), _p(self.gnew),
             _p(self.sr), _p(self.up1["BS"]), _p(self.smag), _p(self.dg), _p(self.delta), _p(self.rho),
             _p(self.nsteps), self._dpar, self._ipar, I(n), _p(None), I(b), _stream())
        torch.sub(self.g, self.g_par, out=self.dg)             # dg = g(new) - g transported along the geodesic
        self.pos, self.geo = pos, geo
        self._update(self.dxf.view(b, 1, n), self.dg.view(b, 1, n), self.up1, None, 1, None)

    def kick(self, dx, diag=False):
        self.ensure_evaluated()
        keep = self.delta.clone()
        self.smag.copy_((dx * self.wmis).abs().max(dim=1).values)
        self._kick_core(dx)
        self.delta.copy_(keep)
        if diag:
            self._run_diag(None)
        return self.rho.clone()

    def converged(self, fmax, cmax=1e-5):
        """InternalPES.get_projected_forces + PES.converged (peswrapper.py:1185-1194, 558-568)."""
        self.ensure_evaluated()
        geo = self.geo
        pg = self._free_project(geo, _row(self.g))
        fc = K.gemm(pg, geo["Bw"]).view(self.batch, self.ncart).contiguous()
        call("sb_converged", _p(fc), I(self.ncart), D(float(fmax)), _p(self.fmax), _p(self.conv), I(self.batch),
             _stream())
        if self.nc:
            self.conv.mul_((geo["res"].norm(dim=1) < cmax).to(torch.int32))
        return self.conv

    def bad_internals(self):
        """check_for_bad_internals (internal.py:3704-3736): [b] bool, an angle within atol of 0 or pi."""
        lo = self.ints.ntrans + self.ints.nbonds
        hi = lo + self.ints.nangles
        if hi == lo:
            return torch.zeros(self.batch, dtype=torch.bool, device=self.dev)
        a = self.x[:, lo:hi]
        return ~((a > self.atol) & (a < np.pi - self.atol)).all(dim=1)

    @property
    def B(self):
        return self._B

    def lowest_evals(self):
        """[b] lowest eigenvalue of H (0 from the complement while not every direction is explicit)."""
        R = max(1, self.sp.rb)
        live = torch.arange(R, device=self.dev)[None, :] < self.sp.mrows[:, None]
        lo = torch.where(live, self.evalsB[:, :R], torch.full_like(self.evalsB[:, :R], float("inf"))).min(dim=1).values
        return torch.where(self.sp.mrows < self.n, torch.clamp(lo, max=0.0), lo)
