"""Batched, device-resident saddle searches: many independent Sella runs advance
in lock step, one CTA (or tile) per system inside every kernel.

Host code is plain Python orchestration of the C-ABI kernels; all state stays in
HBM between steps and the only host<->device traffic on the hot path is one
4-byte "does any system re-diagonalise / keep expanding" flag per decision.

Mirrors, per system, the reference's
  Sella.step / _predict_step      sella/optimize/optimize.py:317-440
  PES.kick / PES.diag             sella/peswrapper.py:508-602
  ApproximateHessian.update       sella/linalg.py:274-304
for the Cartesian, unconstrained case (Ufree = I: `proj_trans=False,
proj_rot=False` in reference terms), quasi-Newton step model, trust-region or
restricted-atomic-step constraint.
"""
import ctypes

import numpy as np
import torch

from . import kernels as K
from ._lib import I, D, LL, _p, _stream, call, check_f64, require_cuda

_DEFAULTS = dict(
    minimum=dict(delta0=1e-1, sigma_inc=1.15, sigma_dec=0.90, rho_inc=1.035,
                 rho_dec=100.0, method="qn", eig=False),
    saddle=dict(delta0=0.1, sigma_inc=1.15, sigma_dec=0.65, rho_inc=1.035,
                rho_dec=5.0, method="prfo", eig=True),
)
# BFGS / BFGS_auto write 2k secant pairs and are served by the one-system mirror
# (sella_b200.hessian_update.update_H); the batched engine keeps k pairs per update
_UPDATE_METHODS = {"TS-BFGS": 0, "PSB": 1, "Greenstadt": 2, "DFP": 3, "SR1": 5}
_EIGENSOLVERS = {"jd0": 0, "jd0_alt": 0, "gd": 1, "lanczos": 2, "mjd0": 3, "mjd0_alt": 3}
_QN_NAMES = ("qn", "quasi-newton", "quasi newton", "newton", "mmf",
             "minimum mode following", "minimum-mode following", "dimer")
_TR_NAMES = ("tr", "trust region", "trust-region", "trust radius", "trust-radius")
_RAS_NAMES = ("ras", "restricted atomic step")
_MIS_NAMES = ("mis", "max internal step")

DAV_EXPAND, DAV_DONE, DAV_IDLE = 0, 1, 2


class QuadraticSurface:
    """Device evaluator of the synthetic surfaces of SURVEY.md 8d
    (f = 1/2 (x-x*)^T A (x-x*)).  Any object with the same `evaluate` signature can
    be plugged into BatchedSella (that is the PES plug-in boundary)."""

    def __init__(self, A, xstar):
        check_f64(A, xstar)
        self.A, self.xstar = A, xstar
        self.batch, self.n = xstar.shape
        self._work = torch.empty_like(xstar)
        self.neval = 0

    def evaluate(self, x, f_out, g_out, active=None):
        self.neval += 1
        call("sb_quadratic_pes", _p(self.A), _p(self.xstar), _p(x), _p(f_out), _p(g_out),
             _p(self._work), _p(active), I(self.batch), I(self.n), _stream())


class _Spec:
    """One compact spectrum: mrows[b] explicit eigenpairs in the first rows of (evals, Vt); every other
    eigenvalue is lam0[b] (shared by the spectra of B and of its projection).  rb / mmin: host-side upper /
    lower bounds on mrows (rows the passes visit; mmin == n: no complement left in any system)."""

    def __init__(self, batch, n, dev, evals=None, Vt=None):
        f64 = dict(dtype=torch.float64, device=dev)
        self.evals = torch.zeros(batch, n, **f64) if evals is None else evals
        self.Vt = torch.zeros(batch, n, n, **f64) if Vt is None else Vt
        self.mrows = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.rb = 0
        self.mmin = 0

    def reset(self):
        self.mrows.zero_()
        self.rb = self.mmin = 0


class BatchedSella:
    def __init__(self, surface, x0, order=1, delta0=None, sigma_inc=None, sigma_dec=None,
                 rho_dec=None, rho_inc=None, eig=None, eta=1e-4, method=None, gamma=0.1,
                 rs=None, nsteps_per_diag=3, diag_every_n=None, diag_maxiter=None,
                 eigensolver="jd0", update_method="TS-BFGS", kcap=16, eig_mode="update",
                 eig_refresh_every=0, constraints=None, threepoint=False, hessian_function=None, v0=None,
                 spectrum=None, track_B=False):
        require_cuda()
        d = _DEFAULTS["minimum" if order == 0 else "saddle"]
        self.surface = surface
        check_f64(x0)
        self.batch, self.n = x0.shape
        b, n = self.batch, self.n
        dev = x0.device
        self.dev = dev
        self.order = int(order)
        method = (d["method"] if method is None else method).lower()
        if method in _QN_NAMES:
            self.method = "qn"
        elif method in ("rfo", "rational function optimization"):
            self.method = "rfo"
        elif method in ("prfo", "p-rfo", "partitioned rational function optimization"):
            self.method = "prfo"
        else:
            raise ValueError("Unknown stepper name: {}".format(method))
        rs = "ras" if rs is None else rs
        if rs in _TR_NAMES:
            self.rs = "tr"
        elif rs in _RAS_NAMES:
            self.rs = "ras"
            if n % 3:
                raise ValueError("restricted atomic step needs 3N coordinates")
        elif rs in _MIS_NAMES:
            if not getattr(self, "_internal", False):          # restricted_step.py:192-196
                raise ValueError("Internal coordinates are required for the MaxInternalStep trust region method")
            self.rs = "mis"
        else:
            raise ValueError("Unknown restricted step name: {}".format(rs))
        self.eig = d["eig"] if eig is None else bool(eig)
        self.eta = float(eta)
        self.gamma = float(gamma)
        self.threepoint = bool(threepoint)
        # hessian_function(x[b,n]) -> B[b,n,n] (device tensors) replaces every Davidson diagonalisation
        # (peswrapper.py:597-606, optimize.py:321-324); v0[b,n]: start vector of the first one (:524)
        self.hessian_function = hessian_function
        self.v0 = v0
        self.diag_maxiter = diag_maxiter
        if eigensolver not in _EIGENSOLVERS:
            raise NotImplementedError("eigensolver %r is not available on the batched path" % eigensolver)
        self.eigensolver = _EIGENSOLVERS[eigensolver]
        self.update_method = _UPDATE_METHODS[update_method]
        self.kcap = int(kcap)
        assert 2 <= self.kcap <= 32
        # "update": keep (evals, Vt) current through the secular-equation update of
        # every low-rank Hessian update (block updates of more than 16 secant pairs -- kcap > 16 --
        # are followed by a full eigensolve instead); "direct": full eigensolve whenever the
        # spectrum is needed (the reference's behaviour).
        if eig_mode not in ("update", "direct"):
            raise ValueError("eig_mode must be 'update' or 'direct'")
        self.eig_mode = eig_mode
        if spectrum not in (None, "compact", "dense"):
            raise ValueError("spectrum must be 'compact' or 'dense'")
        self._spectrum_request = spectrum
        self.eig_refresh_every = int(eig_refresh_every)
        self._updates_since_refresh = 0

        delta0 = d["delta0"] if delta0 is None else delta0
        delta_init = delta0 if self.rs in ("ras", "mis") else delta0 * n      # optimize.py:183-187
        self._dpar = (ctypes.c_double * 5)(
            d["rho_inc"] if rho_inc is None else rho_inc,
            d["rho_dec"] if rho_dec is None else rho_dec,
            d["sigma_inc"] if sigma_inc is None else sigma_inc,
            d["sigma_dec"] if sigma_dec is None else sigma_dec,
            self.eta)
        self._ipar = (ctypes.c_int * 4)(self.order, int(self.eig), int(nsteps_per_diag),
                                        -1 if diag_every_n is None else int(diag_every_n))

        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        z = lambda *s: torch.zeros(*s, **f64)      # noqa: E731
        zi = lambda *s: torch.zeros(*s, **i32)     # noqa: E731
        kc = self.kcap
        # geometry / surface values
        self.x = x0.clone()
        self.f, self.g = z(b), z(b, n)
        self.xnew, self.fnew, self.gnew = z(b, n), z(b), z(b, n)
        self.xdisp, self.fplus, self.gplus = z(b, n), z(b), z(b, n)
        if self.threepoint:
            self.xminus, self.fminus, self.gminus = z(b, n), z(b), z(b, n)
        # step state
        self.s, self.dg, self.Vg, self.coef = z(b, n), z(b, n), z(b, n), z(b, n)
        self.delta = torch.full((b,), float(delta_init), **f64)
        self.rho = torch.ones(b, **f64)
        self.smag, self.alpha = z(b), z(b)
        self.nsteps, self.since_diag, self.ev = zi(b), zi(b), zi(b)
        self.status = zi(b)
        self.fmax, self.conv = z(b), zi(b)
        # approximate Hessian and its spectrum
        # "compact": B = lam0 I + VR^T diag(theta - lam0) VR with mrows[b] explicit eigenpairs in the first rows
        # of (evals, Vt) and lam0 on the rest (csrc/compact.cu); the dense matrix is only materialised on
        # request (property B).  "dense": B [b,n,n] and a full (evals, Vt), as in round 1.
        want = self._spectrum_request
        fixed = self._single_coordinate_constraints(constraints, x0)
        can = (eig_mode == "update" and (constraints is None or (fixed is not None and hessian_function is None))
               and self.eigensolver != 3 and self.kcap <= 16 and n <= 1536)
        if want == "compact" and not can:
            raise NotImplementedError("spectrum='compact' needs eig_mode='update', kcap <= 16, an eigensolver other "
                                      "than mjd0 and either no constraints or fixed Cartesian coordinates that hold "
                                      "at x0 (everything else runs on the dense representation)")
        self.compact = can and want != "dense"
        # fixed Cartesian coordinates (what Constraints.fix_translation(i) yields; peswrapper.py:51-69 makes
        # Ufree a signed permutation then): the projection onto the free space is a 0/1 mask
        self.fmask = None
        if self.compact and fixed is not None and len(fixed):
            mk = np.ones(n)
            mk[fixed] = 0.0
            self.fmask = torch.from_numpy(mk).to(dev)
            self.nfree = n - len(fixed)
            constraints = None
        self._B = None if self.compact else z(b, n, n)
        # compact only: ALSO carry the dense matrix through every update (sb_update_apply), as an
        # independent check of the carried spectrum (tests); off on the hot path
        self.tracked_B = z(b, n, n) if (self.compact and track_B) else None
        self.evals, self.Vt = z(b, n), z(b, n, n)
        self.eig_ws = K.EighWorkspace(b, n, dev)
        self.H_initialized = False
        self.eig_valid = False
        # Davidson state
        self.V, self.AV, self.Vs, self.AVs, self.Yw = (z(b, kc, n) for _ in range(5))
        self.ksz, self.ninit, self.nhist, self.dav_state, self.nvec = (zi(b) for _ in range(5))
        self.lams = z(b, kc)
        self.rv, self.rvhat = z(b, 2, n), z(b, 2, n)
        self.theta, self.that, self.t, self.vnew, self.signnorm = z(b), z(b, n), z(b, n), z(b, n), z(b)
        # update work space: one-pair (step) and kcap-pair (post-diagonalisation)
        self.up1 = {k: z(b, 1, n) for k in ("Ytil", "BS", "VtS", "aC", "aBS", "U", "J", "W", "Xw")}
        self.upk = {k: z(b, kc, n) for k in ("Ytil", "BS", "VtS", "aC", "aBS", "U", "J", "W", "Xw")}
        self.lam0, self.skip = torch.ones(b, **f64), zi(b)
        self.c2, self.s2 = z(b, 2, n), z(b, 2, n)
        if self.eig_mode == "update":
            self.sec1 = dict(P=z(b, 2, n), Z=z(b, 2, n), sig=z(b, 2))
            self.seck = dict(P=z(b, 2 * kc, n), Z=z(b, 2 * kc, n), sig=z(b, 2 * kc)) if kc <= 16 else None
            self.Cmat = z(b, 32 * 33)
            self.nterm = zi(b)
            self.qwork = z(b, n, n)
        if self.compact:
            # sp: spectrum of the step model (B projected onto the free coordinates); spB: spectrum of B itself
            # (|B| S and B S of the secant update, the "lowest mode went positive" test) -- one object when
            # nothing is fixed
            self.sp = _Spec(b, n, dev, self.evals, self.Vt)
            self.spB = self.sp if self.fmask is None else _Spec(b, n, dev)
            self.gm = self.g if self.fmask is None else z(b, n)               # P_f g
            # rotation of the eigenvector rows per rank-one term: a batched DMMA GEMM over the whole GPU
            # (split mode of sb_secular_update_c) once there are enough rows for 64 x 64 tiles to pay
            import os
            self.split_rotation = os.environ.get("SB_SPLIT_ROTATION", "1") != "0"
            self.split_min_rows = int(os.environ.get("SB_SPLIT_MIN_ROWS", "24"))
            self.sec_aux = zi(b, n + 4)
            self.lowk = z(b, max(1, self.order))
            self.cev, self.cvg, self.ccoef = z(b * n), z(b * n), z(b * n)      # pole lists, stride = width
            self.rowmap = zi(b * n)
            self.gperp, self.Wg, self.gam, self.kappa = z(b, n), z(b, n), z(b), z(b)
            self.C4, self.T4 = z(b, 4, n), z(b, 4, n)
            self.jd_ed = z(b, 2)
            self.ncand = zi(b)
            for sec, zc in ((self.sec1, 2), (self.seck, 2 * kc)):
                for k in ("W1", "Qc", "D2", "W2"):
                    sec[k] = z(b, zc, n)
            if self.fmask is not None:
                for bufs, kk in ((self.up1, 1), (self.upk, kc)):
                    bufs["Um"], bufs["Jm"] = z(b, kk, n), z(b, kk, n)
                if self.rs == "tr":
                    # optimize.py:183-187: the spherical radius scales with the number of free coordinates
                    self.delta.fill_(float(delta0 * self.nfree))
        self._setup_constraints(constraints)
        if self.cons is not None and self.rs == "tr":
            # optimize.py:183-187: the spherical radius scales with the number of free coordinates
            self.delta.fill_(float(delta0 * self.cons["nfree"]))
        self.initialized = False
        self._evaluated = False
        self.ndiag = 0
        self.prof = None          # set to {} to collect CUDA-event timings of selected kernels

    # ------------------------------------------------------------------ constraints
    def _setup_constraints(self, constraints):
        """Linear constraints C x = c (what Constraints.fix_translation yields).
        `constraints` = (C, c) with C [nc, n] shared by the batch or [b, nc, n], c [b, nc] or
        None (= the constraints hold at x0).  Bases follow peswrapper.py:51-69."""
        self.cons = None
        self.evalsB, self.VtB = self.evals, self.Vt            # spectrum of B itself
        if constraints is None:
            return
        from scipy.linalg import qr
        b, n, dev = self.batch, self.n, self.dev
        if len(constraints) == 4:
            return self._setup_nonlinear_constraints(*constraints)
        C, c = constraints
        C = np.asarray(C, dtype=np.float64)
        shared = C.ndim == 2
        Cs = C[None] if shared else C
        nc = Cs.shape[1]
        if nc == 0:
            return
        if self.eig_mode != "update":
            raise NotImplementedError("constraints need eig_mode='update'")
        Uc_l, M_l, Q_l, ranks = [], [], [], []
        for Ci in Cs:
            Q, R, _ = qr(Ci.T, mode="full", pivoting=True, check_finite=False)
            dg = np.abs(np.diag(R))
            rk = int(np.sum(dg > 1e-6 * dg[0])) if (dg.size and dg[0] > 0) else 0
            Ucons, Ufree = Q[:, :rk], Q[:, rk:]
            M = Ucons @ np.linalg.pinv(Ci @ Ucons)               # scons = -M res  (peswrapper.py:429-438)
            Uc_l.append(Ucons.T.copy()); M_l.append(M.T.copy())
            Q_l.append(np.vstack([Ufree.T, Ucons.T])); ranks.append(rk)
        if len(set(ranks)) != 1:
            raise NotImplementedError("all systems of a batch must have the same constraint rank")
        rk = ranks[0]
        up = lambda lst: torch.from_numpy(np.ascontiguousarray(np.stack(lst))).to(dev)   # noqa: E731
        f64 = dict(dtype=torch.float64, device=dev)
        cons = dict(shared=shared, nc=nc, rank=rk, nfree=n - rk,
                    C=up(list(Cs)), Uc=up(Uc_l), Mr=up(M_l), Q=up(Q_l),
                    cstride=0 if shared else nc * n, ustride=0 if shared else rk * n)
        x0 = self.x
        if c is None:
            ctar = torch.einsum("bjn,bn->bj", cons["C"].expand(b, nc, n), x0) if shared else \
                torch.einsum("bjn,bn->bj", cons["C"], x0)
        else:
            ctar = torch.from_numpy(np.ascontiguousarray(np.asarray(c, dtype=np.float64))).to(dev).reshape(b, nc)
        cons["c"] = ctar.contiguous()
        cons["res"] = torch.zeros(b, nc, **f64)
        cons["uw"] = torch.zeros(b, max(rk, 1), **f64)
        for k in ("scons", "gp", "pg", "slift"):
            cons[k] = torch.zeros(b, n, **f64)
        for k in ("scons2", "consval"):
            cons[k] = torch.zeros(b, **f64)
        cons["naive"] = torch.zeros(b, dtype=torch.int32, device=dev)
        cons["regular"] = torch.ones(b, dtype=torch.int32, device=dev)
        cons["cmax"] = torch.zeros(b, **f64)
        self.cons = cons
        # the step model lives on the projected Hessian Bp = P_f B P_f + sigma P_c; B keeps its own spectrum
        self.evalsB = torch.zeros(b, n, **f64)
        self.VtB = torch.zeros(b, n, n, **f64)

    def _setup_nonlinear_constraints(self, C_lin, c_lin, internals, targets):
        """Position-dependent constraints (peswrapper.py:395-407, 429-438, 467-481): the internal
        coordinates of `internals` (a BatchedInternals: bonds, angles, dihedrals, single-atom
        translations) held at `targets` [b, nnl] (None: their values at x0), optionally together with
        linear rows C_lin x = c_lin.  The constraint basis, the multipliers, the constraint Hessian
        Hc = sum_i L_i d2q_i/dx2 and the spectrum of the projected Lagrangian Hessian are rebuilt
        at every geometry (`_refresh_constraints`), as the reference does."""
        b, n, dev = self.batch, self.n, self.dev
        f64 = dict(dtype=torch.float64, device=dev)
        nlin = 0 if C_lin is None else int(np.asarray(C_lin).shape[-2])
        nnl = internals.nint
        nc = nlin + nnl
        if nc == 0:
            return
        if nc > 32:
            raise NotImplementedError("at most 32 constraints when some are position dependent")
        cons = dict(shared=False, nc=nc, rank=nc, nfree=n - nc, cstride=nc * n, ustride=nc * n,
                    C=torch.zeros(b, nc, n, **f64), Uc=torch.zeros(b, nc, n, **f64), Mr=torch.zeros(b, nc, n, **f64),
                    c=torch.zeros(b, nc, **f64), res=torch.zeros(b, nc, **f64), uw=torch.zeros(b, nc, **f64))
        if nlin:
            Cl = np.asarray(C_lin, dtype=np.float64)
            Cl = np.array(np.broadcast_to(Cl if Cl.ndim == 3 else Cl[None], (b, nlin, n)))
            cons["C"][:, :nlin] = torch.from_numpy(Cl).to(dev)
            if c_lin is None:
                cons["c"][:, :nlin] = torch.einsum("bjn,bn->bj", cons["C"][:, :nlin], self.x)
            else:
                cons["c"][:, :nlin] = torch.from_numpy(np.ascontiguousarray(np.asarray(c_lin, dtype=np.float64))).to(dev).reshape(b, nlin)
        q0 = internals.calc(self.x)
        if targets is None:
            tgt = q0.clone()
        else:
            tgt = torch.from_numpy(np.ascontiguousarray(np.asarray(targets, dtype=np.float64))).to(dev).reshape(b, nnl).clone()
        for k in ("scons", "gp", "pg", "slift"):
            cons[k] = torch.zeros(b, n, **f64)
        for k in ("scons2", "consval", "cmax"):
            cons[k] = torch.zeros(b, **f64)
        cons["naive"] = torch.zeros(b, dtype=torch.int32, device=dev)
        cons["regular"] = torch.ones(b, dtype=torch.int32, device=dev)
        ndih0 = internals.nstd - internals.ndihedrals
        cons["nl"] = dict(ints=internals, target=tgt, nlin=nlin, dih0=ndih0, Lmul=torch.zeros(b, nc, **f64),
                          Hc=torch.zeros(b, n, n, **f64), HL=torch.zeros(b, n, n, **f64), u=torch.zeros(b, nc, **f64),
                          evalsHL=torch.zeros(b, n, **f64), hcv=torch.zeros(b, 1, n, **f64), hcv2=torch.zeros(b, 1, n, **f64),
                          HcVs=None, x_basis=None, x_model=None)
        self.cons = cons
        self.evalsB = torch.zeros(b, n, **f64)          # B keeps its own (updated) spectrum; (evals, Vt) hold the
        self.VtB = torch.zeros(b, n, n, **f64)          # projected Lagrangian Hessian, rebuilt per geometry
        if self.eig_mode != "update":
            raise NotImplementedError("constraints need eig_mode='update'")

    def _refresh_bases(self):
        """Constraint Jacobian, residual, Ucons, the scons map and the multipliers at self.x / self.g."""
        cn = self.cons
        nl = cn["nl"]
        b, n, nc, nlin = self.batch, self.n, cn["nc"], nl["nlin"]
        ints = nl["ints"]
        q, Bm = ints.calc(self.x, jacobian=True)
        cn["C"][:, nlin:] = Bm
        res = q - nl["target"]
        if ints.ndihedrals:                             # dihedrals live on a circle
            lo, hi = nl["dih0"], nl["dih0"] + ints.ndihedrals
            res[:, lo:hi] = torch.remainder(res[:, lo:hi] + np.pi, 2.0 * np.pi) - np.pi
        # c is chosen so that C x - c is the true residual: the linear-constraint kernels then apply as they are
        call("sb_rect_dots", _p(cn["C"]), LL(cn["cstride"]), I(nc), _p(self.x), LL(n), _p(None), _p(cn["res"]),
             I(n), _p(None), I(b), _stream())
        cn["c"][:, nlin:] = cn["res"][:, nlin:] - res
        cn["Uc"].copy_(cn["C"])
        nkept, st = K.mgs(cn["Uc"])                     # rows of drdx, orthonormalised (subspace of peswrapper.py:51-69)
        self.status |= st
        self.status |= ((nkept != nc).to(torch.int32) * 16)          # rank-deficient constraint Jacobian
        G = K.gemm(cn["C"], cn["Uc"], transB=True)      # drdx = G Uc
        call("sb_rect_dots", _p(cn["Uc"]), LL(cn["ustride"]), I(nc), _p(self.g), LL(n), _p(None), _p(nl["u"]),
             I(n), _p(None), I(b), _stream())
        call("sb_cons_solve", _p(G), _p(cn["Uc"]), _p(nl["u"]), I(nc), I(n), _p(cn["Mr"]), _p(nl["Lmul"]),
             _p(self.status), _p(None), I(b), _stream())
        nl["x_basis"] = self.x.clone()

    def _refresh_constraints(self):
        """Everything that depends on the geometry in the position-dependent case: bases, multipliers,
        Hc (peswrapper.py:343-352) and -- once the Hessian exists -- the spectrum of
        Bp = P_f (B - Hc) P_f + sigma P_c, which is the model of restricted_step.py:57-62."""
        cn = self.cons
        nl = cn["nl"]
        b, n, nlin = self.batch, self.n, nl["nlin"]
        self._refresh_bases()
        nl["Hc"] = nl["ints"].ldot(self.x, nl["Lmul"][:, nlin:].contiguous())
        if not self.H_initialized:
            # B is None in the reference: the projected Lagrangian Hessian is None as well (peswrapper.py:363-
            # 386 skips Hc then), i.e. the step model is the identity on the free space (stepper.py:76-80):
            # Bp = P_f + sigma P_c at the current geometry
            Pc = K.gemm(cn["Uc"], cn["Uc"], transA=True)
            call("sb_add_scaled_identity", _p(Pc), D(7.0), D(1.0), I(n), I(b), _stream())      # I + 7 Pc, in place
            K.eigh(Pc, evals=self.evals, Vt=self.Vt, ws=self.eig_ws, status=self.status)
            return
        torch.sub(self._B, nl["Hc"], out=nl["HL"])
        if self._projected_spectrum_by_update():
            return
        Pc = K.gemm(cn["Uc"], cn["Uc"], transA=True)                  # Ucons Ucons^T
        Pf = torch.eye(n, dtype=torch.float64, device=self.dev).expand(b, n, n) - Pc
        Pf = Pf.contiguous()
        sigma = 1.0 + 8.0 * torch.maximum(self.evalsB[:, 0].abs(), self.evalsB[:, -1].abs())
        Bp = K.gemm(Pf, K.gemm(nl["HL"], Pf))
        Bp = 0.5 * (Bp + Bp.transpose(1, 2)) + sigma[:, None, None] * Pc
        K.eigh(Bp.contiguous(), evals=self.evals, Vt=self.Vt, ws=self.eig_ws, status=self.status)
        nl["x_model"] = True

    def _projected_spectrum_by_update(self):
        """Spectrum of Bp = P_f (B - Hc) P_f + sigma P_c WITHOUT a full eigensolve, for the reference's
        default projection of molecules (linear rows + the three rotation coordinates).  With
        A = (B - Hc) Ucons and G = Ucons^T A,
            Bp - B = -Hc - Ucons A^T - A Ucons^T + Ucons (G + sigma I) Ucons^T,
        and the rotation block of Hc is a sum of outer products of the per-coordinate 4-vectors that
        sb_rotation leaves in its work buffer: Hc = 1/2 sum_c (x_c y_c^T + y_c x_c^T) with
        (x, y) = (p_k - dFw_k, dq_k), k = 1..4, and (2 dE, dq.w).  That is nc + 5 secant-like pairs
        (U_i, J_i) with C = -(G + sigma I): numerical rank 2 nc, one eigen-update of the CARRIED
        spectrum of B.  Returns False when the situation is not this one (the caller then builds Bp
        densely and calls the eigensolver)."""
        import os
        cn = self.cons
        nl = cn["nl"]
        ints = nl["ints"]
        nc = cn["nc"]
        if os.environ.get("SB_NL_SECULAR", "1") == "0" or ints.nstd != 0 or ints.nrotations != 3 or nc + 5 > 16 \
                or self.eig_mode != "update" or not self.eig_valid:
            return False
        b, n = self.batch, self.n
        f64 = dict(dtype=torch.float64, device=self.dev)
        if "U16" not in nl:
            nl["U16"], nl["J16"] = torch.zeros(b, 16, n, **f64), torch.zeros(b, 16, n, **f64)
            nl["A"], nl["A2"] = torch.zeros(b, nc, n, **f64), torch.zeros(b, nc, n, **f64)
            nl["sec"] = dict(P=torch.zeros(b, 32, n, **f64), Z=torch.zeros(b, 32, n, **f64), sig=torch.zeros(b, 32, **f64))
            nl["C16"] = torch.zeros(b, 32, 33, **f64)
            nl["k16"] = torch.full((b,), nc + 5, dtype=torch.int32, device=self.dev)
            nl["skip0"] = torch.zeros(b, dtype=torch.int32, device=self.dev)
            nl["nterm"] = torch.zeros(b, dtype=torch.int32, device=self.dev)
        U, J, A, A2, sec = nl["U16"], nl["J16"], nl["A"], nl["A2"], nl["sec"]
        K.hv_ld(self._B, cn["Uc"], A, nc)
        K.hv_ld(nl["Hc"], cn["Uc"], A2, nc)
        A.sub_(A2)                                                     # rows a_i = (B - Hc) u_i
        G = K.gemm(cn["Uc"], A, transB=True)                           # G_ij = u_i . a_j
        sigma = 1.0 + 8.0 * torch.maximum(self.evalsB[:, 0].abs(), self.evalsB[:, -1].abs())
        U[:, :nc] = cn["Uc"]
        J[:, :nc] = -A
        w = ints._rot_work.view(b, 14 * n)
        dc = w[:, :4 * n].view(b, n, 4)
        pv = w[:, 4 * n:8 * n].view(b, n, 4)
        dfw = w[:, 8 * n:12 * n].view(b, n, 4)
        U[:, nc:nc + 4] = (pv - dfw).transpose(1, 2)
        J[:, nc:nc + 4] = -0.5 * dc.transpose(1, 2)
        U[:, nc + 4] = 2.0 * w[:, 12 * n:13 * n]
        J[:, nc + 4] = -0.5 * w[:, 13 * n:14 * n]
        C = nl["C16"]
        C.zero_()
        C[:, :nc, :nc] = -(0.5 * (G + G.transpose(1, 2)) + sigma[:, None, None] * torch.eye(nc, **f64))
        call("sb_lowrank_factor", _p(U), _p(J), _p(C), I(16), _p(nl["k16"]), I(n), _p(sec["P"]), _p(sec["sig"]),
             _p(nl["nterm"]), _p(nl["skip0"]), I(b), _stream())
        self.evals.copy_(self.evalsB)
        self.Vt.copy_(self.VtB)
        K.hv_ld(self.Vt, sec["P"], sec["Z"], 2 * (nc + 5))
        call("sb_secular_update", _p(self.evals), _p(self.Vt), _p(sec["Z"]), I(32), _p(sec["sig"]), _p(nl["nterm"]),
             I(n), _p(self.eig_ws.work), _p(self.qwork), _p(self.status), _p(nl["skip0"]), I(b), _stream())
        return True

    def _identity_model(self):
        b, n = self.batch, self.n
        if self.compact:                   # lam0 = 1 and no explicit pairs IS the identity
            self.eig_valid = True
            return
        one = torch.ones(b, dtype=torch.float64, device=self.dev)
        call("sb_fill_scaled_identity", _p(self._B), _p(self.evalsB), _p(self.VtB), _p(one), I(n), I(n), _p(None),
             I(b), _stream())
        if self.cons is not None:
            cn = self.cons
            if "nl" in cn:
                self.eig_valid = True          # the model spectrum is rebuilt per geometry (_refresh_constraints)
                return
            self.Vt.copy_(cn["Q"].expand(b, n, n) if cn["shared"] else cn["Q"])
            self.evals.fill_(1.0)
            self.evals[:, cn["nfree"]:] = 8.0
        self.eig_valid = True

    def _project_free(self, X, nvec, ld, active=None):
        """X[b, :nvec, :] <- P_f X  (remove the components along Ucons), in place."""
        cn = self.cons
        b, n, rk = self.batch, self.n, cn["rank"]
        if rk == 0:
            return
        for v in range(nvec):
            xv = X[:, v]                                  # view [b, n], stride ld*n
            call("sb_rect_dots", _p(cn["Uc"]), LL(cn["ustride"]), I(rk), _p(xv), LL(ld * n), _p(None),
                 _p(cn["uw"]), I(n), _p(active), I(b), _stream())
            call("sb_rect_comb", _p(cn["Uc"]), LL(cn["ustride"]), I(rk), _p(cn["uw"]), D(-1.0), _p(xv), LL(ld * n),
                 D(1.0), _p(xv), LL(ld * n), I(n), _p(active), I(b), _stream())

    def _timed(self, name, fn):
        if self.prof is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.prof.setdefault(name, []).append((e0, e1))
        return out

    def prof_summary(self):
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(z) for a, z in v) / len(v)) for k, v in (self.prof or {}).items()}

    # ------------------------------------------------------------------ compact representation
    @staticmethod
    def _single_coordinate_constraints(constraints, x0):
        """Indices of the fixed Cartesian coordinates if `constraints` = (C, c) holds single coordinates at
        their current values (rows of the identity, one matrix for the whole batch), else None."""
        if constraints is None or len(constraints) != 2:
            return None
        C, c = constraints
        C = np.asarray(C, dtype=np.float64)
        if C.ndim != 2 or C.shape[0] == 0:
            return None
        nz = C != 0.0
        if not ((nz.sum(axis=1) == 1).all() and (C[nz] == 1.0).all() and (nz.sum(axis=0) <= 1).all()):
            return None
        idx = np.argmax(nz, axis=1)
        if c is not None:
            cur = x0[:, torch.from_numpy(idx).to(x0.device)].cpu().numpy()
            if not np.array_equal(np.broadcast_to(np.asarray(c, dtype=np.float64), cur.shape), cur):
                return None             # a target away from the current value: the scons machinery (dense path)
        return idx

    @property
    def mrows(self):
        return self.spB.mrows

    @property
    def _rb(self):
        return self.spB.rb

    @property
    def B(self):
        """Dense approximate Hessian [b, n, n] (materialised from the compact representation on request)."""
        if not self.compact:
            return self._B
        sp = self.spB
        b, n = self.batch, self.n
        eye = torch.eye(n, dtype=torch.float64, device=self.dev)
        R = sp.rb
        out = self.lam0[:, None, None] * eye
        if R > 0 and self.H_initialized:
            live = (torch.arange(R, device=self.dev)[None, :] < sp.mrows[:, None]).to(torch.float64)
            d = (sp.evals[:, :R] - self.lam0[:, None]) * live
            VR = sp.Vt[:, :R].contiguous()
            out = out + K.gemm(VR, (d[:, :, None] * VR).contiguous(), transA=True)
        return 0.5 * (out + out.transpose(1, 2))

    def _hvr(self, sp, X, Y, nvec, transposed=False, active=None):
        """Pass over the explicit rows: Y[:, :, :R] = VR X (or Y = VR^T X[:, :, :R]); R = 0 leaves zeros."""
        if sp.rb > 0:
            K.hv_rect(sp.Vt, sp.rb, X, Y, nvec, transposed=transposed, active=active)
        else:
            Y[:, :nvec].zero_()

    def _mask(self, X, out, nvec):
        """out[:, :nvec] = P_f X[:, :nvec] for fixed Cartesian coordinates (a 0/1 mask)."""
        call("sb_mask_vec", _p(X), _p(self.fmask), _p(out), I(X.shape[1]), I(nvec), I(self.n), I(self.batch), _stream())

    def _spectral_apply(self, sp, S, bufs, nv, active):
        """bufs['BS'] = B S and bufs['aBS'] = |B| S from a compact spectrum (linalg.py:293 -> 174-195 and
        hessian_update.py:118-125 need eigh(B) for this in the reference)."""
        b, n, kc, R = self.batch, self.n, S.shape[1], sp.rb
        es = LL(n)
        self._hvr(sp, S, bufs["VtS"], nv, active=active)
        for mode, key in ((0, "aBS"), (1, "BS")):
            if R > 0:
                call("sb_compact_scale", _p(bufs["VtS"]), _p(sp.evals), es, _p(sp.mrows), _p(self.lam0), I(kc), I(nv),
                     I(n), I(R), I(mode), _p(bufs["aC"]), _p(self.skip), I(b), _stream())
            self._hvr(sp, bufs["aC"], bufs["Xw"], nv, transposed=True, active=active)
            call("sb_compact_axpy", _p(S), _p(bufs["Xw"]), _p(self.lam0), I(kc), I(nv), I(n), I(mode), _p(bufs[key]),
                 _p(self.skip), I(b), _stream())

    def _eigen_update(self, sp, U, J, kc, kvec, nv, active):
        """(lam0, theta, VR) of M  ->  of M + U J^T + J U^T - U sym(C) U^T (C = self.Cmat): the part of the
        update outside span(VR) joins VR as new rows with eigenvalue lam0, then ONE secular-equation
        update of the explicit pairs."""
        b, n = self.batch, self.n
        sec = self.sec1 if kc == 1 else self.seck
        zc, T = 2 * kc, 2 * nv
        es, vs = LL(n), LL(n * n)
        call("sb_lowrank_factor", _p(U), _p(J), _p(self.Cmat), I(kc), _p(kvec), I(n),
             _p(sec["P"]), _p(sec["sig"]), _p(self.nterm), _p(self.skip), I(b), _stream())
        R = sp.rb
        self._hvr(sp, sec["P"], sec["Z"], T, active=active)
        if sp.mmin < n:
            self._hvr(sp, sec["Z"], sec["W1"], T, transposed=True, active=active)
            call("sb_compact_append_a", _p(sec["P"]), _p(sec["W1"]), I(zc), _p(self.nterm), _p(sp.mrows), I(n),
                 _p(sec["Qc"]), _p(self.ncand), _p(self.skip), I(b), _stream())
            self._hvr(sp, sec["Qc"], sec["D2"], T, active=active)
            self._hvr(sp, sec["D2"], sec["W2"], T, transposed=True, active=active)
            call("sb_compact_append_b", _p(sec["P"]), _p(sec["Qc"]), _p(sec["W2"]), I(zc), _p(self.nterm),
                 _p(self.ncand), I(n), _p(sp.evals), es, _p(sp.Vt), vs, _p(sp.mrows), _p(self.lam0),
                 _p(sec["Z"]), _p(self.skip), I(b), _stream())
            sp.rb = min(n, R + T)
        split = self.split_rotation and sp.rb >= self.split_min_rows
        call("sb_secular_update_c", _p(sp.evals), _p(sp.Vt), _p(sec["Z"]), I(zc), _p(sec["sig"]),
             _p(self.nterm), I(n), _p(self.eig_ws.work), _p(self.qwork), _p(self.status), _p(self.skip),
             _p(sp.mrows), I(sp.rb), es, vs, I(T if split else 0), _p(self.sec_aux if split else None), I(b),
             _stream())

    def _update_compact(self, S, Y, bufs, kvec, nv, active, bs_ready=False, abs_ready=False):
        """ApproximateHessian.update (linalg.py:274-304) on the compact representation."""
        b, n = self.batch, self.n
        kc = S.shape[1]
        first = not self.H_initialized
        call("sb_update_prep", _p(S), _p(Y), _p(bufs["Ytil"]), I(kc), _p(kvec), I(n), I(n), I(int(first)),
             I(2), _p(self.lam0), _p(self.skip), _p(self.status), _p(active), I(b), _stream())
        if first:
            # B = lam0 I (hessian_update.py:58-67): no explicit pairs yet; systems whose first update is a
            # no-op keep the identity model (lam0 = 1)
            self.sp.reset()
            self.spB.reset()
            self.H_initialized = True
            bs_ready = abs_ready = False
        if not (bs_ready and (abs_ready or self.update_method != 0)):
            self._spectral_apply(self.spB, S, bufs, nv, active)
        call("sb_update_mid", _p(S), _p(bufs["Ytil"]), _p(bufs["BS"]),
             _p(bufs["aBS"] if self.update_method == 0 else None), _p(bufs["U"]), _p(bufs["J"]), _p(bufs["W"]),
             _p(bufs["Xw"]), I(kc), _p(kvec), I(n), I(self.update_method), _p(self.skip), _p(self.status),
             _p(self.Cmat), _p(None), _p(None), I(b), _stream())
        if self.tracked_B is not None:
            if first:
                call("sb_fill_scaled_identity", _p(self.tracked_B), _p(None), _p(None), _p(self.lam0), I(n), I(n),
                     _p(self.skip), I(b), _stream())
            call("sb_update_apply", _p(self.tracked_B), _p(bufs["U"]), _p(bufs["J"]), _p(bufs["W"]), I(kc), _p(kvec),
                 I(n), _p(self.skip), I(b), _stream())

        def run():
            self._eigen_update(self.spB, bufs["U"], bufs["J"], kc, kvec, nv, active)
            if self.fmask is not None:
                # the same update seen by the projected Hessian: (B + Delta)_ff = B_ff + Delta_ff
                self._mask(bufs["U"], bufs["Um"], nv)
                self._mask(bufs["J"], bufs["Jm"], nv)
                self._eigen_update(self.sp, bufs["Um"], bufs["Jm"], kc, kvec, nv, active)
        self._timed("eigen_update_k%d" % kc, run)
        self._updates_since_refresh += 1
        self.eig_valid = True

    def _refresh_poles(self, active=None):
        """Vg = VR (P_f g), g_perp and the merged pole list of the step model at the current gradient."""
        b, n, sp = self.batch, self.n, self.sp
        width = min(n, sp.rb + 1)
        if self.fmask is not None:
            self._mask(self.g.view(b, 1, n), self.gm.view(b, 1, n), 1)
        self._hvr(sp, self.gm.view(b, 1, n), self.Vg.view(b, 1, n), 1, active=active)
        self._hvr(sp, self.Vg.view(b, 1, n), self.Wg.view(b, 1, n), 1, transposed=True, active=active)
        call("sb_compact_prepare", _p(self.gm), _p(self.Vg), _p(self.Wg), _p(sp.evals), LL(n), _p(sp.mrows),
             _p(self.lam0), I(n), I(width), _p(self.gperp), _p(self.gam), _p(self.cev), _p(self.cvg), _p(self.rowmap),
             _p(active), I(b), _stream())
        return width

    def _predict_compact(self, active):
        """Restricted step from the compact spectral model; returns True when B s and |B| s of the FULL
        Hessian are already in up1 (no fixed coordinates: the model is B itself)."""
        b, n, sp = self.batch, self.n, self.sp
        width = self._refresh_poles(active)
        if self.rs == "tr":
            if self.method == "qn":
                call("sb_qn_tr", _p(self.cvg), _p(self.cev), _p(self.delta), I(self.order), I(width), _p(self.ccoef),
                     _p(self.smag), _p(self.alpha), _p(self.status), _p(active), _p(None), I(b), _stream())
            else:
                call("sb_rfo_tr", _p(self.cvg), _p(self.cev), _p(self.delta), I(self.order), I(width),
                     I(1 if self.method == "prfo" else 0), _p(self.ccoef), _p(self.smag), _p(self.alpha),
                     _p(self.status), _p(active), _p(None), I(b), _stream())
            call("sb_compact_finish", _p(self.ccoef), _p(self.rowmap), I(width), _p(sp.evals), LL(n), _p(self.gam),
                 I(n), I(sp.rb), _p(self.C4), _p(self.kappa), _p(active), I(b), _stream())
            self._hvr(sp, self.C4, self.T4, 3, transposed=True, active=active)
            # s, |B| s, B s and x + s from the one transposed pass
            call("sb_compact_finish2", _p(self.T4), _p(self.gperp), _p(self.kappa), _p(self.lam0), _p(self.x), I(n),
                 _p(self.s), _p(self.up1["aBS"]), _p(self.up1["BS"]), _p(self.xnew), _p(active), I(b), _stream())
            return self.fmask is None
        if self.method == "qn":
            call("sb_qn_ras_c", _p(self.cvg), _p(self.cev), _p(sp.Vt), _p(self.delta), I(self.order), I(n),
                 _p(self.s), _p(self.smag), _p(self.alpha), _p(self.status), _p(active), _p(None), I(width),
                 _p(self.rowmap), _p(self.gperp), _p(self.gam), LL(n * n), I(b), _stream())
        else:
            call("sb_rfo_ras_c", _p(self.cvg), _p(self.cev), _p(sp.Vt), _p(self.delta), I(self.order), I(n),
                 I(1 if self.method == "prfo" else 0), _p(self.s), _p(self.smag), _p(self.alpha), _p(self.status),
                 _p(active), _p(None), I(width), _p(self.rowmap), _p(self.gperp), _p(self.gam), LL(n * n), I(b),
                 _stream())
        call("sb_axpy", _p(self.x), _p(self.s), _p(self.xnew), I(n), _p(active), I(b), _stream())
        self.skip.zero_()
        self._spectral_apply(self.spB, self.s.view(b, 1, n), self.up1, 1, active)
        return True

    def _sync_rows(self):
        """ONE small device->host read: the exact bounds on the explicit-row counts (and whether any system
        wants a re-diagonalisation)."""
        stats = [self.ev.max(), self.spB.mrows.max(), self.spB.mrows.min()]
        if self.sp is not self.spB:
            stats += [self.sp.mrows.max(), self.sp.mrows.min()]
        vals = torch.stack(stats).tolist()
        nev, self.spB.rb, self.spB.mmin = vals[:3]
        if self.sp is not self.spB:
            self.sp.rb, self.sp.mmin = vals[3:]
        return nev

    def _step_compact(self, active):
        b, n = self.batch, self.n
        ready = self._predict_compact(active)
        # optimize.py:369-371 looks at the lowest `order` eigenvalues of the Hessian itself (Unred = I)
        call("sb_compact_lowest", _p(self.spB.evals), LL(n), _p(self.spB.mrows), _p(self.lam0), I(n),
             I(max(1, self.order)), _p(self.lowk), I(b), _stream())
        call("sb_ev_decide", _p(self.lowk), I(max(1, self.order)), I(1), _p(self.since_diag), _p(self.ev),
             self._dpar, self._ipar, _p(active), I(b), _stream())
        self.surface.evaluate(self.xnew, self.fnew, self.gnew, active=active)
        # rho needs s.(B s): s lives in the free space, where the model's B s is the Hessian's
        call("sb_kick_finish", _p(self.x), _p(self.f), _p(self.g), _p(self.xnew), _p(self.fnew), _p(self.gnew),
             _p(self.s), _p(self.up1["BS"]), _p(self.smag), _p(self.dg), _p(self.delta), _p(self.rho),
             _p(self.nsteps), self._dpar, self._ipar, I(n), _p(active), I(b), _stream())
        self._update_compact(self.s.view(b, 1, n), self.dg.view(b, 1, n), self.up1, None, 1, active,
                             bs_ready=ready, abs_ready=ready)
        # ONE small read per step: does any system re-diagonalise, and how many explicit rows do the
        # passes have to visit (the bound kept on the host grows by the number of TERMS per update, the
        # true count by the number of NEW directions, usually half of that)
        nev = self._sync_rows()
        if nev > 0:
            if self.hessian_function is not None:
                self._calculate_hessian(self.ev)
            else:
                self._diag(self.ev)

    # ------------------------------------------------------------------ helpers
    def _eigh(self, active=None):
        K.eigh(self._B, active=active, evals=self.evals, Vt=self.Vt, ws=self.eig_ws, status=self.status)

    def _update(self, S, Y, bufs, kvec, nv, active, bs_ready=False, abs_ready=False):
        """ApproximateHessian.update (linalg.py:274-304) for S, Y of shape [b,kc,n]."""
        if self.compact:
            return self._update_compact(S, Y, bufs, kvec, nv, active, bs_ready, abs_ready)
        b, n = self.batch, self.n
        kc = S.shape[1]
        first = not self.H_initialized
        call("sb_update_prep", _p(S), _p(Y), _p(bufs["Ytil"]), I(kc), _p(kvec), I(n), I(n), I(int(first)),
             I(2), _p(self.lam0), _p(self.skip), _p(self.status), _p(active), I(b), _stream())
        if first:
            call("sb_fill_scaled_identity", _p(self._B), _p(self.evalsB), _p(self.VtB), _p(self.lam0), I(n),
                 I(n), _p(self.skip), I(b), _stream())
            if self.cons is not None and "nl" not in self.cons:
                # Bp = lam0 P_f + sigma P_c: eigenvectors = [Ufree; Ucons] rows, sigma > lam0
                cn = self.cons
                self.Vt.copy_(cn["Q"].expand(b, n, n) if cn["shared"] else cn["Q"])
                self.evals[:, :cn["nfree"]] = self.lam0[:, None]
                self.evals[:, cn["nfree"]:] = 8.0 * torch.clamp(self.lam0, min=1e-3)[:, None]
            self.H_initialized = True
            bs_ready = False
        if not bs_ready:
            K.hv_ld(self._B, S, bufs["BS"], nv, active=active)
        if first:
            abs_ready = False
        if self.update_method == 0 and not abs_ready:
            K.hv_ld(self.VtB, S, bufs["VtS"], nv, active=active)
            call("sb_abs_scale", _p(bufs["VtS"]), _p(self.evalsB), _p(bufs["aC"]), I(kc), I(n), _p(self.skip),
                 I(b), _stream())
            K.hv_ld(self.VtB, bufs["aC"], bufs["aBS"], nv, transposed=True, active=active)
        wide = kc > 16                      # more terms than one eigen-update takes: full eigensolve afterwards
        track = self.eig_mode == "update" and (self.eig_valid or first) and not wide
        call("sb_update_mid", _p(S), _p(bufs["Ytil"]), _p(bufs["BS"]),
             _p(bufs["aBS"] if self.update_method == 0 else None), _p(bufs["U"]), _p(bufs["J"]), _p(bufs["W"]),
             _p(bufs["Xw"]), I(kc), _p(kvec), I(n), I(self.update_method), _p(self.skip), _p(self.status),
             _p(self.Cmat if track else None), _p(None), _p(None), I(b), _stream())
        self._timed("update_apply_k%d" % kc, lambda: call(
            "sb_update_apply", _p(self._B), _p(bufs["U"]), _p(bufs["J"]), _p(bufs["W"]), I(kc), _p(kvec),
            I(n), _p(self.skip), I(b), _stream()))
        self._updates_since_refresh += 1
        if track and not (self.eig_refresh_every and self._updates_since_refresh >= self.eig_refresh_every):
            # B+ = B + Delta: carry the eigenpairs along instead of a fresh eigensolve
            sec = self.sec1 if kc == 1 else self.seck
            call("sb_lowrank_factor", _p(bufs["U"]), _p(bufs["J"]), _p(self.Cmat), I(kc), _p(kvec), I(n),
                 _p(sec["P"]), _p(sec["sig"]), _p(self.nterm), _p(self.skip), I(b), _stream())
            K.hv_ld(self.VtB, sec["P"], sec["Z"], 2 * nv, active=active)
            self._timed("secular_update_k%d" % kc, lambda: call(
                "sb_secular_update", _p(self.evalsB), _p(self.VtB), _p(sec["Z"]), I(2 * kc), _p(sec["sig"]),
                _p(self.nterm), I(n), _p(self.eig_ws.work), _p(self.qwork), _p(self.status), _p(self.skip),
                I(b), _stream()))
            if self.cons is not None and "nl" not in self.cons:
                # same update seen through the projector: Bp+ = Bp + P_f Delta P_f
                self._project_free(bufs["U"], nv, kc, active)
                self._project_free(bufs["J"], nv, kc, active)
                call("sb_lowrank_factor", _p(bufs["U"]), _p(bufs["J"]), _p(self.Cmat), I(kc), _p(kvec), I(n),
                     _p(sec["P"]), _p(sec["sig"]), _p(self.nterm), _p(self.skip), I(b), _stream())
                K.hv_ld(self.Vt, sec["P"], sec["Z"], 2 * nv, active=active)
                call("sb_secular_update", _p(self.evals), _p(self.Vt), _p(sec["Z"]), I(2 * kc), _p(sec["sig"]),
                     _p(self.nterm), I(n), _p(self.eig_ws.work), _p(self.qwork), _p(self.status), _p(self.skip),
                     I(b), _stream())
            self.eig_valid = True
        elif wide and self.eig_mode == "update":
            self._direct_spectra()
            self.eig_valid = True
        else:
            self.eig_valid = False
            self._updates_since_refresh = 0

    def _calculate_hessian(self, part=None):
        """PES.calculate_hessian (peswrapper.py:604-606): B <- hessian_function at the current
        geometry (for the systems with part[b] != 0), spectra by full eigensolves."""
        Bnew = self.hessian_function(self.x)
        check_f64(Bnew)
        Bnew = 0.5 * (Bnew + Bnew.transpose(1, 2))
        if self.compact:
            # a dense Hessian has no complement left: all n eigenpairs become explicit (rows of Vt)
            sp = self.sp
            K.eigh(Bnew.contiguous(), active=part, evals=sp.evals, Vt=sp.Vt, ws=self.eig_ws, status=self.status)
            if part is None:
                sp.mrows.fill_(self.n)
                sp.mmin = self.n
            else:
                sp.mrows.copy_(torch.where(part > 0, torch.full_like(sp.mrows, self.n), sp.mrows))
            sp.rb = self.n
            self.H_initialized = True
            self.eig_valid = True
            self._updates_since_refresh = 0
            self.ndiag += 1
            return
        if part is None:
            self._B.copy_(Bnew)
        else:
            m = part.to(torch.bool)
            self._B[m] = Bnew[m]
        self.H_initialized = True
        if self.eig_mode == "update":
            self._direct_spectra()
            self.eig_valid = True
        else:
            self.eig_valid = False
        self.ndiag += 1

    def _direct_spectra(self):
        """Full eigensolves of B (and, with linear constraints, of Bp = P_f B P_f + sigma P_c)."""
        b, n = self.batch, self.n
        K.eigh(self._B, evals=self.evalsB, Vt=self.VtB, ws=self.eig_ws, status=self.status)
        cn = self.cons
        if cn is not None and "nl" not in cn:
            Uc = cn["Uc"][0] if cn["shared"] else cn["Uc"]
            Pc = K.gemm(Uc, Uc, transA=True)                         # [1 or b, n, n]
            eye = torch.eye(n, dtype=torch.float64, device=self.dev)
            Pf = (eye - Pc[0]).contiguous() if cn["shared"] else (eye.expand(b, n, n) - Pc).contiguous()
            sigma = 1.0 + 8.0 * torch.maximum(self.evalsB[:, 0].abs(), self.evalsB[:, -1].abs())
            Bp = K.gemm(Pf, K.gemm(self._B, Pf))
            Bp = 0.5 * (Bp + Bp.transpose(1, 2)) + sigma[:, None, None] * Pc
            K.eigh(Bp.contiguous(), evals=self.evals, Vt=self.Vt, ws=self.eig_ws, status=self.status)
        self._updates_since_refresh = 0

    def _hvp(self, vec, vstride, mask, maskval, active):
        """One finite-difference Hessian-vector product per participating system."""
        b, n = self.batch, self.n
        call("sb_hvp_prepare", _p(vec), LL(vstride), _p(self.x), _p(self.g), D(self.eta), _p(self.xdisp),
             _p(self.signnorm), I(n), _p(mask), I(maskval), I(b), _stream())
        self.surface.evaluate(self.xdisp, self.fplus, self.gplus, active=active)
        gbase, eta_eff = self.g, self.eta
        if self.threepoint:
            # central difference (linalg.py:82-85): second evaluation at x0 - eta v/(sign |v|)
            call("sb_hvp_prepare", _p(vec), LL(vstride), _p(self.x), _p(self.g), D(-self.eta), _p(self.xminus),
                 _p(self.signnorm), I(n), _p(mask), I(maskval), I(b), _stream())
            self.surface.evaluate(self.xminus, self.fminus, self.gminus, active=active)
            gbase, eta_eff = self.gminus, 2.0 * self.eta
        if self.cons is not None:
            kslot = self.ksz.clone()                   # slot that hvp_finish is about to fill
        call("sb_hvp_finish", _p(vec), LL(vstride), _p(self.gplus), _p(gbase), _p(self.signnorm), D(eta_eff),
             _p(self.AV), _p(self.Vs), _p(self.AVs), I(self.kcap), _p(self.ksz), _p(self.nhist), I(n),
             _p(mask), I(maskval), I(b), _stream())
        if self.compact and self.fmask is not None:
            # Uproj^T (H v), linalg.py:92-93: the subspace image lives in the free space (a 0/1 mask here);
            # the operator history (Vs, AVs) keeps the unprojected product, as the reference does
            self._mask(self.AV, self.AV, self.kcap)
        if self.cons is not None:
            # Uproj^T (H v): the subspace image lives in the free space (linalg.py:92-93); the
            # operator history (Vs, AVs) keeps the unprojected vector, as the reference does
            kmax = int(kslot.max().item())
            nl = self.cons.get("nl")
            if nl is not None:
                # operator of peswrapper.py:537: Hproj - Ufree^T Hc Ufree
                nl["hcv"].copy_(vec.unsqueeze(1))
                K.hv_ld(nl["Hc"], nl["hcv"], nl["hcv2"], 1, active=active)
                if "one" not in nl:
                    nl["one"] = torch.ones(b, 1, dtype=torch.float64, device=self.dev)
            for k in range(kmax + 1):
                m = ((kslot == k) & (active > 0)).to(torch.int32) if active is not None else (kslot == k).to(torch.int32)
                if int(m.sum().item()):
                    if nl is not None:
                        call("sb_rect_comb", _p(nl["hcv2"]), LL(n), I(1), _p(nl["one"]), D(-1.0), _p(self.AV[:, k]),
                             LL(self.kcap * n), D(1.0), _p(self.AV[:, k]), LL(self.kcap * n), I(n), _p(m), I(b), _stream())
                    self._project_free(self.AV[:, k:k + 1], 1, self.kcap, m)

    def _diag(self, part=None):
        """PES.diag (peswrapper.py:508-556) for the systems with part[b] != 0."""
        b, n, kc = self.batch, self.n, self.kcap
        first = not self.H_initialized          # P = identity, v0 = g
        nl = self.cons.get("nl") if self.cons is not None else None
        if nl is not None:
            self._refresh_constraints()         # bases, Hc and the preconditioner at the current geometry
        if not first and not self.eig_valid and not self.compact:
            self._eigh(active=part)             # spectrum of the preconditioner P = B
        v0 = self.g
        if first and self.v0 is not None:
            if self.cons is not None:
                raise NotImplementedError("v0 lives in the reference's free-space basis; only without constraints")
            check_f64(self.v0)
            v0 = self.v0
        if self.cons is not None and first:
            v0 = self.cons["pg"]
            v0.copy_(self.g)
            self._project_free(v0.view(b, 1, n), 1, 1, part)            # Ufree^T g, lifted (peswrapper.py:524)
        use_v0 = first
        v0s = self._diag_start_vector(part)
        if v0s is not None:
            v0, use_v0 = v0s, True
        if self.compact:
            if not first:
                self._refresh_poles(part)          # g_perp of the CURRENT complement (fallback start vector)
            if first and self.fmask is not None:
                self._mask(self.g.view(b, 1, n), self.gm.view(b, 1, n), 1)
                v0 = self.gm                       # Ufree^T g (peswrapper.py:524)
            call("sb_davidson_init_c", _p(v0), _p(self.sp.evals), _p(self.sp.Vt), I(0 if first else 1), _p(self.V),
                 I(kc), I(n), _p(self.ksz), _p(self.ninit), _p(self.nhist), _p(self.dav_state), _p(self.status),
                 _p(part), _p(self.sp.mrows), _p(self.lam0), _p(self.gperp), LL(n), LL(n * n), I(b), _stream())
        else:
            call("sb_davidson_init", _p(v0), _p(self.evals), _p(self.Vt), I(0 if use_v0 else 1), _p(self.V),
                 I(kc), I(n), _p(self.ksz), _p(self.ninit), _p(self.nhist), _p(self.dav_state), _p(self.status),
                 _p(part), I(b), _stream())
        nstart = 1 if use_v0 else int(self.ninit.max().item())
        for j in range(nstart):
            m = ((self.dav_state == DAV_EXPAND) & (self.ninit > j)).to(torch.int32)
            self._hvp(self.V[:, j], kc * n, m, 1, m)
        # rayleigh_ritz stops at min(n, maxiter) vectors with n the dimension of the FREE space
        # (eigensolvers.py:31-66: A is the projected operator of peswrapper.py:531-537)
        nfree = self._free_dim()
        maxiter_eff = nfree if self.diag_maxiter is None else min(nfree, int(self.diag_maxiter))
        rounds = nstart                 # host-side bound on the operator products so far
        restarted = False               # after a thick restart ksz no longer counts the products (see below)
        kbound = nstart                 # host-side bound on the current subspace sizes ksz[b]
        hbound = nstart                 # ... and on the history lengths nhist[b]
        while True:
            call("sb_davidson_rr", _p(self.V), _p(self.AV), I(kc), _p(self.ksz), I(n), D(self.gamma),
                 I(maxiter_eff), _p(self.lams), _p(self.rv), _p(self.theta), _p(self.dav_state),
                 _p(self.status), I(b), _stream())
            m = (self.dav_state == DAV_EXPAND).to(torch.int32)
            # sb_davidson_rr stops each system at ksz[b] >= maxiter on its own (systems start with different
            # numbers of vectors, so the host count `rounds` is only an upper bound); once a thick restart has
            # clamped ksz the host count is what ends the run at the reference's total
            if int(m.sum().item()) == 0 or (restarted and rounds >= maxiter_eff):
                break
            if kbound >= kc:
                # The device subspace is full where the reference would go on (up to min(n, maxiter) vectors,
                # eigensolvers.py:65-66).  Thick restart: sb_davidson_rr has just rotated (V, AV) to Ritz
                # vectors in ascending order, so keeping the lowest `keep` of them loses nothing about the pairs
                # being sought; the operator products collected so far go into the Hessian as one block
                # update (they would all enter the single update at the end of PES.diag, peswrapper.py:541-551)
                keep = max(2, kc // 2)
                self._flush_history(part, min(hbound, kc), nl)
                self.nhist.zero_()
                self.ksz.copy_(torch.where(m > 0, torch.clamp(self.ksz, max=keep), self.ksz))
                kbound, hbound, first, restarted = keep, 0, False, True
            lanczos = int(self.eigensolver == 2)
            if first or lanczos:
                tin = None
            elif self.compact:
                # (P - theta)^-1 through the compact spectrum: explicit rows + the complement's 1/(lam0 - theta)
                self._hvr(self.sp, self.rv, self.rvhat, 2, active=m)
                call("sb_compact_jd_coeff", _p(self.rvhat), _p(self.rv), _p(self.sp.evals), LL(n), _p(self.sp.mrows),
                     _p(self.lam0), _p(self.theta), I(n), I(self.sp.rb), I(self.eigensolver), _p(self.that),
                     _p(self.jd_ed), _p(self.dav_state), I(b), _stream())
                self._hvr(self.sp, self.that.view(b, 1, n), self.t.view(b, 1, n), 1, transposed=True, active=m)
                call("sb_compact_jd_finish", _p(self.t), _p(self.rv), _p(self.jd_ed), I(n), I(self.eigensolver),
                     _p(self.dav_state), I(b), _stream())
                tin = self.t
            elif self.eigensolver == 3:
                if not hasattr(self, "Vhat"):
                    self.Vhat = torch.zeros_like(self.V)
                K.hv_ld(self.Vt, self.rv, self.rvhat, 1, active=m)
                K.hv_ld(self.Vt, self.V, self.Vhat, min(rounds, kc), active=m)
                call("sb_davidson_mjd_coeff", _p(self.Vhat), I(kc), _p(self.ksz), _p(self.rvhat), _p(self.evals),
                     _p(self.theta), _p(self.that), I(n), _p(self.dav_state), _p(self.status), I(b), _stream())
                K.hv_ld(self.Vt, self.that.view(b, 1, n), self.t.view(b, 1, n), 1, transposed=True, active=m)
                tin = self.t
            else:
                K.hv_ld(self.Vt, self.rv, self.rvhat, 2, active=m)
                call("sb_davidson_jd_coeff", _p(self.rvhat), _p(self.evals), _p(self.theta), _p(self.that),
                     I(n), I(self.eigensolver), _p(self.dav_state), I(b), _stream())
                K.hv_ld(self.Vt, self.that.view(b, 1, n), self.t.view(b, 1, n), 1, transposed=True, active=m)
                tin = self.t
            call("sb_davidson_expand", _p(tin), _p(self.rv), _p(self.theta), _p(self.V), _p(self.Yw), I(kc),
                 _p(self.ksz), I(n), I(int(first)), I(lanczos), _p(self.vnew), _p(self.dav_state),
                 _p(self.status), I(b), _stream())
            m = (self.dav_state == DAV_EXPAND).to(torch.int32)
            self._hvp(self.vnew, n, self.dav_state, DAV_EXPAND, m)
            rounds += 1
            kbound += 1
            hbound += 1
        if hbound > 0:
            self._flush_history(part, min(hbound, kc), nl)
        self.ndiag += 1

    def _free_dim(self):
        """Dimension of the space the Davidson operator acts in (Ufree.shape[1], peswrapper.py:513-514)."""
        if self.cons is not None:
            return self.cons["nfree"]
        return self.nfree if getattr(self, "fmask", None) is not None else self.n

    def _diag_start_vector(self, part):
        """Hook: a start vector that overrides the choice made from the model (internal coordinates: the first
        diagonalisation starts from Ufree^T g although a model Hessian exists, peswrapper.py:521-528)."""
        return None

    def _flush_history(self, part, nv, nl):
        """PES.diag tail (peswrapper.py:541-551): Ritz-rotate the operator history and feed it to the
        Hessian as one block update."""
        b, n, kc = self.batch, self.n, self.kcap
        hcvs = None
        if nl is not None:
            if nl["HcVs"] is None:
                nl["HcVs"] = torch.zeros_like(self.Vs)
            hcvs = nl["HcVs"]
            K.hv_ld(nl["Hc"], self.Vs, hcvs, nv, active=part)
        call("sb_history_ritz", _p(self.Vs), _p(self.AVs), I(kc), _p(self.nhist), I(n), _p(self.nvec),
             _p(self.dav_state), _p(self.status), _p(hcvs), I(b), _stream())
        self._update(self.Vs, self.AVs, self.upk, self.nvec, nv, part)

    # ------------------------------------------------------------------ public
    def step(self, active=None):
        """One Sella.step for every (active) system."""
        b, n = self.batch, self.n
        if not self.initialized:
            self.ensure_evaluated()
            if self.eig:
                if self.hessian_function is not None:
                    self._calculate_hessian()
                else:
                    self._diag(None)
                self.since_diag.fill_(-1)
            self.initialized = True
        # ---- _predict_step: restricted step from the spectral model
        if not self.H_initialized:
            # B is None in the reference: the step model is the identity (linalg.py:319-334,
            # stepper.py:76-80) until the first update scales it (hessian_update.py:58-67)
            self._identity_model()
        if self.compact:
            return self._step_compact(active)
        cn = self.cons
        nl = cn.get("nl") if cn is not None else None
        if nl is not None:
            self._refresh_constraints()
        elif not self.eig_valid:
            self._eigh(active)
            self.eig_valid = True
        gvec, extra2, sadd, act_rs = self.g, None, None, active
        if cn is not None:
            # scons = -Ucons lstsq(C Ucons, res);  g' = P_f (g + B scons)   (restricted_step.py:28-37)
            nc = cn["nc"]
            call("sb_rect_dots", _p(cn["C"]), LL(cn["cstride"]), I(nc), _p(self.x), LL(n), _p(cn["c"]),
                 _p(cn["res"]), I(n), _p(active), I(b), _stream())
            call("sb_rect_comb", _p(cn["Mr"]), LL(cn["cstride"]), I(nc), _p(cn["res"]), D(-1.0), _p(None), LL(0),
                 D(0.0), _p(cn["scons"]), LL(n), I(n), _p(active), I(b), _stream())
            call("sb_scons_measure", _p(cn["scons"]), _p(self.delta), I(0 if self.rs == "tr" else 1), I(n),
                 _p(cn["scons2"]), _p(cn["consval"]), _p(cn["naive"]), _p(cn["regular"]), I(b), _stream())
            K.hv_ld(self._B, cn["scons"].view(b, 1, n), cn["gp"].view(b, 1, n), 1, active=active)
            call("sb_axpy", _p(self.g), _p(cn["gp"]), _p(cn["gp"]), I(n), _p(active), I(b), _stream())
            self._project_free(cn["gp"].view(b, 1, n), 1, 1, active)
            gvec, extra2, sadd = cn["gp"], cn["scons2"], cn["scons"]
            act_rs = cn["regular"] if active is None else (cn["regular"] * active)
        K.hv_ld(self.Vt, gvec.view(b, 1, n), self.Vg.view(b, 1, n), 1, active=active)
        sdst = self.s if cn is None else cn["slift"]
        if self.rs == "tr":
            if self.method == "qn":
                call("sb_qn_tr", _p(self.Vg), _p(self.evals), _p(self.delta), I(self.order), I(n), _p(self.coef),
                     _p(self.smag), _p(self.alpha), _p(self.status), _p(act_rs), _p(extra2), I(b), _stream())
            else:
                call("sb_rfo_tr", _p(self.Vg), _p(self.evals), _p(self.delta), I(self.order), I(n),
                     I(1 if self.method == "prfo" else 0), _p(self.coef), _p(self.smag), _p(self.alpha),
                     _p(self.status), _p(act_rs), _p(extra2), I(b), _stream())
            if cn is None:
                # s = V c and |B| s = V(|lam| c) (needed by the TS-BFGS update) in one pass over Vt
                call("sb_pack_coef", _p(self.coef), _p(self.evals), _p(self.c2), I(n), _p(active), I(b), _stream())
                K.hv_ld(self.Vt, self.c2, self.s2, 2, transposed=True, active=active)
                call("sb_unpack2", _p(self.s2), _p(self.s), _p(self.up1["aBS"]), I(n), _p(active), I(b), _stream())
                abs_ready = True
            else:
                K.hv_ld(self.Vt, self.coef.view(b, 1, n), sdst.view(b, 1, n), 1, transposed=True, active=active)
                abs_ready = False
        elif self.method != "qn":
            abs_ready = False
            call("sb_rfo_ras", _p(self.Vg), _p(self.evals), _p(self.Vt), _p(self.delta), I(self.order), I(n),
                 I(1 if self.method == "prfo" else 0), _p(sdst), _p(self.smag), _p(self.alpha),
                 _p(self.status), _p(act_rs), _p(sadd), I(b), _stream())
        else:
            abs_ready = False
            call("sb_qn_ras", _p(self.Vg), _p(self.evals), _p(self.Vt), _p(self.delta), I(self.order), I(n),
                 _p(sdst), _p(self.smag), _p(self.alpha), _p(self.status), _p(act_rs), _p(sadd), I(b), _stream())
        if cn is not None:
            if self.rs != "tr":
                # the ras kernels return the total step (s_free + scons): take scons out again so that
                # combine_step handles both branches uniformly
                sdst.sub_(cn["scons"])
            call("sb_combine_step", _p(sdst), _p(cn["scons"]), _p(cn["consval"]), _p(self.delta), _p(cn["naive"]),
                 _p(self.s), _p(self.smag), I(n), _p(active), I(b), _stream())
        # ---- re-diagonalise?  (spectrum of the Hessian itself, before this step's update)
        ev_evals = self.evalsB
        if nl is not None and self.eig and int((self.since_diag >= int(self._ipar[2])).any().item()):
            # optimize.py:369-371 looks at the Hessian of the Lagrangian (B - Hc), not at B
            K.eigvalsh(nl["HL"], evals=nl["evalsHL"], status=self.status)
            ev_evals = nl["evalsHL"]
        call("sb_ev_decide", _p(ev_evals), I(n), I(1), _p(self.since_diag), _p(self.ev), self._dpar,
             self._ipar, _p(active), I(b), _stream())
        # ---- kick
        call("sb_axpy", _p(self.x), _p(self.s), _p(self.xnew), I(n), _p(active), I(b), _stream())
        self.surface.evaluate(self.xnew, self.fnew, self.gnew, active=active)
        S1 = self.s.view(b, 1, n)
        K.hv_ld(self._B, S1, self.up1["BS"], 1, active=active)
        call("sb_kick_finish", _p(self.x), _p(self.f), _p(self.g), _p(self.xnew), _p(self.fnew), _p(self.gnew),
             _p(self.s), _p(self.up1["BS"]), _p(self.smag), _p(self.dg), _p(self.delta), _p(self.rho),
             _p(self.nsteps), self._dpar, self._ipar, I(n), _p(active), I(b), _stream())
        self._update(S1, self.dg.view(b, 1, n), self.up1, None, 1, active, bs_ready=True, abs_ready=abs_ready)
        # kick(dx, diag=ev) diagonalises whenever ev is set, whatever `eig` says (optimize.py:362-380:
        # diag_every_n fires for eig=False runs too)
        if int(self.ev.sum().item()) > 0:
            if self.hessian_function is not None:
                self._calculate_hessian(self.ev)
            else:
                self._diag(self.ev)

    def kick(self, dx, diag=False):
        """PES.kick (peswrapper.py:578-602) with a caller-supplied displacement dx [b, n]: move, evaluate,
        rho = (f1 - f0) / (g0.dx + 1/2 dx.B dx), secant update with (dx, g1 - g0) and, if `diag`, a
        re-diagonalisation.  The trust radius is NOT touched (that is Sella.step's business, optimize.py:
        412-434).  Returns rho [b]."""
        b, n = self.batch, self.n
        check_f64(dx)
        self.ensure_evaluated()
        if not self.H_initialized:
            self._identity_model()
        self.s.copy_(dx)
        S1 = self.s.view(b, 1, n)
        call("sb_axpy", _p(self.x), _p(self.s), _p(self.xnew), I(n), _p(None), I(b), _stream())
        if self.compact:
            self.skip.zero_()
            self._spectral_apply(self.spB, S1, self.up1, 1, None)
        else:
            K.hv_ld(self._B, S1, self.up1["BS"], 1)
        self.surface.evaluate(self.xnew, self.fnew, self.gnew)
        keep = self.delta.clone()
        self.smag.copy_(self.s.norm(dim=1))
        call("sb_kick_finish", _p(self.x), _p(self.f), _p(self.g), _p(self.xnew), _p(self.fnew), _p(self.gnew),
             _p(self.s), _p(self.up1["BS"]), _p(self.smag), _p(self.dg), _p(self.delta), _p(self.rho),
             _p(self.nsteps), self._dpar, self._ipar, I(n), _p(None), I(b), _stream())
        self.delta.copy_(keep)
        self._update(S1, self.dg.view(b, 1, n), self.up1, None, 1, None, bs_ready=True, abs_ready=self.compact)
        if self.compact:
            self._sync_rows()
        if diag:
            if self.hessian_function is not None:
                self._calculate_hessian()
            else:
                self._diag(None)
        return self.rho.clone()

    def ensure_evaluated(self):
        """Energy and gradient at the current geometry, evaluated at most once per geometry
        (PES._update, peswrapper.py:440-465): converged()/log() before the first step and the
        first step itself share one surface call."""
        if not self._evaluated:
            self.surface.evaluate(self.x, self.f, self.g)
            self._evaluated = True

    def converged(self, fmax, cmax=1e-5):
        """PES.converged (peswrapper.py:558-568): max atomic |P_f g| < fmax and |res| < cmax."""
        b, n = self.batch, self.n
        cn = self.cons
        if cn is None:
            gsrc = self.g
            if self.compact and self.fmask is not None:
                self._mask(self.g.view(b, 1, n), self.gm.view(b, 1, n), 1)      # |P_f g| per atom
                gsrc = self.gm
            call("sb_converged", _p(gsrc), I(n), D(float(fmax)), _p(self.fmax), _p(self.conv), I(b), _stream())
            return self.conv
        if "nl" in cn and (cn["nl"]["x_basis"] is None or not torch.equal(cn["nl"]["x_basis"], self.x)):
            self._refresh_bases()
        cn["pg"].copy_(self.g)
        self._project_free(cn["pg"].view(b, 1, n), 1, 1)
        call("sb_rect_dots", _p(cn["C"]), LL(cn["cstride"]), I(cn["nc"]), _p(self.x), LL(n), _p(cn["c"]),
             _p(cn["res"]), I(n), _p(None), I(b), _stream())
        call("sb_converged_cons", _p(cn["pg"]), _p(cn["res"]), I(cn["nc"]), I(n), D(float(fmax)), D(float(cmax)),
             _p(self.fmax), _p(cn["cmax"]), _p(self.conv), I(b), _stream())
        return self.conv

    def lowest_evals(self):
        """[b] lowest eigenvalue of the approximate Hessian B of every system (refreshes a stale spectrum)."""
        if not self.H_initialized:
            return torch.ones(self.batch, dtype=torch.float64, device=self.dev)
        if self.compact:
            out = torch.empty(self.batch, dtype=torch.float64, device=self.dev)
            call("sb_compact_lowest", _p(self.spB.evals), LL(self.n), _p(self.spB.mrows), _p(self.lam0), I(self.n),
                 I(1), _p(out), I(self.batch), _stream())
            return out
        if not self.eig_valid:
            self._eigh(None)
            self.eig_valid = True
        return self.evalsB[:, 0].clone()

    def explicit_pairs(self, i):
        """(theta [m], VR [m, n], lam0, m) of system i as numpy arrays: the explicit eigenpairs of its
        approximate Hessian; every other eigenvalue equals lam0 (dense representation: m = n)."""
        if self.compact:
            sp = self.spB
            m = int(sp.mrows[i])
            th, VR = sp.evals[i, :m].cpu().numpy(), sp.Vt[i, :m].cpu().numpy()
            order = np.argsort(th, kind="stable")            # the explicit pairs are stored in any order
            return th[order], VR[order], float(self.lam0[i]), m
        if not self.eig_valid:
            self._eigh(None)
            self.eig_valid = True
        return self.evalsB[i].cpu().numpy(), self.VtB[i].cpu().numpy(), float(self.lam0[i]), self.n

    def rank_bound(self):
        """Upper bound (host-side count) on the number of distinct non-cluster eigenpairs of the model."""
        return self.spB.rb if self.compact else self.n

    def check_status(self):
        st = self.status.cpu().numpy()
        if st.any():
            bad = np.nonzero(st)[0]
            raise RuntimeError("sella_b200: %d system(s) reported errors, first: system %d status %d"
                               % (len(bad), bad[0], st[bad[0]]))
