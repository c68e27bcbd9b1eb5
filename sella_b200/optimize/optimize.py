"""`Sella(atoms, ...).run(fmax, steps)` — the reference's user-facing optimiser
(sella/optimize/optimize.py:42-502) on top of the CUDA engine, for ONE search.

The calculator stays where ASE puts it (on the host): every surface evaluation moves the
positions to the host, calls ``atoms.get_potential_energy()/get_forces()`` and moves
energy and gradient back; everything else (Davidson, Hessian update, restricted step)
runs on the device through ``sella_b200.batched.BatchedSella`` with a batch of one.

If ASE is importable the class derives from ``ase.optimize.optimize.Optimizer`` (so
``run``/``irun``/logging/trajectory attachment are ASE's); otherwise a minimal base
class with the same ``run(fmax, steps)`` loop is used.

Scope (raises NotImplementedError otherwise, never falls back to the CPU): Cartesian
coordinates; translation constraints (``sella_b200.Constraints.fix_translation``, incl.
the centre-of-geometry projection the reference adds by default, peswrapper.py:233-244) and
bond / angle / dihedral constraints (``fix_bond``, ``fix_angle``, ``fix_dihedral``);
the rotation projection of non-periodic systems (``fix_rotation`` of the whole configuration,
peswrapper.py:246-253); ``hessian_function`` and (without constraints) ``v0``; no cell optimisation.
"""
import warnings
from time import localtime, strftime

import numpy as np
import torch

from ..batched import BatchedSella
from .._host import dev

try:                                            # pragma: no cover - ASE is not in this image
    from ase.optimize.optimize import Optimizer as _Base
    _HAVE_ASE = True
except Exception:
    _HAVE_ASE = False

    class _Base:
        """Stand-in for ase.optimize.optimize.Optimizer: the run loop only."""

        def __init__(self, atoms, restart=None, logfile=None, trajectory=None, master=None, **kw):
            self.atoms = atoms
            self.optimizable = atoms
            self.nsteps = 0
            self.max_steps = 0
            self.fmax = None
            self.logfile = None
            if logfile == '-':
                import sys
                self.logfile = sys.stdout
            elif logfile is not None and hasattr(logfile, "write"):
                self.logfile = logfile

        def closelater(self, f):
            return f

        def run(self, fmax=0.05, steps=100000000):
            self.fmax = fmax
            self.max_steps = steps
            self.log()
            while not self.converged() and self.nsteps < steps:
                self.step()
                self.nsteps += 1
                self.log()
            return self.converged()


def _open_trajectory(trajectory, atoms, append, master):
    """optimize.py:144-150: a file name opens a trajectory attached to the atoms -- ASE's `Trajectory` when ASE is
    installed (the reference's format), else / for *.xyz names the extended-XYZ writer of
    sella_b200.utilities.trajectory; an object with a write() method is used as it is."""
    if trajectory is None:
        return None
    if isinstance(trajectory, str):
        mode = "a" if append else "w"
        if _HAVE_ASE and not trajectory.lower().endswith((".xyz", ".extxyz")):
            from ase.io.trajectory import Trajectory
            return Trajectory(trajectory, mode=mode, atoms=atoms, master=master)
        from ..utilities.trajectory import XYZTrajectory
        return XYZTrajectory(trajectory, mode=mode, atoms=atoms)
    if not hasattr(trajectory, "write"):
        raise TypeError("trajectory must be a file name or an object with a write() method")
    return trajectory


class _CalculatorSurface:
    """PES plug-in that evaluates the user's ASE calculator on the host (batch of one)."""

    def __init__(self, atoms, traj=None):
        self.atoms = atoms
        self.neval = 0
        self.traj = traj

    def evaluate(self, x, f_out, g_out, active=None):
        self.neval += 1
        old = self.atoms.positions.copy()
        self.atoms.positions = x[0].cpu().numpy().reshape((-1, 3))
        f = float(self.atoms.get_potential_energy())
        g = -np.asarray(self.atoms.get_forces(), dtype=np.float64).ravel()
        if self.traj is not None:        # PES.eval writes every evaluated geometry (peswrapper.py:409-418)
            from ..utilities.trajectory import XYZTrajectory
            if isinstance(self.traj, XYZTrajectory):
                self.traj.write(self.atoms, energy=f, forces=-g.reshape(-1, 3))
            else:
                self.traj.write()
        self.atoms.positions = old
        f_out.copy_(torch.tensor([f], dtype=torch.float64))
        g_out.copy_(torch.from_numpy(g).view(1, -1))


class _HessianView:
    """Duck type of the reference's ApproximateHessian (sella/linalg.py:143-353) over a host copy of a
    dense matrix (B is None: uninitialised = identity, linalg.py:319-334).  `evals` / `evecs` come from
    the CUDA eigensolver through the sella._gpu seam (sella_b200/_gpu.py), lazily."""

    def __init__(self, dim, B=None, initialized=None):
        self.dim = dim
        self.shape = (dim, dim)
        self.B = None if B is None else np.array(B, dtype=np.float64)
        self.initialized = (B is not None) if initialized is None else initialized
        self._evals = self._evecs = None

    def _spectrum(self):
        if self._evals is None and self.B is not None:
            from .._gpu import gpu_eigh
            self._evals, self._evecs = gpu_eigh(self.B)

    @property
    def evals(self):
        self._spectrum()
        return self._evals

    @property
    def evecs(self):
        self._spectrum()
        return self._evecs

    def asarray(self):
        return self.B if self.B is not None else np.eye(self.dim)

    def project(self, U):
        """linalg.py:306-317: U^T B U (None stays None)."""
        if self.B is None:
            return _HessianView(U.shape[1], None)
        from .._gpu import gpu_project
        return _HessianView(U.shape[1], gpu_project(self.B, np.ascontiguousarray(U)))

    def dot(self, v):
        return v if self.B is None else self.B @ v

    __matmul__ = dot

    def __add__(self, other):
        ob = other.B if isinstance(other, _HessianView) else other
        if self.B is None or ob is None:
            return _HessianView(self.dim, None, initialized=False)
        return _HessianView(self.dim, self.B + ob)

    def __sub__(self, other):
        ob = other.B if isinstance(other, _HessianView) else other
        if self.B is None or ob is None:
            return _HessianView(self.dim, None, initialized=False)
        return _HessianView(self.dim, self.B - ob)


class _PESView:
    """The reference's `dyn.pes` duck type (sella/peswrapper.py:214-606) as views over the
    device-resident engine state of this one search: names, argument meaning and return conventions of the
    reference; every array is a fresh host copy."""
    n_cell_dof = 0
    int = None
    dummies = None

    def __init__(self, opt):
        self._o = opt
        self._saved = None

    # -- state
    @property
    def _e(self):
        return self._o._eng

    @property
    def atoms(self):
        return self._o.atoms

    @property
    def cons(self):
        return self._o.constraints

    @property
    def dim(self):
        return self._e.n

    ncart = dim

    @property
    def neval(self):
        return self._o._surface.neval

    @property
    def eta(self):
        return self._e.eta

    @property
    def hessian_function(self):
        return self._o._hessian_function

    @property
    def traj(self):
        return self._o._surface.traj

    def get_x(self):
        return self._e.x[0].cpu().numpy()

    def get_f(self):
        self._e.ensure_evaluated()
        return float(self._e.f[0])

    def get_g(self):
        self._e.ensure_evaluated()
        return self._e.g[0].cpu().numpy()

    @property
    def curr(self):
        e = self._e
        ev = e._evaluated
        return dict(x=self.get_x(), f=float(e.f[0]) if ev else None, g=e.g[0].cpu().numpy() if ev else None,
                    L=self._multipliers() if ev else None)

    def save(self):
        """peswrapper.py:305-312 keeps the atomic positions of a point to come back to."""
        self._saved = self._e.x.clone()

    def restore(self):
        if self._saved is None:
            raise RuntimeError("PES.restore() without a save()")
        self._e.x.copy_(self._saved)
        self._e._evaluated = False
        self.atoms.positions = self.get_x().reshape((-1, 3))

    # -- constraints (peswrapper.py:388-438, 467-481)
    def _bases(self):
        """(drdx, Ucons, Unred, Ufree) at the current geometry, as _calc_basis (peswrapper.py:395-407)."""
        from scipy.linalg import qr
        e = self._e
        n = e.n
        drdx = self.get_drdx()
        if drdx.shape[0] == 0:
            return drdx, np.zeros((n, 0)), np.eye(n), np.eye(n)
        Q, R, _ = qr(drdx.T, mode="full", pivoting=True, check_finite=False)
        d = np.abs(np.diag(R))
        nc = int(np.sum(d > 1e-6 * d[0])) if (d.size and d[0] > 0) else 0
        return drdx, Q[:, :nc], np.eye(n), Q[:, nc:]

    def get_drdx(self):
        e = self._e
        if getattr(e, "fmask", None) is not None:
            fixed = np.nonzero(e.fmask.cpu().numpy() == 0.0)[0]
            return np.eye(e.n)[fixed]
        if e.cons is None:
            return np.zeros((0, e.n))
        if "nl" in e.cons:
            e.ensure_evaluated()
            e.converged(0.0)                      # refreshes the bases if the geometry moved
            return e.cons["C"][0].cpu().numpy()
        C = e.cons["C"]
        return (C[0] if C.dim() == 3 else C).cpu().numpy()

    def get_res(self):
        e = self._e
        if e.cons is None:
            return np.zeros(self.get_drdx().shape[0])
        e.ensure_evaluated()
        e.converged(0.0)
        return e.cons["res"][0].cpu().numpy()

    def get_Ucons(self):
        return self._bases()[1]

    def get_Unred(self):
        return self._bases()[2]

    def get_Ufree(self):
        return self._bases()[3]

    def get_scons(self):
        """peswrapper.py:429-438: minimum-norm step that restores the constraints."""
        drdx, Ucons = self._bases()[:2]
        if drdx.shape[0] == 0:
            return np.zeros(self._e.n)
        return -Ucons @ np.linalg.lstsq(drdx @ Ucons, self.get_res(), rcond=None)[0]

    def _multipliers(self):
        drdx = self.get_drdx()
        if drdx.shape[0] == 0:
            return np.zeros(0)
        return np.linalg.lstsq(drdx.T, self.get_g(), rcond=None)[0]

    # -- Hessians (peswrapper.py:335-386)
    @property
    def H(self):
        e = self._e
        return _HessianView(e.n, e.B[0].cpu().numpy() if e.H_initialized else None)

    def get_H(self):
        return self.H

    def get_Hc(self):
        e = self._e
        if e.cons is not None and "nl" in e.cons:
            e.ensure_evaluated()
            e._refresh_bases()
            nl = e.cons["nl"]
            Hc = nl["ints"].ldot(e.x, nl["Lmul"][:, nl["nlin"]:].contiguous())
            return Hc[0].cpu().numpy()
        return np.zeros((e.n, e.n))

    def get_HL(self):
        return self.get_H() - self.get_Hc()

    def get_HL_projected(self, U):
        return self.get_HL().project(U)

    # -- actions
    def converged(self, fmax, cmax=1e-5):
        e = self._e
        e.ensure_evaluated()
        e.converged(fmax, cmax)
        f1 = float(e.fmax[0])
        c1 = float(e.cons["cmax"][0]) if e.cons is not None else 0.0
        return bool(e.conv[0]), f1, c1

    def diag(self, gamma=0.1, threepoint=False, maxiter=None):
        """PES.diag (peswrapper.py:508-556): Davidson diagonalisation at the current geometry followed by
        the block update of the approximate Hessian with every operator product collected."""
        e = self._e
        e.ensure_evaluated()
        old = (e.gamma, e.threepoint, e.diag_maxiter)
        if threepoint and not e.threepoint:
            raise NotImplementedError("threepoint must be chosen when the optimiser is created")
        e.gamma, e.diag_maxiter = float(gamma), maxiter
        try:
            if e.hessian_function is not None:
                e._calculate_hessian()
            else:
                e._diag(None)
        finally:
            e.gamma, e.threepoint, e.diag_maxiter = old
        self._o._check()

    def kick(self, dx, diag=False, **diag_kwargs):
        """PES.kick (peswrapper.py:578-602): displace by dx, update the Hessian with the secant pair, optionally
        re-diagonalise; returns rho = actual / predicted energy change."""
        e = self._e
        if getattr(e, "cons", None) is not None:
            raise NotImplementedError("PES.kick with a caller-supplied displacement: unconstrained searches only")
        old = (e.gamma, e.diag_maxiter)
        e.gamma = float(diag_kwargs.get("gamma", e.gamma))
        e.diag_maxiter = diag_kwargs.get("maxiter", e.diag_maxiter)
        try:
            d = torch.from_numpy(np.asarray(dx, dtype=np.float64).reshape(1, -1).copy()).to(e.x.device)
            rho = float(e.kick(d, diag=diag)[0])
        finally:
            e.gamma, e.diag_maxiter = old
        self._o._check()
        return rho

    def set_x(self, target):
        """peswrapper.py:313-322 (Cartesian): move to `target`; returns (dx_initial, dx_final, g_par)."""
        e = self._e
        target = np.asarray(target, dtype=np.float64).ravel()
        diff = target - self.get_x()
        g0 = e.g[0].cpu().numpy() if e._evaluated else np.zeros_like(diff)
        e.x.copy_(torch.from_numpy(target[None].copy()).to(e.x.device))
        e._evaluated = False
        self.atoms.positions = target.reshape((-1, 3))
        return diff, diff, g0


class _InternalPESView:
    """`dyn.pes` of an internal-coordinate search: the reference's InternalPES duck type
    (sella/peswrapper.py:609-1288) as views over the engine state of this one search."""
    n_cell_dof = 0
    dummies = None

    def __init__(self, opt):
        self._o = opt
        self._saved = None

    _e = property(lambda self: self._o._eng)
    atoms = property(lambda self: self._o.atoms)
    int = property(lambda self: self._o.internal)
    cons = property(lambda self: self._o.internal.cons)
    dim = property(lambda self: self._e.n)
    ncart = property(lambda self: self._e.ncart)
    neval = property(lambda self: self._o._surface.neval)
    eta = property(lambda self: self._e.eta)
    traj = property(lambda self: self._o._surface.traj)
    hessian_function = None
    apos = property(lambda self: self._e.pos[0].cpu().numpy().reshape((-1, 3)))

    def get_x(self):
        return self._e.x[0].cpu().numpy()

    def get_f(self):
        self._e.ensure_evaluated()
        return float(self._e.f[0])

    def get_g(self):
        self._e.ensure_evaluated()
        return self._e.g[0].cpu().numpy()

    @property
    def H(self):
        return _HessianView(self._e.n, self._e.B[0].cpu().numpy())

    def get_H(self):
        return self.H

    def _np(self, key):
        return self._e.geo[key][0].cpu().numpy()

    def get_Unred(self):
        return self._np("Q")

    def get_Ucons(self):
        if not self._e.nc:
            return np.zeros((self._e.n, 0))
        return self._np("Q") @ self._np("Vc")

    def get_Ufree(self):
        """An orthonormal basis of the free space (peswrapper.py:1050-1082); any basis of that space serves the
        callers (column order and rotation within the space are not defined by the reference either)."""
        Q = self._np("Q")
        if not self._e.nc:
            return Q
        from scipy.linalg import qr
        Vc = self._np("Vc")
        full = qr(Vc, mode="full")[0]
        return Q @ full[:, Vc.shape[1]:]

    def get_drdx(self):
        if not self._e.nc:
            return np.zeros((0, self._e.n))
        return self._np("red") @ self._np("Q").T

    def get_res(self):
        return self._np("res") if self._e.nc else np.zeros(0)

    def get_Hc(self):
        e = self._e
        if not e.nc:
            return np.zeros((e.n, e.n))
        e.ensure_evaluated()
        if not e.geo.get("model"):
            e._model()
        Q = self._np("Q")
        return Q @ self._np("HcR") @ Q.T

    def get_HL(self):
        return self.get_H() - self.get_Hc()

    def get_HL_projected(self, U):
        return self.get_HL().project(U)

    def get_projected_forces(self):
        self.converged(0.0)
        g, Uf = self.get_g(), self.get_Ufree()
        return -((Uf @ (Uf.T @ g)) @ self._np("Bw")).reshape((-1, 3))

    @property
    def curr(self):
        e = self._e
        ev = e._evaluated
        return dict(x=self.get_x(), f=float(e.f[0]) if ev else None, g=e.g[0].cpu().numpy() if ev else None)

    def converged(self, fmax, cmax=1e-5):
        e = self._e
        e.converged(fmax, cmax)
        c1 = float(np.linalg.norm(self.get_res())) if e.nc else 0.0
        return bool(e.conv[0]), float(e.fmax[0]), c1

    def diag(self, gamma=0.1, threepoint=False, maxiter=None):
        e = self._e
        e.ensure_evaluated()
        old = (e.gamma, e.diag_maxiter)
        e.gamma, e.diag_maxiter = float(gamma), maxiter
        try:
            e._run_diag(None)
        finally:
            e.gamma, e.diag_maxiter = old
        self._o._check()

    def kick(self, dx, diag=False, **diag_kwargs):
        e = self._e
        d = torch.from_numpy(np.asarray(dx, dtype=np.float64).reshape(1, -1).copy()).to(e.x.device)
        rho = float(e.kick(d, diag=diag)[0])
        self._o._check()
        return rho


class Sella(_Base):
    def __init__(self, atoms, restart=None, logfile='-', trajectory=None, master=None, delta0=None,
                 sigma_inc=None, sigma_dec=None, rho_dec=None, rho_inc=None, order=1, eig=None, eta=1e-4,
                 method=None, gamma=0.1, threepoint=False, constraints=None, constraints_tol=1e-5, v0=None,
                 internal=False, append_trajectory=False, rs=None, nsteps_per_diag=3, diag_every_n=None,
                 hessian_function=None, optimize_cell=False, **kwargs):
        if optimize_cell:
            raise NotImplementedError("optimize_cell is not on the CUDA path")
        if internal:
            return self._init_internal(atoms, restart, logfile, trajectory, master, order, eig, eta, method, gamma,
                                       threepoint, constraints, constraints_tol, v0, internal, append_trajectory, rs,
                                       nsteps_per_diag, diag_every_n, hessian_function,
                                       dict(delta0=delta0, sigma_inc=sigma_inc, sigma_dec=sigma_dec, rho_dec=rho_dec,
                                            rho_inc=rho_inc), kwargs)
        pbc = np.asarray(getattr(atoms, "pbc", [False] * 3))
        proj_trans = kwargs.pop("proj_trans", None)
        proj_rot = kwargs.pop("proj_rot", None)
        from ..constraints import Constraints
        if constraints is None:
            constraints = Constraints(atoms)
        if not isinstance(constraints, Constraints):
            raise NotImplementedError("constraints must be a sella_b200.Constraints (translation constraints)")
        if proj_trans is None:                      # peswrapper.py:233-244
            proj_trans = not constraints.internals['translations']
        if proj_trans:
            try:
                constraints.fix_translation(replace_ok=False)
            except ValueError:
                pass
        if proj_rot is None:                        # peswrapper.py:246-253
            proj_rot = not pbc.any()
        if proj_rot and not constraints.internals['rotations']:
            constraints.fix_rotation()
        self.constraints = constraints
        lin = constraints.linear_system() if len(constraints._targets) else None
        nonlin = constraints.nonlinear_system()
        eigensolver = kwargs.pop("eigensolver", "jd0")
        # not a keyword of the reference: bounds the Davidson vectors per diagonalisation (this shell
        # holds at most 32; the reference goes on to 2n+1)
        diag_maxiter = kwargs.pop("diag_maxiter", None)
        if kwargs:
            raise TypeError("unsupported keyword arguments: %s" % sorted(kwargs))
        trajectory = _open_trajectory(trajectory, atoms, append_trajectory, master)
        if restart is not None and not _HAVE_ASE:
            # the reference hands `restart` to ase.optimize.Optimizer (it has no read() of its own)
            raise NotImplementedError("restart files are handled by ASE's Optimizer, which is not installed")
        _Base.__init__(self, atoms, restart=restart, logfile=logfile, trajectory=None, master=master)
        if trajectory is not None and isinstance(trajectory, object) and hasattr(self, "closelater") \
                and hasattr(trajectory, "close"):
            self.closelater(trajectory)
        self._surface = _CalculatorSurface(atoms, traj=trajectory)
        x0 = torch.from_numpy(np.asarray(atoms.positions, dtype=np.float64).reshape(1, -1).copy()).to(dev())
        if rs is None:
            rs = 'ras'
        if order != 0 and eig is False:
            warnings.warn("Saddle point optimizations with eig=False will most likely fail!\n Proceeding anyway, "
                          "but you shouldn't be optimistic.")
        self._eng = BatchedSella(self._surface, x0, order=order, delta0=delta0, sigma_inc=sigma_inc,
                                 sigma_dec=sigma_dec, rho_dec=rho_dec, rho_inc=rho_inc, eig=eig, eta=eta,
                                 method=method, gamma=gamma, rs=rs, nsteps_per_diag=nsteps_per_diag,
                                 diag_every_n=diag_every_n, eigensolver=eigensolver, kcap=32, threepoint=threepoint,
                                 diag_maxiter=diag_maxiter, hessian_function=self._wrap_hessian(hessian_function),
                                 v0=None if v0 is None else torch.from_numpy(
                                     np.asarray(v0, dtype=np.float64).reshape(1, -1).copy()).to(dev()),
                                 constraints=self._engine_constraints(lin, nonlin, x0))
        self._hessian_function = hessian_function
        self.pes = _PESView(self)
        self.ord = order
        self.eta = eta
        self.constraints_tol = constraints_tol
        self.fmax = None

    # ------------------------------------------------------------------ internal coordinates
    def _init_internal(self, atoms, restart, logfile, trajectory, master, order, eig, eta, method, gamma, threepoint,
                       constraints, constraints_tol, v0, internal, append_trajectory, rs, nsteps_per_diag,
                       diag_every_n, hessian_function, trust, kwargs):
        """optimize.py:237-280: `internal=True` builds the coordinate list with the three finders on a copy,
        `internal=<Internals>` takes the list as given (constraints then belong to the Internals object)."""
        from ..topology import Internals
        if hessian_function is not None or v0 is not None:
            raise NotImplementedError("hessian_function / v0 together with internal coordinates are not on the CUDA path")
        if isinstance(internal, Internals):
            auto = False
            if constraints is not None:
                raise ValueError("Internals object and Constraint object cannot both be provided to Sella. "
                                 "Instead, you must pass the Constraints object to the constructor of the "
                                 "Internals object.")
        else:
            auto = True
            internal = Internals(atoms, cons=constraints)
        self.user_internal = internal
        self._auto_internals = auto
        self.constraints = None
        eigensolver = kwargs.pop("eigensolver", "jd0")
        diag_maxiter = kwargs.pop("diag_maxiter", None)
        exact_geodesic = kwargs.pop("exact_geodesic", None)
        if kwargs.pop("iterative_stepper", 0):
            raise NotImplementedError("iterative_stepper is not on the CUDA path (the geodesic integrator is)")
        if kwargs:
            raise TypeError("unsupported keyword arguments: %s" % sorted(kwargs))
        trajectory = _open_trajectory(trajectory, atoms, append_trajectory, master)
        if restart is not None and not _HAVE_ASE:
            raise NotImplementedError("restart files are handled by ASE's Optimizer, which is not installed")
        _Base.__init__(self, atoms, restart=restart, logfile=logfile, trajectory=None, master=master)
        self._surface = _CalculatorSurface(atoms, traj=trajectory)
        if order != 0 and eig is False:
            warnings.warn("Saddle point optimizations with eig=False will most likely fail!\n Proceeding anyway, "
                          "but you shouldn't be optimistic.")
        self._ikw = dict(order=order, eig=eig, eta=eta, method=method, gamma=gamma, rs='mis' if rs is None else rs,
                         nsteps_per_diag=nsteps_per_diag, diag_every_n=diag_every_n, eigensolver=eigensolver,
                         threepoint=threepoint, diag_maxiter=diag_maxiter,
                         exact_geodesic=True if exact_geodesic is None else exact_geodesic, **trust)
        self._hessian_function = None
        self._build_internal_engine()
        self.ord = order
        self.eta = eta
        self.constraints_tol = constraints_tol
        self.fmax = None

    def _build_internal_engine(self):
        """InternalPES.__init__ (peswrapper.py:609-661): coordinate list (copy + finders), model Hessian."""
        from ..batched_internal import BatchedInternalSella
        ints = self.user_internal.copy()
        if self._auto_internals:
            ints.find_all_bonds()
            ints.find_all_angles()
            ints.find_all_dihedrals()
        rows, targets = ints.constraint_rows()
        self.internal = ints
        x0 = torch.from_numpy(np.asarray(self.atoms.positions, dtype=np.float64).reshape(1, -1).copy()).to(dev())
        self._eng = BatchedInternalSella(self._surface, x0, ints.device_coordinates(), cons_rows=rows,
                                         cons_targets=targets if len(rows) else None,
                                         h0=np.diag(ints.guess_hessian()), atol=ints.atol * 180.0 / np.pi,
                                         kcap=16, **self._ikw)
        self.pes = _InternalPESView(self)

    def _wrap_hessian(self, fn):
        """hessian_function(atoms) -> (3N, 3N) ndarray, as in the reference (optimize.py:76)."""
        if fn is None:
            return None

        def device_fn(x):
            old = self.atoms.positions.copy()
            self.atoms.positions = x[0].cpu().numpy().reshape((-1, 3))
            H = np.asarray(fn(self.atoms), dtype=np.float64)
            self.atoms.positions = old
            return torch.from_numpy(np.ascontiguousarray(H[None])).to(x.device)
        return device_fn

    @staticmethod
    def _engine_constraints(lin, nonlin, x0):
        if nonlin is None:
            return None if lin is None else (lin[0], lin[1][None, :])
        ints, tg = nonlin
        q0 = ints.calc(x0)[0].cpu().numpy()
        targets = np.where(np.isnan(tg), q0, tg)[None, :]
        return (None if lin is None else lin[0], None if lin is None else lin[1][None, :], ints, targets)

    # attributes the reference exposes
    delta = property(lambda self: float(self._eng.delta[0]))
    rho = property(lambda self: float(self._eng.rho[0]))
    nsteps_since_diag = property(lambda self: int(self._eng.since_diag[0]))

    def step(self):
        self._eng.step()
        self._check()
        if getattr(self, "internal", None) is not None and bool(self._eng.bad_internals()[0]):
            # optimize.py:382-410: an angle has come close to 0 or pi -- new coordinate list, new PES object,
            # the trust radius is kept
            if not self._auto_internals:
                raise RuntimeError("an angle of the user-supplied Internals has become linear; re-detection of the "
                                   "coordinate list needs internal=True")
            delta = self._eng.delta.clone()
            self._build_internal_engine()
            self._eng.delta.copy_(delta)

    def _check(self):
        st = int(self._eng.status[0])
        if st & 8:
            # the Davidson subspace filled its 32 slots before the reference's criterion was met: the
            # Hessian is updated with the vectors found and the search goes on (the reference would
            # keep expanding up to 2n+1 vectors)
            warnings.warn("sella_b200: Davidson stopped at the subspace capacity (32 vectors)")
            self._eng.status &= ~8
        self._eng.check_status()
        cart = self._eng.pos if getattr(self, "internal", None) is not None else self._eng.x
        self.atoms.positions = cart[0].cpu().numpy().reshape((-1, 3))

    def converged(self, forces=None):
        fmax = self.fmax if self.fmax is not None else 0.05
        # the reference evaluates the surface once before the first convergence test; the engine
        # reuses that evaluation in its first step (PES._update, peswrapper.py:440-465)
        return bool(self.pes.converged(fmax)[0])

    def gradient_converged(self, gradient=None):
        return self.converged()

    def log(self, forces=None):
        if self.logfile is None:
            return
        fmax = self.fmax if self.fmax is not None else 0.05
        _, f1, c1 = self.pes.converged(fmax)
        name = self.__class__.__name__
        T = strftime("%H:%M:%S", localtime())
        if self.nsteps == 0:
            self.logfile.write(" " * len(name) + "{:>4s} {:>8s} {:>15s} {:>12s} {:>12s} {:>12s} {:>12s}\n"
                               .format("Step", "Time", "Energy", "fmax", "cmax", "rtrust", "rho"))
        self.logfile.write("{} {:>3d} {:>8s} {:>15.6f} {:>12.4f} {:>12.4f} {:>12.4f} {:>12.4f}\n"
                           .format(name, self.nsteps, T, self.pes.get_f(), f1, c1, self.delta, self.rho))
