"""TEST INFRASTRUCTURE ONLY -- duck-typed ``Internals`` / ``Constraints`` / ``atoms`` objects that let the
reference's OWN ``InternalPES`` (sella/peswrapper.py:609-1288), ``MaxInternalStep``
(sella/optimize/restricted_step.py:186-243) and ``Sella`` (sella/optimize/optimize.py) run without JAX or ASE.

The reference's ``sella/internal.py`` cannot be imported here (JAX), so the *coordinate values and
derivatives* these objects hand to the reference code come from ``oracle.intcoords.CoordinateSet`` (hyper-dual
derivatives of the reference's primal formulas, checked by finite differences and against the reference's
``SparseInternalHessians`` assembly).  Everything that is compared afterwards -- the Wilson-matrix algebra,
the geodesic ODE and its LSODA integration, the constraint projection, Hc, the restricted step, the trust-radius
policy, the Hessian updates, Davidson -- is executed by the reference's unmodified files.

Used by tests/golden/make_golden.py (build container only) to produce tests/golden/internal_loop.npz, which
pins ``oracle/internal_pes.py``.
"""
import numpy as np


class _Dummies:
    """ase.Atoms of dummy atoms: none."""

    def __init__(self):
        self.positions = np.zeros((0, 3))

    def __len__(self):
        return 0


class _Curvature:
    def __init__(self, cs, atoms):
        self.cs, self.atoms = cs, atoms

    def ldot(self, v):
        return self.cs.ldot(self.atoms.positions.ravel(), np.asarray(v, float))


class DuckCoordinates:
    """The part of BaseInternals (sella/internal.py:1209-2745) that peswrapper.py / restricted_step.py /
    optimize.py call, evaluated at ``atoms.positions`` at call time."""

    def __init__(self, atoms, cs):
        self.atoms, self.cs = atoms, cs
        self.dummies = _Dummies()
        self.ndof = cs.ndof
        self.ntrans, self.nbonds, self.nangles = cs.ntrans, cs.nbonds, cs.nangles
        self.ndihedrals, self.nother, self.nrotations = cs.ndihedrals, 0, 0
        self.nint = cs.nint

    def _x(self):
        return np.asarray(self.atoms.positions, float).ravel()

    def calc(self):
        return self.cs.calc(self._x())

    def jacobian(self):
        return self.cs.jacobian(self._x())

    def hessian(self):
        return _Curvature(self.cs, self.atoms)

    def hessian_rdot(self, v):
        return self.cs.rdot(self._x(), np.asarray(v, float))

    def wrap(self, vec):
        return self.cs.wrap(vec)


class DuckConstraints(DuckCoordinates):
    """sella.internal.Constraints over a coordinate subset held at `targets`."""

    def __init__(self, atoms, cs, targets=None):
        DuckCoordinates.__init__(self, atoms, cs)
        self.targets = self.calc() if targets is None else np.asarray(targets, float)
        self.internals = dict(translations=[True] * cs.ntrans)

    def residual(self):
        return self.wrap(self.calc() - self.targets)

    def disable_satisfied_inequalities(self):
        pass

    def has_inequalities(self):
        return False

    def validate_inequalities(self):
        return True


class DuckInternals(DuckCoordinates):
    """sella.internal.Internals: an explicit coordinate list (no topology search, no dummies)."""

    def __init__(self, atoms, cs, cons):
        DuckCoordinates.__init__(self, atoms, cs)
        self.cons = cons
        self.allow_fragments = False

    def copy(self):
        return DuckInternals(self.atoms, self.cs, self.cons)

    def validate_basis(self):
        pass

    def check_for_bad_internals(self):
        bad = self.cs.bad_angles(self._x())
        return None if bad is None else dict(bonds=[], angles=list(bad))

    def guess_hessian(self, h0cart=70.0):
        return np.diag(self.cs.guess_hessian(self._x(), h0cart))


class _Cell:
    """ase.cell.Cell as far as PES._state_hash reads it (peswrapper.py:297-303)."""

    def __init__(self, array):
        self.array = np.asarray(array, float)

    def any(self):
        return bool(self.array.any())


class SurfaceAtoms:
    def __init__(self, func, pos, cell=None, pbc=(False, False, False)):
        self.func = func
        self.positions = np.array(pos, dtype=float).reshape((-1, 3))
        self.pbc = np.array(pbc)
        self.cell = None if cell is None else _Cell(cell)
        self.constraints = []
        self.calc = None

    def __len__(self):
        return len(self.positions)

    def get_potential_energy(self):
        return self.func(self.positions.ravel())[0]

    def get_forces(self):
        return -self.func(self.positions.ravel())[1].reshape((-1, 3))


def make_reference_internal_sella(ref, func, pos, cs, csc, cell=None, pbc=(False,) * 3, **kw):
    """The reference's Sella(atoms, internal=<Internals>) on a synthetic / EMT-form surface.  `cs`: the
    coordinate list, `csc`: its constrained subset (None: no constraints)."""
    from .intcoords import CoordinateSet
    atoms = SurfaceAtoms(func, pos, cell, pbc)
    if csc is None:
        csc = CoordinateSet(cs.natoms)
    cons = DuckConstraints(atoms, csc)
    ints = DuckInternals(atoms, cs, cons)
    # `isinstance(internal, Internals)` (optimize.py:238): the name both modules imported from the stubbed
    # sella.internal becomes the duck class
    ref.optimize.Internals = DuckInternals
    ref.peswrapper.Internals = DuckInternals
    kw.setdefault("logfile", None)
    return ref.optimize.Sella(atoms, internal=ints, **kw)
