"""TEST INFRASTRUCTURE ONLY -- duck-typed ``atoms`` / ``constraints`` objects that
let the *reference's own* ``PES`` (peswrapper.py:214-606) and ``Sella``
(optimize/optimize.py:42-502) classes run on a synthetic ``x -> (f, g)`` surface
without ASE or JAX being installed.  Nothing here computes any optimiser
arithmetic: the objects only hand positions, energies, forces and a constant
linear constraint Jacobian to the reference code.

Used by tests/golden/make_golden.py (build container only).
"""
import numpy as np


class SurfaceAtoms:
    """What peswrapper.PES touches on ``atoms``: positions, pbc, cell, len(),
    constraints, get_potential_energy(), get_forces()."""

    def __init__(self, func, x0):
        self.func = func
        self.positions = np.array(x0, dtype=float).reshape((-1, 3))
        self.pbc = np.array([True, True, True])     # -> proj_rot defaults to False
        self.cell = None
        self.constraints = []
        self.calc = None

    def __len__(self):
        return len(self.positions)

    def get_potential_energy(self):
        return self.func(self.positions.ravel())[0]

    def get_forces(self):
        return -self.func(self.positions.ravel())[1].reshape((-1, 3))


class _ZeroCurvature:
    def __init__(self, n):
        self.n = n

    def ldot(self, L):
        return np.zeros((self.n, self.n))


class LinearConstraints:
    """Stand-in for sella.internal.Constraints holding constant rows C x = c
    (exactly what per-atom ``fix_translation`` yields: rows of the identity)."""

    def __init__(self, atoms, C=None, c=None):
        self.atoms = atoms
        n = 3 * len(atoms)
        self.C = np.zeros((0, n)) if C is None else np.asarray(C, float)
        self.c = np.zeros(self.C.shape[0]) if c is None else np.asarray(c, float)
        self.internals = dict(translations=[True])   # truthy -> proj_trans False

    def jacobian(self):
        return self.C

    def residual(self):
        return self.C @ self.atoms.positions.ravel() - self.c

    def hessian(self):
        return _ZeroCurvature(self.C.shape[1])

    def disable_satisfied_inequalities(self):
        pass

    def has_inequalities(self):
        return False

    def validate_inequalities(self):
        return True


def make_reference_sella(ref, func, x0, C=None, c=None, **kw):
    """Instantiate the reference's Sella on a synthetic surface."""
    atoms = SurfaceAtoms(func, x0)
    cons = LinearConstraints(atoms, C, c)
    kw.setdefault("logfile", None)
    dyn = ref.optimize.Sella(atoms, constraints=cons, proj_trans=False,
                             proj_rot=False, **kw)
    return dyn
