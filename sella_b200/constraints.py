"""`Constraints(atoms)` with the part of the reference API (sella/internal.py:2748-3030)
that yields LINEAR constraints, which is what the CUDA path supports:

    cons = Constraints(atoms)
    cons.fix_translation(i)              # atom i held fixed (three rows of the identity)
    cons.fix_translation(i, dim=2)       # one Cartesian component
    cons.fix_translation()               # mean position of all atoms (Translation = pos[:, dim].mean(),
                                         #   sella/internal.py:466-470)
    cons.fix_translation((i, j, k))      # mean position of a group

`fix_rotation`, `fix_bond`, `fix_angle`, `fix_dihedral`, `fix_other` are nonlinear
(geometry-dependent Jacobian, non-zero constraint Hessian) and raise NotImplementedError.
"""
import numpy as np


class DuplicateInternalError(ValueError):
    pass


class DuplicateConstraintError(DuplicateInternalError):
    pass


class Constraints:
    def __init__(self, atoms):
        self.atoms = atoms
        self.natoms = len(atoms)
        self.internals = dict(translations=[], bonds=[], angles=[], dihedrals=[], other=[], rotations=[])
        self._targets = []

    def fix_translation(self, index=None, dim=None, target=None, replace_ok=True):
        if index is None:
            index = np.arange(self.natoms, dtype=np.int32)
        if np.isscalar(index):
            index = np.array((index,), dtype=np.int32)
        index = np.asarray(index, dtype=np.int32)
        if dim is None:
            if target is not None:
                raise ValueError('"target" keyword requires explicit "dim"!')
            for d in range(3):
                self.fix_translation(index, dim=d, replace_ok=replace_ok)
            return
        key = (frozenset(int(i) for i in index), int(dim))
        if target is None:
            target = float(np.asarray(self.atoms.positions)[index, dim].mean())
        for k, (kk, _) in enumerate(self.internals['translations']):
            if kk == key:
                if replace_ok:
                    self._targets[k] = target
                    return
                raise DuplicateConstraintError("Coordinate {} is already fixed".format(key))
        self.internals['translations'].append((key, index.copy()))
        self._targets.append(target)

    def _nonlinear(self, *a, **k):
        raise NotImplementedError("only translation constraints (linear) are on the CUDA path yet")

    fix_rotation = fix_bond = fix_angle = fix_dihedral = fix_other = _nonlinear

    # -- what the engine needs
    @property
    def ncons(self):
        return len(self._targets)

    def linear_system(self):
        """(C [nc, 3N], c [nc]) with C x = c."""
        n = 3 * self.natoms
        C = np.zeros((self.ncons, n))
        for r, ((_, dim), index) in enumerate(self.internals['translations']):
            C[r, 3 * index + dim] = 1.0 / len(index)
        return C, np.asarray(self._targets, dtype=np.float64)

    def residual(self):
        C, c = self.linear_system()
        return C @ np.asarray(self.atoms.positions, dtype=np.float64).ravel() - c

    def jacobian(self):
        return self.linear_system()[0]
