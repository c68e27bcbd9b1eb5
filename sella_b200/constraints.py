"""`Constraints(atoms)` with the reference API (sella/internal.py:2748-3030) for the constraint
kinds the CUDA path supports:

    cons = Constraints(atoms)
    cons.fix_translation(i)              # atom i held fixed (three rows of the identity)
    cons.fix_translation(i, dim=2)       # one Cartesian component
    cons.fix_translation()               # mean position of all atoms (Translation = pos[:, dim].mean(),
                                         #   sella/internal.py:466-470)
    cons.fix_translation((i, j, k))      # mean position of a group
    cons.fix_bond((i, j), target=None)           # Angstrom
    cons.fix_angle((i, j, k), target=None)       # degrees, as in the reference (:2946)
    cons.fix_dihedral((i, j, k, l), target=None) # degrees

Translations are linear (`linear_system`); bonds, angles and dihedrals are position dependent
(`nonlinear_system`: their Jacobian rows and Hessians come from the internal-coordinate kernels at
every geometry), and so is `fix_rotation()` (the three rotation coordinates of the whole
configuration, internal.py:507-800).  `fix_other`, rotations of atom subsets, inequality comparators
and `ncvecs`/`mic` periodic images are not on the CUDA path and raise NotImplementedError.
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
        self._nl = dict(bonds=[], angles=[], dihedrals=[])
        self._rot_ref = None
        # sella/internal.py:2760-2762: ASE constraints attached to the atoms become Sella constraints
        for ase_cons in (getattr(atoms, "constraints", None) or []):
            self.merge_ase_constraint(ase_cons)

    def merge_ase_constraint(self, ase_cons):
        """sella/internal.py:2981-3030.  ASE is an optional dependency: the classes are recognised by
        name and by the attributes the reference reads (FixAtoms.index, FixCartesian.a/.mask,
        FixBondLengths.pairs/.bondlengths, FixInternals.bonds/.angles/.dihedrals/.bondcombos)."""
        kind = type(ase_cons).__name__
        if kind == "FixAtoms":
            for index in np.atleast_1d(ase_cons.index):
                try:
                    self.fix_translation(int(index), replace_ok=False)
                except DuplicateConstraintError:
                    pass
        elif kind == "FixCom":
            try:
                self.fix_translation(replace_ok=False)
            except DuplicateConstraintError:
                pass
        elif kind == "FixBondLengths":
            lengths = getattr(ase_cons, "bondlengths", None)
            for i, indices in enumerate(ase_cons.pairs):
                try:
                    self.fix_bond(indices, mic=True, target=None if lengths is None else lengths[i],
                                  replace_ok=False)
                except DuplicateConstraintError:
                    pass
        elif kind == "FixCartesian":
            a = getattr(ase_cons, "a", getattr(ase_cons, "index", None))
            for dim, relaxed in enumerate(ase_cons.mask):
                if relaxed:
                    continue
                try:
                    self.fix_translation(a, dim=dim, replace_ok=False)
                except DuplicateConstraintError:
                    pass
        elif kind == "FixInternals":
            for lst, adder in ((ase_cons.bonds, self.fix_bond), (ase_cons.angles, self.fix_angle),
                               (ase_cons.dihedrals, self.fix_dihedral)):
                for target, indices in lst:
                    try:
                        adder(indices, target=target, replace_ok=False)
                    except DuplicateInternalError:
                        pass
            if getattr(ase_cons, "bondcombos", None):
                raise RuntimeError("Sella currently does not support combination constraints.")
        else:
            raise RuntimeError("Sella does not currently implement the ASE {} Constraint class.".format(kind))

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

    def fix_other(self, *a, **k):
        raise NotImplementedError("fix_other is not on the CUDA path yet")

    def fix_rotation(self, indices=None, axis=None, replace_ok=True):
        """sella/internal.py:2825-2859: the three rotation coordinates of the whole configuration
        relative to its current geometry are held at zero (what the reference adds by default for
        non-periodic systems, peswrapper.py:246-253).  Subsets of atoms / single axes are not on
        the CUDA path yet."""
        if indices is not None and sorted(int(i) for i in np.atleast_1d(indices)) != list(range(self.natoms)):
            raise NotImplementedError("fix_rotation of a subset of atoms is not on the CUDA path yet")
        if axis is not None:
            raise NotImplementedError("fix_rotation of a single axis is not on the CUDA path yet")
        if self._rot_ref is not None and not replace_ok:
            raise DuplicateConstraintError("This rotation has already been constrained!")
        self._rot_ref = np.array(self.atoms.positions, dtype=np.float64)
        self.internals['rotations'] = [('all', k) for k in range(3)]

    def _fix_internal(self, name, width, conv, indices, ncvecs=None, mic=None, target=None, comparator='eq',
                      replace_ok=True):
        """sella/internal.py:2906-2947."""
        periodic = bool(np.any(np.asarray(getattr(self.atoms, "pbc", False))))
        if ncvecs is not None or (mic and periodic):
            raise NotImplementedError("periodic-image constraints (ncvecs / mic) are not on the CUDA path yet")
        if comparator != 'eq':
            raise NotImplementedError("inequality constraints are not on the CUDA path yet")
        key = tuple(int(i) for i in indices)
        if len(key) != width:
            raise ValueError("{} needs {} atom indices".format(name, width))
        if key[0] > key[-1]:                      # a coordinate and its reverse are the same coordinate
            key = key[::-1]
        target = None if target is None else float(target) * conv
        store = self._nl[name]
        for k, (kk, _) in enumerate(store):
            if kk == key:
                if replace_ok:
                    store[k] = (key, target)
                    return
                raise DuplicateConstraintError("Coordinate {} is already fixed".format(key))
        store.append((key, target))
        self.internals[name].append(key)

    def fix_bond(self, indices, **kw):
        self._fix_internal('bonds', 2, 1.0, indices, **kw)

    def fix_angle(self, indices, **kw):
        self._fix_internal('angles', 3, np.pi / 180.0, indices, **kw)

    def fix_dihedral(self, indices, **kw):
        self._fix_internal('dihedrals', 4, np.pi / 180.0, indices, **kw)

    # -- what the engine needs
    @property
    def ncons(self):
        return len(self._targets) + self.nnonlinear

    @property
    def nnonlinear(self):
        return sum(len(v) for v in self._nl.values()) + (3 if self._rot_ref is not None else 0)

    def nonlinear_system(self):
        """(BatchedInternals over the fixed bonds/angles/dihedrals, targets [nnl] with NaN where the
        value at the start geometry is to be held), or None."""
        if not self.nnonlinear:
            return None
        from .internal import BatchedInternals
        ints = BatchedInternals(self.natoms, bonds=[k for k, _ in self._nl['bonds']],
                                angles=[k for k, _ in self._nl['angles']],
                                dihedrals=[k for k, _ in self._nl['dihedrals']], rotation_ref=self._rot_ref)
        tg = [t for name in ('bonds', 'angles', 'dihedrals') for _, t in self._nl[name]]
        if self._rot_ref is not None:
            tg += [0.0, 0.0, 0.0]
        return ints, np.array([np.nan if t is None else t for t in tg], dtype=np.float64)

    def linear_system(self):
        """(C [nc, 3N], c [nc]) with C x = c."""
        n = 3 * self.natoms
        C = np.zeros((len(self._targets), n))
        for r, ((_, dim), index) in enumerate(self.internals['translations']):
            C[r, 3 * index + dim] = 1.0 / len(index)
        return C, np.asarray(self._targets, dtype=np.float64)

    def residual(self):
        C, c = self.linear_system()
        return C @ np.asarray(self.atoms.positions, dtype=np.float64).ravel() - c

    def jacobian(self):
        return self.linear_system()[0]
