"""`Internals(atoms, ...)`: the host-side coordinate list of an internal-coordinate search, with the
reference's constructor and builder methods (sella/internal.py:3033-3830):

    ints = Internals(atoms, cons=cons)        # constraint coordinates join the list first (:3058-3063)
    ints.add_bond((i, j)); ints.add_angle((i, j, k)); ints.add_dihedral((i, j, k, l))
    ints.add_translation(i)                   # the three Cartesian coordinates of atom i
    ints.find_all_bonds(); ints.find_all_angles(); ints.find_all_dihedrals()
    Sella(atoms, internal=ints)               # or internal=True: the three finders run on a copy

It is pure index bookkeeping (numpy, no device code): the values, the Wilson matrix and the second
derivatives of the coordinates it lists are computed by the CUDA kernels behind
`sella_b200.internal.BatchedInternals` (csrc/internals.cu), which `device_coordinates()` builds.

Follows the reference's algorithms: covalent-radius bond search with a growing scale factor until the
bond graph is connected (find_all_bonds :3366-3455, _find_bonds_vectorized :3260-3332), angles from
pairs of bonds at an atom with near-linear ones replaced by improper dihedrals (find_all_angles
:3457-3573), proper dihedrals from pairs of angles sharing a bond plus impropers at 3-/4-coordinate
centres without one (find_all_dihedrals :3575-3671), the Lindh-type diagonal model Hessian
(guess_hessian :3738-3830) and the linear-angle test (check_for_bad_internals :3704-3736).

Not built (raise NotImplementedError): dummy atoms for two-coordinate linear centres (:3478-3545),
fragments with translation/rotation coordinates (`allow_fragments`), `add_rotation`, `add_other`,
translations of atom groups.  Periodic images are searched over the 27 neighbouring cells of the
GIVEN cell (the reference Minkowski-reduces it first with ASE, :3273-3279), which is the same set for
the moderately skewed cells of slabs and bulk supercells.
"""
from itertools import combinations, product

import numpy as np

# covalent radii (Angstrom), Cordero et al., Dalton Trans. 2008, 2832: the table ASE ships as
# ase.data.covalent_radii, read by the reference at internal.py:3371 and :3744; index = atomic number
COVALENT_RADII = np.array([
    0.20, 0.31, 0.28, 1.28, 0.96, 0.84, 0.76, 0.71, 0.66, 0.57, 0.58, 1.66, 1.41, 1.21, 1.11, 1.07, 1.05, 1.02,
    1.06, 2.03, 1.76, 1.70, 1.60, 1.53, 1.39, 1.39, 1.32, 1.26, 1.24, 1.32, 1.22, 1.22, 1.20, 1.19, 1.20, 1.20,
    1.16, 2.20, 1.95, 1.90, 1.75, 1.64, 1.54, 1.47, 1.46, 1.42, 1.39, 1.45, 1.44, 1.42, 1.39, 1.39, 1.38, 1.39,
    1.40, 2.44, 2.15, 2.07, 2.04, 2.03, 2.01, 1.99, 1.98, 1.98, 1.96, 1.94, 1.92, 1.92, 1.89, 1.90, 1.87, 1.87,
    1.75, 1.70, 1.62, 1.51, 1.44, 1.41, 1.36, 1.36, 1.32, 1.45, 1.46, 1.48, 1.40, 1.50, 1.50])
BOHR = 0.5291772105638411          # ase.units.Bohr / Hartree (CODATA 2014, ASE's default set)
HARTREE = 27.211386024367243


class NoValidInternalError(ValueError):
    pass


class DuplicateInternalError(ValueError):
    pass


def _numbers(atoms):
    if hasattr(atoms, "numbers"):
        return np.asarray(atoms.numbers, dtype=int)
    if hasattr(atoms, "get_atomic_numbers"):
        return np.asarray(atoms.get_atomic_numbers(), dtype=int)
    raise ValueError("Internals needs atomic numbers (atoms.numbers) for the covalent radii")


def _cell(atoms):
    c = getattr(atoms, "cell", None)
    if c is None:
        return np.zeros((3, 3))
    c = np.asarray(getattr(c, "array", c), dtype=float)
    return c if c.shape == (3, 3) else np.zeros((3, 3))


class Coord:
    """indices (m,) and integer cell offsets ncvecs (m-1, 3): atom a+1 is displaced by ncvecs[a] @ cell
    relative to atom a (internal.py:331-399).  A coordinate equals its reverse."""
    __slots__ = ("indices", "ncvecs")

    def __init__(self, indices, ncvecs=None):
        self.indices = tuple(int(i) for i in indices)
        m = len(self.indices)
        nc = np.zeros((m - 1, 3), dtype=int) if ncvecs is None else np.asarray(ncvecs, dtype=int).reshape(m - 1, 3)
        self.ncvecs = tuple(tuple(int(v) for v in row) for row in nc)

    @property
    def key(self):
        return (self.indices, self.ncvecs)

    def reverse(self):
        return Coord(self.indices[::-1], [tuple(-v for v in row) for row in self.ncvecs[::-1]])

    def same(self, other):
        return len(self.indices) == len(other.indices) and (self.key == other.key or self.reverse().key == other.key)

    def join(self, other):
        """The coordinate spanning both (bond + bond -> angle, angle + angle -> dihedral), internal.py:371-392."""
        if self.same(other):
            raise NoValidInternalError("Cannot add a coordinate to itself.")
        for s, o in product((self, self.reverse()), (other, other.reverse())):
            if s.indices[1:] == o.indices[:-1] and s.ncvecs[1:] == o.ncvecs[:-1]:
                return Coord(s.indices + (o.indices[-1],), s.ncvecs + (o.ncvecs[-1],))
        raise NoValidInternalError("indices do not overlap!")

    def split(self):
        return Coord(self.indices[:-1], self.ncvecs[:-1]), Coord(self.indices[1:], self.ncvecs[1:])

    def vectors(self, pos, cell):
        nc = np.asarray(self.ncvecs, dtype=float).reshape(-1, 3)
        idx = np.asarray(self.indices)
        return pos[idx[1:]] - pos[idx[:-1]] + nc @ cell

    def value(self, pos, cell):
        """internal.py:58-80."""
        d = self.vectors(pos, cell)
        if len(d) == 1:
            return float(np.linalg.norm(d[0]))
        if len(d) == 2:
            d1, d2 = -d[0], d[1]
            return float(np.arccos(np.clip(d1 @ d2 / (np.linalg.norm(d1) * np.linalg.norm(d2)), -1.0, 1.0)))
        c12, c23 = np.cross(d[0], d[1]), np.cross(d[1], d[2])
        return float(np.arctan2(d[1] @ np.cross(c12, c23), np.linalg.norm(d[1]) * (c12 @ c23)))


class Internals:
    _names = ("translations", "bonds", "angles", "dihedrals", "other", "rotations")

    def __init__(self, atoms, dummies=None, atol=15.0, dinds=None, cons=None, allow_fragments=False):
        if dummies is not None or dinds is not None:
            raise NotImplementedError("dummy atoms are not on the CUDA path")
        if allow_fragments:
            raise NotImplementedError("allow_fragments (per-fragment translation/rotation coordinates) is not "
                                      "on the CUDA path")
        self.atoms = atoms
        self.natoms = len(atoms)
        self.ndof = 3 * self.natoms
        self.atol = atol * np.pi / 180.0
        self.allow_fragments = False
        self.internals = {k: [] for k in self._names}
        self._keys = {k: set() for k in self._names}
        self.forbidden = {k: [] for k in self._names}
        if cons is None:
            from .constraints import Constraints
            cons = Constraints(atoms)
        self.cons = cons
        # internal.py:3058-3063: every constrained coordinate is one of the internal coordinates
        for (key, index) in cons.internals["translations"]:
            if len(index) != 1:
                raise NotImplementedError("translations of atom groups (fix_translation() of a centre of mass) "
                                          "together with internal coordinates are not on the CUDA path")
            self.add_translation(int(index[0]), key[1])
        for name, adder in (("bonds", self.add_bond), ("angles", self.add_angle), ("dihedrals", self.add_dihedral)):
            for idx in cons.internals[name]:
                adder(idx)
        if cons.internals["rotations"] or cons.internals["other"]:
            raise NotImplementedError("rotation / other constraints together with internal coordinates are not "
                                      "on the CUDA path")

    # ------------------------------------------------------------------ bookkeeping
    @property
    def positions(self):
        return np.asarray(self.atoms.positions, dtype=float)

    ntrans = property(lambda self: len(self.internals["translations"]))
    nbonds = property(lambda self: len(self.internals["bonds"]))
    nangles = property(lambda self: len(self.internals["angles"]))
    ndihedrals = property(lambda self: len(self.internals["dihedrals"]))
    nother = property(lambda self: 0)
    nrotations = property(lambda self: 0)
    nint = property(lambda self: self.ntrans + self.nbonds + self.nangles + self.ndihedrals)
    ndummies = 0

    def copy(self):
        new = Internals(self.atoms, atol=self.atol * 180.0 / np.pi, cons=self.cons)
        for name in self._names:
            new.internals[name] = list(self.internals[name])
            new._keys[name] = set(self._keys[name])
            new.forbidden[name] = list(self.forbidden[name])
        return new

    def add_translation(self, index=None, dim=None):
        """internal.py:3117-3144 for single atoms: coordinate = positions[index, dim]."""
        if index is None or not np.isscalar(index):
            raise NotImplementedError("translations of atom groups are not on the CUDA path")
        if dim is None:
            for d in range(3):
                self.add_translation(index, d)
            return
        key = (int(index), int(dim))
        if key in self._keys["translations"]:
            raise DuplicateInternalError
        self.internals["translations"].append(key)
        self._keys["translations"].add(key)

    def _mic(self, indices):
        """internal.py:2651-2668: cell offsets of the minimum-image chain through `indices`."""
        pos, cell = self.positions, _cell(self.atoms)
        pbc = np.asarray(getattr(self.atoms, "pbc", (False,) * 3), dtype=bool) & (np.abs(cell).sum(axis=1) > 0)
        out = []
        for a, b in zip(indices[:-1], indices[1:]):
            dx = pos[b] - pos[a]
            best, bestd = (0, 0, 0), np.inf
            for t in product(*[(-1, 0, 1) if p else (0,) for p in pbc]):
                d = np.linalg.norm(dx + np.asarray(t, float) @ cell)
                if d < bestd - 1e-12:
                    best, bestd = t, d
            out.append(best)
        return out

    def _add(self, name, width, indices, ncvecs=None, mic=None):
        if isinstance(indices, Coord):
            new = indices
        else:
            if len(indices) != width:
                raise ValueError("{} need {} atom indices".format(name, width))
            if ncvecs is None and mic:
                ncvecs = self._mic(tuple(indices))
            new = Coord(indices, ncvecs)
        if new.key in self._keys[name] or any(new.same(f) for f in self.forbidden[name]):
            raise DuplicateInternalError
        self.internals[name].append(new)
        self._keys[name].add(new.key)

    def add_bond(self, indices, ncvecs=None, mic=None):
        self._add("bonds", 2, indices, ncvecs, mic)

    def add_angle(self, indices, ncvecs=None, mic=None):
        self._add("angles", 3, indices, ncvecs, mic)

    def add_dihedral(self, indices, ncvecs=None, mic=None):
        self._add("dihedrals", 4, indices, ncvecs, mic)

    def forbid_angle(self, coord):
        if not any(coord.same(f) for f in self.forbidden["angles"]):
            self.forbidden["angles"].append(coord)

    def add_rotation(self, *a, **k):
        raise NotImplementedError("rotation coordinates of fragments are not on the CUDA path")

    def add_other(self, *a, **k):
        raise NotImplementedError("user-defined coordinates are not on the CUDA path")

    # ------------------------------------------------------------------ topology search
    def _bond_candidates(self, labels, scale, rcov):
        """internal.py:3260-3332 without the Minkowski reduction."""
        pos, cell = self.positions, _cell(self.atoms)
        pbc = np.asarray(getattr(self.atoms, "pbc", (False,) * 3), dtype=bool) & (np.abs(cell).sum(axis=1) > 0)
        ii, jj = np.triu_indices(self.natoms, k=0)
        keep = ~((labels[ii] == labels[jj]) & (labels[ii] != -1))
        ii, jj = ii[keep], jj[keep]
        if len(ii) == 0:
            return []
        dx = pos[jj] - pos[ii]
        offset = np.zeros(dx.shape, dtype=int)
        if pbc.any():
            full = cell.copy()
            for d in range(3):                       # complete a partly periodic cell so that it can be inverted
                if not pbc[d] or np.abs(full[d]).sum() == 0:
                    full[d] = 0.0
                    full[d, d] = 1.0
            frac = dx @ np.linalg.inv(full)
            for _ in range(2):
                offset += (pbc * np.floor(frac - offset)).astype(int)
        base = np.array(list(product(*[np.arange(-1 * p, p + 1) for p in pbc])), dtype=int)
        shifted = base[None, :, :] - offset[:, None, :]
        dist = np.linalg.norm(dx[:, None, :] + shifted @ cell, axis=2)
        mask = dist <= (scale * (rcov[ii] + rcov[jj]))[:, None]
        mask &= ~((ii == jj)[:, None] & np.all(shifted == 0, axis=2))
        out = []
        for p, t in zip(*np.nonzero(mask)):
            out.append((int(ii[p]), int(jj[p]), tuple(int(v) for v in shifted[p, t])))
        return out

    def find_all_bonds(self, nbond_cart_thr=6, max_bonds=20, scale=1.25):
        """internal.py:3366-3418."""
        rcov = COVALENT_RADII[_numbers(self.atoms)]
        n = self.natoms
        neigh = [[] for _ in range(n)]
        for b in self.internals["bonds"]:
            i, j = b.indices
            neigh[i].append(j)
            neigh[j].append(i)
        first = True
        while True:
            labels = -np.ones(n, dtype=int)
            nlabels = 0
            for i in range(n):                        # flood fill over the bond graph
                if labels[i] == -1:
                    stack = [i]
                    labels[i] = nlabels
                    while stack:
                        a = stack.pop()
                        for c in neigh[a]:
                            if labels[c] != nlabels:
                                labels[c] = nlabels
                                stack.append(c)
                    nlabels += 1
            if nlabels == 1:
                break
            labels[np.array([len(v) == 0 for v in neigh])] = -1
            for i, j, ts in self._bond_candidates(labels, scale, rcov):
                try:
                    self.add_bond((i, j), [ts])
                except DuplicateInternalError:
                    continue
                if len(neigh[i]) < max_bonds and len(neigh[j]) < max_bonds:
                    neigh[i].append(j)
                    neigh[j].append(i)
            first = False
            scale *= 1.05
        return first

    def find_all_angles(self):
        """internal.py:3457-3573 (no dummy atoms)."""
        pos, cell = self.positions, _cell(self.atoms)
        at = [[] for _ in range(self.natoms)]
        for b in self.internals["bonds"]:
            i, j = b.indices
            at[i].append(b)
            at[j].append(b.reverse())
        for j, jb in enumerate(at):
            linear = []
            for b1, b2 in combinations(jb, 2):
                new = b1.join(b2)
                assert new.indices[1] == j
                if self.atol < new.value(pos, cell) < np.pi - self.atol:
                    try:
                        self.add_angle(new)
                    except DuplicateInternalError:
                        pass
                else:
                    self.forbid_angle(new)
                    linear.append((b1, b2))
            if not linear:
                continue
            if len(jb) == 2:
                raise NotImplementedError("a linear two-coordinate centre (atom %d) needs a dummy atom "
                                          "(internal.py:3478-3545), which is not on the CUDA path" % j)
            for b1, b2 in linear:
                for b3 in jb:
                    if b3 is b1 or b3 is b2:
                        continue
                    n1, n3, n2 = np.array(b1.ncvecs[0]), np.array(b3.ncvecs[0]), np.array(b2.ncvecs[0])
                    try:
                        self.add_dihedral((b1.indices[1], j, b3.indices[1], b2.indices[1]), (-n1, n3, n2 - n3))
                    except DuplicateInternalError:
                        pass
                    break
                else:
                    raise RuntimeError("Unable to find improper dihedral to replace linear angle!")

    def find_all_dihedrals(self):
        """internal.py:3575-3671."""
        edges = {}
        for a in self.internals["angles"]:
            i, j, k = a.indices
            for e in ((min(i, j), max(i, j)), (min(j, k), max(j, k))):
                edges.setdefault(e, []).append(a)
        seen = set()
        for lst in edges.values():
            for a1, a2 in combinations(lst, 2):
                if (id(a1), id(a2)) in seen:
                    continue
                seen.add((id(a1), id(a2)))
                try:
                    new = a1.join(a2)
                except NoValidInternalError:
                    continue
                if new.indices[0] == new.indices[3] and not np.any(np.sum(np.array(new.ncvecs), axis=0)):
                    continue
                try:
                    self.add_dihedral(new)
                except DuplicateInternalError:
                    continue
        centres = set()
        for d in self.internals["dihedrals"]:
            centres.update(d.indices[1:3])
        neigh = [[] for _ in range(self.natoms)]
        for b in self.internals["bonds"]:
            i, j = b.indices
            neigh[i].append((j, np.array(b.ncvecs[0])))
            neigh[j].append((i, -np.array(b.ncvecs[0])))
        for c in range(self.natoms):
            if len(neigh[c]) not in (3, 4) or c in centres:
                continue
            (n0, v0), (n1, v1), (n2, v2) = neigh[c][:3]
            try:
                self.add_dihedral((n0, c, n1, n2), (-v0, v1, v2 - v1))
            except DuplicateInternalError:
                pass

    # ------------------------------------------------------------------ what the engine needs
    def lists(self):
        """(translations, bonds, angles, dihedrals) index lists and the matching Cartesian shift vectors."""
        cell = _cell(self.atoms)
        out, tv = {}, {}
        for name in ("bonds", "angles", "dihedrals"):
            out[name] = [c.indices for c in self.internals[name]]
            width = dict(bonds=1, angles=2, dihedrals=3)[name]
            tv[name] = np.array([np.asarray(c.ncvecs, dtype=float).reshape(-1, 3) @ cell
                                 for c in self.internals[name]]).reshape(len(out[name]), width, 3)
        return list(self.internals["translations"]), out["bonds"], out["angles"], out["dihedrals"], tv

    def device_coordinates(self):
        """The CUDA evaluator of this coordinate list (values, Wilson matrix, second derivatives)."""
        from .internal import BatchedInternals
        tr, b, a, d, tv = self.lists()
        return BatchedInternals(self.natoms, tr, b, a, d, tvec_bonds=tv["bonds"] if len(b) else None,
                                tvec_angles=tv["angles"] if len(a) else None,
                                tvec_dihedrals=tv["dihedrals"] if len(d) else None)

    def constraint_rows(self):
        """Positions (within this coordinate list) of the constrained coordinates and their targets
        (NaN: hold the value at the start geometry)."""
        rows, targets = [], []
        cons = self.cons
        for k, (key, index) in enumerate(cons.internals["translations"]):
            rows.append(self.internals["translations"].index((int(index[0]), key[1])))
            targets.append(cons._targets[k])
        off = self.ntrans
        for name in ("bonds", "angles", "dihedrals"):
            for key, target in cons._nl[name]:
                c = Coord(key)
                pos = [i for i, x in enumerate(self.internals[name]) if x.same(c)]
                rows.append(off + pos[0])
                targets.append(np.nan if target is None else target)
            off += len(self.internals[name])
        return np.asarray(rows, dtype=int), np.asarray(targets, dtype=float)

    def check_for_bad_internals(self, positions=None):
        """internal.py:3704-3736: the angles that have come within `atol` of 0 or pi, or None."""
        pos = self.positions if positions is None else np.asarray(positions, float).reshape(-1, 3)
        cell = _cell(self.atoms)
        bad = [a for a in self.internals["angles"] if not (self.atol < a.value(pos, cell) < np.pi - self.atol)]
        return dict(bonds=[], angles=bad) if bad else None

    def guess_hessian(self, h0cart=70.0):
        """internal.py:3738-3830: diagonal matrix of model force constants."""
        pos, cell = self.positions, _cell(self.atoms)
        rc = COVALENT_RADII[_numbers(self.atoms)]
        nb = np.zeros(self.natoms, dtype=int)
        h0 = [h0cart] * self.ntrans
        for b in self.internals["bonds"]:
            i, j = b.indices
            nb[i] += 1
            nb[j] += 1
            h0.append(0.3601 * np.exp(-1.944 * (b.value(pos, cell) - rc[i] - rc[j]) / BOHR) * HARTREE / BOHR ** 2)
        for a in self.internals["angles"]:
            ab, bc = a.split()
            cab, cbc = rc[list(ab.indices)].sum(), rc[list(bc.indices)].sum()
            r = ab.value(pos, cell) + bc.value(pos, cell)
            h0.append((0.089 + 0.11 * np.exp(-0.44 * (r - cab - cbc) / BOHR) / (cab * cbc / BOHR ** 2) ** -0.42)
                      * HARTREE)
        for d in self.internals["dihedrals"]:
            bc = d.split()[0].split()[1]
            cbc, rbc = rc[list(bc.indices)].sum(), bc.value(pos, cell)
            L = nb[list(bc.indices)].sum() - 2
            h0.append((0.0015 + 14.0 * L ** 0.57 * np.exp(-2.85 * (rbc - cbc) / BOHR)
                       / (rbc * cbc / BOHR ** 2) ** 4.0) * HARTREE)
        return np.diag(np.abs(np.asarray(h0, dtype=float)))
