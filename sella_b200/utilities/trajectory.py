"""A trajectory writer that needs no ASE: one extended-XYZ frame per evaluated geometry.

The reference writes every geometry the calculator evaluates into an ASE ``Trajectory``
(sella/peswrapper.py:409-418, sella/optimize/optimize.py:144-150).  ASE's binary ``.traj`` container is a
third-party format; when ASE is installed ``Sella(trajectory="run.traj")`` uses it as the reference does.
Without ASE -- or whenever the file name ends in ``.xyz`` / ``.extxyz`` -- frames go to the extended-XYZ text
format, which ``ase.io.read(name, ":")`` and most visualisers read back:

    <natoms>
    Lattice="ax ay az bx by bz cx cy cz" Properties=species:S:1:pos:R:3:forces:R:3 energy=<E> pbc="T T F"
    Cu x y z fx fy fz
    ...
"""
import numpy as np

_SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr "
            "Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W "
            "Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()


class XYZTrajectory:
    def __init__(self, filename, mode="w", atoms=None):
        if mode not in ("w", "a"):
            raise ValueError("mode must be 'w' or 'a'")
        self.atoms = atoms
        self.fh = open(filename, mode)
        self.nframes = 0

    def _symbols(self, atoms):
        if hasattr(atoms, "get_chemical_symbols"):
            return list(atoms.get_chemical_symbols())
        if hasattr(atoms, "numbers"):
            return [_SYMBOLS[int(z)] if 0 <= int(z) < len(_SYMBOLS) else "X" for z in atoms.numbers]
        return ["X"] * len(atoms)

    def write(self, atoms=None, energy=None, forces=None):
        atoms = self.atoms if atoms is None else atoms
        pos = np.asarray(atoms.positions, dtype=float).reshape(-1, 3)
        head = []
        cell = getattr(atoms, "cell", None)
        if cell is not None:
            c = np.asarray(getattr(cell, "array", cell), dtype=float)
            if c.shape == (3, 3) and np.abs(c).sum() > 0:
                head.append('Lattice="%s"' % " ".join("%.10f" % v for v in c.ravel()))
        props = "species:S:1:pos:R:3"
        if forces is not None:
            forces = np.asarray(forces, dtype=float).reshape(-1, 3)
            props += ":forces:R:3"
        head.append("Properties=" + props)
        if energy is not None:
            head.append("energy=%.12f" % float(energy))
        pbc = getattr(atoms, "pbc", None)
        if pbc is not None:
            head.append('pbc="%s"' % " ".join("T" if p else "F" for p in np.broadcast_to(np.asarray(pbc, dtype=bool), 3)))
        lines = ["%d" % len(pos), " ".join(head)]
        for i, (s, p) in enumerate(zip(self._symbols(atoms), pos)):
            row = "%-2s %18.10f %18.10f %18.10f" % (s, p[0], p[1], p[2])
            if forces is not None:
                row += " %18.10f %18.10f %18.10f" % tuple(forces[i])
            lines.append(row)
        self.fh.write("\n".join(lines) + "\n")
        self.fh.flush()
        self.nframes += 1

    def close(self):
        if self.fh is not None:
            self.fh.close()
            self.fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
