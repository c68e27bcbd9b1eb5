"""Host-side logic that needs no GPU: the Constraints bookkeeping (sella/internal.py:2748-3030
semantics), the bench workload builders and the synthetic geometries."""
import numpy as np
import pytest


class _Atoms:
    def __init__(self, pos):
        self.positions = np.array(pos, dtype=float)
    def __len__(self):
        return len(self.positions)


def test_constraints_bookkeeping():
    from sella_b200.constraints import Constraints, DuplicateConstraintError
    rng = np.random.RandomState(0)
    atoms = _Atoms(rng.normal(size=(6, 3)))
    cons = Constraints(atoms)
    cons.fix_translation(2)                               # three rows of the identity
    cons.fix_translation((0, 1, 3), dim=1)                # mean y of a group
    cons.fix_translation()                                # centre of geometry
    assert cons.ncons == 3 + 1 + 3
    C, c = cons.linear_system()
    assert C.shape == (7, 18)
    np.testing.assert_allclose(C @ atoms.positions.ravel(), c)            # satisfied at the start geometry
    np.testing.assert_allclose(C[3, [1, 4, 10]], 1.0 / 3)
    np.testing.assert_allclose(C[4:].sum(axis=1), 1.0)
    np.testing.assert_allclose(cons.residual(), 0.0, atol=1e-15)
    # replace_ok semantics (internal.py:2894-2904)
    cons.fix_translation(2, dim=0, target=1.5)
    assert cons.linear_system()[1][0] == 1.5 and cons.ncons == 7
    with pytest.raises(DuplicateConstraintError):
        cons.fix_translation(2, dim=0, replace_ok=False)
    with pytest.raises(ValueError):
        cons.fix_translation(2, target=0.0)               # "target" needs an explicit "dim"
    # position-dependent kinds: bookkeeping only (the kernels need a GPU)
    cons.fix_bond((4, 1), target=2.0)
    cons.fix_bond((1, 4), target=2.2)                     # same bond, reversed: replaces the target
    cons.fix_angle((0, 1, 2), target=90.0)                # degrees, as in the reference
    cons.fix_dihedral((0, 1, 2, 3))
    assert cons.nnonlinear == 3 and cons.ncons == 10
    assert cons._nl["bonds"] == [((1, 4), 2.2)]
    np.testing.assert_allclose(cons._nl["angles"][0][1], np.pi / 2)
    with pytest.raises(DuplicateConstraintError):
        cons.fix_angle((2, 1, 0), replace_ok=False)
    with pytest.raises(NotImplementedError):
        cons.fix_bond((0, 5), comparator="lt")
    with pytest.raises(NotImplementedError):
        cons.fix_rotation((0, 1, 2))                      # fragments: not on the CUDA path
    cons.fix_rotation()
    assert cons.ncons == 13 and len(cons.internals["rotations"]) == 3


def test_constraints_merge_ase_constraints():
    """Constraints.__init__ merges atoms.constraints (sella/internal.py:2760-2762, 2981-3030); the ASE
    classes are matched by name because ASE is an optional dependency."""
    from sella_b200.constraints import Constraints

    class FixAtoms:
        def __init__(self, indices):
            self.index = np.asarray(indices)

    class FixCartesian:
        def __init__(self, a, mask):
            self.a, self.mask = a, mask          # mask[d] True = relaxed (as the reference reads it)

    class FixBondLengths:
        def __init__(self, pairs, bondlengths=None):
            self.pairs, self.bondlengths = pairs, bondlengths

    class Hookean:
        pass

    rng = np.random.RandomState(1)
    atoms = _Atoms(rng.normal(size=(5, 3)))
    atoms.constraints = [FixAtoms([1, 3]), FixCartesian(0, (True, False, True)), FixBondLengths([(2, 4)], [1.7])]
    cons = Constraints(atoms)
    C, c = cons.linear_system()
    assert C.shape == (7, 15) and cons.ncons == 8
    rows = sorted(int(np.nonzero(r)[0][0]) for r in C)
    assert rows == [1, 3, 4, 5, 9, 10, 11]                   # atom 0 y; atoms 1 and 3 xyz
    np.testing.assert_allclose(C @ atoms.positions.ravel(), c)
    assert cons._nl["bonds"] == [((2, 4), 1.7)]
    cons.fix_translation(1)                                  # already there: replaces, no new rows
    assert cons.ncons == 8
    atoms.constraints = [Hookean()]
    with pytest.raises(RuntimeError):
        Constraints(atoms)


def test_synthetic_geometries():
    from sella_b200.synthetic import fcc_cluster, fcc111_slab, fcc111_with_adatom, quadratic_system
    a = 3.61
    nn = a / np.sqrt(2)
    pos = fcc_cluster(64, seed=0, rattle=0.0)
    d = np.linalg.norm(pos[:, None] - pos[None], axis=-1) + 10 * np.eye(64)
    np.testing.assert_allclose(d.min(), nn, rtol=1e-12)
    np.testing.assert_allclose(pos.mean(0), 0, atol=1e-12)
    p, cell, pbc = fcc111_slab(4, 2, 8)
    assert len(p) == 128 and pbc == (True, True, False)
    p, cell, pbc = fcc111_with_adatom()
    assert len(p) == 151
    # every slab atom has 6 in-plane neighbours at the nearest-neighbour distance (through the cell)
    top = p[125:150]
    cnt = 0
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            dd = np.linalg.norm(top[:, None] - (top[None] + i * cell[0] + j * cell[1]), axis=-1)
            cnt += (np.abs(dd - nn) < 1e-9).sum(axis=1)
    assert (cnt == 6).all()
    # the adatom sits 2 A above the bridge between two top-layer atoms
    np.testing.assert_allclose(p[-1, 2] - top[:, 2].max(), 2.0)
    np.testing.assert_allclose(sorted(np.linalg.norm(top[:, :2] - p[-1, :2], axis=1))[:2], [nn / 2, nn / 2])
    assert (p[:, 2] < cell[2, 2] / 2).sum() == 75
    A, xs, x0 = quadratic_system(3, 24)
    w = np.linalg.eigvalsh(A)
    assert (w < 0).sum() == 1 and abs(w[0] + 0.5) < 1e-12


def test_bench_workload_builders():
    import argparse
    import bench
    ns = argparse.Namespace(workload="emt-slab", n=384)
    X0, C, cell, pbc = bench.emt_problem(ns, 5, 2)
    assert X0.shape == (2, 384) and C.shape == (192, 384) and pbc == (True, True, False)
    assert (C.sum(axis=1) == 1).all() and (C.sum(axis=0) <= 1).all()
    ns = argparse.Namespace(workload="emt-cluster", n=192)
    X0, C, cell, pbc = bench.emt_problem(ns, 0, 3)
    assert X0.shape == (3, 192) and C.shape == (3, 192) and cell is None
    np.testing.assert_allclose(C @ X0[0], X0[0].reshape(-1, 3).mean(0))


def test_bench_reference_arm_contract():
    """bench.py --impl reference: rank 0 prints ONE JSON line with the contract's keys; other ranks
    exit 0 without output (the driver launches it under torchrun for N > 1)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--n", "24", "--batch", "8",
            "--steps", "2", "--warmup", "1", "--cpu-systems", "2"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run(base + ["--gpus", "2"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    out = subprocess.run(base, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] > 0
    # "reference": the reference's own Sella + PES classes (reference tree or its staged copy
    # baseline/_ref present); "port": the oracle restatement
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]


def test_bench_internal_problem_and_reference_arm():
    """BASELINE config C3 as named (internal coordinates): the host-side problem builder shared by both arms
    (one coordinate list per batch, per-system model Hessian) and the `--impl reference --internal` line."""
    import argparse
    import json
    import os
    import subprocess
    import sys
    import bench
    ns = argparse.Namespace(workload="emt-slab", n=384)
    X0, cell, pbc, ints, rows, cs, h0 = bench.internal_problem(ns, 5, 2)
    assert X0.shape == (2, 384) and ints.nbonds == 720 and ints.ntrans == 192 and ints.nint == 912
    np.testing.assert_array_equal(rows, np.arange(192))
    assert h0.shape == (2, 912) and (h0 > 0).all() and not np.array_equal(h0[0], h0[1])
    B = cs.jacobian(X0[0])
    assert np.linalg.matrix_rank(B) == 384
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "emt-slab", "--internal",
           "--n", "96", "--batch", "4"]
    out = subprocess.run(cmd, env=dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0"), capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["config"]["coordinates"] == "internal" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port"


def test_internals_bookkeeping_and_sella_internal_arguments():
    """Host-side `Internals` (sella/internal.py:3033-3260) and the argument checks of `Sella(internal=...)`
    (optimize.py:237-252) -- nothing here touches the device."""
    from sella_b200 import Sella, Internals, Constraints
    from sella_b200.topology import DuplicateInternalError, Coord
    from sella_b200.synthetic import fcc111_slab

    class Atoms:
        def __init__(self, pos, cell, pbc):
            self.positions, self.cell, self.pbc = np.array(pos, float), cell, np.array(pbc)
            self.numbers = np.full(len(pos), 29)

        def __len__(self):
            return len(self.positions)
    pos, cell, pbc = fcc111_slab(3, 2, 2, seed=1, rattle=0.02)
    at = Atoms(pos, cell, pbc)
    ints = Internals(at)
    ints.add_bond((0, 1), mic=True)
    with pytest.raises(DuplicateInternalError):
        ints.add_bond((0, 1), mic=True)
    # a coordinate equals its reverse (internal.py:359-369); the minimum-image bond is the short one
    assert Coord((0, 1), ints.internals["bonds"][0].ncvecs).same(ints.internals["bonds"][0].reverse())
    assert ints.internals["bonds"][0].value(pos, cell) < 3.0
    ints.add_translation(4)
    assert ints.ntrans == 3 and ints.nint == 4
    with pytest.raises(NotImplementedError):
        ints.add_translation((0, 1, 2), 0)
    with pytest.raises(NotImplementedError):
        Internals(at, allow_fragments=True)
    cons = Constraints(at)
    cons.fix_translation(2)
    cons.fix_bond((0, 3))
    ic = Internals(at, cons=cons)
    assert ic.ntrans == 3 and ic.nbonds == 1                      # constraint coordinates join the list first
    ic.find_all_bonds()
    rows, targets = ic.constraint_rows()
    np.testing.assert_array_equal(rows, [0, 1, 2, 3])
    np.testing.assert_allclose(targets[:3], pos[2])
    assert np.isnan(targets[3])                                    # "hold the current value"
    cp = ic.copy()
    cp.add_angle((0, 1, 2))
    assert cp.nangles == 1 and ic.nangles == 0
    with pytest.raises(ValueError):                                # Internals AND Constraints (optimize.py:241-247)
        Sella(at, internal=ic, constraints=Constraints(at), logfile=None)
    with pytest.raises(NotImplementedError):
        Sella(at, internal=True, hessian_function=lambda a: None, logfile=None)
    with pytest.raises(NotImplementedError):
        Sella(at, internal=True, optimize_cell=True, logfile=None)


def test_xyz_trajectory_writer(tmp_path):
    """f4: a frame per evaluated geometry without ASE (peswrapper.py:409-418 writes an ASE Trajectory)."""
    from sella_b200.utilities.trajectory import XYZTrajectory
    from sella_b200.optimize.optimize import _open_trajectory, _CalculatorSurface

    class Atoms:
        def __init__(self):
            self.positions = np.array([[0.0, 0.0, 0.0], [1.5, 0.2, -0.1]])
            self.numbers = np.array([29, 8])
            self.cell = np.diag([5.0, 6.0, 7.0])
            self.pbc = np.array([True, True, False])

        def __len__(self):
            return 2

        def get_potential_energy(self):
            return float((self.positions ** 2).sum())

        def get_forces(self):
            return -2.0 * self.positions
    at = Atoms()
    name = str(tmp_path / "run.xyz")
    traj = _open_trajectory(name, at, False, None)
    assert isinstance(traj, XYZTrajectory)
    torch = pytest.importorskip("torch")
    surf = _CalculatorSurface(at, traj=traj)
    f, g = torch.zeros(1, dtype=torch.float64), torch.zeros(1, 6, dtype=torch.float64)
    for k in range(3):
        surf.evaluate(torch.from_numpy(at.positions.reshape(1, -1) + 0.1 * k), f, g)
    traj.close()
    assert surf.neval == 3 and traj.nframes == 3
    lines = open(name).read().splitlines()
    assert len(lines) == 3 * 4 and lines[0] == "2" and lines[2].split()[0] == "Cu" and lines[3].split()[0] == "O"
    assert 'Lattice="5.0000000000 0.0000000000' in lines[1] and 'pbc="T T F"' in lines[1] and "forces:R:3" in lines[1]
    e2 = float([t for t in lines[9].split() if t.startswith("energy=")][0].split("=")[1])
    np.testing.assert_allclose(e2, ((at.positions + 0.2) ** 2).sum(), rtol=1e-11)
    row = np.array(lines[10].split()[1:], dtype=float)
    np.testing.assert_allclose(row[:3], at.positions[0] + 0.2, atol=1e-9)
    np.testing.assert_allclose(row[3:], -2.0 * (at.positions[0] + 0.2), atol=1e-9)
    # append mode continues the file; an object with write() is used as it is; anything else is rejected
    t2 = _open_trajectory(name, at, True, None)
    t2.write(energy=1.0)
    t2.close()
    assert len(open(name).read().splitlines()) == 4 * 4
    assert _open_trajectory(t2, at, False, None) is t2
    with pytest.raises(TypeError):
        _open_trajectory(3.5, at, False, None)


def test_topology_finders_on_small_molecules():
    """find_all_angles / find_all_dihedrals (internal.py:3457-3671) on hand-made geometries."""
    from sella_b200.topology import Internals

    class Mol:
        def __init__(self, pos, numbers):
            self.positions = np.array(pos, float)
            self.numbers = np.array(numbers)
            self.cell, self.pbc = None, np.array([False] * 3)

        def __len__(self):
            return len(self.positions)

    def full(mol):
        im = Internals(mol)
        im.find_all_bonds(); im.find_all_angles(); im.find_all_dihedrals()
        return im
    # planar star (NO3-like): 3 bonds, 3 angles, no proper dihedral through the centre -> ONE improper (n0, c, n1, n2)
    r = 1.25
    star = Mol([[0, 0, 0]] + [[r * np.cos(a), r * np.sin(a), 0] for a in (0, 2 * np.pi / 3, 4 * np.pi / 3)], [7, 8, 8, 8])
    im = full(star)
    assert (im.nbonds, im.nangles, im.ndihedrals) == (3, 3, 1)
    assert im.internals["dihedrals"][0].indices[1] == 0
    # butane-like zig-zag chain: 3 bonds, 2 angles, 1 proper dihedral; no improper (the centres carry a proper one)
    chain = Mol([[0, 0, 0], [1.5, 0, 0], [2.1, 1.4, 0], [3.6, 1.5, 0.4]], [6, 6, 6, 6])
    ic = full(chain)
    assert (ic.nbonds, ic.nangles, ic.ndihedrals) == (3, 2, 1)
    assert ic.internals["dihedrals"][0].indices in ((0, 1, 2, 3), (3, 2, 1, 0))
    np.testing.assert_allclose(np.diag(ic.guess_hessian()) > 0, True)
    # a near-linear angle at a centre with a third neighbour is replaced by an improper dihedral (:3546-3573) ...
    tee = Mol([[0, 0, 0], [1.5, 0, 0], [-1.5, 0.02, 0], [0, 1.5, 0]], [6, 6, 6, 6])
    it = Internals(tee)
    it.find_all_bonds(); it.find_all_angles()
    assert it.nangles == 2 and it.ndihedrals == 1 and len(it.forbidden["angles"]) == 1
    # ... and at a centre with only two neighbours it needs a dummy atom, which the CUDA path does not have
    co2 = Mol([[0, 0, 0], [1.16, 0, 0], [-1.16, 0, 0]], [6, 8, 8])
    i2 = Internals(co2)
    i2.find_all_bonds()
    with pytest.raises(NotImplementedError):
        i2.find_all_angles()
    # disconnected fragments are joined by growing the covalent scale factor (:3392-3418)
    dimer = Mol([[0, 0, 0], [0.96, 0, 0], [4.0, 0, 0], [4.96, 0, 0]], [8, 1, 8, 1])
    idm = Internals(dimer)
    idm.find_all_bonds()
    assert idm.nbonds >= 3
