"""Internal-coordinate searches (reference: InternalPES, sella/peswrapper.py:609-1288; MaxInternalStep,
optimize/restricted_step.py:186-243; Internals topology, internal.py:3033-3830).

CPU: the exact-derivative coordinate oracle against the finite-difference one, the host-side topology
search, the oracle InternalPES (reference integrator LSODA vs the Dormand-Prince scheme the engine runs),
and the engine's Wilson-matrix algebra / geodesic integrator with the device operators replaced by
torch-CPU stand-ins (tests only -- the product has no CPU path).
GPU: the CUDA engine against the oracle loop, step by step, and `Sella(atoms, internal=True)`.
"""
import numpy as np
import pytest

from oracle import emt as oemt
from oracle import internals as oi
from oracle.intcoords import CoordinateSet
from oracle.internal_pes import InternalPES
from oracle.driver import SaddleSearch
from sella_b200.synthetic import fcc111_slab, fcc_cluster
from sella_b200.topology import Internals
from sella_b200.constraints import Constraints


class _Atoms:
    """The part of ase.Atoms the optimiser touches."""

    def __init__(self, pos, cell, pbc, func=None, number=29):
        self.positions = np.array(pos, dtype=float)
        self.cell, self.pbc = cell, np.array(pbc)
        self.numbers = np.full(len(pos), number)
        self.func = func

    def __len__(self):
        return len(self.positions)

    def get_potential_energy(self):
        return self.func(self.positions.ravel())[0]

    def get_forces(self):
        return -self.func(self.positions.ravel())[1].reshape(-1, 3)


def slab_problem(seed, nx=2, ny=1, nl=4, rattle=0.08, angles=False):
    if angles:
        return cluster_problem(31, variant=seed % 10)
    """C3-style: Cu(111) slab, bottom half held by fix_translation, nearest-neighbour bonds (+ the fixed atoms'
    Cartesian coordinates) as internal coordinates."""
    pos, cell, pbc = fcc111_slab(nx, ny, nl, seed=seed, rattle=rattle)
    at = _Atoms(pos, cell, pbc, oemt.emt_func(cell, pbc))
    cons = Constraints(at)
    for i in np.nonzero(pos[:, 2] < pos[:, 2].mean())[0]:
        cons.fix_translation(int(i))
    ints = Internals(at, cons=cons)
    ints.find_all_bonds()
    return at, cons, ints


def cluster_problem(seed, natoms=10, rattle=0.08, variant=0):
    """A Cu cluster with three atoms held by fix_translation (so that the Wilson matrix has full column rank)
    and the full automatic coordinate list: bonds, angles, improper dihedrals for the near-linear angles."""
    pos = fcc_cluster(natoms, seed=seed, rattle=rattle)
    at = _Atoms(pos, None, (False,) * 3, oemt.emt_func(None, (False,) * 3))
    cons = Constraints(at)
    for i in range(3):
        cons.fix_translation(i)
    ints = Internals(at, cons=cons)
    ints.find_all_bonds()
    ints.find_all_angles()
    ints.find_all_dihedrals()
    if variant:                       # same coordinate list, slightly different start geometry (a batch shares its list)
        at.positions += 0.02 * np.random.RandomState(1000 + variant).normal(size=at.positions.shape)
        cons2 = Constraints(at)
        for i in range(3):
            cons2.fix_translation(i)
        ints.cons = cons = cons2
    return at, cons, ints


def molecule_problem(variant=0, natoms=8):
    """A free Cu cluster (no constraints): the Wilson matrix has rank 3N - 6, the reference's SVD branch."""
    pos = fcc_cluster(natoms, seed=77, rattle=0.08)
    at = _Atoms(pos, None, (False,) * 3, oemt.emt_func(None, (False,) * 3))
    cons = Constraints(at)
    ints = Internals(at, cons=cons)
    ints.find_all_bonds()
    ints.find_all_angles()
    ints.find_all_dihedrals()
    if variant:
        at.positions += 0.02 * np.random.RandomState(2000 + variant).normal(size=at.positions.shape)
    return at, cons, ints


def constrained_molecule_problem(variant=0):
    """A free Cu cluster with one bond and one angle held (fix_bond / fix_angle): the constraint rows are ordinary
    internal coordinates with non-zero second derivatives (the D_cons term of peswrapper.py:1011-1031)."""
    pos = fcc_cluster(8, seed=77, rattle=0.08)
    at = _Atoms(pos, None, (False,) * 3, oemt.emt_func(None, (False,) * 3))
    cons = Constraints(at)
    cons.fix_bond((0, 1))
    cons.fix_angle((1, 0, 2))
    ints = Internals(at, cons=cons)
    ints.find_all_bonds()
    ints.find_all_angles()
    ints.find_all_dihedrals()
    if variant:
        at.positions += 0.02 * np.random.RandomState(3000 + variant).normal(size=at.positions.shape)
    return at, cons, ints


def oracle_sets(ints):
    tr, b, a, d, tv = ints.lists()
    cs = CoordinateSet(ints.natoms, tr, b, a, d, tvecs=tv, numbers=ints.atoms.numbers)
    rows, tg = ints.constraint_rows()
    csc = subset(ints.natoms, (tr, b, a, d, tv), rows) if len(rows) else None
    return cs, csc, rows


def subset(natoms, lists, rows):
    """The coordinates `rows` (positions in the full list, kind by kind in the reference's order) as their own set."""
    tr, b, a, d, tv = lists
    sel = dict(translations=[], bonds=[], angles=[], dihedrals=[])
    tvs = dict(bonds=[], angles=[], dihedrals=[])
    off = [0, len(tr), len(tr) + len(b), len(tr) + len(b) + len(a)]
    for r in rows:
        if r < off[1]:
            sel["translations"].append(tr[r])
        else:
            kind, lst, o = (("bonds", b, off[1]) if r < off[2] else ("angles", a, off[2]) if r < off[3]
                            else ("dihedrals", d, off[3]))
            sel[kind].append(lst[r - o])
            tvs[kind].append(tv[kind][r - o])
    return CoordinateSet(natoms, sel["translations"], sel["bonds"], sel["angles"], sel["dihedrals"],
                         tvecs={k: (np.array(v) if len(v) else None) for k, v in tvs.items()})


def test_exact_coordinate_derivatives_match_the_finite_difference_oracle():
    rng = np.random.RandomState(1)
    n = 9
    pos = rng.normal(size=(n, 3)) * 1.2 + np.arange(n)[:, None] * np.array([0.9, 0.2, -0.1])
    bonds = [(i, i + 1) for i in range(n - 1)]
    angles = [(i, i + 1, i + 2) for i in range(n - 2)]
    diheds = [(i, i + 1, i + 2, i + 3) for i in range(n - 3)]
    trans = [(0, 0), (0, 1), (n - 1, 2)]
    tv = dict(bonds=rng.normal(size=(len(bonds), 1, 3)) * .3, angles=rng.normal(size=(len(angles), 2, 3)) * .3,
              dihedrals=rng.normal(size=(len(diheds), 3, 3)) * .3)
    for t in (None, tv):
        q, B, H = oi.evaluate(pos, trans, bonds, angles, diheds, tvecs=t)
        cs = CoordinateSet(n, trans, bonds, angles, diheds, tvecs=t)
        np.testing.assert_allclose(cs.calc(pos), q, atol=1e-14)
        np.testing.assert_allclose(cs.jacobian(pos), B, atol=1e-13)
        v, w = rng.normal(size=cs.nint), rng.normal(size=3 * n)
        np.testing.assert_allclose(cs.ldot(pos, v), sum(a * h for a, h in zip(v, H)), atol=2e-8)
        np.testing.assert_allclose(cs.rdot(pos, w), np.array([h @ w for h in H]), atol=2e-8)
        # tight: central differences of the analytic Jacobian
        D, x, h = cs.ldot(pos, v), pos.ravel(), 1e-6
        fd = np.array([(v @ cs.jacobian(x + h * e) - v @ cs.jacobian(x - h * e)) / (2 * h) for e in np.eye(3 * n)])
        np.testing.assert_allclose(D, fd, atol=5e-9)


def test_topology_of_a_slab_and_model_hessian():
    at, cons, ints = slab_problem(20, nx=2, ny=2)
    assert ints.ntrans == 48 and ints.natoms == 32
    # every atom of the ideal slab has 6 in-plane + 3 (surface) or 6 (inner) out-of-plane neighbours
    ideal, _, _ = fcc111_slab(2, 2, 4, seed=20, rattle=0.0)
    at0 = _Atoms(ideal, at.cell, at.pbc)
    i0 = Internals(at0)
    i0.find_all_bonds()
    assert i0.nbonds == 168
    lengths = [b.value(ideal, at.cell) for b in i0.internals["bonds"]]
    np.testing.assert_allclose(lengths, 3.61 / np.sqrt(2), atol=1e-9)
    cs, csc, rows = oracle_sets(ints)
    B = cs.jacobian(at.positions)
    assert np.linalg.matrix_rank(B) == 96                 # full column rank: the QR branch of the reference
    np.testing.assert_allclose(np.diag(ints.guess_hessian()), cs.guess_hessian(at.positions), rtol=1e-13)
    np.testing.assert_array_equal(rows, np.arange(48))    # constraint coordinates come first (internal.py:3058)
    # angles and dihedrals of a chain molecule
    chain = _Atoms([[0, 0, 0], [1.5, 0, 0], [2.2, 1.3, 0], [3.7, 1.4, 0.3], [4.3, 2.8, 0.5]], None, (False,) * 3, number=6)
    im = Internals(chain)
    im.find_all_bonds(); im.find_all_angles(); im.find_all_dihedrals()
    assert (im.nbonds, im.nangles, im.ndihedrals) == (4, 3, 2)
    assert im.check_for_bad_internals() is None
    chain.positions[2] = [3.0, 0.01, 0]
    assert im.check_for_bad_internals() is not None


def test_oracle_internal_search_lsoda_and_rk_reach_the_same_saddle():
    at, cons, ints = slab_problem(20)
    cs, csc, rows = oracle_sets(ints)
    out = {}
    for integ in ("lsoda", "rk"):
        p = InternalPES(at.func, at.positions.ravel(), cs, csc, integrator=integ)
        o = SaddleSearch(p, rs="mis")
        assert o.run(2e-4, 400)
        assert np.abs(p.get_res()).max() < 1e-7
        out[integ] = (p.pos.copy(), o.nsteps)
        # a first-order saddle of the constrained surface: one negative mode in the free space
        p.diag(gamma=1e-12)
        Uf = p.get_Ufree()
        ev = np.linalg.eigvalsh(Uf.T @ (p.H.B - p.get_Hc()) @ Uf)
        assert (ev < 0).sum() == 1
    np.testing.assert_allclose(out["lsoda"][0], out["rk"][0], atol=2e-4)


# ----------------------------------------------------------------------------------------------
# the engine's new algebra on torch-CPU stand-ins for the device operators
class _FakeK:
    @staticmethod
    def gemm(A, B, transA=False, transB=False, **kw):
        import torch
        A = A.transpose(-1, -2) if transA else A
        B = B.transpose(-1, -2) if transB else B
        return torch.matmul(A, B).contiguous()

    @staticmethod
    def qr(A, want_q=True, **kw):
        import torch
        Q, R = torch.linalg.qr(A, mode="reduced")
        return (Q.contiguous() if want_q else None), R.contiguous()

    @staticmethod
    def potrf(A, **kw):
        import torch
        return torch.linalg.cholesky(A, upper=True).contiguous(), torch.zeros(A.shape[0], dtype=torch.int32)

    @staticmethod
    def trtri(R, **kw):
        import torch
        return torch.linalg.inv(R).contiguous(), torch.zeros(R.shape[0], dtype=torch.int32)

    @staticmethod
    def eigh(A, evals=None, Vt=None, status=None, **kw):
        import torch
        w, V = torch.linalg.eigh(A)
        if evals is not None:
            evals.copy_(w)
        return w, V.transpose(1, 2).contiguous(), status

    @staticmethod
    def eigvalsh(A, evals=None, status=None, **kw):
        import torch
        w = torch.linalg.eigvalsh(A)
        if evals is not None:
            evals.copy_(w)
        return w

    @staticmethod
    def hv_ld(A, X, Y, nvec, transposed=False, active=None):
        import torch
        Y[:, :nvec] = torch.matmul(X[:, :nvec], A if transposed else A.transpose(1, 2))
        return Y


class _FakeInts:
    def __init__(self, cs):
        self.cs = cs
        self.ntrans, self.nbonds, self.nangles, self.ndihedrals = cs.ntrans, cs.nbonds, cs.nangles, cs.ndihedrals
        self.nstd = self.nint = cs.nint
        self.nrotations = 0

    def calc(self, x, jacobian=False):
        import torch
        q = torch.from_numpy(np.stack([self.cs.calc(p) for p in x.numpy()]))
        if not jacobian:
            return q
        return q, torch.from_numpy(np.stack([self.cs.jacobian(p) for p in x.numpy()]))

    def jacobian(self, x):
        return self.calc(x, jacobian=True)[1]

    def ldot(self, x, v):
        import torch
        return torch.from_numpy(np.stack([self.cs.ldot(p, w) for p, w in zip(x.numpy(), v.numpy())]))

    def rdot(self, x, w):
        import torch
        return torch.from_numpy(np.stack([self.cs.rdot(p, u) for p, u in zip(x.numpy(), w.numpy())]))


def _cpu_engine(monkeypatch, problems):
    """A BatchedInternalSella with only the attributes its Wilson-matrix / geodesic methods use."""
    import torch
    import sella_b200.batched_internal as bi
    monkeypatch.setattr(bi, "K", _FakeK)
    at, cons, ints = problems[0]
    cs, csc, rows = oracle_sets(ints)
    eng = object.__new__(bi.BatchedInternalSella)
    b = len(problems)
    pos = torch.from_numpy(np.stack([p[0].positions.ravel() for p in problems]))
    eng.ints, eng.pos, eng.batch, eng.n, eng.ncart = _FakeInts(cs), pos, b, cs.nint, cs.ndof
    eng.dev = pos.device
    eng.nc, eng.rows = len(rows), torch.from_numpy(rows)
    eng.nfree_int = cs.ndof - len(rows)
    eng.status = torch.zeros(b, dtype=torch.int32)
    eng.dih = (cs.ntrans + cs.nbonds + cs.nangles, cs.nint)
    eng.exact_geodesic = True
    eng.ode_steps = 0
    eng.targets = eng.ints.calc(pos)[:, eng.rows].clone() if eng.nc else None
    sv = np.linalg.svd(cs.jacobian(problems[0][0].positions), compute_uv=False)
    eng.nnull = int((sv <= 1e-6).sum())
    eng.svd_path = eng.nnull > 0
    eng.nfree_int = cs.ndof - len(rows) - eng.nnull
    eng.geo = eng._geometry(pos)
    eng.x = eng.geo["q"].clone()
    return eng, cs, csc


@pytest.mark.parametrize("angles", [False, True, "free", "free+cons"])
def test_engine_algebra_matches_oracle_on_cpu_stand_ins(monkeypatch, angles):
    torch = pytest.importorskip("torch")
    problems = [molecule_problem(s) for s in range(2)] if angles == "free" else \
        [constrained_molecule_problem(s) for s in range(2)] if angles == "free+cons" else \
        [slab_problem(30 + s, angles=angles) for s in range(2)]
    eng, cs, csc = _cpu_engine(monkeypatch, problems)
    b, n, ncart, nc = eng.batch, eng.n, eng.ncart, eng.nc
    rng = np.random.RandomState(5)
    oracles = []
    for at, cons, ints in problems:
        p = InternalPES(at.func, at.positions.ravel(), cs, csc, integrator="rk")
        p.get_g()
        oracles.append(p)
    # engine state: H, g, f as the oracle has them
    eng._B = torch.from_numpy(np.stack([p.H.B for p in oracles]))
    eng.evalsB = torch.from_numpy(np.stack([np.linalg.eigvalsh(p.H.B) for p in oracles]))
    from types import SimpleNamespace
    eng.sp = SimpleNamespace(rb=n, mrows=torch.full((b,), n, dtype=torch.int32))
    eng.g = torch.from_numpy(np.stack([p.curr["g"] for p in oracles]))
    eng._evaluated = True
    eng.evals, eng.Vt = torch.zeros(b, n), torch.zeros(b, n, n)
    eng.evals, eng.Vt = eng.evals.double(), eng.Vt.double()
    # gradient conversion
    gc = torch.from_numpy(np.stack([at.func(at.positions.ravel())[1] for at, _, _ in problems]))
    np.testing.assert_allclose(eng._to_internal(eng.geo, gc).numpy(), eng.g.numpy(), atol=1e-12)
    eng._model()
    geo = eng.geo
    for i, p in enumerate(oracles):
        Uf, Uc = p.get_Ufree(), p.get_Ucons()
        nfree = Uf.shape[1]
        assert nfree == eng.nfree_int
        Hc = p.get_Hc()
        V = rng.normal(size=(3, n))
        if nc:
            np.testing.assert_allclose(geo["L"][i].numpy(), p.curr["L"], atol=1e-10)
            np.testing.assert_allclose(eng._hc_apply(geo, torch.from_numpy(V[None].repeat(b, 0)))[i].numpy(), V @ Hc,
                                       atol=1e-9)
        np.testing.assert_allclose(eng._free_project(geo, torch.from_numpy(V[None].repeat(b, 0)))[i].numpy(),
                                   V @ Uf @ Uf.T, atol=1e-11)
        ref = np.linalg.eigvalsh(Uf.T @ (p.H.B - Hc) @ Uf)
        np.testing.assert_allclose(geo["evr"][i, :nfree].numpy(), ref, atol=1e-9)
        assert (geo["evr"][i, nfree:] > ref.max() + 1).all()
        W = eng.Vt[i, :ncart].numpy()
        np.testing.assert_allclose(W[:nfree] @ W[:nfree].T, np.eye(nfree), atol=1e-11)
        np.testing.assert_allclose(Uf @ Uf.T @ W[:nfree].T, W[:nfree].T, atol=1e-11)       # span = free space
        if nc:
            np.testing.assert_allclose(Uc @ Uc.T @ W[nfree:nfree + nc].T, W[nfree:nfree + nc].T, atol=1e-10)   # = Ucons
        ev2 = np.linalg.eigvalsh(geo["HLr"][i].numpy())
        ref2 = p.get_HL_projected(p.get_Unred()).evals
        np.testing.assert_allclose(ev2[:len(ref2)], ref2, atol=1e-9)      # null directions sit at sigma, above
    # geodesic + projection: the engine's integrator against the oracle's restatement of it
    s = 0.05 * rng.normal(size=(b, n))
    s[1] *= 14.0           # a long step: more integrator iterations than system 0 (which then leaves the working set)
    for i, p in enumerate(oracles):
        Uf = p.get_Ufree()
        s[i] = Uf @ (Uf.T @ s[i])
    pos, g2, dx_i, dx_f, g_par = eng._set_x(eng.x + torch.from_numpy(s))
    nfev = [p.ode_nfev for p in oracles]
    for i, p in enumerate(oracles):
        a, c, d = p.set_x(p.get_x() + s[i])
        np.testing.assert_allclose(pos[i].numpy(), p.pos, atol=1e-10)
        np.testing.assert_allclose(dx_i[i].numpy(), a, atol=1e-12)
        np.testing.assert_allclose(dx_f[i].numpy(), c, atol=1e-9)
        np.testing.assert_allclose(g_par[i].numpy(), d, atol=1e-9)
        nfev[i] = p.ode_nfev - nfev[i]
        # the fixed atoms have stayed where they were
        rr = eng.rows.numpy()
        np.testing.assert_allclose(cs.calc(p.pos)[rr], eng.x[i].numpy()[rr], atol=2e-7)
    assert int(eng.status.max()) == 0
    if angles is not True:                      # (the dihedral-rich cluster takes both steps in one iteration)
        assert nfev[1] > nfev[0] and eng.ode_steps == (nfev[1] - 1) // 6
    if not nc:
        return
    # Newton projection onto the constraint manifold after a displaced start
    posd = eng.pos.clone()
    posd[:, :3] += 1e-4
    g3 = eng._geometry(posd)
    assert float(g3["res"].abs().max()) > 5e-5
    pos4, g4, _ = eng._project_to_constraints(posd, g3, torch.zeros(b, n).double())
    assert float(g4["res"].abs().max()) < 1e-7


# ----------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("method,angles", [("qn", False), ("prfo", False), ("prfo", True), ("prfo", "free")])
def test_cuda_internal_engine_matches_oracle_loop(method, angles):
    """Every step of three searches in internal coordinates on the CUDA engine equals the oracle InternalPES loop
    run with the same integrator: C3-style slabs (bonds, bottom half fixed), clusters with three held atoms and the
    full automatic list (bonds, angles, dihedrals), and FREE clusters (rank-deficient Wilson matrix: the
    reference's SVD branch)."""
    torch = pytest.importorskip("torch")
    from sella_b200.batched_internal import BatchedInternalSella
    from sella_b200.emt import EMTSurface
    dev = torch.device("cuda:0")
    problems = [molecule_problem(s) for s in range(3)] if angles == "free" else \
        [slab_problem(40 + s, angles=angles) for s in range(3)]
    at, cons, ints = problems[0]
    cs, csc, rows = oracle_sets(ints)
    x0 = np.stack([p[0].positions.ravel() for p in problems])
    surf = EMTSurface(len(problems), ints.natoms, dev, cell=at.cell, pbc=tuple(at.pbc))
    kw = dict(method=method, diag_maxiter=6)
    h0 = np.stack([cs.guess_hessian(a.positions) for a, _, _ in problems])      # per system: it depends on the geometry
    eng = BatchedInternalSella(surf, torch.from_numpy(x0).to(dev), ints.device_coordinates(), cons_rows=rows,
                               h0=h0, **kw)
    oracles = []
    for a, _, _ in problems:
        p = InternalPES(a.func, a.positions.ravel(), cs, csc, integrator="rk")
        oracles.append((p, SaddleSearch(p, rs="mis", **kw)))
    for t in range(6):
        eng.step()
        pos, delta = eng.pos.cpu().numpy(), eng.delta.cpu().numpy()
        for i, (p, o) in enumerate(oracles):
            o.step()
            np.testing.assert_allclose(pos[i], p.pos, rtol=0, atol=2e-7, err_msg="system %d step %d" % (i, t))
            np.testing.assert_allclose(delta[i], o.delta, rtol=1e-6)
    eng.check_status()
    Hd = eng.B.cpu().numpy()
    mrows = eng.sp.mrows.cpu().numpy()
    for i, (p, o) in enumerate(oracles):
        np.testing.assert_allclose(Hd[i], p.H.B, rtol=1e-5, atol=1e-5)
        # the carried eigenpairs of H (explicit rows + the eigenvalue 0 on their complement) are those of the
        # dense copy that sb_update_apply carries independently
        m = int(mrows[i])
        th, VR = eng.evalsB[i, :m].cpu().numpy(), eng.VtB[i, :m].cpu().numpy()
        np.testing.assert_allclose(VR @ VR.T, np.eye(m), atol=1e-10)
        np.testing.assert_allclose(VR.T @ (th[:, None] * VR), Hd[i], atol=1e-8 * max(1.0, np.abs(th).max()))
    conv = eng.converged(1e-3).cpu().numpy()
    fm = eng.fmax.cpu().numpy()
    for i, (p, o) in enumerate(oracles):
        c, f1, c1 = p.converged(1e-3)
        assert bool(conv[i]) == bool(c)
        np.testing.assert_allclose(fm[i], f1, rtol=1e-5)


@pytest.mark.gpu
def test_sella_internal_true_on_a_slab():
    """`Sella(slab, internal=ints, constraints via Internals)` (BASELINE config C3 as named, small): converges to
    the saddle the oracle loop with the reference's LSODA integrator finds (1e-6 Angstrom is the north star; the
    two integrators differ by their tolerances, so the saddle is compared at the convergence threshold)."""
    pytest.importorskip("torch")
    from sella_b200 import Sella
    at, cons, ints = slab_problem(20)
    cs, csc, rows = oracle_sets(ints)
    ref = InternalPES(at.func, at.positions.ravel(), cs, csc, integrator="lsoda")
    o = SaddleSearch(ref, rs="mis")
    assert o.run(2e-4, 400)
    dyn = Sella(at, internal=ints, logfile=None)
    assert dyn.run(2e-4, 400)
    np.testing.assert_allclose(at.positions.ravel(), ref.pos, rtol=0, atol=2e-4)
    assert abs(dyn.nsteps - o.nsteps) <= max(3, o.nsteps // 4)
    fixed = np.nonzero(ref.pos.reshape(-1, 3)[:, 2] < ref.pos.reshape(-1, 3)[:, 2].mean())[0]
    assert dyn.pes.int is not None and dyn.pes.dim == ints.nint


@pytest.mark.gpu
def test_sella_internal_true_on_a_free_cluster():
    """`Sella(cluster, internal=True, order=0)`: automatic coordinate list, rank-deficient Wilson matrix, a
    minimisation without Davidson (eig=False); ends at a minimum of the surface with the rigid-body motions
    untouched (no net force or torque enters the internal gradient)."""
    pytest.importorskip("torch")
    from sella_b200 import Sella
    at, cons, ints = molecule_problem(0)
    e0 = at.get_potential_energy()
    dyn = Sella(at, internal=True, order=0, logfile=None)
    assert dyn.pes.int.nbonds > 0 and dyn.pes.int.nangles > 0
    assert dyn.run(1e-3, 300)
    assert at.get_potential_energy() < e0
    assert np.linalg.norm(at.get_forces(), axis=1).max() < 1e-3


# ----------------------------------------------------------------------------------------------
# more of the oracle's InternalPES (CPU): the variants of set_x the reference offers lead to the same place
def test_oracle_set_x_variants_agree():
    """peswrapper.py:749-903: the Newton ("iterative") stepper, the geodesic with the exact B+ and the geodesic with
    B+ frozen at the start reach the same geometry for a moderate step (to the integrators' tolerances), keep the
    constrained coordinates, and report consistent (dx_initial, dx_final, g_par)."""
    at, cons, ints = slab_problem(31)
    cs, csc, rows = oracle_sets(ints)
    rng = np.random.RandomState(2)
    out = {}
    for name, kw in (("ode", {}), ("frozen", dict(exact_geodesic=False)), ("newton", dict(iterative_stepper=1)),
                     ("rk", dict(integrator="rk"))):
        p = InternalPES(at.func, at.positions.ravel(), cs, csc, **kw)
        p.get_g()
        Uf = p.get_Ufree()
        s = Uf @ (Uf.T @ (0.03 * np.random.RandomState(2).normal(size=cs.nint)))
        x0 = p.get_x()
        dx_i, dx_f, g_par = p.set_x(x0 + s)
        out[name] = (p.pos.copy(), dx_i, dx_f, g_par, p.get_x() - x0)
        np.testing.assert_allclose(cs.calc(p.pos)[:len(rows)], x0[:len(rows)], atol=1e-6)     # fixed atoms stay
        np.testing.assert_allclose(dx_i, s, atol=1e-12)
        # the realised change of the NON-redundant part of q equals the requested one to second order
        Q = p._jacobian_qr()[0]
        assert np.linalg.norm(Q.T @ ((p.get_x() - x0) - s)) < 0.05 * np.linalg.norm(s)
    for name in ("frozen", "newton", "rk"):
        np.testing.assert_allclose(out[name][0], out["ode"][0], atol=5e-4, err_msg=name)
    np.testing.assert_allclose(out["rk"][0], out["ode"][0], atol=2e-5)
    np.testing.assert_allclose(out["rk"][3], out["ode"][3], atol=1e-3)                         # transported gradient


def test_oracle_constraint_projection_and_dihedral_unwrapping():
    """_project_to_constraints (peswrapper.py:928-994) pulls a displaced constrained coordinate back without
    touching the free internal coordinates to first order; get_x continues dihedrals across +-pi (:996-1008)."""
    at, cons, ints = cluster_problem(31)
    cs, csc, rows = oracle_sets(ints)
    p = InternalPES(at.func, at.positions.ravel(), cs, csc)
    p.get_g()
    q0 = p.get_x()
    p.pos = p.pos.copy()
    p.pos[:3] += np.array([2e-3, -1e-3, 1.5e-3])             # atom 0 is held: residual 2e-3
    assert np.abs(p.get_res()).max() > 1e-3
    q1 = cs.calc(p.pos)
    Uf = p._calc_basis()[3]
    assert p._project_to_constraints()
    assert np.abs(p.get_res()).max() < 1e-7
    # the correction lives in the constraint subspace: the free internal coordinates of the displaced geometry
    # are unchanged to first order in the residual (2e-3)
    moved = cs.wrap(cs.calc(p.pos) - q1)
    assert np.linalg.norm(Uf.T @ moved) < 0.05 * np.linalg.norm(moved)
    # a step larger than the safety limit is refused (the optimiser's step must not be overridden)
    p.pos[:3] += 0.2
    before = p.pos.copy()
    assert not p._project_to_constraints()
    np.testing.assert_array_equal(p.pos, before)
    # dihedral continuation
    lo = cs.ntrans + cs.nbonds + cs.nangles
    p2 = InternalPES(at.func, at.positions.ravel(), cs, csc)
    p2.get_g()
    x = p2.get_x()
    k = lo + int(np.argmax(np.abs(x[lo:])))                   # the dihedral closest to +-pi
    p2.curr["x"] = x.copy()
    p2.curr["x"][k] = x[k] + 2 * np.pi * np.sign(x[k]) - 1e-3 * np.sign(x[k])     # "previous" value just across the cut
    assert abs(p2.get_x()[k] - p2.curr["x"][k]) < 2e-3
