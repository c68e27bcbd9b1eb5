"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, loaded piecewise by oracle/ref_loader.py) on seeded inputs.

Run in the build container only:   python tests/golden/make_golden.py
The fixtures are committed; tests never need /root/reference at run time.

Every file stores inputs *and* the reference's outputs, so both the CPU oracle
(``-m "not gpu"``) and the CUDA path (``-m gpu``) are checked on identical data.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, ref_harness          # noqa: E402
from sella_b200.synthetic import quadratic_system, quadratic_func   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def sym(n, rng, pd=False):
    """Same construction as the reference's tests/test_utils/matrix_factory.py:3-15."""
    A = rng.normal(size=(n, n))
    A = 0.5 * (A + A.T)
    if pd:
        w, v = np.linalg.eigh(A)
        A = v @ (np.abs(w)[:, None] * v.T)
    return A


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name, {k: np.shape(v) for k, v in arrays.items()})


def golden_mgs(ref):
    rng = np.random.RandomState(11)
    out = {}
    n = 37
    cases = []
    # 0: plain X vs Y
    cases.append((rng.normal(size=(n, 3)), rng.normal(size=(n, 4)), 1e-15, 1e-6, 100))
    # 1: no Y
    cases.append((rng.normal(size=(n, 5)), None, 1e-15, 1e-6, 100))
    # 2: duplicate column -> rank drop
    X = rng.normal(size=(n, 4)); X[:, 2] = X[:, 0]
    cases.append((X, rng.normal(size=(n, 2)), 1e-15, 1e-6, 100))
    # 3: X column inside span(Y) -> dropped
    Y = rng.normal(size=(n, 3)); X = rng.normal(size=(n, 2)); X[:, 0] = Y @ [1., -2., .5]
    cases.append((X, Y, 1e-15, 1e-6, 100))
    # 4: Davidson shape: one vector vs k=5 (nearly dependent: small new component)
    Y = rng.normal(size=(384, 5)); X = (Y @ rng.normal(size=(5, 1))) + 1e-3 * rng.normal(size=(384, 1))
    cases.append((X, Y, 1e-15, 1e-6, 100))
    # 5: eps2 huge -> everything dropped
    cases.append((rng.normal(size=(n, 2)), None, 1e-15, 1e10, 100))
    for i, (X, Y, e1, e2, mi) in enumerate(cases):
        res = ref.math.modified_gram_schmidt(X, Y, eps1=e1, eps2=e2, maxiter=mi)
        out["X%d" % i] = X
        out["Y%d" % i] = np.zeros((X.shape[0], 0)) if Y is None else Y
        out["hasY%d" % i] = np.array(Y is not None)
        out["par%d" % i] = np.array([e1, e2, mi])
        out["out%d" % i] = res
    out["ncases"] = np.array(len(cases))
    save("mgs", **out)


def golden_symmetrize(ref):
    rng = np.random.RandomState(12)
    out = {}
    i = 0
    for n, k in [(20, 1), (20, 2), (20, 4), (384, 5)]:
        S = rng.normal(size=(n, k))
        H = sym(n, rng)
        Y = H @ S + 1e-3 * rng.normal(size=(n, k))      # slightly non-symmetric S^T Y
        for symm in (0, 1, 2):
            out["S%d" % i], out["Y%d" % i] = S, Y
            out["symm%d" % i] = np.array(symm)
            out["out%d" % i] = ref.hessian_update.symmetrize_Y(S, Y, symm)
            i += 1
    out["ncases"] = np.array(i)
    save("symmetrize_Y", **out)


def golden_update_H(ref):
    rng = np.random.RandomState(13)
    out = {}
    i = 0
    grp = 0
    methods = ["TS-BFGS", "PSB", "Greenstadt", "BFGS", "DFP", "SR1", "BFGS_auto"]
    for n, k in [(24, 1), (24, 2), (24, 3), (96, 1), (96, 2)]:
        for pd in ((False, True) if n < 50 else (False,)):
            B = sym(n, rng, pd)
            H = sym(n, rng, pd)
            S = rng.normal(size=(n, k))
            Y = H @ S
            out["B_g%d" % grp], out["S_g%d" % grp], out["Y_g%d" % grp] = B, S, Y
            for method in methods:
                if n > 50 and method not in ("TS-BFGS", "PSB", "Greenstadt"):
                    continue
                for symm in ((0, 1, 2) if (k > 1 and n < 50) else (2,)):
                    for useB in (True, False):
                        Bin = B if useB else None
                        flat = (k == 1 and i % 2 == 1)
                        Sin, Yin = (S.ravel(), Y.ravel()) if flat else (S, Y)
                        res = ref.hessian_update.update_H(Bin, Sin, Yin, method=method, symm=symm)
                        out["meta%d" % i] = np.array([str(grp), method, str(symm), str(int(useB)), str(int(flat))])
                        out["out%d" % i] = res
                        i += 1
            grp += 1
    # tiny step: must hand back B itself (reference tests/test_hessian_update.py:43-45)
    out["ncases"] = np.array(i)
    save("update_H", **out)


def golden_rayleigh_ritz(ref):
    rng = np.random.RandomState(14)
    out = {}
    i = 0
    for n in (40, 96):
        A = sym(n, rng)
        w, v = np.linalg.eigh(A)
        w = np.abs(w) + 0.1
        w[:2] = [-1.3, -0.4]
        A = v @ (w[:, None] * v.T)
        A = 0.5 * (A + A.T)
        P = A + 0.05 * sym(n, rng)
        v0 = rng.normal(size=n)
        out["A_%d" % n], out["P_%d" % n], out["v0_%d" % n] = A, P, v0
        methods = ["jd0", "lanczos", "gd", "jd0_alt", "mjd0", "mjd0_alt"] if n < 100 else ["jd0", "lanczos"]
        for method in methods:
            for gamma, maxiter, use_v0 in [(0.1, None, True), (0.1, None, False),
                                           (1e-32, 3, True), (0.1, 5, True),
                                           (1e-3, None, True), (0.1, 2, True)]:
                lams, V, AV = ref.eigensolvers.rayleigh_ritz(
                    A, gamma, P, v0=v0 if use_v0 else None, method=method, maxiter=maxiter)
                out["meta%d" % i] = np.array([str(n), method, repr(gamma), str(maxiter), str(int(use_v0))])
                out["lams%d" % i], out["V%d" % i], out["AV%d" % i] = lams, V, AV
                i += 1
    out["ncases"] = np.array(i)
    save("rayleigh_ritz", **out)


def golden_fd_hessian(ref):
    """NumericalHessian._matvec incl. the canonical-sign rule and Uproj."""
    rng = np.random.RandomState(15)
    n = 30
    A, xs, x0 = quadratic_system(3, n)
    func = quadratic_func(A, xs)
    _, g0 = func(x0)
    out = dict(A=A, xstar=xs, x0=x0, g0=g0)
    U, _ = np.linalg.qr(rng.normal(size=(n, n - 4)))
    out["U"] = U
    i = 0
    for threepoint in (False, True):
        for useU in (False, True):
            H = ref.linalg.NumericalHessian(func, x0, g0, 1e-4, threepoint, U if useU else None)
            m = U.shape[1] if useU else n
            vs = [rng.normal(size=m), -rng.normal(size=m) * 3.0, np.zeros(m)]
            # vector orthogonal to g0 and x0 (sign falls through to first element)
            if not useU:
                q, _ = np.linalg.qr(np.column_stack([g0, x0, rng.normal(size=n)]))
                vs.append(q[:, 2] * (-1 if q[np.argmax(np.abs(q[:, 2]) > 1e-4), 2] > 0 else 1))
            for v in vs:
                out["v%d" % i] = v
                out["meta%d" % i] = np.array([int(threepoint), int(useU)])
                out["out%d" % i] = H.dot(v)
                i += 1
    out["ncases"] = np.array(i)
    save("fd_hessian", **out)


class _DuckPES:
    """what BaseRestrictedStep reads from a PES."""
    int = None
    n_cell_dof = 0

    def __init__(self, ref, g, B, Ufree, scons):
        self.ref, self.g, self.B, self.Ufree, self.scons = ref, g, B, Ufree, scons
        n = len(g)
        self.H = ref.linalg.ApproximateHessian(n, n, B.copy())

    def get_g(self): return self.g.copy()
    def get_scons(self): return self.scons.copy()
    def get_H(self): return self.H
    def get_Ufree(self): return self.Ufree
    def get_Unred(self): return np.eye(len(self.g))
    def get_HL_projected(self, U): return self.H.project(U)


def golden_restricted(ref):
    rng = np.random.RandomState(16)
    out = {}
    i = 0
    for n in (30, 96):
        B = sym(n, rng)
        w, v = np.linalg.eigh(B)
        w = np.abs(w) + 0.05; w[0] = -0.8
        B = v @ (w[:, None] * v.T); B = 0.5 * (B + B.T)
        g = rng.normal(size=n)
        # free space: fix the first two atoms (6 coordinates)
        Ufree_c = np.eye(n)[:, 6:]
        for cons_case, (Ufree, scons) in enumerate([(np.eye(n), np.zeros(n)),
                                                    (Ufree_c, np.concatenate([1e-3 * rng.normal(size=6), np.zeros(n - 6)]))]):
            out["B_%d_%d" % (n, cons_case)] = B
            out["g_%d_%d" % (n, cons_case)] = g
            out["Ufree_%d_%d" % (n, cons_case)] = Ufree
            out["scons_%d_%d" % (n, cons_case)] = scons
            for rs in ("tr", "ras"):
                for method in ("qn", "rfo", "prfo"):
                    for order in (0, 1, 2):
                        for delta in (0.05, 0.5, 50.0):
                            if n > 50 and (order == 2 or method == "rfo"):
                                continue
                            pes = _DuckPES(ref, g, B, Ufree, scons)
                            cls = ref.restricted_step.get_restricted_step(rs)
                            obj = cls(pes, order, delta, method=method)
                            s, smag = obj.get_s()
                            out["meta%d" % i] = np.array([str(n), str(cons_case), rs, method, str(order), repr(delta)])
                            out["s%d" % i] = s
                            out["smag%d" % i] = np.array(smag)
                            i += 1
    out["ncases"] = np.array(i)
    save("restricted_step", **out)


def golden_steppers(ref):
    rng = np.random.RandomState(17)
    out = {}
    n = 20
    B = sym(n, rng)
    g = rng.normal(size=n)
    out["B"], out["g"] = B, g
    i = 0
    for name in ("qn", "rfo", "prfo"):
        for order in (0, 1, 2):
            H = ref.linalg.ApproximateHessian(n, 0, B.copy())
            st = ref.stepper.get_stepper(name)(g, H, order)
            for alpha in (0.0, 0.3, 0.9, 1.0, 7.5) if name == "qn" else (0.05, 0.3, 0.9, 1.0):
                s, dsda = st.get_s(alpha)
                out["meta%d" % i] = np.array([name, str(order), repr(alpha)])
                out["s%d" % i], out["dsda%d" % i] = s, dsda
                i += 1
    out["ncases"] = np.array(i)
    save("steppers", **out)


def golden_loop(ref):
    """The reference's own Sella.step / PES.kick / PES.diag on quadratic surfaces."""
    out = {}
    i = 0
    for n, b in [(30, 0), (48, 1), (96, 2)]:
        A, xs, x0 = quadratic_system(b, n)
        func = quadratic_func(A, xs)
        # constraint variants: none / first two atoms fixed at their start position
        Cfix = np.eye(n)[:6]
        for cons_case, (C, c) in enumerate([(None, None), (Cfix, Cfix @ x0)]):
            for method, rs in [("qn", "tr"), ("qn", "ras"), ("prfo", "ras"), ("rfo", "tr")]:
                if n > 50 and method == "rfo":
                    continue
                for kw in (dict(), dict(nsteps_per_diag=1), dict(gamma=1e-3, delta0=0.05)):
                    if n > 50 and kw:
                        continue
                    dyn = ref_harness.make_reference_sella(ref, func, x0, C, c, method=method, rs=rs, **kw)
                    nsteps = 15
                    X = np.empty((nsteps, n)); D = np.empty(nsteps); R = np.empty(nsteps)
                    F = np.empty(nsteps); NE = np.empty(nsteps)
                    for t in range(nsteps):
                        dyn.step()
                        X[t] = dyn.pes.get_x(); D[t] = dyn.delta; R[t] = dyn.rho
                        F[t] = dyn.pes.get_f(); NE[t] = dyn.pes.neval
                    out["meta%d" % i] = np.array([str(n), str(b), str(cons_case), method, rs, repr(sorted(kw.items()))])
                    out["x%d" % i], out["delta%d" % i], out["rho%d" % i] = X, D, R
                    out["f%d" % i], out["neval%d" % i] = F, NE
                    out["B%d" % i] = dyn.pes.H.B
                    i += 1
    out["ncases"] = np.array(i)
    save("loop", **out)


def reference_rotation_functions():
    """The rotation coordinate of sella/internal.py is pure numpy but lives in a module that imports
    jax and ase at the top: take the function definitions (unmodified) out of the source."""
    import ast
    src = open("/root/reference/sella/internal.py").read()
    want = {"_build_F_matrix_np", "_stabilize_quaternion", "_stabilize_quaternion_from_eigh", "_asinc_np",
            "_expmap_np", "_rotation_3axis_jacobian_np", "_apply_dF", "_rotation_hessian_single"}
    ns = {"np": np}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module([node], []), "sella/internal.py", "exec"), ns)
    assert want <= set(ns)
    return ns


def golden_rotation():
    fn = reference_rotation_functions()
    rng = np.random.RandomState(21)
    out = {}
    i = 0
    for N in (3, 7, 20):
        ref = rng.normal(size=(N, 3)) * 1.5
        ref -= ref.mean(0)
        for amp in (0.0, 1e-6, 2e-4, 1e-3, 0.05, 0.4, 1.5):
            th = amp * np.array([0.3, -0.5, 0.8])
            ang = np.linalg.norm(th)
            if ang > 0:
                k = th / ang
                Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
                Rm = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
            else:
                Rm = np.eye(3)
            pos = ref @ Rm.T + (0.05 * rng.normal(size=(N, 3)) if amp > 0 else 0.0) + rng.normal(size=3)
            F = fn["_build_F_matrix_np"](pos - pos.mean(0), ref)
            q = fn["_stabilize_quaternion"](F, None)
            out["ref%d" % i], out["pos%d" % i], out["q%d" % i] = ref, pos, q
            out["val%d" % i] = fn["_expmap_np"](q)
            out["jac%d" % i] = fn["_rotation_3axis_jacobian_np"](pos, ref, q).reshape(3, -1)
            out["hess%d" % i] = np.stack([fn["_rotation_hessian_single"](pos, k, ref, q_stable=q).reshape(3 * N, 3 * N)
                                          for k in range(3)])
            i += 1
    out["ncases"] = np.array(i)
    save("rotation", **out)


def golden_sparse_hessians(ref):
    """Assembly of per-coordinate second derivatives by the reference's own SparseInternalHessians.ldot / rdot
    (sella/linalg.py:540-646; importable without JAX).  The per-coordinate (m,3,m,3) blocks fed to it are
    the oracle's (oracle/internals.py: the reference obtains them from JAX, which is not installed), so this
    pins the scatter / contraction logic of `ldot` and `rdot` -- everything but the derivative values."""
    from oracle import internals as oi
    rng = np.random.RandomState(31)
    natoms = 9
    pos = rng.normal(size=(natoms, 3)) * 1.2 + np.arange(natoms)[:, None] * np.array([0.9, 0.2, -0.1])
    bonds = [(i, i + 1) for i in range(natoms - 1)] + [(0, 5)]
    angles = [(i, i + 1, i + 2) for i in range(natoms - 2)]
    diheds = [(i, i + 1, i + 2, i + 3) for i in range(natoms - 3)]
    q, B, H = oi.evaluate(pos, (), bonds, angles, diheds)
    coords = bonds + angles + diheds
    hess = []
    blocks = []
    for atoms, Hd in zip(coords, H):
        m = len(atoms)
        vals = np.zeros((m, 3, m, 3))
        for ia, a in enumerate(atoms):
            for ja, a2 in enumerate(atoms):
                vals[ia, :, ja, :] = Hd[3 * a:3 * a + 3, 3 * a2:3 * a2 + 3]
        hess.append(ref.linalg.SparseInternalHessian(natoms, list(atoms), vals))
        blocks.append(vals)
    D = ref.linalg.SparseInternalHessians(hess, 3 * natoms)
    v = rng.normal(size=len(coords))
    w = rng.normal(size=3 * natoms)
    out = dict(pos=pos, bonds=np.array(bonds), angles=np.array(angles), dihedrals=np.array(diheds), v=v, w=w,
               ldot=D.ldot(v), rdot=D.rdot(w), ddot=D.ddot(w, w))
    for i, b in enumerate(blocks):
        out["vals%d" % i] = b
    save("sparse_hessians", **out)


def golden_internal_loop(ref):
    """The reference's own InternalPES + MaxInternalStep + Sella.step (peswrapper.py:609-1288,
    restricted_step.py:186-243, optimize.py:317-440) on the EMT-form surface, driven through
    oracle/ref_internal_harness.py: coordinate values and derivatives come from oracle.intcoords (JAX is absent),
    every line of the internal-coordinate search itself is the reference's."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_internal_pes import (slab_problem, cluster_problem, molecule_problem, constrained_molecule_problem,
                                   oracle_sets)
    from oracle.ref_internal_harness import make_reference_internal_sella
    cases = [("slab", slab_problem(40), dict(method="prfo"), 8),
             ("slab", slab_problem(41), dict(method="qn"), 8),
             ("slab", slab_problem(42), dict(method="prfo", exact_geodesic=False), 8),
             ("slab", slab_problem(43), dict(method="prfo", iterative_stepper=1), 8),
             ("cluster", cluster_problem(31), dict(method="prfo"), 6),
             ("free", molecule_problem(0), dict(method="prfo"), 6),
             ("free", molecule_problem(1), dict(order=0), 10),
             ("free+bond+angle", constrained_molecule_problem(), dict(method="prfo"), 6)]
    out = {}
    for i, (kind, (at, cons, ints), kw, nsteps) in enumerate(cases):
        cs, csc, rows = oracle_sets(ints)
        tr, bd, an, dh, tv = ints.lists()
        dyn = make_reference_internal_sella(ref, at.func, at.positions, cs, csc, cell=at.cell, pbc=at.pbc, **kw)
        dyn.diagkwargs["maxiter"] = 6
        n = at.positions.size
        X = np.empty((nsteps, n)); D = np.empty(nsteps); R = np.empty(nsteps); F = np.empty(nsteps)
        for t in range(nsteps):
            dyn.step()
            X[t] = dyn.pes.atoms.positions.ravel(); D[t] = dyn.delta; R[t] = dyn.rho; F[t] = dyn.pes.get_f()
        out["meta%d" % i] = np.array([kind, repr(sorted(kw.items()))])
        out["pos0_%d" % i] = np.array(at.positions)       # (the harness works on its own copy)
        out["numbers%d" % i] = at.numbers
        out["cell%d" % i] = np.zeros((0, 3)) if at.cell is None else np.asarray(at.cell)
        out["pbc%d" % i] = np.asarray(at.pbc)
        out["trans%d" % i] = np.asarray(tr, dtype=int).reshape(-1, 2)
        out["bonds%d" % i] = np.asarray(bd, dtype=int).reshape(-1, 2)
        out["angles%d" % i] = np.asarray(an, dtype=int).reshape(-1, 3)
        out["diheds%d" % i] = np.asarray(dh, dtype=int).reshape(-1, 4)
        for k in ("bonds", "angles", "dihedrals"):
            out["tv_%s%d" % (k, i)] = tv[k]
        out["rows%d" % i] = rows
        out["x%d" % i], out["delta%d" % i], out["rho%d" % i], out["f%d" % i] = X, D, R, F
        out["H%d" % i] = dyn.pes.H.B
        out["neval%d" % i] = np.array(dyn.pes.neval)
    out["ncases"] = np.array(len(cases))
    save("internal_loop", **out)


def main():
    if "--internal-only" in sys.argv:
        return golden_internal_loop(ref_loader.load())
    golden_rotation()
    if "--rotation-only" in sys.argv:
        return
    ref = ref_loader.load()
    golden_mgs(ref)
    golden_symmetrize(ref)
    golden_update_H(ref)
    golden_rayleigh_ritz(ref)
    golden_fd_hessian(ref)
    golden_steppers(ref)
    golden_restricted(ref)
    golden_loop(ref)
    golden_sparse_hessians(ref)
    golden_internal_loop(ref)


if __name__ == "__main__":
    main()
