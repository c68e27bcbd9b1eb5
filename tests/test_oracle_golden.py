"""CPU: the oracle restatement (oracle/*.py) against the committed outputs of
the reference itself (tests/golden/*.npz, produced by make_golden.py)."""
import numpy as np
import pytest

from oracle import orth, hessian, davidson, stepper, restricted, pes, driver
from sella_b200.synthetic import quadratic_system, quadratic_func

TIGHT = dict(rtol=1e-11, atol=1e-12)


def subspace_gap(V1, V2):
    """sin of the largest principal angle between two orthonormal column sets."""
    if V1.shape != V2.shape:
        return np.inf
    return np.linalg.norm(V1 - V2 @ (V2.T @ V1), 2)


def test_mgs(golden):
    G = golden("mgs")
    for i in range(int(G["ncases"])):
        X, Y = G["X%d" % i], G["Y%d" % i]
        e1, e2, mi = G["par%d" % i]
        out = orth.modified_gram_schmidt(X, Y if bool(G["hasY%d" % i]) else None,
                                         eps1=e1, eps2=e2, maxiter=int(mi))
        ref = G["out%d" % i]
        assert out.shape == ref.shape, i
        np.testing.assert_allclose(out, ref, **TIGHT)


def test_mgs_error_codes():
    # reference tests/utilities/test_math.py:144-165: maxiter=1 -> failure,
    # mismatched Y -> failure code
    rng = np.random.RandomState(0)
    X = rng.normal(size=(10, 3))
    assert orth.mgs(X.copy(), None, maxiter=1) == orth.MGS_MAXITER
    assert orth.mgs(X.copy(), rng.normal(size=(9, 2))) == orth.MGS_SHAPE_MISMATCH
    with pytest.raises(RuntimeError):
        orth.modified_gram_schmidt(X, None, maxiter=1)


def test_symmetrize_Y(golden):
    G = golden("symmetrize_Y")
    for i in range(int(G["ncases"])):
        out = hessian.symmetrize_Y(G["S%d" % i], G["Y%d" % i], int(G["symm%d" % i]))
        np.testing.assert_allclose(out, G["out%d" % i], rtol=1e-10, atol=1e-11)


def test_update_H(golden):
    G = golden("update_H")
    for i in range(int(G["ncases"])):
        grp, method, symm, useB, flat = G["meta%d" % i]
        B, S, Y = G["B_g" + grp], G["S_g" + grp], G["Y_g" + grp]
        if flat == "1":
            S, Y = S.ravel(), Y.ravel()
        out = hessian.update_H(B if useB == "1" else None, S, Y, method=method, symm=int(symm))
        ref = G["out%d" % i]
        scale = np.abs(ref).max()
        np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-10 * scale, err_msg=str(G["meta%d" % i]))
        # reference invariant tests/test_hessian_update.py:33-37 (secant condition)
        S2 = S.reshape(len(S), -1)
        Yt = hessian.symmetrize_Y(S2, Y.reshape(len(S), -1), int(symm))
        np.testing.assert_allclose(out @ S2, Yt, rtol=1e-6, atol=1e-6 * scale)


def test_update_H_tiny_step_is_identity():
    # reference tests/test_hessian_update.py:43-45
    rng = np.random.RandomState(1)
    B = rng.normal(size=(10, 10)); B = B + B.T
    s = rng.normal(size=10) / 1e12
    assert hessian.update_H(B, s, B @ s) is B


def test_rayleigh_ritz(golden):
    G = golden("rayleigh_ritz")
    for i in range(int(G["ncases"])):
        n, method, gamma, maxiter, use_v0 = G["meta%d" % i]
        A, P, v0 = G["A_" + n], G["P_" + n], G["v0_" + n]
        lams, V, AV = davidson.rayleigh_ritz(
            A, float(gamma), P, v0=v0 if use_v0 == "1" else None, method=method,
            maxiter=None if maxiter == "None" else int(maxiter))
        ref = G["lams%d" % i]
        if len(ref) <= 8:
            # the regime Sella runs in (a handful of expansions per diagonalisation)
            assert lams.shape == ref.shape, G["meta%d" % i]
            np.testing.assert_allclose(lams, ref, rtol=1e-10, atol=1e-11)
            assert subspace_gap(V, G["V%d" % i]) < 1e-9
        else:
            # dozens of expansions with a nearly singular correction equation
            # amplify round-off (even reference-vs-restatement differ at 1e-5
            # in the *unconverged* Ritz values); the converged target is stable
            np.testing.assert_allclose(lams[0], ref[0], rtol=1e-3)
        # reference invariant tests/test_eigensolvers.py:67
        np.testing.assert_allclose(lams, np.linalg.eigh(V.T @ AV)[0], atol=1e-4)


def test_fd_hessian(golden):
    G = golden("fd_hessian")
    func = quadratic_func(G["A"], G["xstar"])
    for i in range(int(G["ncases"])):
        threepoint, useU = G["meta%d" % i]
        H = pes.FiniteDifferenceHessian(func, G["x0"], G["g0"], 1e-4, bool(threepoint),
                                        G["U"] if useU else None)
        np.testing.assert_allclose(H.dot(G["v%d" % i]), G["out%d" % i], rtol=1e-10, atol=1e-12)


def test_steppers(golden):
    G = golden("steppers")
    B, g = G["B"], G["g"]
    n = len(g)
    for i in range(int(G["ncases"])):
        name, order, alpha = G["meta%d" % i]
        H = pes.ApproxHessian(n, 0, B.copy())
        st = stepper.get_stepper(name)(g, H, int(order))
        s, dsda = st.get_s(float(alpha))
        np.testing.assert_allclose(s, G["s%d" % i], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(dsda, G["dsda%d" % i], rtol=1e-7, atol=1e-9)


class _Duck:
    int = None
    n_cell_dof = 0

    def __init__(self, g, B, Ufree, scons):
        self.g, self.Ufree, self.scons = g, Ufree, scons
        self.H = pes.ApproxHessian(len(g), len(g), B.copy())

    def get_g(self): return self.g.copy()
    def get_scons(self): return self.scons.copy()
    def get_H(self): return self.H
    def get_Ufree(self): return self.Ufree
    def get_Unred(self): return np.eye(len(self.g))
    def get_HL_projected(self, U): return self.H.project(U)


def test_restricted_step(golden):
    G = golden("restricted_step")
    for i in range(int(G["ncases"])):
        n, cc, rs, method, order, delta = G["meta%d" % i]
        key = "_%s_%s" % (n, cc)
        duck = _Duck(G["g" + key], G["B" + key], G["Ufree" + key], G["scons" + key])
        obj = restricted.get_restricted_step(rs)(duck, int(order), float(delta), method=method)
        s, smag = obj.get_s()
        np.testing.assert_allclose(smag, float(G["smag%d" % i]), rtol=1e-12)
        np.testing.assert_allclose(s, G["s%d" % i], rtol=1e-8, atol=1e-10, err_msg=str(G["meta%d" % i]))


def test_loop_matches_reference_sella(golden):
    """oracle.driver.SaddleSearch + oracle.pes.CartesianPES reproduce the
    reference's own Sella.step/PES.kick/PES.diag trajectories."""
    G = golden("loop")
    for i in range(int(G["ncases"])):
        n, b, cc, method, rs, kw = G["meta%d" % i]
        n, b = int(n), int(b)
        A, xs, x0 = quadratic_system(b, n)
        func = quadratic_func(A, xs)
        C = c = None
        if cc == "1":
            C = np.eye(n)[:6]; c = C @ x0
        p = pes.CartesianPES(func, x0, C, c)
        dyn = driver.SaddleSearch(p, method=method, rs=rs, **dict(eval(kw)))
        X = G["x%d" % i]
        for t in range(X.shape[0]):
            dyn.step()
            np.testing.assert_allclose(p.get_x(), X[t], rtol=0, atol=1e-9, err_msg="%s step %d" % (G["meta%d" % i], t))
            np.testing.assert_allclose(dyn.delta, G["delta%d" % i][t], rtol=1e-9)
            assert p.neval == int(G["neval%d" % i][t])
        np.testing.assert_allclose(p.H.B, G["B%d" % i], rtol=1e-7, atol=1e-8)


def test_rotation_coordinate(golden):
    """oracle/rotation.py against the reference's own rotation functions (internal.py:507-800)."""
    from oracle import rotation as orot
    G = golden("rotation")
    for i in range(int(G["ncases"])):
        ref, pos = G["ref%d" % i], G["pos%d" % i]
        L = np.array([0.7, -1.1, 0.4])
        vals, J, q, H = orot.rotation(pos, ref, None, L)
        np.testing.assert_allclose(q, G["q%d" % i], atol=1e-14)
        np.testing.assert_allclose(vals, G["val%d" % i], atol=1e-14)
        np.testing.assert_allclose(J, G["jac%d" % i], atol=1e-12)
        np.testing.assert_allclose(H, np.tensordot(L, G["hess%d" % i], axes=1), atol=1e-11)
    # and against finite differences of the value (what tests/internal/test_get_internal.py does)
    ref, pos = G["ref8"], G["pos8"]
    vals, J, q = orot.rotation(pos, ref)
    h = 1e-6
    for a in range(0, pos.size, 5):
        pp = pos.ravel().copy(); pp[a] += h
        pm = pos.ravel().copy(); pm[a] -= h
        fd = (orot.rotation(pp, ref, q)[0] - orot.rotation(pm, ref, q)[0]) / (2 * h)
        np.testing.assert_allclose(J[:, a], fd, rtol=1e-6, atol=1e-8)


def test_rotation_coordinate_meaning():
    """A rigid rotation by the vector theta gives the coordinate values theta (or -theta, by the
    reference's convention); translations give zero; the Jacobian rows are orthogonal to translations."""
    from oracle import rotation as orot
    rng = np.random.RandomState(3)
    ref = rng.normal(size=(9, 3))
    th = np.array([0.02, -0.05, 0.03])
    ang = np.linalg.norm(th); k = th / ang
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    Rm = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    vals, J, q = orot.rotation((ref - ref.mean(0)) @ Rm.T + 0.7, ref)
    assert min(np.abs(vals - th).max(), np.abs(vals + th).max()) < 1e-12
    vals0, J0, _ = orot.rotation(ref + np.array([0.3, -1.0, 2.0]), ref)
    np.testing.assert_allclose(vals0, 0.0, atol=1e-14)
    for d in range(3):
        np.testing.assert_allclose(J0[:, d::3].sum(axis=1), 0.0, atol=1e-12)


def test_nonlinear_pes_constraint_hessian_is_the_derivative_of_the_jacobian():
    """Hc = sum_i L_i d2r_i/dx2 (peswrapper.py:343-352): finite differences of drdx^T L at fixed L,
    for a bond + angle + rotation constraint set; and the oracle loop holds such constraints."""
    from oracle.pes import NonlinearPES
    from oracle.driver import SaddleSearch
    from sella_b200.synthetic import quadratic_system, quadratic_func
    n = 24
    A, xs, x0 = quadratic_system(2, n)
    coords = dict(bonds=[(0, 1)], angles=[(2, 3, 5)], rotation_ref=x0.reshape(-1, 3))
    p = NonlinearPES(quadratic_func(A, xs), x0, coords)
    p.get_g()
    L = p.curr["L"]
    assert L.shape == (5,)
    Hc = p.get_Hc()
    np.testing.assert_allclose(Hc, Hc.T, atol=1e-9)
    h = 1e-6
    rng = np.random.RandomState(0)
    for _ in range(3):
        v = rng.normal(size=n); v /= np.linalg.norm(v)
        xp, xm = x0 + h * v, x0 - h * v
        qp = NonlinearPES(quadratic_func(A, xs), xp, coords, targets=np.zeros(5)); qp.q_prev = p.q_prev
        qm = NonlinearPES(quadratic_func(A, xs), xm, coords, targets=np.zeros(5)); qm.q_prev = p.q_prev
        fd = (qp.get_drdx().T @ L - qm.get_drdx().T @ L) / (2 * h)
        np.testing.assert_allclose(Hc @ v, fd, rtol=1e-5, atol=1e-7)
    o = SaddleSearch(p, method="prfo", rs="ras", diag_maxiter=6)
    for _ in range(6):
        o.step()
    assert np.abs(p.get_res()).max() < 1e-3


@pytest.mark.parametrize("order", [0, 1])
def test_morse_cluster_like_the_reference_integration_test(order):
    """tests/integration/test_morse_cluster.py:11-46 of the reference (Cartesian variants) on the
    oracle: 4-atom Morse cluster, fix_translation() + fix_rotation(), gamma=1e-3, run to fmax 1e-3;
    the projected gradient vanishes and the projected Hessian of the Lagrangian has exactly `order`
    negative eigenvalues.  Same inputs (RandomState(4), r0 = 4.73, rho0 = 4.73 * 1.099; the `alpha`
    keyword of the reference's call is not a MorsePotential parameter, so epsilon stays 1)."""
    from oracle.pes import NonlinearPES
    from oracle.driver import SaddleSearch
    r0, rho0 = 4.73, 4.73 * 1.099

    def morse(x):
        pos = x.reshape(-1, 3)
        e, g = 0.0, np.zeros_like(pos)
        for i in range(len(pos)):
            for j in range(i + 1, len(pos)):
                d = pos[i] - pos[j]
                r = np.linalg.norm(d)
                ex = np.exp(-rho0 * (r / r0 - 1.0))
                e += ex * ex - 2.0 * ex
                de = (-2.0 * rho0 / r0) * (ex * ex - ex)
                g[i] += de * d / r
                g[j] -= de * d / r
        return e, g.ravel()
    rng = np.random.RandomState(4)
    nat = 4
    x0 = rng.normal(size=(nat, 3), scale=3.0).ravel()
    C = np.zeros((3, 3 * nat))
    for d in range(3):
        C[d, d::3] = 1.0 / nat
    p = NonlinearPES(morse, x0, dict(rotation_ref=x0.reshape(-1, 3)), np.zeros(3), C, C @ x0)
    o = SaddleSearch(p, order=order, gamma=1e-3)
    assert o.run(fmax=1e-3, steps=500)
    Ufree = p.get_Ufree()
    assert Ufree.shape == (12, 6)
    np.testing.assert_allclose(p.get_g() @ Ufree, 0, atol=5e-3)
    p.diag(gamma=1e-16)
    H = p.get_HL_projected(Ufree)
    assert int(np.sum(H.evals < 0)) == order, H.evals
    # the held coordinates are where they were put
    assert np.abs(p.get_res()).max() < 1e-5


def test_projected_spectrum_identity_used_by_the_engine():
    """The algebra behind BatchedSella._projected_spectrum_by_update: with Ucons orthonormal,
    A = (B - Hc) Ucons, G = Ucons^T A,
        Bp - B = P_f (B - Hc) P_f + sigma P_c - B = -Hc - Ucons A^T - A Ucons^T + Ucons (G + sigma I) Ucons^T,
    the rotation block of Hc is 1/2 sum_c (x_c y_c^T + y_c x_c^T) over the five pairs built from the
    kernel's 4-vectors, and the whole correction has rank 2 nc for translation + rotation constraints."""
    from oracle import rotation as orot
    from sella_b200.synthetic import fcc_cluster
    rng = np.random.RandomState(0)
    N = 16; n = 3 * N
    ref = fcc_cluster(N, seed=1)
    pos = ref + 0.05 * rng.normal(size=ref.shape)
    L = rng.normal(size=3)
    vals, J, q, Hc, f = orot.rotation(pos, ref, None, L, factors=True)
    X = np.hstack([f["Pv"] - f["dFw"], 2 * f["dE"][:, None]])           # n x 5
    Y = np.hstack([f["dc"], f["wdc"][:, None]])
    np.testing.assert_allclose(0.5 * (X @ Y.T + Y @ X.T), Hc, atol=1e-12 * max(1.0, np.abs(Hc).max()))
    B = rng.normal(size=(n, n)); B = B + B.T
    C = np.vstack([np.kron(np.ones(N) / N, np.eye(3)[d]) for d in range(3)] + [J[k] for k in range(3)])
    Uc = np.linalg.qr(C.T)[0].T                                          # nc x n, orthonormal rows
    nc, sigma = 6, 7.0
    Pc = Uc.T @ Uc; Pf = np.eye(n) - Pc
    Bp = Pf @ (B - Hc) @ Pf + sigma * Pc
    A = (B - Hc) @ Uc.T
    G = Uc @ A
    U = np.hstack([Uc.T, X]); Jp = np.hstack([-A, -0.5 * Y])             # the engine's (U, J) pairs
    Cm = np.zeros((nc + 5, nc + 5)); Cm[:nc, :nc] = -(G + sigma * np.eye(nc))
    delta = U @ Jp.T + Jp @ U.T - U @ Cm @ U.T
    np.testing.assert_allclose(delta, Bp - B, atol=1e-10)
    sv = np.linalg.svd(Bp - B, compute_uv=False)
    assert int((sv > 1e-10 * sv[0]).sum()) == 2 * nc


def _internal_case(G, i):
    """Rebuild the inputs of case i of internal_loop.npz: coordinate sets, surface, keyword arguments."""
    import ast
    from oracle.intcoords import CoordinateSet
    from oracle import emt as oemt
    pos0 = G["pos0_%d" % i]
    cell = G["cell%d" % i]
    cell = None if cell.shape[0] == 0 else cell
    pbc = tuple(bool(p) for p in G["pbc%d" % i])
    tv = {k: G["tv_%s%d" % (k, i)] for k in ("bonds", "angles", "dihedrals")}
    tr = [tuple(t) for t in G["trans%d" % i]]
    lists = (tr, G["bonds%d" % i], G["angles%d" % i], G["diheds%d" % i], tv)
    cs = CoordinateSet(len(pos0), *lists[:4], tvecs=tv, numbers=G["numbers%d" % i])
    rows = G["rows%d" % i]
    from test_internal_pes import subset
    csc = subset(len(pos0), lists, rows) if len(rows) else None
    kw = dict(ast.literal_eval(str(G["meta%d" % i][1])))
    return pos0, cs, csc, oemt.emt_func(cell, pbc), kw


@pytest.mark.parametrize("case", range(8))
def test_internal_pes_oracle_matches_reference_internal_pes(golden, case):
    """oracle/internal_pes.py + oracle/driver.py against the trajectory of the reference's OWN InternalPES,
    MaxInternalStep and Sella.step (tests/golden/internal_loop.npz, produced by make_golden.py through
    oracle/ref_internal_harness.py): slabs in bond coordinates with held atoms (prfo, qn, frozen B+, Newton
    stepper), a cluster with the full automatic list, free clusters (SVD branch; saddle search, minimisation, and with a bond
    and an angle held -- constraint rows with non-zero second derivatives).
    Both sides integrate the geodesic with scipy's LSODA, so every step agrees to round-off amplification."""
    from oracle.internal_pes import InternalPES
    from oracle.driver import SaddleSearch
    G = golden("internal_loop")
    assert int(G["ncases"]) == 8
    pos0, cs, csc, func, kw = _internal_case(G, case)
    pes_kw = {k: kw.pop(k) for k in ("exact_geodesic", "iterative_stepper") if k in kw}
    p = InternalPES(func, pos0.ravel(), cs, csc, integrator="lsoda", **pes_kw)
    o = SaddleSearch(p, rs="mis", diag_maxiter=6, **kw)
    X, D, R, F = G["x%d" % case], G["delta%d" % case], G["rho%d" % case], G["f%d" % case]
    for t in range(len(X)):
        o.step()
        msg = "case %d (%s) step %d" % (case, G["meta%d" % case], t)
        np.testing.assert_allclose(p.pos, X[t], rtol=0, atol=2e-8, err_msg=msg)
        np.testing.assert_allclose(o.delta, D[t], rtol=1e-8, err_msg=msg)
        np.testing.assert_allclose(o.rho, R[t], rtol=1e-6, atol=1e-8, err_msg=msg)
        np.testing.assert_allclose(p.curr["f"], F[t], rtol=0, atol=2e-8, err_msg=msg)
    H = G["H%d" % case]
    np.testing.assert_allclose(p.H.B, H, rtol=0, atol=1e-6 * max(1.0, np.abs(H).max()))
    assert p.neval == int(G["neval%d" % case])
