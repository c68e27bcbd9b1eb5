"""CPU test pinning the algebra of the engine's compact Hessian representation
(tests/compact_proto.py, the numpy blueprint of sella_b200/csrc/compact.cu) against the oracle's
dense path: the carried (lam0, theta, VR) reproduce B, the restricted steps, |B|s, B s and the jd0
correction that the reference computes from a fresh eigh(B) (linalg.py:174-195, stepper.py:75-185,
eigensolvers.py:123-139)."""
import numpy as np
import pytest

from compact_proto import CompactSpectrum


def _lowrank_terms(D, tol=1e-13):
    w, V = np.linalg.eigh(0.5 * (D + D.T))
    keep = np.abs(w) > tol * max(1.0, np.abs(w).max())
    return V[:, keep].T.copy(), w[keep]


@pytest.mark.parametrize("method", ["qn", "prfo", "rfo"])
@pytest.mark.parametrize("n", [12, 30])
def test_compact_representation_follows_oracle_loop(n, method):
    from oracle.pes import CartesianPES
    from oracle.driver import SaddleSearch
    from oracle.restricted import TrustRegion
    from oracle.stepper import get_stepper
    from sella_b200.synthetic import quadratic_system, quadratic_func
    A, xs, x0 = quadratic_system(3, n)
    p = CartesianPES(quadratic_func(A, xs), x0)
    o = SaddleSearch(p, method=method, rs="tr", diag_every_n=3, diag_maxiter=4)
    cs = None
    Bprev = None
    nsteps = 2 * n            # long enough for the explicit rank to reach n (dense limit)
    for t in range(nsteps):
        o.step()
        B = p.H.B
        if cs is None:
            # first update: scaled identity + low rank (hessian_update.py:58-67)
            w = np.linalg.eigvalsh(B)
            lam0 = np.median(w)
            cs = CompactSpectrum(n, lam0)
            P, sig = _lowrank_terms(B - lam0 * np.eye(n))
        else:
            P, sig = _lowrank_terms(B - Bprev)
        if len(sig):
            cs.update(P, sig)
        Bprev = B.copy()
        assert cs.m <= n
        np.testing.assert_allclose(cs.dense(), B, atol=1e-11, err_msg="step %d" % t)
        np.testing.assert_allclose(cs.VR @ cs.VR.T, np.eye(cs.m), atol=1e-12)
        # the step the reference takes next from (g, eigh(B)) ...
        g = p.get_g()
        rs = TrustRegion(p, 1, o.delta, method=method)
        s_ref, smag_ref = rs.get_s()
        # ... and the same from the compact pole list
        ev, vg, rowmap, gperp = cs.poles(g, width=min(n, cs.m + 3))

        class _H:                      # duck type of ApproximateHessian in the pole basis
            evals, evecs = ev, np.eye(len(ev))

            @staticmethod
            def asarray():
                return np.diag(ev)

            @staticmethod
            def project(U):
                class _P:
                    pass
                hp = _P()
                hp.evals, hp.evecs = None, None
                Bp = U.T @ np.diag(ev) @ U
                hp.asarray = lambda: Bp
                return hp

        class _Pes:
            int, n_cell_dof = None, 0
            get_g = staticmethod(lambda: vg.copy())
            get_scons = staticmethod(lambda: np.zeros(len(ev)))
            get_H = staticmethod(lambda: np.diag(ev))
            get_Ufree = staticmethod(lambda: np.eye(len(ev)))
            get_Unred = staticmethod(lambda: np.eye(len(ev)))
            get_HL_projected = staticmethod(lambda U: _H)

        c_pole, smag = TrustRegion(_Pes, 1, o.delta, method=method).get_s()
        s = cs.lift(c_pole, rowmap, gperp)
        np.testing.assert_allclose(smag, smag_ref, rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(s, s_ref, atol=1e-9 * max(1.0, np.abs(s_ref).max()), err_msg="step %d" % t)
        # spectral functions
        wB, VB = np.linalg.eigh(B)
        absB = (VB * np.abs(wB)) @ VB.T
        np.testing.assert_allclose(cs.apply(np.abs, s), absB @ s, atol=1e-10)
        np.testing.assert_allclose(cs.apply(lambda x: x, s), B @ s, atol=1e-10)
    assert cs.m == n                   # the run ended in the dense limit


def test_compact_jd0_correction_matches_bordered_solve():
    """jd0 of eigensolvers.py:133-139 solves [[P - theta I, v], [v^T, 0]] z = -[r; 0]."""
    rng = np.random.RandomState(5)
    n = 40
    cs = CompactSpectrum(n, 0.7)
    for k in (3, 2, 4):
        Q, _ = np.linalg.qr(rng.normal(size=(n, k)))
        cs.update(Q.T.copy(), rng.normal(size=k))
    assert cs.m == 9
    B = cs.dense()
    v = rng.normal(size=n); v /= np.linalg.norm(v)
    theta = v @ B @ v
    r = B @ v - theta * v
    K = np.block([[B - theta * np.eye(n), v[:, None]], [v[None, :], np.zeros((1, 1))]])
    t_ref = np.linalg.solve(K, -np.concatenate([r, [0.0]]))[:n]
    np.testing.assert_allclose(cs.jd0(r, v, theta), t_ref, rtol=1e-9, atol=1e-11)
