"""GPU parity tests of the individual CUDA operators (through the C ABI) against
numpy/scipy (the arithmetic the reference itself calls) and the CPU oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def rand_sym(rng, b, n):
    A = rng.normal(size=(b, n, n))
    return 0.5 * (A + A.transpose(0, 2, 1))


@pytest.mark.parametrize("n", [16, 30, 33, 96, 192, 384, 768, 1536])
@pytest.mark.parametrize("nvec", [1, 2, 3, 5])
def test_hv_and_transpose(n, nvec):
    from sella_b200 import kernels as K
    rng = np.random.RandomState(n + nvec)
    b = 3 if n >= 768 else 7
    A = rng.normal(size=(b, n, n))
    X = rng.normal(size=(b, nvec, n))
    Y = K.hv(to_dev(A), to_dev(X)).cpu().numpy()
    ref = np.einsum("bij,bvj->bvi", A, X)
    np.testing.assert_allclose(Y, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    Yt = K.hv(to_dev(A), to_dev(X), transposed=True).cpu().numpy()
    reft = np.einsum("bji,bvj->bvi", A, X)
    np.testing.assert_allclose(Yt, reft, rtol=1e-12, atol=1e-12 * np.abs(reft).max())


def test_hv_many_systems_and_mask():
    from sella_b200 import kernels as K
    rng = np.random.RandomState(5)
    b, n = 301, 192           # > #SMs: exercises the persistent tile loop and ring wrap
    A = rng.normal(size=(b, n, n)); x = rng.normal(size=(b, n))
    act = (rng.rand(b) > 0.3).astype(np.int32)
    out = torch.full((b, 1, n), 7.0, dtype=torch.float64, device=dev())
    K.hv(to_dev(A), to_dev(x[:, None, :]), active=to_dev(act), out=out)
    got = out.cpu().numpy()[:, 0]
    ref = np.einsum("bij,bj->bi", A, x)
    np.testing.assert_allclose(got[act == 1], ref[act == 1], rtol=1e-12, atol=1e-11)
    assert np.all(got[act == 0] == 7.0)
    out.fill_(7.0)
    K.hv(to_dev(A), to_dev(x[:, None, :]), transposed=True, active=to_dev(act), out=out)
    got = out.cpu().numpy()[:, 0]
    ref = np.einsum("bji,bj->bi", A, x)
    np.testing.assert_allclose(got[act == 1], ref[act == 1], rtol=1e-12, atol=1e-11)
    assert np.all(got[act == 0] == 7.0)


def test_quadratic_pes_matches_host_surface():
    from sella_b200 import kernels as K
    from sella_b200.synthetic import quadratic_batch
    A, xs, x0 = quadratic_batch(5, 96)
    f, g = K.quadratic_pes(to_dev(A), to_dev(xs), to_dev(x0))
    d = x0 - xs
    gref = np.einsum("bij,bj->bi", A, d)
    np.testing.assert_allclose(g.cpu().numpy(), gref, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(f.cpu().numpy(), 0.5 * np.einsum("bi,bi->b", d, gref), rtol=1e-12)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 24, 96, 192, 384])
def test_eigh_against_lapack(n):
    from sella_b200 import kernels as K
    rng = np.random.RandomState(100 + n)
    b = 6
    A = rand_sym(rng, b, n)
    # quasi-Newton-like member: scaled identity + low rank (massively degenerate)
    u = rng.normal(size=(n, min(2, n)))
    A[1] = 0.7 * np.eye(n) + u @ u.T - 2.0 * np.outer(u[:, 0], u[:, 0])
    A[2] = np.diag(rng.normal(size=n))            # already diagonal
    # exactly rank deficient: P diag(h) P with a projector of rank n/3 (the model Hessian of an
    # internal-coordinate search, peswrapper.py:641-652) -- a (2n/3)-fold eigenvalue 0 +- 1e-16 |A|
    if n >= 6:
        qq = np.linalg.qr(rng.normal(size=(n, n // 3)))[0]
        A[3] = qq @ (qq.T * (1.0 + 50.0 * rng.rand(n))[None, :]) @ qq @ qq.T
        A[3] = 0.5 * (A[3] + A[3].T)
    w, Vt, status = K.eigh(to_dev(A))
    w, Vt = w.cpu().numpy(), Vt.cpu().numpy()
    assert int(status.abs().sum()) == 0
    for i in range(b):
        wref = np.linalg.eigvalsh(A[i])
        scale = max(1.0, np.abs(wref).max())
        np.testing.assert_allclose(w[i], wref, rtol=0, atol=5e-13 * scale * max(1, n / 32))
        V = Vt[i].T
        np.testing.assert_allclose(V.T @ V, np.eye(n), atol=5e-13 * max(1, n / 32))
        np.testing.assert_allclose(A[i] @ V, V * w[i][None, :], atol=5e-12 * scale * max(1, n / 32))


@pytest.mark.parametrize("n", [24, 96, 384])
def test_eigvalsh_values_only(n):
    from sella_b200 import kernels as K
    rng = np.random.RandomState(500 + n)
    A = rand_sym(rng, 5, n)
    st = torch.zeros(5, dtype=torch.int32, device=dev())
    w = K.eigvalsh(to_dev(A), status=st).cpu().numpy()
    assert int(st.abs().sum()) == 0
    for i in range(5):
        wref = np.linalg.eigvalsh(A[i])
        np.testing.assert_allclose(w[i], wref, rtol=0, atol=5e-13 * max(1.0, np.abs(wref).max()) * max(1, n / 32))


def test_eigh_mask_leaves_inactive_untouched():
    from sella_b200 import kernels as K
    rng = np.random.RandomState(3)
    A = rand_sym(rng, 4, 48)
    act = np.array([1, 0, 1, 0], dtype=np.int32)
    w0 = torch.full((4, 48), -5.0, dtype=torch.float64, device=dev())
    V0 = torch.full((4, 48, 48), -5.0, dtype=torch.float64, device=dev())
    K.eigh(to_dev(A), active=to_dev(act), evals=w0, Vt=V0)
    assert torch.all(w0[1] == -5.0) and torch.all(V0[3] == -5.0)
    np.testing.assert_allclose(w0[0].cpu().numpy(), np.linalg.eigvalsh(A[0]), atol=1e-12)


@pytest.mark.parametrize("n,k", [(24, 1), (96, 1), (96, 3), (384, 1), (384, 5)])
def test_eigh_update_secular(n, k):
    """Secular-equation update of an eigendecomposition after a symmetric rank-2k
    change, chained three times, against numpy.linalg.eigh of the updated matrix."""
    from sella_b200 import kernels as K
    rng = np.random.RandomState(7 * n + k)
    b = 4
    B = rand_sym(rng, b, n)
    B[0] = 0.8 * np.eye(n)                          # fully degenerate start (first Sella update)
    u = rng.normal(size=(n, 3))
    B[1] = 0.5 * np.eye(n) + u @ np.diag([1.0, -2.0, 0.3]) @ u.T     # identity + low rank
    w0 = np.empty((b, n)); V0 = np.empty((b, n, n))
    for i in range(b):
        w0[i], v = np.linalg.eigh(B[i]); V0[i] = v.T
    evals, Vt = to_dev(w0), to_dev(V0)
    for rep in range(3):
        U = rng.normal(size=(b, k, n)) / np.sqrt(n)
        J = rng.normal(size=(b, k, n)) / np.sqrt(n)
        C = rng.normal(size=(b, k, k))
        if rep == 1:
            J[2] = 2.0 * U[2]                       # rank-deficient [U J]
        Csym = 0.5 * (C + C.transpose(0, 2, 1))
        Delta = (np.einsum("bki,bkj->bij", U, J) + np.einsum("bki,bkj->bij", J, U)
                 - np.einsum("bki,bkl,blj->bij", U, Csym, U))
        B = B + Delta
        _, _, status = K.eigh_update(evals, Vt, to_dev(U), to_dev(J), to_dev(C))
        assert int(status.abs().sum()) == 0
        w, V = evals.cpu().numpy(), Vt.cpu().numpy()
        for i in range(b):
            wref = np.linalg.eigvalsh(B[i])
            scale = max(1.0, np.abs(wref).max())
            np.testing.assert_allclose(w[i], wref, rtol=0, atol=2e-12 * scale, err_msg="rep %d sys %d" % (rep, i))
            Vi = V[i].T
            np.testing.assert_allclose(Vi.T @ Vi, np.eye(n), atol=2e-12, err_msg="orth rep %d sys %d" % (rep, i))
            np.testing.assert_allclose(B[i] @ Vi, Vi * w[i][None, :], atol=5e-12 * scale)


@pytest.mark.parametrize("n", [768, 1536])
def test_eigh_large_sizes(n):
    """BASELINE configs C4/C5 (3N = 768, 1536): eigenpairs through size-independent properties."""
    from sella_b200 import kernels as K
    rng = np.random.RandomState(n)
    A = rng.normal(size=(2, n, n)); A = A + A.transpose(0, 2, 1)
    w, Vt, st = K.eigh(torch.from_numpy(A).cuda())
    assert int(st.sum()) == 0
    w, Vt = w.cpu().numpy(), Vt.cpu().numpy()
    for i in range(2):
        scale = np.abs(w[i]).max()
        assert (np.diff(w[i]) >= 0).all()
        np.testing.assert_allclose(Vt[i] @ Vt[i].T, np.eye(n), atol=5e-13)
        np.testing.assert_allclose(A[i] @ Vt[i].T, Vt[i].T * w[i][None, :], atol=5e-12 * scale)
        np.testing.assert_allclose(w[i].sum(), np.trace(A[i]), rtol=0, atol=1e-11 * scale * n ** 0.5)


def p_B(pes):
    return pes.H.B


@pytest.mark.parametrize("n", [768, 1536])
def test_engine_large_sizes_match_oracle(n):
    """C4/C5 sizes: a few steps of two searches against the oracle, and the carried eigenpairs
    against the stored Hessian."""
    from oracle.pes import CartesianPES
    from oracle.driver import SaddleSearch
    from sella_b200.batched import BatchedSella, QuadraticSurface
    from sella_b200.synthetic import quadratic_system, quadratic_func
    data = [quadratic_system(b, n) for b in (0, 1)]
    dev = torch.device("cuda:0")
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    eng = BatchedSella(QuadraticSurface(up(np.stack([d[0] for d in data])), up(np.stack([d[1] for d in data]))),
                       up(np.stack([d[2] for d in data])), method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3, kcap=8)
    orc = []
    for (A, xs, x0) in data:
        p = CartesianPES(quadratic_func(A, xs), x0)
        orc.append((p, SaddleSearch(p, method="prfo", rs="tr", diag_maxiter=5, diag_every_n=3)))
    for t in range(4):
        eng.step()
        x = eng.x.cpu().numpy()
        for i, (p, o) in enumerate(orc):
            o.step()
            np.testing.assert_allclose(x[i], p.get_x(), rtol=0, atol=1e-8)
    eng.check_status()
    B = eng.B.cpu().numpy()
    for i in range(2):
        th, VR, lam0, m = eng.explicit_pairs(i)
        assert eng.compact and m < n                      # a handful of steps: far from the dense limit
        np.testing.assert_allclose(B[i] @ VR.T, VR.T * th[None, :], atol=1e-10)
        np.testing.assert_allclose(p_B(orc[i][0]), B[i], rtol=1e-6, atol=1e-7)
