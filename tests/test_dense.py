"""Dense fp64 algebra of the internal-coordinate path (csrc/dense.cu) against numpy/LAPACK:
DMMA GEMM, Householder QR, triangular inverse, the gpu_qr / gpu_project seam and the
Wilson-matrix algebra of InternalPES (peswrapper.py:674-736, 1011-1082, 1124-1127, 1176-1183)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def up(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape", [(5, 7, 3), (64, 64, 16), (70, 130, 37), (384, 768, 384), (1, 384, 384)])
@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_gemm(shape, ta, tb):
    from sella_b200 import kernels as K
    M, N, Kd = shape
    rng = np.random.RandomState(M + 3 * N + 7 * Kd + ta + 2 * tb)
    A = rng.normal(size=(3, Kd, M) if ta else (3, M, Kd))
    B = rng.normal(size=(3, N, Kd) if tb else (3, Kd, N))
    C0 = rng.normal(size=(3, M, N))
    ref = 0.7 * np.matmul(A.transpose(0, 2, 1) if ta else A, B.transpose(0, 2, 1) if tb else B) - 0.3 * C0
    out = up(C0.copy())
    K.gemm(up(A), up(B), transA=bool(ta), transB=bool(tb), alpha=0.7, beta=-0.3, out=out)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=2e-13 * Kd ** 0.5 * 3)
    # shared right operand, beta = 0 must ignore whatever is in the output buffer
    out2 = K.gemm(up(A), up(B[0]), transA=bool(ta), transB=bool(tb))
    ref2 = np.matmul(A.transpose(0, 2, 1) if ta else A, B[0].T if tb else B[0])
    np.testing.assert_allclose(out2.cpu().numpy(), ref2, rtol=0, atol=2e-13 * Kd ** 0.5 * 3)


@pytest.mark.parametrize("m,n", [(9, 4), (50, 50), (195, 96), (768, 384)])
def test_qr_and_trtri(m, n):
    from sella_b200 import kernels as K
    rng = np.random.RandomState(m + n)
    A = rng.normal(size=(3, m, n))
    A[1, :, 2] = 0.0                                   # a zero column: tau = 0 path
    Q, R = K.qr(up(A))
    Q, R = Q.cpu().numpy(), R.cpu().numpy()
    for i in range(3):
        np.testing.assert_allclose(Q[i].T @ Q[i] if i != 1 else (Q[i].T @ Q[i]), np.eye(n), atol=1e-13 * n ** 0.5 * 5)
        np.testing.assert_allclose(Q[i] @ R[i], A[i], atol=1e-13 * n)
        assert np.abs(np.tril(R[i], -1)).max() == 0.0
        if i != 1:
            q_ref, r_ref = np.linalg.qr(A[i], mode="reduced")          # LAPACK: same Householder signs
            np.testing.assert_allclose(R[i], r_ref, atol=1e-12 * n)
            np.testing.assert_allclose(Q[i], q_ref, atol=1e-12 * n)
    Rinv, st = K.trtri(up(R[[0, 2]]))
    assert int(st.sum()) == 0
    for k, i in enumerate((0, 2)):
        scale = np.abs(np.linalg.inv(R[i])).max()
        np.testing.assert_allclose(Rinv[k].cpu().numpy() @ R[i], np.eye(n), atol=1e-12 * n * max(1.0, scale))
    _, st = K.trtri(up(R[[1]]))
    assert int(st[0]) & 16                             # singular diagonal reported, not hidden


def test_gpu_seam_qr_project():
    from sella_b200._gpu import gpu_qr, gpu_project
    rng = np.random.RandomState(5)
    A = rng.normal(size=(120, 40))
    Q, R = gpu_qr(A)
    q_ref, r_ref = np.linalg.qr(A, mode="reduced")
    np.testing.assert_allclose(Q, q_ref, atol=1e-12)
    np.testing.assert_allclose(R, r_ref, atol=1e-12)
    H = rng.normal(size=(120, 120)); H = H + H.T
    np.testing.assert_allclose(gpu_project(H, Q), Q.T @ H @ Q, atol=1e-12)


def _periodic_fcc_internals(seed):
    """2x2x2 fcc cell (32 atoms), nearest-neighbour bonds through the periodic boundary plus the
    three mean-position translations: a full-rank Wilson matrix (nint = 195 >= ncart = 96)."""
    from sella_b200.internal import BatchedInternals
    a = 3.61
    base = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
    pos = np.array([(np.array([i, j, k]) + bvec) * a for i in range(2) for j in range(2) for k in range(2) for bvec in base])
    cell = np.eye(3) * 2 * a
    bonds, tv = [], []
    for i in range(len(pos)):
        for j in range(i + 1, len(pos)):
            for sh in np.ndindex(3, 3, 3):
                t = (np.array(sh) - 1) @ cell
                if abs(np.linalg.norm(pos[j] + t - pos[i]) - a / np.sqrt(2)) < 1e-6:
                    bonds.append((i, j)); tv.append(t)
    rng = np.random.RandomState(seed)
    x = np.stack([(pos + 0.05 * rng.normal(size=pos.shape)).ravel() for _ in range(3)])
    # a mean-position translation is the average of the per-atom coordinates; with one (atom, dim)
    # entry per coordinate the three rows below pin atom 0, which fixes the same null space
    ints = BatchedInternals(len(pos), translations=[(0, 0), (0, 1), (0, 2)], bonds=bonds, tvec_bonds=np.array(tv))
    return ints, x


def test_wilson_algebra():
    from sella_b200.internal import WilsonAlgebra
    ints, x = _periodic_fcc_internals(1)
    xd = up(x)
    q, Bm = ints.calc(xd, jacobian=True)
    W = WilsonAlgebra(Bm)
    assert not bool(W.rank_deficient.any()) and int(W.status.sum()) == 0
    Bn = Bm.cpu().numpy()
    b, nint, ncart = Bn.shape
    rng = np.random.RandomState(2)
    g = rng.normal(size=(b, ncart)); dx = rng.normal(size=(b, nint)); gi = rng.normal(size=(b, nint))
    J = rng.normal(size=(b, 4, ncart)); h0 = np.abs(rng.normal(size=nint)) + 0.1
    Dc = rng.normal(size=(b, ncart, ncart)); Dq = rng.normal(size=(b, ncart, ncart)); H = rng.normal(size=(b, nint, nint))
    H = H + H.transpose(0, 2, 1)
    g_int = W.gradient(up(g)).cpu().numpy()
    red, dint = (t.cpu().numpy() for t in W.drdx(up(J)))
    Hc = W.Hc(up(Dc), up(Dq)).cpu().numpy()
    dfp = W.df_pred(up(dx), up(gi), up(H)).cpu().numpy()
    H0 = W.H0(up(h0)).cpu().numpy()
    for i in range(b):
        Qr, Rr = np.linalg.qr(Bn[i], mode="reduced")               # peswrapper.py:691
        Binv = np.linalg.solve(Rr, Qr.T)                            # :726 (solve_triangular)
        scale = np.abs(Binv).max()
        np.testing.assert_allclose(W.Binv[i].cpu().numpy(), Binv, atol=1e-11 * scale)
        np.testing.assert_allclose(W.Binv[i].cpu().numpy() @ Bn[i], np.eye(ncart), atol=1e-11 * scale)
        np.testing.assert_allclose(g_int[i], g[i] @ Binv, atol=1e-11 * scale)                      # :1127
        rn = J[i] @ np.linalg.inv(Rr)                                                               # :1071
        np.testing.assert_allclose(red[i], rn, atol=1e-10 * scale)
        np.testing.assert_allclose(dint[i], rn @ Qr.T, atol=1e-10 * scale)                          # :1078
        np.testing.assert_allclose(Hc[i], Binv.T @ (Dc[i] - Dq[i]) @ Binv, atol=1e-10 * scale ** 2)  # :1031
        ref_df = (gi[i] @ Qr) @ (dx[i] @ Qr) + 0.5 * (dx[i] @ Qr) @ (Qr.T @ H[i] @ Qr) @ (dx[i] @ Qr)  # :1176-1183
        np.testing.assert_allclose(dfp[i], ref_df, rtol=1e-11, atol=1e-11)
        P = Qr @ Qr.T                                                                                # :72-82
        np.testing.assert_allclose(H0[i], P @ np.diag(h0) @ P, atol=1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [5, 32, 47, 96, 384])
def test_potrf_against_lapack(n):
    """sb_potrf: upper Cholesky factor of B^T B (the R factor of a tall matrix B up to row signs), and the
    non-positive-pivot flag on an indefinite input."""
    torch = pytest.importorskip("torch")
    from sella_b200 import kernels as K
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(n)
    b = 5
    Bm = rng.normal(size=(b, 2 * n + 3, n))
    G = np.einsum("bki,bkj->bij", Bm, Bm)
    R, st = K.potrf(torch.from_numpy(G).to(dev))
    R = R.cpu().numpy()
    assert int(st.abs().sum()) == 0
    for i in range(b):
        ref = np.linalg.cholesky(G[i]).T
        np.testing.assert_allclose(R[i], ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max())
        assert np.array_equal(np.tril(R[i], -1), np.zeros((n, n)))
        Rq = np.linalg.qr(Bm[i], mode="r")
        np.testing.assert_allclose(np.abs(R[i]), np.abs(Rq), rtol=1e-9, atol=1e-9 * np.abs(Rq).max())
    G[2] -= 2.0 * np.linalg.eigvalsh(G[2])[0] * np.eye(n) + np.eye(n)       # indefinite
    _, st = K.potrf(torch.from_numpy(G).to(dev))
    assert int(st[2]) & 16 and int(st[0]) == 0
