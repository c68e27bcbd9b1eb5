"""Internal coordinates: CPU check of the oracle against finite differences (the
reference's own test method, tests/internal/test_get_internal.py:25-55) and GPU parity
of the CUDA hyper-dual kernels against the oracle."""
import numpy as np
import pytest

from oracle import internals as oi


def _molecule(seed=0, natoms=9):
    rng = np.random.RandomState(seed)
    pos = rng.normal(size=(natoms, 3)) * 1.2 + np.arange(natoms)[:, None] * np.array([0.9, 0.2, -0.1])
    bonds = [(i, i + 1) for i in range(natoms - 1)]
    angles = [(i, i + 1, i + 2) for i in range(natoms - 2)]
    diheds = [(i, i + 1, i + 2, i + 3) for i in range(natoms - 3)]
    trans = [(0, 0), (0, 1), (0, 2), (natoms - 1, 2)]
    return pos, trans, bonds, angles, diheds


def test_oracle_jacobian_and_hessians_vs_finite_differences():
    pos, trans, bonds, angles, diheds = _molecule()
    q, B, H = oi.evaluate(pos, trans, bonds, angles, diheds)
    n = pos.size
    h = 1e-6
    Bfd = np.zeros_like(B)
    for i in range(n):
        pp = pos.ravel().copy(); pp[i] += h
        pm = pos.ravel().copy(); pm[i] -= h
        qp = oi.evaluate(pp.reshape(-1, 3), trans, bonds, angles, diheds)[0]
        qm = oi.evaluate(pm.reshape(-1, 3), trans, bonds, angles, diheds)[0]
        Bfd[:, i] = (qp - qm) / (2 * h)
    np.testing.assert_allclose(B, Bfd, rtol=1e-7, atol=1e-7)
    for Hc in H:
        np.testing.assert_allclose(Hc, Hc.T, atol=1e-12)
    # translation of all atoms leaves bonds/angles/dihedrals unchanged: B rows sum to 0 per Cartesian dim
    nt = len(trans)
    for d in range(3):
        np.testing.assert_allclose(B[nt:, d::3].sum(axis=1), 0.0, atol=1e-12)


@pytest.mark.gpu
def test_cuda_internals_match_oracle():
    torch = pytest.importorskip("torch")
    from sella_b200.internal import BatchedInternals
    dev = torch.device("cuda:0")
    natoms, batch = 9, 5
    geoms = [_molecule(seed=s, natoms=natoms) for s in range(batch)]
    _, trans, bonds, angles, diheds = geoms[0]
    rng = np.random.RandomState(3)
    tb = rng.normal(size=(len(bonds), 1, 3)) * 0.3            # PBC shift vectors
    ta = rng.normal(size=(len(angles), 2, 3)) * 0.3
    td = rng.normal(size=(len(diheds), 3, 3)) * 0.3
    for tv in (None, dict(bonds=tb, angles=ta, dihedrals=td)):
        ints = BatchedInternals(natoms, trans, bonds, angles, diheds,
                                **({} if tv is None else dict(tvec_bonds=tb, tvec_angles=ta, tvec_dihedrals=td)))
        x = torch.from_numpy(np.stack([g[0].ravel() for g in geoms])).to(dev)
        q, B = ints.calc(x, jacobian=True)
        v = rng.normal(size=(batch, ints.nint)); w = rng.normal(size=(batch, 3 * natoms))
        D = ints.ldot(x, torch.from_numpy(v).to(dev))
        R = ints.rdot(x, torch.from_numpy(w).to(dev))
        for i, g in enumerate(geoms):
            qr, Br, Hr = oi.evaluate(g[0], trans, bonds, angles, diheds, tvecs=tv)
            np.testing.assert_allclose(q[i].cpu().numpy(), qr, rtol=1e-13, atol=1e-13)
            np.testing.assert_allclose(B[i].cpu().numpy(), Br, rtol=1e-11, atol=1e-12)
            Dref = sum(vc * Hc for vc, Hc in zip(v[i], Hr))
            np.testing.assert_allclose(D[i].cpu().numpy(), Dref, rtol=1e-7, atol=1e-7)     # oracle Hessians are FD
            Rref = np.array([Hc @ w[i] for Hc in Hr])
            np.testing.assert_allclose(R[i].cpu().numpy(), Rref, rtol=1e-7, atol=1e-7)


@pytest.mark.gpu
def test_cuda_internals_hessian_vs_fd_of_cuda_jacobian():
    """Tight check of the hyper-dual second derivatives: central differences of the CUDA
    B-matrix itself (error ~1e-9), independent of the oracle's FD Hessians."""
    torch = pytest.importorskip("torch")
    from sella_b200.internal import BatchedInternals
    dev = torch.device("cuda:0")
    pos, trans, bonds, angles, diheds = _molecule(seed=11, natoms=8)
    ints = BatchedInternals(8, trans, bonds, angles, diheds)
    n = pos.size
    x0 = pos.ravel()
    h = 1e-5
    xs = np.stack([x0] + [x0 + h * np.eye(n)[i] for i in range(n)] + [x0 - h * np.eye(n)[i] for i in range(n)])
    B = ints.jacobian(torch.from_numpy(xs).to(dev)).cpu().numpy()
    rng = np.random.RandomState(5)
    v = rng.normal(size=ints.nint)
    # d/dx_i (v^T B) = row i of sum_c v_c H_c
    Dfd = np.stack([(v @ B[1 + i] - v @ B[1 + n + i]) / (2 * h) for i in range(n)])
    D = ints.ldot(torch.from_numpy(x0[None]).to(dev), torch.from_numpy(v[None]).to(dev))[0].cpu().numpy()
    np.testing.assert_allclose(D, Dfd, rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(D, D.T, atol=1e-12)


@pytest.mark.gpu
def test_cuda_rotation_coordinate_matches_reference_golden(golden):
    """sb_rotation against outputs of the reference's own rotation functions (tests/golden/rotation.npz)."""
    torch = pytest.importorskip("torch")
    from sella_b200.internal import BatchedInternals
    G = golden("rotation")
    dev = torch.device("cuda:0")
    L = np.array([0.7, -1.1, 0.4])
    for i in range(int(G["ncases"])):
        ref, pos = G["ref%d" % i], G["pos%d" % i]
        N = len(ref)
        ints = BatchedInternals(N, bonds=[(0, 1)], rotation_ref=ref)
        x = torch.from_numpy(np.stack([pos.ravel(), pos.ravel() + 0.3])).to(dev)       # 2nd system: translated copy
        q, B = ints.calc(x, jacobian=True)
        q, B = q.cpu().numpy(), B.cpu().numpy()
        np.testing.assert_allclose(ints.qprev.cpu().numpy()[0], G["q%d" % i], atol=1e-13)
        for s in range(2):
            np.testing.assert_allclose(q[s, 1:], G["val%d" % i], atol=1e-13)
            np.testing.assert_allclose(B[s, 1:], G["jac%d" % i], atol=1e-9, rtol=1e-9)
        np.testing.assert_allclose(q[0, 0], np.linalg.norm(pos[1] - pos[0]), rtol=1e-14)
        v = np.zeros((2, 4)); v[:, 1:] = L
        D = ints.ldot(x, torch.from_numpy(v).to(dev)).cpu().numpy()
        Href = np.tensordot(L, G["hess%d" % i], axes=1)
        scale = max(1.0, np.abs(Href).max())
        np.testing.assert_allclose(D[0], Href, atol=2e-8 * scale)
        np.testing.assert_allclose(D[1], Href, atol=2e-8 * scale)


def test_oracle_ldot_rdot_assembly_matches_reference_golden(golden):
    """tests/golden/sparse_hessians.npz: outputs of the reference's own SparseInternalHessians.ldot / rdot / ddot
    (sella/linalg.py:540-646) fed with the oracle's per-coordinate blocks.  Pins the assembly (scatter and
    contraction) of the second derivatives; the derivative VALUES remain pinned by finite differences only
    (the reference takes them from JAX)."""
    G = golden("sparse_hessians")
    pos = G["pos"]
    bonds, angles, diheds = [tuple(r) for r in G["bonds"]], [tuple(r) for r in G["angles"]], [tuple(r) for r in G["dihedrals"]]
    q, B, H = oi.evaluate(pos, (), bonds, angles, diheds)
    H = np.array(H)
    # the blocks the golden was generated from are reproduced bit for bit (same code, same inputs) ...
    coords = bonds + angles + diheds
    for i, atoms in enumerate(coords):
        vals = G["vals%d" % i]
        for ia, a in enumerate(atoms):
            for ja, a2 in enumerate(atoms):
                np.testing.assert_array_equal(vals[ia, :, ja, :], H[i][3 * a:3 * a + 3, 3 * a2:3 * a2 + 3])
    # ... and the dense restatement agrees with the reference's sparse assembly
    np.testing.assert_allclose(np.einsum("i,ijk->jk", G["v"], H), G["ldot"], atol=1e-13)
    np.testing.assert_allclose(H @ G["w"], G["rdot"], atol=1e-13)
    np.testing.assert_allclose(np.einsum("j,ijk,k->i", G["w"], H, G["w"]), G["ddot"], atol=1e-12)


@pytest.mark.gpu
def test_cuda_ldot_rdot_match_reference_golden(golden):
    """The CUDA hyper-dual kernels (sb_internals_hess) against the reference's SparseInternalHessians outputs."""
    torch = pytest.importorskip("torch")
    from sella_b200.internal import BatchedInternals
    G = golden("sparse_hessians")
    dev = torch.device("cuda:0")
    natoms = G["pos"].shape[0]
    ints = BatchedInternals(natoms, (), [tuple(r) for r in G["bonds"]], [tuple(r) for r in G["angles"]],
                            [tuple(r) for r in G["dihedrals"]])
    x = torch.from_numpy(G["pos"].reshape(1, -1).copy()).to(dev)
    D = ints.ldot(x, torch.from_numpy(G["v"][None].copy()).to(dev))
    R = ints.rdot(x, torch.from_numpy(G["w"][None].copy()).to(dev))
    # the golden blocks come from central differences of analytic gradients (h = 1e-5): ~1e-9 accurate
    np.testing.assert_allclose(D[0].cpu().numpy(), G["ldot"], atol=2e-8)
    np.testing.assert_allclose(R[0].cpu().numpy(), G["rdot"], atol=2e-8)
