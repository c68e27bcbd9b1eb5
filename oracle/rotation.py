"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the reference's rotation coordinate (sella/internal.py:507-1007, 1031-1078):
the rotation vector (exponential map of the unit quaternion q that best superimposes the
current positions on a reference geometry; q = top eigenvector of the 4x4 matrix F built from
the correlation matrix, :534-551), its Jacobian by first-order eigenvector perturbation
(:607-648) and its Hessian by second-order perturbation (:703-800).

PINNED: tests/golden/rotation.npz holds outputs of the reference's own pure-numpy functions
(_build_F_matrix_np, _stabilize_quaternion, _rotation_3axis_jacobian_np,
_rotation_hessian_single), extracted from the reference source by tests/golden/make_golden.py
(the module itself cannot be imported: it needs jax and ase).
"""
import numpy as np


def build_F(dx, y):
    """:534-551.  dx: centred positions, y: centred reference."""
    R = dx.T @ y
    t = np.trace(R)
    top = np.array([R[1, 2] - R[2, 1], R[2, 0] - R[0, 2], R[0, 1] - R[1, 0]])
    F = np.empty((4, 4))
    F[0, 0] = t
    F[0, 1:] = top
    F[1:, 0] = top
    F[1:, 1:] = -t * np.eye(3) + R + R.T
    return F


def stabilize(ws, vecs, q_prev):
    """:569-585: the vector of the top eigenspace closest to q_prev, q[0] >= 0."""
    if q_prev is None:
        q_prev = np.array([1.0, 0.0, 0.0, 0.0])
    top = vecs[:, (ws[-1] - ws) < 1e-10]
    q = top @ (top.T @ q_prev)
    nrm = np.linalg.norm(q)
    q = vecs[:, -1].copy() if nrm < 1e-14 else q / nrm
    return -q if q[0] < 0 else q


def asinc(x):
    """:588-597."""
    if x < 0.97:
        return np.arccos(x) / np.sqrt(1.0 - x * x)
    y = x - 1.0
    return (1.0 - y / 3 + 2 * y**2 / 15 - 2 * y**3 / 35 + 8 * y**4 / 315 - 8 * y**5 / 693 + 16 * y**6 / 3003
            - 16 * y**7 / 6435 + 128 * y**8 / 109395 - 128 * y**9 / 230945)


def asinc_derivs(q0):
    """(asinc, asinc', asinc'') with the branches of :739-764 (the Hessian's own asinc value)."""
    if abs(q0 - 1.0) < 1e-8:
        y = q0 - 1.0
        return 1 - y / 3 + 2 * y**2 / 15, -1.0 / 3 + 4 * y / 15, 4.0 / 15
    if abs(q0) < 1.0 - 1e-12:
        s2 = 1 - q0**2
        s = np.sqrt(s2)
        ac = np.arccos(q0)
        return ac / s, -1.0 / s2 + q0 * ac / (s * s2), (3 * q0 / s2 - (1 + 2 * q0**2) * ac / (s * s2)) * (-1.0 / s2)
    return (np.pi / 2 if q0 > 0 else -np.pi / 2), 0.0, 0.0


def apply_dF(y, v):
    """(dF/dx_{k,d}) v for every atom k and Cartesian direction d -> (N, 3, 4)   (:651-700)."""
    N = len(y)
    out = np.zeros((N, 3, 4))
    v0, v3 = v[0], v[1:]
    yv = y @ v3
    for d in range(3):
        d1, d2 = (d + 1) % 3, (d + 2) % 3
        top = np.zeros((N, 3))
        top[:, d1] = -y[:, d2]
        top[:, d2] = y[:, d1]
        out[:, d, 0] = y[:, d] * v0 + top @ v3
        for i in range(3):
            val = -y[:, d] * v3[i] + y[:, i] * v3[d]
            if i == d:
                val = val + yv
            out[:, d, 1 + i] = top[:, i] * v0 + val
    return out


def rotation(pos, refpos, q_prev=None, L=None, factors=False):
    """values (3,), Jacobian (3, 3N), q, and sum_k L_k Hessian_k (3N x 3N) if L is given."""
    pos = np.asarray(pos, float).reshape(-1, 3)
    y = np.asarray(refpos, float).reshape(-1, 3)
    y = y - y.mean(0)
    N = len(pos)
    F = build_F(pos - pos.mean(0), y)
    ws, vecs = np.linalg.eigh(F)
    q = stabilize(ws, vecs, q_prev)
    vals = 2.0 * q[1:] * asinc(q[0])
    gaps = ws - ws[-1]
    inv = np.where(np.abs(gaps) > 1e-14, 1.0 / np.where(np.abs(gaps) > 1e-14, gaps, 1.0), 0.0)
    Minv = vecs @ (inv[:, None] * vecs.T)
    dFq = apply_dF(y, q).reshape(3 * N, 4)
    dc = -dFq @ Minv
    # Jacobian (:629-648): its own branches for asinc'
    q0 = q[0]
    a_j = asinc(q0)
    if abs(q0 - 1.0) < 1e-8:
        da_j = -1.0 / 3 + 4 * (q0 - 1.0) / 15
    elif abs(q0) < 1.0 - 1e-12:
        s2 = 1 - q0**2
        da_j = -1.0 / s2 + q0 * np.arccos(q0) / (np.sqrt(s2) * s2)
    else:
        da_j = 0.0
    J = np.stack([2 * (dc[:, k + 1] * a_j + q[k + 1] * da_j * dc[:, 0]) for k in range(3)])
    if L is None:
        return vals, J, q
    a, da, d2a = asinc_derivs(q0)
    df = np.zeros(4)
    d2f = np.zeros((4, 4))
    for k in range(3):
        df[0] += L[k] * 2 * q[k + 1] * da
        df[k + 1] += L[k] * 2 * a
        d2f[0, 0] += L[k] * 2 * q[k + 1] * d2a
        d2f[0, k + 1] += L[k] * 2 * da
        d2f[k + 1, 0] += L[k] * 2 * da
    w = Minv @ df
    dE = dFq @ q
    wdc = dc @ w
    dFw = apply_dF(y, w).reshape(3 * N, 4)
    H = dc @ d2f @ dc.T
    cross = dFw @ dc.T
    H += (dE[:, None] * wdc[None, :] + dE[None, :] * wdc[:, None] + 2 * (dFq @ dc.T) * (w @ q)
          - cross - cross.T - (df @ q) * (dc @ dc.T))
    if factors:
        # the per-coordinate 4-vectors the CUDA kernel keeps (sella_b200/csrc/rotation.cu):
        #   H = Pv dc^T - dc dFw^T + dE wdc^T + wdc dE^T,  Pv = dc d2f + 2 (w.q) dFq - dFw - (df.q) dc
        Pv = dc @ d2f + 2 * (w @ q) * dFq - dFw - (df @ q) * dc
        return vals, J, q, H, dict(dc=dc, Pv=Pv, dFw=dFw, dE=dE, wdc=wdc)
    return vals, J, q, H
