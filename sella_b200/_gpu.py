"""CUDA mirror of the reference's device seam sella/_gpu.py (gpu_eigh :70-84,
gpu_eigh_t :87-97, gpu_qr :100-111, gpu_project :114-132, to_gpu :55-67): same names and return
conventions, backed by the hand-written kernels instead of torch.linalg.

There is deliberately no CPU fallback and no size threshold: the reference's
``SELLA_GPU_MIN_DIM`` / OOM bookkeeping exists to fall back to LAPACK, which this
package does not do."""
import numpy as np
import torch

from . import kernels as K
from ._host import up, up_mat, raise_status


def to_gpu(A):
    return up(A)


def gpu_eigh_t(A_gpu):
    """(evals, evecs) as CUDA tensors; eigenvectors in COLUMNS like torch.linalg.eigh."""
    w, Vt, status = K.eigh(A_gpu.unsqueeze(0).contiguous())
    raise_status(status, "eigh")
    return w[0], Vt[0].T.contiguous()


def gpu_eigh(A, A_gpu=None):
    w, V = gpu_eigh_t(A_gpu if A_gpu is not None else up(A))
    return w.cpu().numpy(), V.cpu().numpy()


def gpu_qr(A):
    """Economy QR (sella/_gpu.py:100-111): (Q, R) as numpy arrays, LAPACK sign convention."""
    Q, R = K.qr(up(np.asarray(A, dtype=np.float64)).unsqueeze(0).contiguous())
    return Q[0].cpu().numpy(), R[0].cpu().numpy()


def gpu_project(H, U, H_gpu=None):
    """U.T @ H @ U (sella/_gpu.py:114-132): two fp64 tensor-core GEMMs."""
    Hd = (H_gpu if H_gpu is not None else up(H)).unsqueeze(0).contiguous()
    Ud = up(np.asarray(U, dtype=np.float64)).unsqueeze(0).contiguous()      # [1, n, m]
    HU = K.gemm(Hd, Ud)                                                       # [1, n, m]
    return K.gemm(Ud, HU, transA=True)[0].cpu().numpy()
