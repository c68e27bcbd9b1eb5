"""Helpers for the numpy-facing (batch = 1) mirror of the reference API: move host
arrays to the device, run the CUDA kernels, bring the result back.  No numerics here."""
import numpy as np
import torch

from ._lib import require_cuda


def dev():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def up(a):
    """numpy (any float layout) -> contiguous float64 CUDA tensor."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev())


def up_mat(a):
    """(n, n) -> [1, n, n]"""
    return up(a).unsqueeze(0).contiguous()


def up_cols(a):
    """(n, k) column block (reference convention) -> [1, k, n] vector-major."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    return up(a.T).unsqueeze(0).contiguous()


def down_cols(t, k=None):
    """[1, kcap, n] vector-major -> numpy (n, k)."""
    x = t[0] if k is None else t[0, :k]
    return x.transpose(0, 1).contiguous().cpu().numpy()


def zeros(*shape, dtype=torch.float64):
    return torch.zeros(*shape, dtype=dtype, device=dev())


def raise_status(status, what):
    st = int(status.max().item()) if status.numel() else 0
    if st & 1:
        raise RuntimeError("MGS failed.")
    if st & 2:
        raise RuntimeError("Restricted step failed to converge!")
    if st:
        raise RuntimeError("%s failed (status %d)" % (what, st))
