"""CPU: the C-ABI shared library loads and exports every symbol declared in
include/sella_b200.h (no compute call is made: there is no GPU here)."""
import ctypes
import os

import pytest

from sella_b200 import _lib


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(_lib.LIB_PATH):
        from sella_b200 import _build
        _build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _lib.declared_symbols()
    assert len(names) >= 4
    for name in names:
        assert hasattr(lib, name), name
    assert lib.sb_version() >= 1


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sella_b200 import kernels
    with pytest.raises(_lib.SellaB200Error):
        kernels.hv(torch.zeros(1, 4, 4, dtype=torch.float64), torch.zeros(1, 1, 4, dtype=torch.float64))
