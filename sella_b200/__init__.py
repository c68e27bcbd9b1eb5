"""sella_b200 — B200-native implementation of the Sella saddle-point inner loop.

Reference-named entry points (numpy in/out, one search, all arithmetic in CUDA):
    Sella                      sella_b200.optimize.optimize
    rayleigh_ritz, exact       sella_b200.eigensolvers
    update_H, symmetrize_Y     sella_b200.hessian_update
    modified_gram_schmidt      sella_b200.utilities.math
    get_restricted_step        sella_b200.optimize.restricted_step
Batched engine (many searches in lock step, state resident in HBM):
    BatchedSella, QuadraticSurface   sella_b200.batched
C ABI: include/sella_b200.h, sella_b200/csrc/libsella_b200.so.
"""


def __getattr__(name):          # lazy: importing the package must not need torch/CUDA
    if name == "Sella":
        from .optimize.optimize import Sella
        return Sella
    if name == "Constraints":
        from .constraints import Constraints
        return Constraints
    if name in ("BatchedSella", "QuadraticSurface"):
        from . import batched
        return getattr(batched, name)
    raise AttributeError(name)
