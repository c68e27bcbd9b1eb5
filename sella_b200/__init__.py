"""sella_b200 — B200-native implementation of the Sella saddle-point inner loop.

Reference-named entry points (numpy in/out, one search, all arithmetic in CUDA):
    Sella                      sella_b200.optimize.optimize
    rayleigh_ritz, exact       sella_b200.eigensolvers
    update_H, symmetrize_Y     sella_b200.hessian_update
    modified_gram_schmidt      sella_b200.utilities.math
    get_restricted_step        sella_b200.optimize.restricted_step
    Constraints, Internals     sella_b200.constraints, sella_b200.topology
Batched engine (many searches in lock step, state resident in HBM):
    BatchedSella, QuadraticSurface   sella_b200.batched
    BatchedInternalSella             sella_b200.batched_internal   (internal coordinates, geodesic steps)
C ABI: include/sella_b200.h, sella_b200/csrc/libsella_b200.so.
"""


def __getattr__(name):          # lazy: importing the package must not need torch/CUDA
    if name == "Sella":
        from .optimize.optimize import Sella
        return Sella
    if name == "Constraints":
        from .constraints import Constraints
        return Constraints
    if name == "Internals":
        from .topology import Internals
        return Internals
    if name == "BatchedInternalSella":
        from .batched_internal import BatchedInternalSella
        return BatchedInternalSella
    if name in ("BatchedSella", "QuadraticSurface"):
        from . import batched
        return getattr(batched, name)
    raise AttributeError(name)
