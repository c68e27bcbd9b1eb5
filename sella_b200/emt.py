"""On-device EMT-form copper surface: the PES plug-in of the EMT configurations
(BASELINE.json C1-C3).  Any object with this ``evaluate`` signature can be
handed to ``BatchedSella``; this one replaces the host ASE calculator call of the reference
(sella/peswrapper.py:413-418) by ``sb_emt_pes`` (sella_b200/csrc/emt.cu).

The functional form is the published effective-medium theory of Jacobsen, Stoltze and
Norskov with their copper constants; it is "EMT-form" -- ASE's own implementation is not
part of the reference tree and numerical equality with it is not claimed.
"""
import ctypes

import numpy as np
import torch

from ._lib import I, LL, _p, _stream, call, check_f64

BOHR = 0.52917721
# E0 [eV], s0 [A], V0 [eV], eta2 [1/A], kappa [1/A], lambda [1/A]
CU_PARAMETERS = (-3.51, 2.67 * BOHR, 2.476, 1.652 / BOHR, 2.74 / BOHR, 1.906 / BOHR)
_BETA = 1.809


def _list_cutoff(par):
    return _BETA * par[1] * 0.5 * (np.sqrt(3.0) + 2.0) + 0.5


class EMTSurface:
    def __init__(self, batch, natoms, device, cell=None, pbc=(False, False, False), parameters=CU_PARAMETERS):
        self.batch, self.natoms, self.n = int(batch), int(natoms), 3 * int(natoms)
        self._par = (ctypes.c_double * 6)(*[float(v) for v in parameters])
        self.neval = 0
        nimg = [0, 0, 0]
        self._cell = None
        self._cellstride = 0
        if cell is not None and any(pbc):
            c = np.asarray(cell, dtype=np.float64)
            per_system = c.ndim == 3
            cells = c if per_system else c.reshape(1, 3, 3)
            rlist = _list_cutoff(parameters)
            for cc in cells:                                   # image ranges that cover every cell of the batch
                vol = abs(np.linalg.det(cc))
                for d in range(3):
                    if pbc[d]:
                        h = vol / np.linalg.norm(np.cross(cc[(d + 1) % 3], cc[(d + 2) % 3]))
                        nimg[d] = max(nimg[d], int(np.ceil(rlist / h)))
            self._cell = torch.from_numpy(np.ascontiguousarray(cells.reshape(-1, 9))).to(device)
            self._cellstride = 9 if per_system else 0
        self._nimg = (ctypes.c_int * 3)(*nimg)

    def evaluate(self, x, f_out, g_out, active=None):
        check_f64(x, f_out, g_out)
        self.neval += 1
        call("sb_emt_pes", _p(x), I(self.natoms), _p(self._cell), LL(self._cellstride), self._nimg, self._par,
             _p(f_out), _p(g_out), _p(active), I(self.batch), _stream())
