"""CUDA mirror of sella/optimize/restricted_step.py + stepper.py at the level the
reference's callers use them: ``get_restricted_step(name)(pes, order, delta,
method=...).get_s() -> (s, smag)`` (optimize.py:339-355).

The whole alpha search runs on the device (csrc/trstep.cu, csrc/rfo.cu).  ``pes`` is the
reference's duck type (restricted_step.py:28-62): get_g, get_scons, get_H (with .B or
.asarray()), get_Ufree, attribute ``int``.  Supported here: Ufree = identity, scons = 0
(Cartesian, no constraints); the batched engine is the route for everything else."""
import numpy as np
import torch

from .. import kernels as K
from .._host import up, up_mat, zeros, raise_status
from .._lib import I, _p, _stream, call

_QN = ('qn', 'quasi-newton', 'quasi newton', 'newton', 'mmf', 'minimum mode following',
       'minimum-mode following', 'dimer')
_RFO = ('rfo', 'rational function optimization')
_PRFO = ('prfo', 'p-rfo', 'partitioned rational function optimization')


def _model(method):
    m = method.lower()
    if m in _QN:
        return "qn"
    if m in _RFO:
        return "rfo"
    if m in _PRFO:
        return "prfo"
    raise ValueError("Unknown stepper name: {}".format(method))


class BaseRestrictedStep:
    synonyms = []
    kind = None

    def __init__(self, pes, order, delta, method='qn', tol=None, maxiter=1000, d1=None, W=None):
        if d1 is not None or W is not None:
            raise NotImplementedError("IRC / weighted steps are not on the CUDA path")
        self.pes, self.order, self.delta = pes, int(order), float(delta)
        self.model = _model(method)
        g = np.asarray(pes.get_g(), dtype=np.float64)
        n = len(g)
        scons = np.asarray(pes.get_scons())
        Ufree = np.asarray(pes.get_Ufree())
        if np.abs(scons).max(initial=0.0) > 0 or Ufree.shape != (n, n) or not np.array_equal(Ufree, np.eye(n)):
            raise NotImplementedError("constrained restricted steps: use sella_b200.batched.BatchedSella")
        H = pes.get_H()
        B = H.asarray() if hasattr(H, "asarray") else np.asarray(H)
        self._B = up_mat(B)
        self._g = up(g).view(1, 1, n)
        self.n = n

    def get_s(self):
        n = self.n
        evals, Vt, status = K.eigh(self._B)
        Vg = K.hv(Vt, self._g)
        d = torch.full((1,), self.delta, dtype=torch.float64, device=Vg.device)
        coef, smag, alpha = zeros(1, n), zeros(1), zeros(1)
        if self.kind == "tr":
            if self.model == "qn":
                call("sb_qn_tr", _p(Vg), _p(evals), _p(d), I(self.order), I(n), _p(coef), _p(smag), _p(alpha),
                     _p(status), _p(None), _p(None), I(1), _stream())
            else:
                call("sb_rfo_tr", _p(Vg), _p(evals), _p(d), I(self.order), I(n), I(int(self.model == "prfo")),
                     _p(coef), _p(smag), _p(alpha), _p(status), _p(None), _p(None), I(1), _stream())
            s = K.hv(Vt, coef.view(1, 1, n), transposed=True).view(1, n)
        else:
            if getattr(self.pes, "int", None) is not None:
                raise ValueError("Internal coordinates are not compatible with the RestrictedAtomicStep "
                                 "trust region method.")
            s = zeros(1, n)
            if self.model == "qn":
                call("sb_qn_ras", _p(Vg), _p(evals), _p(Vt), _p(d), I(self.order), I(n), _p(s), _p(smag), _p(alpha),
                     _p(status), _p(None), _p(None), I(1), _stream())
            else:
                call("sb_rfo_ras", _p(Vg), _p(evals), _p(Vt), _p(d), I(self.order), I(n),
                     I(int(self.model == "prfo")), _p(s), _p(smag), _p(alpha), _p(status), _p(None), _p(None),
                     I(1), _stream())
        raise_status(status, "restricted step")
        self.alpha = float(alpha[0])
        return s[0].cpu().numpy(), float(smag[0])

    @classmethod
    def match(cls, name):
        return name in cls.synonyms


class TrustRegion(BaseRestrictedStep):
    synonyms = ['tr', 'trust region', 'trust-region', 'trust radius', 'trust-radius']
    kind = "tr"


class RestrictedAtomicStep(BaseRestrictedStep):
    synonyms = ['ras', 'restricted atomic step']
    kind = "ras"


_all_restricted_step = [TrustRegion, RestrictedAtomicStep]


def get_restricted_step(name):
    for rs in _all_restricted_step:
        if rs.match(name):
            return rs
    raise ValueError("Unknown restricted step name: {}".format(name))
