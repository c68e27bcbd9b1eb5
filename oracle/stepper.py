"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the reference's step models, ``sella/optimize/stepper.py``:

  NaiveStepper                    stepper.py:44-55
  QuasiNewton                     stepper.py:58-96
  RationalFunctionOptimization    stepper.py:114-157
  PartitionedRFO                  stepper.py:160-185
  get_stepper                     stepper.py:188-199

Each model maps a scalar damping parameter alpha to a step ``s(alpha)`` and its
derivative ``ds/dalpha``; class attributes ``alpha0/alphamin/alphamax/slope/
newton_safe`` drive the root-finder in ``oracle/restricted.py``.

``H`` is any object with ``evals``, ``evecs``, ``asarray()``, ``project(U)`` (the
duck type of the reference's ``ApproximateHessian``; see ``oracle/pes.py``).

Pinned against the reference through ``tests/golden`` (make_golden.py).
"""
import numpy as np
from scipy.linalg import eigh


class StepModel:
    alpha0 = None
    alphamin = None
    alphamax = None
    slope = None
    newton_safe = True
    names = ()

    def __init__(self, g, H, order=0, d1=None):
        self.g, self.H, self.order, self.d1 = g, H, order, d1
        self.setup()


class Naive(StepModel):
    """Straight-line constraint-restoring step (stepper.py:44-55)."""
    alpha0, alphamin, alphamax, slope = 0.5, 0.0, 1.0, 1.0

    def __init__(self, dx):
        self.dx = dx

    def get_s(self, alpha):
        return alpha * self.dx, self.dx


class QuasiNewton(StepModel):
    """s(a) = -V [ (V^T g) / (L + a*sigma) ], L=|lambda| with the lowest `order`
    entries negated, sigma=+-1 likewise (stepper.py:75-96)."""
    alpha0, alphamin, alphamax, slope = 0.0, 0.0, np.inf, -1
    names = ("qn", "quasi-newton", "quasi newton", "newton", "mmf",
             "minimum mode following", "minimum-mode following", "dimer")

    def setup(self):
        if self.H.evals is None:
            self.H.evals, self.H.evecs = eigh(self.H.asarray())
        self.L = np.abs(self.H.evals)
        self.L[:self.order] *= -1
        self.sigma = np.ones_like(self.L)
        self.sigma[:self.order] = -1
        self.V = self.H.evecs
        self.Vg = self.V.T @ self.g

    def get_s(self, alpha):
        den = self.L + alpha * self.sigma
        c = self.Vg / den
        return -self.V @ c, self.V @ (c / den)


class RFO(StepModel):
    """Eigenvector `order` of the scaled bordered matrix
    [[a^2 H, a g], [a g^T, 0]]  ->  s = a * v[:-1] / v[-1]; derivative from
    first-order eigenvector perturbation theory (stepper.py:122-157)."""
    alpha0, alphamin, alphamax, slope = 1.0, 0.0, 1.0, 1.0
    newton_safe = False
    names = ("rfo", "rational function optimization")

    def setup(self):
        g = self.g
        self.K = np.block([[self.H.asarray(), g[:, None]],
                           [g[None, :], np.zeros((1, 1))]])

    def get_s(self, alpha):
        o = self.order
        Ka = self.K * alpha
        Ka[:-1, :-1] *= alpha
        w, Z = eigh(Ka)
        z = Z[:, o]
        last = z[-1]
        if abs(last) < 1e-12:
            last = np.sign(last) * 1e-12 if last != 0 else 1e-12
        s = z[:-1] * alpha / last

        dK = self.K.copy()
        dK[:-1, :-1] *= 2 * alpha
        Zo = np.delete(Z, o, 1)
        gap = np.delete(w, o) - w[o]
        gap = np.where(gap >= 0, np.maximum(gap, 1e-12), np.minimum(gap, -1e-12))
        dz = Zo @ ((Zo.T @ (dK @ z)) / gap)
        dsda = (z[:-1] / last + (alpha / last) * dz[:-1]
                - (z[:-1] * alpha / last ** 2) * dz[-1])
        return s, dsda


class PartitionedRFO(RFO):
    """RFO maximising along the lowest `order` modes and minimising along the
    rest, both in H's eigenbasis (stepper.py:163-185)."""
    names = ("prfo", "p-rfo", "partitioned rational function optimization")

    def setup(self):
        o = self.order
        self.Vmax = self.H.evecs[:, :o]
        self.Vmin = self.H.evecs[:, o:]
        self.up = RFO(self.Vmax.T @ self.g, self.H.project(self.Vmax), order=o)
        self.down = RFO(self.Vmin.T @ self.g, self.H.project(self.Vmin), order=0)

    def get_s(self, alpha):
        su, dsu = self.up.get_s(alpha)
        sd, dsd = self.down.get_s(alpha)
        return self.Vmax @ su + self.Vmin @ sd, self.Vmax @ dsu + self.Vmin @ dsd


_MODELS = (QuasiNewton, RFO, PartitionedRFO)


def get_stepper(name):
    for cls in _MODELS:
        if name in cls.names:
            return cls
    raise ValueError("Unknown stepper name: {}".format(name))
