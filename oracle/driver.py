"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the reference's optimiser policy, ``sella/optimize/optimize.py``:

  defaults by order            optimize.py:20-39, 181-194
  _predict_step                optimize.py:317-357
  step (re-diagonalise test,   optimize.py:359-440
        kick, trust radius)
  converged                    optimize.py:445-455

and the ASE ``Optimizer.run`` loop it plugs into (third-party, not in the
reference tree): ``while not converged(): step(); nsteps += 1``.

``SaddleSearch`` drives an ``oracle.pes.CartesianPES``; the benchmark's
``cpu_baseline`` leg times exactly this loop.

Pinned by running the reference's own ``Sella`` + ``PES`` classes on the same
inputs (``tests/golden/make_golden.py`` -> ``tests/golden/loop_*.npz``).
"""
import numpy as np

from .restricted import get_restricted_step

DEFAULTS = dict(
    minimum=dict(delta0=1e-1, sigma_inc=1.15, sigma_dec=0.90, rho_inc=1.035,
                 rho_dec=100, method="qn", eig=False),
    saddle=dict(delta0=0.1, sigma_inc=1.15, sigma_dec=0.65, rho_inc=1.035,
                rho_dec=5.0, method="prfo", eig=True),
)


class SaddleSearch:
    def __init__(self, pes, order=1, delta0=None, sigma_inc=None, sigma_dec=None,
                 rho_dec=None, rho_inc=None, eig=None, eta=1e-4, method=None,
                 gamma=0.1, threepoint=False, rs=None, nsteps_per_diag=3,
                 diag_every_n=None, diag_maxiter=None):
        d = DEFAULTS["minimum" if order == 0 else "saddle"]
        self.pes, self.ord = pes, order
        if rs is None:
            rs = "ras"
        self.rs = get_restricted_step(rs)
        delta0 = d["delta0"] if delta0 is None else delta0
        self.delta = delta0 if rs in ("mis", "ras") else delta0 * pes.get_Ufree().shape[1]
        self.sigma_inc = d["sigma_inc"] if sigma_inc is None else sigma_inc
        self.sigma_dec = d["sigma_dec"] if sigma_dec is None else sigma_dec
        self.rho_inc = d["rho_inc"] if rho_inc is None else rho_inc
        self.rho_dec = d["rho_dec"] if rho_dec is None else rho_dec
        self.method = d["method"] if method is None else method
        self.eig = d["eig"] if eig is None else eig
        self.delta_min = eta
        self.diagkwargs = dict(gamma=gamma, threepoint=threepoint)
        if diag_maxiter is not None:
            # not a Sella kwarg: fixes the Davidson work per diagonalisation for
            # the benchmark, forwarded to PES.diag(maxiter=...)
            self.diagkwargs["maxiter"] = diag_maxiter
        self.rho = 1.0
        self.initialized = False
        self.nsteps_per_diag = nsteps_per_diag
        self.nsteps_since_diag = 0
        self.diag_every_n = np.inf if diag_every_n is None else diag_every_n
        self.nsteps = 0
        self.history = []

    def _predict_step(self):
        pes = self.pes
        if not self.initialized:
            pes.get_g()
            if self.eig:
                if getattr(pes, "hessian_function", None) is not None:     # optimize.py:321-324
                    pes.calculate_hessian()
                else:
                    pes.diag(**self.diagkwargs)
                self.nsteps_since_diag = -1
            self.initialized = True
        pes._update_basis()
        return self.rs(pes, self.ord, self.delta, method=self.method).get_s()

    def step(self):
        pes = self.pes
        s, smag = self._predict_step()

        if self.nsteps_since_diag >= self.diag_every_n:
            ev = True
        elif self.eig and self.nsteps_since_diag >= self.nsteps_per_diag:
            if pes.H.evals is None:
                ev = True
            else:
                ev = bool((pes.get_HL_projected(pes.get_Unred()).evals[:self.ord] > 0).any())
        else:
            ev = False
        self.nsteps_since_diag = 0 if ev else self.nsteps_since_diag + 1

        rho = pes.kick(s, ev, **self.diagkwargs)

        if rho is None:
            self.rho = 1.0
        else:
            if rho < 1.0 / self.rho_dec or rho > self.rho_dec:
                self.delta = max(smag * self.sigma_dec, self.delta_min)
            elif 1.0 / self.rho_inc < rho < self.rho_inc:
                self.delta = max(self.sigma_inc * smag, self.delta)
            self.rho = rho
        self.history.append(dict(s=s.copy(), smag=smag, ev=ev, rho=self.rho,
                                 delta=self.delta, x=pes.get_x(), f=pes.curr["f"]))

    def run(self, fmax=0.05, steps=100000):
        while self.nsteps < steps and not self.pes.converged(fmax)[0]:
            self.step()
            self.nsteps += 1
        return self.pes.converged(fmax)[0]
