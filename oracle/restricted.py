"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the reference's restricted-step solve,
``sella/optimize/restricted_step.py``:

  BaseRestrictedStep.__init__   restricted_step.py:14-67   (reduced problem)
  eval / get_s                  restricted_step.py:72-121  (safeguarded Newton on alpha)
  TrustRegion.cons              restricted_step.py:136-142
  RestrictedAtomicStep.cons     restricted_step.py:172-183
  MaxInternalStep.cons          restricted_step.py:206-216

``pes`` is the duck type documented in SURVEY.md section 8b (``oracle/pes.py``
provides one).

Pinned against the reference through ``tests/golden`` (make_golden.py).
"""
import numpy as np

from .stepper import get_stepper, Naive, StepModel


class RestrictedStep:
    names = ()

    def __init__(self, pes, order, delta, method="qn", tol=None, maxiter=1000,
                 W=None):
        self.pes, self.delta = pes, delta
        g0 = pes.get_g()
        self.scons = pes.get_scons()
        g = g0 + pes.get_H() @ self.scons

        model = method if (isinstance(method, type) and issubclass(method, StepModel)) \
            else get_stepper(method.lower())

        if self.cons(self.scons) - delta > 1e-8:
            # constraint violation alone exceeds the radius: scale it (:44-48)
            self.P = pes.get_Unred().T
            self.stepper = Naive(self.P @ self.scons)
            self.scons[:] *= 0
        else:
            self.P = pes.get_Ufree().T if W is None else pes.get_Ufree().T @ W
            self.stepper = model(self.P @ g, pes.get_HL_projected(self.P.T), order)

        self.tol = tol if tol is not None else (1e-10 if self.stepper.newton_safe else 1e-15)
        self.maxiter = maxiter

    def cons(self, s, dsda=None):
        raise NotImplementedError

    def eval(self, alpha):
        s, dsda = self.stepper.get_s(alpha)
        stot = self.P.T @ s + self.scons
        val, dval = self.cons(stot, self.P.T @ dsda)
        return stot, val, dval

    def get_s(self, trace=None):
        st = self.stepper
        alpha = st.alpha0
        s, val, dval = self.eval(alpha)
        if val < self.delta:
            return s, val                             # interior step
        err = val - self.delta
        lo, hi = st.alphamin, st.alphamax
        for it in range(self.maxiter):
            if abs(err) <= self.tol:
                break
            if np.nextafter(lo, hi) >= hi:
                break
            if err * st.slope > 0:
                hi = alpha
            else:
                lo = alpha
            newton = alpha - err / dval
            if (np.isnan(newton) or newton <= lo or newton >= hi
                    or (it > 4 and not st.newton_safe)):
                mid = (lo + hi) / 2.0
                if np.isinf(mid):
                    alpha = alpha + max(1, 0.5 * alpha) * np.sign(mid)
                else:
                    alpha = mid
            else:
                alpha = newton
            s, val, dval = self.eval(alpha)
            err = val - self.delta
            if trace is not None:
                trace.append((alpha, val))
        else:
            raise RuntimeError("Restricted step failed to converge!")
        self.alpha = alpha
        return s, self.delta


class TrustRegion(RestrictedStep):
    names = ("tr", "trust region", "trust-region", "trust radius", "trust-radius")

    def cons(self, s, dsda=None):
        val = np.linalg.norm(s)
        if dsda is None:
            return val
        return val, dsda @ s / max(val, 1e-12)


class RestrictedAtomicStep(RestrictedStep):
    names = ("ras", "restricted atomic step")

    def __init__(self, pes, *a, **kw):
        if pes.int is not None:
            raise ValueError("Internal coordinates are not compatible with the "
                             "RestrictedAtomicStep trust region method.")
        super().__init__(pes, *a, **kw)

    def cons(self, s, dsda=None):
        per_atom = s.reshape((-1, 3))
        norms = np.linalg.norm(per_atom, axis=1)
        a = int(np.argmax(norms))
        val = norms[a]
        if dsda is None:
            return val
        return val, dsda.reshape((-1, 3))[a] @ per_atom[a] / max(val, 1e-12)


class MaxInternalStep(RestrictedStep):
    names = ("mis", "max internal step")

    def __init__(self, pes, *a, wx=1., wb=1., wa=1., wd=1., wo=1., wc=1., **kw):
        if pes.int is None:
            raise ValueError("Internal coordinates are required for the "
                             "MaxInternalStep trust region method")
        i = pes.int
        self.w = np.array([wx] * i.ntrans + [wb] * i.nbonds + [wa] * i.nangles
                          + [wd] * i.ndihedrals + [wo] * i.nother
                          + [wx] * i.nrotations)
        super().__init__(pes, *a, **kw)

    def cons(self, s, dsda=None):
        sw = np.abs(s * self.w)
        i = int(np.argmax(sw))
        if dsda is None:
            return sw[i]
        return sw[i], np.sign(s[i]) * dsda[i] * self.w[i]


_ALL = (TrustRegion, RestrictedAtomicStep, MaxInternalStep)


def get_restricted_step(name):
    for cls in _ALL:
        if name in cls.names:
            return cls
    raise ValueError("Unknown restricted step name: {}".format(name))
