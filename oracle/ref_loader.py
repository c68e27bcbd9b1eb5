"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference (zadorlab/sella)
piecewise from /root/reference, without copying any of its sources.

Why a loader: ``import sella`` itself is impossible in this image because
``sella/__init__.py`` imports jax and ``optimize.py`` / ``peswrapper.py`` /
``internal.py`` import ase, neither of which is installed.
The arithmetic on the hot path is pure numpy/scipy (+ one Cython module), so we

  * pre-insert an empty ``sella`` package object whose ``__path__`` points at
    ``/root/reference/sella`` (the real ``__init__`` is never executed),
  * compile ``sella/utilities/math.pyx`` *from where it lies* into
    ``oracle/_ref/`` (binary only; the generated C goes to a temp dir),
  * stub the names of ``ase`` and ``sella.internal`` that ``peswrapper.py`` and
    ``optimize/optimize.py`` import at module scope (only type names and the
    ``Optimizer`` base class; no numerics live in those stubs).

After ``load()`` the following reference modules are the genuine article:
``sella.eigensolvers``, ``sella.hessian_update``, ``sella.linalg``, ``sella._gpu``
(forced to its CPU path by SELLA_DISABLE_GPU=1), ``sella.utilities.math``,
``sella.optimize.stepper``, ``sella.optimize.restricted_step``,
``sella.peswrapper`` (class ``PES``), ``sella.optimize.optimize`` (class ``Sella``).

This only works inside the build container (``/root/reference`` is not shipped
to the GPU box).  It is used by ``tests/golden/make_golden.py`` to produce the
committed fixtures and by the ``not gpu`` tests that cross-check the oracle port
when the reference is present.
"""
from __future__ import annotations

import importlib
import os
import subprocess
import sys
import sysconfig
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref")
# git-ignored staging area that travels to the GPU box (the base contract's baseline/_ref): the
# reference's own files on the benchmarked path, copied there by stage() at build time
STAGED_ROOT = os.path.join(os.path.dirname(HERE), "baseline", "_ref")
STAGED_FILES = ["eigensolvers.py", "hessian_update.py", "linalg.py", "_gpu.py", "peswrapper.py",
                "utilities/math.pyx", "utilities/math.pxd", "optimize/stepper.py",
                "optimize/restricted_step.py", "optimize/optimize.py"]


def _pick_root():
    env = os.environ.get("SELLA_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile(os.path.join("/root/reference", "sella", "eigensolvers.py")):
        return "/root/reference"
    return STAGED_ROOT


REF_ROOT = _pick_root()


def stage(force: bool = False) -> str:
    """Copy the reference's files on the benchmarked path (unmodified) from /root/reference into the
    git-ignored baseline/_ref/, so that `bench.py --impl reference` and the cpu_baseline leg can run the
    reference's OWN classes on the GPU box, where /root/reference does not exist."""
    import shutil
    src_root = os.path.join("/root/reference", "sella")
    if not os.path.isdir(src_root):
        raise RuntimeError("reference tree not present; nothing to stage")
    for rel in STAGED_FILES:
        src = os.path.join(src_root, rel)
        dst = os.path.join(STAGED_ROOT, "sella", rel)
        if not os.path.isfile(src):
            continue
        if force or not os.path.isfile(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return STAGED_ROOT


def source() -> str:
    """Where load() takes the reference from: 'reference tree' or 'staged copy' (baseline/_ref)."""
    return "staged copy (baseline/_ref)" if os.path.abspath(REF_ROOT) == os.path.abspath(STAGED_ROOT) \
        else "reference tree (%s)" % REF_ROOT

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "sella", "eigensolvers.py"))


def build_ref_math(force: bool = False) -> str:
    """Cython-compile the reference's utilities/math.pyx in place -> oracle/_ref/.

    Only the shared object is kept; nothing from the reference is copied.
    """
    outdir = os.path.join(REF_BIN, "sella", "utilities")
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(outdir, "math" + suffix)
    if os.path.isfile(target) and not force:
        return target
    if not available():
        raise RuntimeError("reference tree not present; cannot build oracle/_ref")
    os.makedirs(outdir, exist_ok=True)
    import numpy as np
    pyx = os.path.join(REF_ROOT, "sella", "utilities", "math.pyx")
    with tempfile.TemporaryDirectory() as tmp:
        cfile = os.path.join(tmp, "math.c")
        # module name must be sella.utilities.math (Cython checks the import name)
        subprocess.check_call([
            sys.executable, "-m", "cython", "-3", "--module-name",
            "sella.utilities.math", "-I", os.path.join(REF_ROOT),
            "-o", cfile, pyx])
        inc = sysconfig.get_paths()["include"]
        subprocess.check_call([
            "gcc", "-O2", "-fPIC", "-shared", "-fwrapv", "-Wno-deprecated-declarations",
            "-I", inc, "-I", np.get_include(), cfile, "-o", target])
    return target


class _Stub(types.ModuleType):
    """Module whose unknown attributes resolve to inert placeholder classes."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {})
        setattr(self, name, cls)
        return cls


def _install_ase_stubs():
    """Names imported at module scope by peswrapper.py:7-12 and optimize.py:9-12."""
    if "ase" in sys.modules and not isinstance(sys.modules["ase"], _Stub):
        return  # a real ase is present; use it
    names = ["ase", "ase.build", "ase.utils", "ase.visualize", "ase.calculators",
             "ase.calculators.singlepoint", "ase.io", "ase.io.trajectory",
             "ase.optimize", "ase.optimize.optimize", "ase.data", "ase.geometry",
             "ase.constraints", "ase.units", "ase.cell"]
    for nm in names:
        mod = _Stub(nm)
        mod.__path__ = []
        sys.modules[nm] = mod
    sys.modules["ase.utils"].basestring = str

    class Optimizer:
        """Minimal stand-in for ase.optimize.optimize.Optimizer: the run loop
        only (`while not converged: step(); nsteps += 1; log()`), which is
        third-party ASE code, not Sella's."""

        def __init__(self, atoms, restart=None, logfile=None, trajectory=None,
                     master=None, **kw):
            self.atoms = atoms
            self.optimizable = atoms
            self.logfile = None
            self.nsteps = 0
            self.max_steps = 0
            self.fmax = None

        def closelater(self, f):
            return f

        def run(self, fmax=0.05, steps=100000):
            self.fmax = fmax
            self.max_steps = steps
            self.log()
            while not self.converged() and self.nsteps < steps:
                self.step()
                self.nsteps += 1
                self.log()
            return self.converged()

    sys.modules["ase.optimize.optimize"].Optimizer = Optimizer


def load():
    """Return a namespace of genuine reference modules (see module docstring)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree %s not present" % REF_ROOT)
    os.environ["SELLA_DISABLE_GPU"] = "1"
    build_ref_math()
    _install_ase_stubs()

    ref_pkg = os.path.join(REF_ROOT, "sella")
    sella = types.ModuleType("sella")
    sella.__path__ = [ref_pkg]
    sys.modules["sella"] = sella
    util = types.ModuleType("sella.utilities")
    util.__path__ = [os.path.join(REF_BIN, "sella", "utilities"),
                     os.path.join(ref_pkg, "utilities")]
    sys.modules["sella.utilities"] = util
    sella.utilities = util
    opt = types.ModuleType("sella.optimize")
    opt.__path__ = [os.path.join(ref_pkg, "optimize")]
    sys.modules["sella.optimize"] = opt
    sella.optimize = opt

    # sella.internal needs jax; peswrapper/optimize only need these names.
    internal = _Stub("sella.internal")

    class DuplicateInternalError(ValueError):
        pass

    internal.DuplicateInternalError = DuplicateInternalError
    sys.modules["sella.internal"] = internal
    sella.internal = internal

    ns = types.SimpleNamespace()
    sys.dont_write_bytecode, old = True, sys.dont_write_bytecode
    try:
        ns.math = importlib.import_module("sella.utilities.math")
        ns.hessian_update = importlib.import_module("sella.hessian_update")
        ns.eigensolvers = importlib.import_module("sella.eigensolvers")
        ns.linalg = importlib.import_module("sella.linalg")
        ns.stepper = importlib.import_module("sella.optimize.stepper")
        ns.peswrapper = importlib.import_module("sella.peswrapper")
        ns.restricted_step = importlib.import_module("sella.optimize.restricted_step")
        ns.optimize = importlib.import_module("sella.optimize.optimize")
    finally:
        sys.dont_write_bytecode = old
    ns.DuplicateInternalError = DuplicateInternalError
    _loaded = ns
    return ns


if __name__ == "__main__":
    r = load()
    print("reference modules loaded:", sorted(k for k in vars(r) if not k.startswith("_")))
