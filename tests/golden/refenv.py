"""
TEST INFRASTRUCTURE ONLY.  Imports the reference's own modules

    thetis/utility.py  equation.py  shallowwater_eq.py  tracer_eq_2d.py  timeintegrator.py  rungekutta.py
    coupled_timeintegrator_2d.py
    physical_constants.py  field_defs.py  log.py

from /root/reference and lets them run on top of `ufl_lite` (the numpy stand-in for `firedrake` / `ufl`): the stand-in
modules are registered in `sys.modules` under the names the reference imports, and `thetis` itself is registered as an
empty package whose `__path__` points at the reference tree, so that `thetis/__init__.py` (which pulls in the whole
model: traitlets, exporters, 3-D solver ...) is NOT executed while `from .utility import *` etc. resolve to the
reference's files.  Nothing of the reference is copied or modified.

Only usable where /root/reference exists (the build container); the GPU box never imports this.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("THETIS_REFERENCE", "/root/reference")

if HERE not in sys.path:
    sys.path.insert(0, HERE)
import ufl_lite as U   # noqa: E402

_STUBBED = ["firedrake", "firedrake.petsc", "ufl", "ufl.algorithms", "mpi4py", "pyop2", "pyop2.profiling", "pyadjoint", "pyadjoint.tape",
            "thetis"]


class _Comm:
    rank, size = 0, 1

    def allreduce(self, x, op=None):
        return x

    def barrier(self):
        pass

    Get_rank = lambda self: 0
    Get_size = lambda self: 1


class _Log:
    @staticmethod
    def EventDecorator(*a, **k):
        return lambda f: f

    @staticmethod
    def Event(*a, **k):
        return types.SimpleNamespace(begin=lambda: None, end=lambda: None)


class _Timed:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def install():
    """Register the stand-in modules; returns the dict of reference modules (utility, equation, ...)."""
    for name in list(sys.modules):
        if name == "thetis" or name.startswith("thetis."):
            del sys.modules[name]
    fd = types.ModuleType("firedrake")
    for k in dir(U):
        if not k.startswith("_"):
            setattr(fd, k, getattr(U, k))
    fd.Dx = U.Dx_
    fd.COMM_WORLD = _Comm()
    fd.firedrake = fd                           # utility.py refers to `firedrake.HDivElement` after `from firedrake import *`
    fd.op2 = types.SimpleNamespace()
    fd.__all__ = [k for k in vars(fd) if not k.startswith("_")]
    petsc = types.ModuleType("firedrake.petsc")
    petsc.PETSc = types.SimpleNamespace(Log=_Log, Sys=types.SimpleNamespace(Print=print))
    fd.petsc = petsc
    ufl = types.ModuleType("ufl")
    for k in dir(U):
        if not k.startswith("_"):
            setattr(ufl, k, getattr(U, k))
    alg = types.ModuleType("ufl.algorithms")
    alg.estimate_total_polynomial_degree = U.estimate_total_polynomial_degree
    ufl.algorithms = alg
    sys.modules["ufl.algorithms"] = alg
    mpi = types.ModuleType("mpi4py")
    mpi.MPI = types.SimpleNamespace(MIN="min", MAX="max", SUM="sum", COMM_WORLD=fd.COMM_WORLD)
    pyop2 = types.ModuleType("pyop2")
    prof = types.ModuleType("pyop2.profiling")
    prof.timed_stage = _Timed
    prof.timed_region = _Timed
    pyop2.profiling = prof
    pyad = types.ModuleType("pyadjoint")
    tape = types.ModuleType("pyadjoint.tape")
    tape.no_annotations = lambda f: f
    pyad.tape = tape
    pkg = types.ModuleType("thetis")
    pkg.__path__ = [os.path.join(REF_ROOT, "thetis")]
    pkg.__package__ = "thetis"
    for name, mod in (("firedrake", fd), ("firedrake.petsc", petsc), ("ufl", ufl), ("mpi4py", mpi), ("pyop2", pyop2),
                      ("pyop2.profiling", prof), ("pyadjoint", pyad), ("pyadjoint.tape", tape), ("thetis", pkg)):
        sys.modules[name] = mod
    mods = {}
    for name in ("utility", "equation", "shallowwater_eq", "tracer_eq_2d", "timeintegrator", "rungekutta",
                 "coupled_timeintegrator_2d"):
        mods[name] = importlib.import_module("thetis." + name)
        assert os.path.realpath(mods[name].__file__).startswith(os.path.realpath(REF_ROOT)), mods[name].__file__
    mods["physical_constants"] = sys.modules["thetis.physical_constants"].physical_constants
    return mods


def uninstall():
    for name in list(sys.modules):
        if name in _STUBBED or name.startswith("thetis."):
            del sys.modules[name]


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "thetis"))
