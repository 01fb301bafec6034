"""
thetis_b200 -- B200-native explicit P1DG shallow-water stepper, drop-in for the
SSPRK33 path of thetisproject/thetis (see DESIGN.md / INTEGRATION.md).

Importing the package does not need a GPU; constructing an integrator does
(there is no CPU fallback on the product path).
"""
__version__ = "0.1.0"


def install(thetis_module=None, sync_policy="every_step", wd_mass=None):
    """
    Rebind `thetis.rungekutta.SSPRK33` (+ the Butcher-form ERK classes and `thetis.timeintegrator.ForwardEuler`
    where the module has them) and `thetis.limiter.VertexBasedP1DGLimiter` to the B200
    implementations so that FlowSolver2d.create_timestepper() picks them up (the `steppers` dict is built from
    module attributes at call time, thetis/solver2d.py:662-672).  See INTEGRATION.md.
    ``wd_mass``: 'plain' | 'displaced' -- mass functional of the explicit wetting-drying step for every integrator built
    afterwards (rungekutta.WD_MASS_DEFAULT); None leaves the current default.
    """
    from . import rungekutta as rk, limiter as lim
    if wd_mass is not None:
        if wd_mass not in ("plain", "displaced"):
            raise ValueError(f"wd_mass must be 'plain' or 'displaced', not {wd_mass!r}")
        rk.WD_MASS_DEFAULT = wd_mass
    if thetis_module is None:
        import thetis as thetis_module          # raises ImportError without a Thetis/Firedrake install
    policy = sync_policy

    class SSPRK33(rk.SSPRK33):
        def __init__(self, equation, solution, fields, dt, options=None, bnd_conditions=None, terms_to_add="all"):
            super().__init__(equation, solution, fields, dt, options, bnd_conditions, terms_to_add,
                             sync_policy=policy)

    cls = SSPRK33
    thetis_module.rungekutta.SSPRK33 = cls
    thetis_module.limiter.VertexBasedP1DGLimiter = lim.VertexBasedP1DGLimiter

    def _bind(base):
        class _Bound(base):
            def __init__(self, equation, solution, fields, dt, options=None, bnd_conditions=None, terms_to_add="all"):
                super().__init__(equation, solution, fields, dt, options, bnd_conditions, terms_to_add,
                                 sync_policy=policy)
        _Bound.__name__ = _Bound.__qualname__ = base.__name__
        return _Bound

    # the Butcher-form explicit schemes (rungekutta.py:959-980) and timeintegrator.ForwardEuler, where present
    for name in ("ERKLSPUM2", "ERKLPUM2", "ERKMidpoint", "ERKEuler"):
        if hasattr(thetis_module.rungekutta, name):
            setattr(thetis_module.rungekutta, name, _bind(getattr(rk, name)))
    ti_mod = getattr(thetis_module, "timeintegrator", None)
    if ti_mod is not None and hasattr(ti_mod, "ForwardEuler"):
        ti_mod.ForwardEuler = _bind(rk.ForwardEuler)
    return cls
