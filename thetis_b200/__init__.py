"""
thetis_b200 -- B200-native explicit P1DG shallow-water stepper, drop-in for the
SSPRK33 path of thetisproject/thetis (see DESIGN.md / INTEGRATION.md).

Importing the package does not need a GPU; constructing an integrator does
(there is no CPU fallback on the product path).
"""
__version__ = "0.1.0"


def install(thetis_module=None, sync_policy="every_step"):
    """
    Rebind `thetis.rungekutta.SSPRK33` and `thetis.limiter.VertexBasedP1DGLimiter` to the B200
    implementations so that FlowSolver2d.create_timestepper() picks them up (the `steppers` dict is built from
    module attributes at call time, thetis/solver2d.py:662-672).  See INTEGRATION.md.
    """
    from . import rungekutta as rk, limiter as lim
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
    return cls
