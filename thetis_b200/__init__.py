"""
thetis_b200 -- B200-native explicit P1DG shallow-water stepper, drop-in for the
SSPRK33 path of thetisproject/thetis (see DESIGN.md / INTEGRATION.md).

Importing the package does not need a GPU; constructing an integrator does
(there is no CPU fallback on the product path).
"""
__version__ = "0.1.0"


def install(thetis_module=None, sync_policy="every_step", wd_mass=None, fallback=False):
    """
    Rebind `thetis.rungekutta.SSPRK33` (+ the Butcher-form ERK classes and `thetis.timeintegrator.ForwardEuler`
    where the module has them) and `thetis.limiter.VertexBasedP1DGLimiter` to the B200
    implementations so that FlowSolver2d.create_timestepper() picks them up (the `steppers` dict is built from
    module attributes at call time, thetis/solver2d.py:662-672).  See INTEGRATION.md.
    ``wd_mass``: 'plain' | 'displaced' -- mass functional of the explicit wetting-drying step for every integrator built
    afterwards (rungekutta.WD_MASS_DEFAULT); None leaves the current default.
    ``fallback``: the same `steppers` entries also serve equations this library does not accelerate (sediment, Exner,
    the 3-D model's explicit parts, turbines, SUPG ...).  By default (fallback off) such a construction raises
    NotImplementedError -- "outside the accelerated path" -- like everything else this library cannot do: it fails
    loudly.  With ``fallback=True`` (opt-in) it instead returns an instance of the REFERENCE class that was bound to the
    name before install(), i.e. the user's own Thetis / Firedrake object, built from the same arguments, and warns once
    per class and reason.  Only NotImplementedError at construction is treated that way: a missing CUDA library, a
    missing GPU, an error inside a kernel launch or an error the reference raises too (invalid boundary tag, ...) always
    propagate, and nothing inside this library ever computes on the CPU.
    """
    import warnings
    from . import rungekutta as rk, limiter as lim
    if wd_mass is not None:
        if wd_mass not in ("plain", "displaced"):
            raise ValueError(f"wd_mass must be 'plain' or 'displaced', not {wd_mass!r}")
        rk.WD_MASS_DEFAULT = wd_mass
    if thetis_module is None:
        import thetis as thetis_module          # raises ImportError without a Thetis/Firedrake install
    policy = sync_policy
    warned = set()

    def _reference(orig, name, err, args, kwargs):
        if not fallback or orig is None or orig is object or getattr(orig, "_thetis_b200_bound", False):
            raise err
        key = (name, str(err))
        if key not in warned:
            warned.add(key)
            warnings.warn(f"thetis_b200: {name} falls back to the reference class ({err})", RuntimeWarning, stacklevel=3)
        return orig(*args, **kwargs)

    def _unwrap(orig):
        # install() called twice: the reference class is the one remembered by the first binding
        return getattr(orig, "_reference_class", None) if getattr(orig, "_thetis_b200_bound", False) else orig

    def _bind(base, orig):
        orig = _unwrap(orig)

        class _Bound(base):
            _thetis_b200_bound = True
            _reference_class = orig

            def __new__(cls, *args, **kwargs):
                obj = object.__new__(cls)
                try:
                    base.__init__(obj, *args, sync_policy=policy, **kwargs)
                except NotImplementedError as err:
                    return _reference(orig, base.__name__, err, args, kwargs)
                return obj

            def __init__(self, *args, **kwargs):
                pass                              # constructed in __new__ (so that a fallback can replace the object)
        _Bound.__name__ = _Bound.__qualname__ = base.__name__
        return _Bound

    cls = _bind(rk.SSPRK33, getattr(thetis_module.rungekutta, "SSPRK33", None))
    thetis_module.rungekutta.SSPRK33 = cls

    orig_lim = _unwrap(getattr(thetis_module.limiter, "VertexBasedP1DGLimiter", None))
    if fallback and orig_lim is not None and orig_lim is not object and not getattr(orig_lim, "_thetis_b200_bound", False):
        class VertexBasedP1DGLimiter(lim.VertexBasedP1DGLimiter):
            _thetis_b200_bound = True
            _reference_class = orig_lim

            def __new__(cls, *args, **kwargs):
                obj = object.__new__(cls)
                try:
                    lim.VertexBasedP1DGLimiter.__init__(obj, *args, **kwargs)
                except NotImplementedError as err:      # vector fields, extruded meshes: the reference's 3-D path
                    return _reference(orig_lim, "VertexBasedP1DGLimiter", err, args, kwargs)
                return obj

            def __init__(self, *args, **kwargs):
                pass
        thetis_module.limiter.VertexBasedP1DGLimiter = VertexBasedP1DGLimiter
    else:
        thetis_module.limiter.VertexBasedP1DGLimiter = lim.VertexBasedP1DGLimiter

    # the Butcher-form explicit schemes (rungekutta.py:959-980) and timeintegrator.ForwardEuler, where present
    for name in ("ERKLSPUM2", "ERKLPUM2", "ERKMidpoint", "ERKEuler"):
        if hasattr(thetis_module.rungekutta, name):
            setattr(thetis_module.rungekutta, name, _bind(getattr(rk, name), getattr(thetis_module.rungekutta, name)))
    ti_mod = getattr(thetis_module, "timeintegrator", None)
    if ti_mod is not None and hasattr(ti_mod, "ForwardEuler"):
        ti_mod.ForwardEuler = _bind(rk.ForwardEuler, ti_mod.ForwardEuler)
    return cls
