"""
Descriptor classes with the reference's names for the stand-alone harness.

With real Thetis the integrator receives thetis.shallowwater_eq.ShallowWaterEquations
/ thetis.tracer_eq_2d.TracerEquation2D objects and reads only
`.function_space`, `.depth`, `.options` and the class name from them.  These
look-alikes carry the same attributes (no UFL forms: the forms ARE the CUDA
kernels).  Reference: thetis/shallowwater_eq.py:893-928, thetis/tracer_eq_2d.py:448-488,
thetis/utility.py:936-996, thetis/physical_constants.py.
"""
from __future__ import annotations

from .shim import Constant

__all__ = ["physical_constants", "DepthExpression", "ShallowWaterEquations", "ModeSplit2DEquations", "TracerEquation2D"]

# mutable Constants, read live at every stage like the UFL forms do
# (thetis/physical_constants.py:37-45; test/swe2d/test_rossby_wave.py:153-155 mutates g_grav)
physical_constants = {
    "g_grav": Constant(9.81),
    "rho0": Constant(1000.0),
    "von_karman": Constant(0.4),       # thetis/physical_constants.py:9
}


class DepthExpression:
    """Holds depth options exactly like thetis/utility.py:936-973."""

    def __init__(self, bathymetry_2d, use_nonlinear_equations=True, use_wetting_and_drying=False,
                 wetting_and_drying_alpha=0.5):
        self.bathymetry_2d = bathymetry_2d
        self.use_nonlinear_equations = use_nonlinear_equations
        self.use_wetting_and_drying = use_wetting_and_drying
        self.wetting_and_drying_alpha = wetting_and_drying_alpha


class ShallowWaterEquations:
    """2D depth-averaged shallow water equations in non-conservative form (descriptor)."""

    def __init__(self, function_space, depth, options, tidal_farms=None):
        self.function_space = function_space
        self.depth = depth
        self.options = options
        self.tidal_farms = tidal_farms
        self.bnd_functions = {}
        self.physical_constants = physical_constants


class ModeSplit2DEquations(ShallowWaterEquations):
    """2D depth-averaged shallow water equations for mode splitting schemes (descriptor of
    thetis/shallowwater_eq.py:931-966): external pressure gradient, Coriolis, momentum source, atmospheric pressure
    and the continuity terms -- no momentum advection, drag, wind stress or viscosity."""


class TracerEquation2D:
    """2D tracer advection-diffusion equation in conservative or non-conservative form (descriptor)."""

    def __init__(self, system, function_space, depth, options, velocity):
        self.system = system
        self.function_space = function_space
        self.depth = depth
        self.options = options
        self.velocity = velocity
        # conservative (depth-integrated) or non-conservative terms per label, like add_conservative_terms /
        # add_nonconservative_terms (tracer_eq_2d.py:470-488)
        tr = getattr(options, "tracer", {}) or {}
        self.labels = system.split(",")
        self.conservative = {label: bool(getattr(tr.get(label), "use_conservative_form", False)) for label in self.labels}
