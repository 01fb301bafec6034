"""
Plain-Python stand-in for the part of `thetis.options.ModelOptions2d`
(thetis/options.py:583-733, 838-1041) the explicit dg-dg path reads.  Same
attribute names and defaults; no traitlets (not installed here).  With real
Thetis the integrator reads the real options object instead.
"""
from __future__ import annotations

from .shim import Constant

__all__ = ["ModelOptions2d", "ExplicitTimeStepperOptions2d", "TracerFieldOptions"]


class ExplicitTimeStepperOptions2d:
    """thetis/options.py:24-26, 140-163"""

    def __init__(self, solver_parameters=None):
        self.use_automatic_timestep = True
        self.solver_parameters = dict(solver_parameters or {})
        self.ad_block_tag = None


class TracerFieldOptions:
    def __init__(self):
        self.function = None
        self.source = None
        self.diffusivity = None
        self.use_conservative_form = False
        self.metadata = {}


class ModelOptions2d:
    def __init__(self):
        # CommonModelOptions (options.py:583-733)
        self.polynomial_degree = 1
        self.element_family = "dg-dg"
        self.use_nonlinear_equations = True
        self.use_lax_friedrichs_velocity = True
        self.lax_friedrichs_velocity_scaling_factor = Constant(1.0)
        self.use_lax_friedrichs_tracer = False
        self.lax_friedrichs_tracer_scaling_factor = Constant(1.0)
        self.use_limiter_for_tracers = True
        self.check_volume_conservation_2d = False
        self.timestep = 10.0
        self.cfl_2d = 1.0
        self.simulation_export_time = 100.0
        self.simulation_end_time = None
        self.horizontal_velocity_scale = Constant(0.1)
        self.output_directory = "outputs"
        self.no_exports = True
        self.fields_to_export = ["elev_2d", "uv_2d"]
        self.verbose = 0
        self.linear_drag_coefficient = None
        self.quadratic_drag_coefficient = None
        self.manning_drag_coefficient = None
        self.nikuradse_bed_roughness = None
        self.norm_smoother = Constant(0.0)
        self.horizontal_viscosity = None
        self.use_grad_div_viscosity_term = False          # options.py:597
        self.use_grad_depth_viscosity_term = True         # options.py:602
        self.sipg_factor = Constant(1.0)                  # options.py:730
        self.sipg_factor_tracer = Constant(1.0)           # options.py:732
        self.horizontal_viscosity_scale = Constant(1.0)   # options.py:655 (only used by the implicit SIPG estimate)
        self.horizontal_diffusivity_scale = Constant(1.0)
        self.coriolis_frequency = None
        self.wind_stress = None
        self.atmospheric_pressure = None
        self.momentum_source_2d = None
        self.volume_source_2d = None
        # ModelOptions2d (options.py:838-949)
        self.swe_timestepper_type = "SSPRK33"        # NB the reference default is 'CrankNicolson' (implicit, out of scope)
        self.swe_timestepper_options = ExplicitTimeStepperOptions2d(
            {"snes_type": "ksponly", "ksp_type": "cg", "pc_type": "bjacobi", "sub_ksp_type": "preonly",
             "sub_pc_type": "ilu", "mat_type": "aij"})
        self.tracer_timestepper_type = "SSPRK33"
        self.tracer_timestepper_options = ExplicitTimeStepperOptions2d({"ksp_type": "gmres", "pc_type": "sor"})
        self.use_tracer_conservative_form = False
        self.use_wetting_and_drying = False
        self.wetting_and_drying_alpha = Constant(0.5)
        self.check_tracer_conservation = False
        self.tracer_advective_velocity_factor = Constant(1.0)
        self.check_tracer_overshoot = False
        self.tracer_only = False
        self.tracer_element_family = "dg"
        self.use_supg_tracer = False
        self.tracer_picard_iterations = 1
        self.tracer = {}
        self.tracer_fields = {}

    def add_tracer_2d(self, label, name, filename, shortname=None, unit="-", **kwargs):
        """options.py:950-985"""
        assert label not in self.tracer, f"Field '{label}' already exists."
        assert " " not in label and "," not in label
        o = TracerFieldOptions()
        o.metadata = {"name": name, "shortname": shortname or name, "unit": unit, "filename": filename}
        o.function = kwargs.get("function")
        o.source = kwargs.get("source")
        o.diffusivity = kwargs.get("diffusivity")
        o.use_conservative_form = kwargs.get("use_conservative_form", False)
        self.tracer[label] = o
        self.tracer_fields[label] = o.function

    def set_timestepper_type(self, timestepper_type, **kwargs):
        """options.py:1018-1041"""
        self.swe_timestepper_type = timestepper_type
        self.tracer_timestepper_type = timestepper_type
        for key, value in kwargs.items():
            for o in (self.swe_timestepper_options, self.tracer_timestepper_options):
                setattr(o, key, value)

    def update(self, d):
        for k, v in d.items():
            if not hasattr(self, k):
                raise AttributeError(f"unknown option {k}")
            setattr(self, k, v)
