"""
The 2-D (external mode) side of `thetis.coupled_timeintegrator.CoupledTwoStageRK`
(thetis/coupled_timeintegrator.py:563-715) on the B200 integrators.

The reference's two-stage coupled loop advances the 3-D fields with SSPRK(2,2) (`SSPRK22ALE`) and calls, once per
stage, ``self.timesteppers.swe2d.solve_stage(i_stage, t, update_forcings)`` on an integrator built over
`ModeSplit2DEquations` (shallowwater_eq.py:931-966) whose momentum source is the coupling term
``split_residual_2d (+ momentum_source_2d)`` (:181-199), re-assigned by the 3-D side after every stage
(`_update_2d_coupling_term`, :65-70).  This module reproduces exactly that contract for the 2-D mode:

* `create_swe_integrator` builds the 2-D integrator with the reference's `fields` dict (:186-199);
* `advance` runs the stage loop in the reference's order: store_elevation -> swe2d.solve_stage ->
  compute_mesh_velocity -> [3-D side: prepare / solve] -> _update_2d_coupling, leaving the stage solution visible to
  the host after every stage (the 3-D side reads uv_2d / elev_2d there, `_copy_uv_2d_to_3d`, :60-63).

The 3-D equations themselves (momentum, tracers, turbulence, ALE mesh) are outside the scope of this library
("the first step toward accelerating the 3-D model"): they stay with the reference and plug in through
the `mode3d` object (`prepare_stage`, `solve_stage`, `update_2d_coupling`).  The reference pairs the loop with the
implicit `ESDIRKTrapezoid` 2-D integrator; the explicit counterpart offered here is `rungekutta.SSPRK22`, the same
scheme and stage times (c = [0, 1]) `SSPRK22ALE` uses for the 3-D fields.
"""
from __future__ import annotations

from . import rungekutta

__all__ = ["CoupledTwoStageRK2D"]


class CoupledTwoStageRK2D:
    """
    :arg solver: object with the attributes the reference's coupled integrator reads from `FlowSolver`:
        `equations.sw` (a `ModeSplit2DEquations`), `fields.solution_2d`, `fields.split_residual_2d`, `options`
        (`coriolis_frequency`, `momentum_source_2d`, `volume_source_2d`, `atmospheric_pressure`,
        `timestepper_options.swe_options` or `swe_timestepper_options`), `dt`, `bnd_functions['shallow_water']`.
    :kwarg mode3d: optional 3-D side: any object with `prepare_stage(i_stage, t, update_forcings3d)`,
        `solve_stage(i_stage)` and `update_2d_coupling(last_stage)`; the last one is expected to re-assign
        `solver.fields.split_residual_2d` like `_update_2d_coupling_term` does.
    """
    integrator_2d = rungekutta.SSPRK22

    def __init__(self, solver, mode3d=None, integrator_2d=None):
        self.solver = solver
        self.options = solver.options
        self.fields = solver.fields
        self.mode3d = mode3d
        if integrator_2d is not None:
            self.integrator_2d = integrator_2d
        self.timesteppers = type(solver.fields)()
        self._initialized = False
        self.create_swe_integrator()
        self.n_stages = self.timesteppers.swe2d.n_stages        # coupled_timeintegrator.py:162
        self.cfl_coeff_2d = self.timesteppers.swe2d.cfl_coeff   # :348

    def create_swe_integrator(self):
        """coupled_timeintegrator.py:181-199"""
        solver = self.solver
        momentum_source_2d = solver.fields.split_residual_2d
        if self.options.momentum_source_2d is not None:
            momentum_source_2d = solver.fields.split_residual_2d + self.options.momentum_source_2d
        fields = {
            "coriolis": self.options.coriolis_frequency,
            "momentum_source": momentum_source_2d,
            "volume_source": self.options.volume_source_2d,
            "atmospheric_pressure": self.options.atmospheric_pressure,
        }
        ts_opts = getattr(getattr(self.options, "timestepper_options", None), "swe_options", None)
        if ts_opts is None:
            ts_opts = getattr(self.options, "swe_timestepper_options", None)
        # every stage solution must be visible to the 3-D side on the host
        self.timesteppers.swe2d = self.integrator_2d(
            solver.equations.sw, self.fields.solution_2d, fields, solver.dt, ts_opts,
            solver.bnd_functions["shallow_water"], sync_policy="every_stage" if self.mode3d is not None else "every_step")

    def set_dt(self, dt, dt_2d=None):
        """coupled_timeintegrator.py:350-366 (the 2-D integrator of this loop runs with the 3-D time step)"""
        self.timesteppers.swe2d.set_dt(dt)

    def initialize(self):
        """coupled_timeintegrator.py:368-394 (2-D part)"""
        self.timesteppers.swe2d.initialize(self.fields.solution_2d)
        self._initialized = True

    # ALE hooks of the reference (:580-620): the mesh geometry belongs to the 3-D side
    def store_elevation(self, i_stage):
        if self.mode3d is not None and hasattr(self.mode3d, "store_elevation"):
            self.mode3d.store_elevation(i_stage)

    def compute_mesh_velocity(self, i_stage):
        if self.mode3d is not None and hasattr(self.mode3d, "compute_mesh_velocity"):
            self.mode3d.compute_mesh_velocity(i_stage)

    def _update_2d_coupling(self, last_stage):
        if self.mode3d is not None:
            self.mode3d.update_2d_coupling(last_stage)

    def advance(self, t, update_forcings=None, update_forcings3d=None):
        """coupled_timeintegrator.py:622-715, the calls that involve the 2-D mode"""
        if not self._initialized:
            self.initialize()
        for i_stage in range(self.n_stages):
            # solve 2D mode
            self.store_elevation(i_stage)
            self.timesteppers.swe2d.solve_stage(i_stage, t, update_forcings)
            self.compute_mesh_velocity(i_stage)
            # 3D mode (with the reference): preprocess in the old mesh, update the mesh, solve
            if self.mode3d is not None:
                self.mode3d.prepare_stage(i_stage, t, update_forcings3d)
                self.mode3d.solve_stage(i_stage)
            last_stage = i_stage == self.n_stages - 1
            # update the variables the explicit solvers depend on: split_residual_2d for the next 2-D stage
            self._update_2d_coupling(last_stage)
        if self.timesteppers.swe2d.sync_policy == "every_step":
            self.timesteppers.swe2d.sync_to_host()
