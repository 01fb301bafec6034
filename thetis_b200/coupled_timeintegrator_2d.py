"""
Stand-alone mirror of `thetis.coupled_timeintegrator_2d.CoupledTimeIntegrator2D`
(thetis/coupled_timeintegrator_2d.py:10-158): SWE -> tracers -> limiter, once per
step.  With real Thetis the reference's own class runs unmodified on top of the
B200 integrators; this copy of the *ordering* exists only because Thetis cannot
be imported here.
"""
from __future__ import annotations

__all__ = ["GeneralCoupledTimeIntegrator2D"]


class GeneralCoupledTimeIntegrator2D:
    def __init__(self, solver, integrators):
        self.solver = solver
        self.options = solver.options
        self.fields = solver.fields
        self.swe_integrator = integrators.get("shallow_water")
        self.tracer_integrator = integrators.get("tracer")
        # attribute access like the reference's AttrDict (tests reach for `.timesteppers.tracer_2d`,
        # test/tracerEq/test_h-advection_mes_2d.py:97)
        self.timesteppers = type(solver.fields)()
        self._initialized = False
        if not self.options.tracer_only:
            self.timesteppers["swe2d"] = solver.get_swe_timestepper(self.swe_integrator)
        for system in self.options.tracer_fields:
            self.timesteppers[system] = solver.get_tracer_timestepper(self.tracer_integrator, system)
        self.cfl_coeff = min(ts.cfl_coeff for ts in self.timesteppers.values())
        self.n_stages = 1

    def set_dt(self, dt):
        for k in sorted(self.timesteppers):
            self.timesteppers[k].set_dt(dt)

    def initialize(self, solution2d):
        assert solution2d is self.fields.solution_2d
        if not self.options.tracer_only:
            self.timesteppers["swe2d"].initialize(self.fields.solution_2d)
        for system in self.options.tracer_fields:
            self.timesteppers[system].initialize(self.fields[system])
        self._initialized = True

    def advance(self, t, update_forcings=None):
        """coupled_timeintegrator_2d.py:94-105"""
        if not self.options.tracer_only:
            self.timesteppers["swe2d"].advance(t, update_forcings=update_forcings)
        for system in self.options.tracer_fields:
            self.timesteppers[system].advance(t, update_forcings=update_forcings)
            if self.options.use_limiter_for_tracers:
                if "," in system:
                    raise NotImplementedError("Slope limiters not supported for mixed systems of tracers")
                self.solver.tracer_limiter.apply(self.fields[system])

    def sync_to_host(self):
        for ts in self.timesteppers.values():
            ts.sync_to_host()
