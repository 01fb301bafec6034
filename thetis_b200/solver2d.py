"""
Stand-alone harness with the surface of `thetis.solver2d.FlowSolver2d`
(thetis/solver2d.py:28-1144) for the explicit dg-dg path: same method names,
option names, `bnd_functions` layout and time loop, so that set-ups read like
the reference's demos/tests.  It exists because Thetis/Firedrake cannot be
imported in this environment; with a live Thetis install the reference's own
FlowSolver2d drives the B200 integrators.

Only data containers live on the host (thetis_b200.shim); every time step runs
in the CUDA library.  Fields are copied back to the host when something on the
host looks: exports, `print_state`, 'timestep' callbacks, the end of
`iterate()`.
"""
from __future__ import annotations

import sys
import time as time_mod

import numpy as np
import torch

from . import rungekutta
from .coupled_timeintegrator_2d import GeneralCoupledTimeIntegrator2D
from .equations import DepthExpression, ShallowWaterEquations, TracerEquation2D, physical_constants
from .limiter import VertexBasedP1DGLimiter
from .mesh import Mesh2D
from .options import ModelOptions2d
from .shim import Constant, Function, FunctionSpace, MixedFunctionSpace, as_shim_mesh

__all__ = ["FlowSolver2d"]


class AttrDict(dict):
    """thetis/utility.py AttrDict: dictionary with attribute access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def print_output(msg):
    print(msg)
    sys.stdout.flush()


def _p1_mass_project(fs, cell_rhs):
    """
    L2 projection into a P1 (CG or DG) shim space: solve M x = l with
    l given per cell node, (nt, 3) = int phi_a * g dx.
    """
    from scipy.sparse import coo_matrix
    from scipy.sparse.linalg import spsolve
    m = fs.mesh().topology_mesh
    area = m.cell_area()
    cmap = fs.cell_node_map().values.astype(np.int64)
    n = fs.node_count()
    mloc = (np.ones((3, 3)) + np.eye(3)) / 12.0
    rows = np.repeat(cmap, 3, axis=1).reshape(-1)
    cols = np.tile(cmap, (1, 3)).reshape(-1)
    vals = (area[:, None, None] * mloc[None]).reshape(-1)
    M = coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc()
    rhs = np.zeros(n)
    np.add.at(rhs, cmap.reshape(-1), cell_rhs.reshape(-1))
    return spsolve(M, rhs)


# Dunavant degree-5, 7-point rule (set-up only: CFL time step projection)
def _dunavant7():
    a1, w1 = 0.059715871789770, 0.132394152788506
    b1 = 0.470142064105115
    a2, w2 = 0.797426985353087, 0.125939180544827
    b2 = 0.101286507323456
    lam = np.array([[1 / 3, 1 / 3, 1 / 3],
                    [a1, b1, b1], [b1, a1, b1], [b1, b1, a1],
                    [a2, b2, b2], [b2, a2, b2], [b2, b2, a2]])
    w = np.array([0.225, w1, w1, w1, w2, w2, w2])
    return lam, w / w.sum()


class FlowSolver2d:
    def __init__(self, mesh2d, bathymetry_2d, options=None):
        self.mesh2d = as_shim_mesh(mesh2d)
        self.comm = None
        self.dt = None
        self.options = ModelOptions2d()
        if options is not None:
            self.options.update(options)
        self.simulation_time = 0.0
        self.iteration = 0
        self.i_export = 0
        self.next_export_t = self.simulation_time + self.options.simulation_export_time
        self.fields = AttrDict()
        self.function_spaces = AttrDict()
        self.fields.bathymetry_2d = bathymetry_2d
        self.export_initial_state = True
        self.sediment_model = None
        self.bnd_functions = {"shallow_water": {}, "tracer": {}, "sediment": {}}
        self.callbacks = {"timestep": [], "export": []}
        self._export_consumers = []
        self._pending_exports = []
        self.solve_tracer = False
        self.keep_log = False
        self._initialized = False
        self.tracer_limiter = None
        self.equations = None
        self.timestepper = None

    # ------------------------------------------------------------ time step (solver2d.py:150-248)
    def compute_time_step(self, u_scale=0.0):
        if "h_elem_size_2d" not in self.fields:
            # get_horizontal_elem_size_2d (utility.py:622-640): P1 projection of sqrt(CellVolume); set-up only,
            # computed lazily because it needs a global P1 mass solve
            m_ = self.mesh2d.topology_mesh
            self.fields.h_elem_size_2d = Function(self.function_spaces.P1_2d)
            area_ = m_.cell_area()
            rhs_ = np.repeat((np.sqrt(area_) * area_ / 3.0)[:, None], 3, axis=1)
            self.fields.h_elem_size_2d.dat.data[:] = _p1_mass_project(self.function_spaces.P1_2d, rhs_)
        csize = self.fields.h_elem_size_2d
        bath = self.fields.bathymetry_2d
        fs = bath.function_space()
        m = self.mesh2d.topology_mesh
        min_depth = 0.05
        bpos = np.array(bath.dat.data_ro, dtype=float)
        bpos[bpos < min_depth] = min_depth
        g = float(physical_constants["g_grav"])
        us = float(u_scale) if not hasattr(u_scale, "values") else float(u_scale.values()[0])
        lam, w = _dunavant7()
        bq = np.einsum("qa,ca->cq", lam, bpos[fs.cell_node_map().values])
        cq = np.einsum("qa,ca->cq", lam, csize.dat.data_ro[csize.function_space().cell_node_map().values])
        gq = cq / (np.sqrt(g * bq) + us)
        cell_rhs = np.einsum("c,q,cq,qa->ca", m.cell_area(), w, gq, lam)
        sol = Function(fs)
        sol.dat.data[:] = _p1_mass_project(fs, cell_rhs)
        return sol

    def set_time_step(self, alpha=0.05):
        automatic = False
        # solver2d.py:222-231 loops over every sub-option object that has the flag; the reference's default tracer
        # stepper is implicit (no flag), so here the tracer options only count when tracers are solved
        ts_options = [self.options.swe_timestepper_options]
        if self.solve_tracer:
            ts_options.append(self.options.tracer_timestepper_options)
        for o in ts_options:
            if getattr(o, "use_automatic_timestep", False):
                automatic = True
        if automatic:
            mesh2d_dt = self.compute_time_step(u_scale=self.options.horizontal_velocity_scale)
            vals = np.asarray(mesh2d_dt.dat.data_ro, dtype=float)
            plan = getattr(self.mesh2d, "halo_plan", None)
            if plan is not None:
                # distributed mesh (solver2d.py:241: dt = comm.allreduce(dt, op=MPI.MIN)): the minimum is taken over
                # the nodes of the cells this rank owns (the projection is polluted on the far side of the ghost
                # layer, where the local mass matrix misses neighbours) and then over the ranks, so that every rank
                # advances with the same time step
                import torch
                import torch.distributed as dist
                nodes = np.unique(mesh2d_dt.function_space().cell_node_map().values[: plan.part.n_owned])
                local = torch.tensor([float(vals[nodes].min())], dtype=torch.float64)
                if dist.get_backend() == "nccl":
                    local = local.cuda()
                dist.all_reduce(local, op=dist.ReduceOp.MIN)
                dt_min = float(local.item())
            else:
                dt_min = float(vals.min())
            self.dt = self.options.cfl_2d * alpha * dt_min
        else:
            assert self.options.timestep is not None and self.options.timestep > 0.0
            self.dt = self.options.timestep
        print_output("dt = {:}".format(self.dt))

    # ------------------------------------------------------------ spaces / fields / equations
    def create_function_spaces(self):
        """solver2d.py:307-352 (dg-dg branch only)"""
        if self.options.element_family != "dg-dg" or self.options.polynomial_degree != 1:
            raise NotImplementedError("only element_family='dg-dg', polynomial_degree=1 is on the accelerated path")
        fs = self.function_spaces
        mesh = self.mesh2d
        fs.P0_2d = FunctionSpace(mesh, "DG", 0, name="P0_2d")
        fs.P1_2d = FunctionSpace(mesh, "CG", 1, name="P1_2d")
        fs.P1v_2d = FunctionSpace(mesh, "CG", 1, value_size=2, name="P1v_2d")
        fs.P1DG_2d = FunctionSpace(mesh, "DG", 1, name="P1DG_2d")
        fs.P1DGv_2d = FunctionSpace(mesh, "DG", 1, value_size=2, name="P1DGv_2d")
        fs.U_2d = FunctionSpace(mesh, "DG", 1, value_size=2, name="U_2d")
        fs.H_2d = FunctionSpace(mesh, "DG", 1, name="H_2d")
        fs.V_2d = MixedFunctionSpace([fs.U_2d, fs.H_2d], name="V_2d")
        fs.Q_2d = FunctionSpace(mesh, "DG", 1, name="Q_2d")

    def create_fields(self):
        """solver2d.py:389-449"""
        if not self.function_spaces:
            self.create_function_spaces()
        self.depth = DepthExpression(self.fields.bathymetry_2d,
                                     use_nonlinear_equations=self.options.use_nonlinear_equations,
                                     use_wetting_and_drying=self.options.use_wetting_and_drying,
                                     wetting_and_drying_alpha=self.options.wetting_and_drying_alpha)
        self.fields.solution_2d = Function(self.function_spaces.V_2d, name="solution_2d")
        uv_2d, elev_2d = self.fields.solution_2d.subfunctions
        self.fields.uv_2d = uv_2d
        self.fields.elev_2d = elev_2d
        self.solve_tracer = len(self.options.tracer_fields) > 0
        for system, parent in list(self.options.tracer_fields.items()):
            if "," in system:
                raise NotImplementedError("mixed tracer systems are outside the accelerated path")
            if parent is None:
                parent = Function(self.function_spaces.Q_2d, name=system)
                self.options.tracer[system].function = parent
                self.options.tracer_fields[system] = parent
            self.fields[system] = parent

    def create_equations(self):
        """solver2d.py:453-539"""
        if "solution_2d" not in self.fields:
            self.create_fields()
        self.equations = AttrDict()
        self.equations.sw = ShallowWaterEquations(self.fields.solution_2d.function_space(), self.depth, self.options)
        self.equations.sw.bnd_functions = self.bnd_functions["shallow_water"]
        uv_2d, _ = self.fields.solution_2d.subfunctions
        for system, parent in self.options.tracer_fields.items():
            self.equations[system] = TracerEquation2D(system, parent.function_space(), self.depth, self.options, uv_2d)
        if self.solve_tracer:
            if self.options.use_limiter_for_tracers and self.options.polynomial_degree > 0:
                self.tracer_limiter = VertexBasedP1DGLimiter(self.function_spaces.Q_2d)
            else:
                self.tracer_limiter = None

    def get_swe_timestepper(self, integrator):
        """solver2d.py:542-573"""
        o = self.options
        fields = {
            "linear_drag_coefficient": o.linear_drag_coefficient,
            "quadratic_drag_coefficient": o.quadratic_drag_coefficient,
            "manning_drag_coefficient": o.manning_drag_coefficient,
            "nikuradse_bed_roughness": o.nikuradse_bed_roughness,
            "viscosity_h": o.horizontal_viscosity,
            "lax_friedrichs_velocity_scaling_factor": o.lax_friedrichs_velocity_scaling_factor,
            "coriolis": o.coriolis_frequency,
            "wind_stress": o.wind_stress,
            "atmospheric_pressure": o.atmospheric_pressure,
            "momentum_source": o.momentum_source_2d,
            "volume_source": o.volume_source_2d,
        }
        return integrator(self.equations.sw, self.fields.solution_2d, fields, self.dt,
                          o.swe_timestepper_options, self.bnd_functions["shallow_water"], sync_policy="manual")

    def get_tracer_timestepper(self, integrator, system):
        """solver2d.py:576-598"""
        uv, elev = self.fields.solution_2d.subfunctions
        o = self.options
        fields = {
            "elev_2d": elev,
            "uv_2d": uv,
            "lax_friedrichs_tracer_scaling_factor": o.lax_friedrichs_tracer_scaling_factor,
            "tracer_advective_velocity_factor": o.tracer_advective_velocity_factor,
        }
        for label in system.split(","):
            fields[f"diffusivity_h-{label}"] = o.tracer[label].diffusivity
            fields[f"source-{label}"] = o.tracer[label].source
        bcs = {}
        if system in self.bnd_functions:
            bcs = self.bnd_functions[system]
        elif system[:-3] in self.bnd_functions:
            bcs = self.bnd_functions[system[:-3]]
        return integrator(self.equations[system], self.fields[system], fields, self.dt,
                          o.tracer_timestepper_options, bcs, sync_policy="manual")

    def create_timestepper(self):
        """solver2d.py:651-701"""
        if self.equations is None:
            self.create_equations()
        self.set_time_step()
        # solver2d.py:662-672 knows 'SSPRK33' and 'ForwardEuler' among the explicit schemes; the Butcher-form ERK
        # classes of rungekutta.py:959-980 are accepted by name as well
        steppers = {"SSPRK33": rungekutta.SSPRK33, "ForwardEuler": rungekutta.ForwardEuler,
                    "ERKLSPUM2": rungekutta.ERKLSPUM2, "ERKLPUM2": rungekutta.ERKLPUM2,
                    "ERKMidpoint": rungekutta.ERKMidpoint, "ERKEuler": rungekutta.ERKEuler}
        for t in (self.options.swe_timestepper_type, self.options.tracer_timestepper_type):
            if t not in steppers:
                raise NotImplementedError(f"time integrator {t!r} is outside the accelerated path "
                                          f"(explicit only: {sorted(steppers)})")
        if self.solve_tracer:
            self.timestepper = GeneralCoupledTimeIntegrator2D(self, {
                "shallow_water": steppers[self.options.swe_timestepper_type],
                "tracer": steppers[self.options.tracer_timestepper_type]})
        else:
            self.timestepper = self.get_swe_timestepper(steppers[self.options.swe_timestepper_type])
        print_output("Using time integrator: {:}".format(self.timestepper.__class__.__name__))

    def initialize(self):
        """solver2d.py:732-744"""
        if not self.function_spaces:
            self.create_function_spaces()
        if "solution_2d" not in self.fields:
            self.create_fields()
        if self.equations is None:
            self.create_equations()
        if self.timestepper is None:
            self.create_timestepper()
        if not getattr(self, "_exporters_created", False):
            self.create_exporters()
        self._initialized = True

    @staticmethod
    def _assign(target, value):
        # the reference projects (solver2d.py:765-768); for P1 data projection onto P1DG is the identity,
        # callables are interpolated nodally
        if value is None:
            return
        if callable(value) and not isinstance(value, (Function, Constant)):
            target.interpolate(value)
        elif isinstance(value, Function):
            src_fs = value.function_space()
            tgt_fs = target.function_space()
            if src_fs.family == tgt_fs.family:
                target.assign(value)
            else:
                # CG -> DG: copy vertex values to every cell's nodes
                target.dat.data[tgt_fs.cell_node_map().values.reshape(-1)] = \
                    value.dat.data_ro[src_fs.cell_node_map().values.reshape(-1)]
        else:
            target.assign(value)

    def assign_initial_conditions(self, elev=None, uv=None, **tracers):
        """solver2d.py:747-785"""
        if not self._initialized:
            self.initialize()
        uv_2d, elev_2d = self.fields.solution_2d.subfunctions
        self._assign(elev_2d, elev)
        self._assign(uv_2d, uv)
        for l, func in tracers.items():
            label = l if len(l) > 3 and l[-3:] == "_2d" else l + "_2d"
            assert label in self.options.tracer, f"Unknown tracer label {label}"
            self._assign(self.fields[label], func)
        self.timestepper.initialize(self.fields.solution_2d)

    def add_callback(self, callback, eval_interval="export"):
        self.callbacks[eval_interval].append(callback)

    def create_exporters(self):
        """solver2d.py:1040-1075 (diagnostic callbacks only; file exporters are outside the accelerated path)"""
        from . import callback
        o = self.options
        if o.check_volume_conservation_2d:
            self.add_callback(callback.VolumeConservation2DCallback(self, append_to_log=True))
        if o.check_tracer_conservation:
            for label, tracer in o.tracer.items():
                cls = (callback.ConservativeTracerMassConservation2DCallback if tracer.use_conservative_form
                       else callback.TracerMassConservation2DCallback)
                self.add_callback(cls(label, self, append_to_log=True), eval_interval="export")
        if o.check_tracer_overshoot:
            for label in o.tracer:
                self.add_callback(callback.TracerOvershootCallBack(label, self, append_to_log=True),
                                  eval_interval="export")
        self._exporters_created = True

    def _run_callbacks(self, mode):
        from .callback import DiagnosticCallback
        host_needed = any(not isinstance(cb, DiagnosticCallback) for cb in self.callbacks[mode])
        if host_needed:
            self.sync_to_host()               # user callbacks read host Functions; the device diagnostics do not
        for cb in self.callbacks[mode]:
            if isinstance(cb, DiagnosticCallback):
                cb.evaluate(index=self.i_export)
            elif hasattr(cb, "evaluate"):
                cb.evaluate(self)
            else:
                cb(self)

    # ------------------------------------------------------------ host visibility
    def sync_to_host(self):
        self.timestepper.sync_to_host()

    def _swe_stepper(self):
        ts = self.timestepper
        if isinstance(ts, GeneralCoupledTimeIntegrator2D):
            return ts.timesteppers.get("swe2d")
        return ts

    def print_state(self, cputime, print_header=False):
        """solver2d.py:923-971; the norms are reduced on the device (tb_swe_integrals)."""
        entries = [("exp", self.i_export, "5d"), ("iter", self.iteration, "5d"),
                   ("time", f"{self.simulation_time:.2f}".rjust(15), "15s")]
        sw = self._swe_stepper()
        if sw is not None:
            out = torch.zeros(4, dtype=torch.float64, device=sw.engine.device)
            sw.engine.swe_integrals(sw.device_state(), out)
            if sw.halo is not None:
                sw.halo.allreduce_sum(out)          # norms are global (Firedrake's norm() reduces over the communicator)
            o = out.cpu().numpy()
            entries += [("eta norm", float(np.sqrt(o[0])), "14.4f"), ("u norm", float(np.sqrt(o[1])), "14.4f")]
            self.last_norms = (float(np.sqrt(o[0])), float(np.sqrt(o[1])))
        entries.append(("Tcpu", cputime, "6.2f"))
        if print_header:
            print_output(" ".join([e[0].rjust(len(f"{e[1]:{e[2]}}")) for e in entries]))
        print_output(" ".join([f"{e[1]:{e[2]}}" for e in entries]))

    def add_export_consumer(self, consumer):
        """
        Register ``consumer(time, i_export, arrays)`` for NON-BLOCKING exports: at every export the device solution
        is staged into pinned host memory on a side stream (`TimeIntegrator.stage_export`) and the consumer runs when
        the copy has landed -- at the next export, or at the end of `iterate()` at the latest -- while the time loop
        keeps the GPU busy.  ``arrays``: {'uv_2d': (n, 2), 'elev_2d': (n,), '<tracer>': (n,)} numpy views in the
        Thetis dof order, valid during the call.  This is what a file exporter (exporter.py) plugs into; exports
        that need the host `Function`s themselves (`export_func`, user callbacks) still synchronise.
        """
        self._export_consumers.append(consumer)

    def _stage_exports(self, time):
        ts = self.timestepper
        handles = {}
        if isinstance(ts, GeneralCoupledTimeIntegrator2D):
            for name, st in ts.timesteppers.items():
                handles[name] = st.stage_export()
        else:
            handles["swe2d"] = ts.stage_export()
        self._pending_exports.append((time, self.i_export, handles))

    def _drain_exports(self, block=False):
        while self._pending_exports:
            time, i_export, handles = self._pending_exports[0]
            if not block and not all(h.ready() for h in handles.values()):
                return
            arrays = {}
            for name, h in handles.items():
                a = h.wait()
                if name == "swe2d":
                    arrays["uv_2d"], arrays["elev_2d"] = a
                else:
                    arrays[name] = a[0]
            for c in self._export_consumers:
                c(time, i_export, arrays)
            self._pending_exports.pop(0)

    def export(self, time=None):
        """Fields become host-visible here; VTK/HDF5 writers are outside the accelerated path (exporter.py).
        With export consumers registered the D2H is staged asynchronously instead (solver2d.py:1132-1142 blocks the
        time loop on every export)."""
        if self._export_consumers:
            self._drain_exports(block=len(self._pending_exports) >= 2)      # two staging buffers per integrator
            self._stage_exports(self.simulation_time if time is None else time)
        else:
            self.sync_to_host()
        self._run_callbacks("export")

    # ------------------------------------------------------------ time loop (solver2d.py:974-1144)
    def iterate(self, update_forcings=None, export_func=None):
        for _ in self.create_iterator(update_forcings=update_forcings, export_func=export_func):
            pass

    def create_iterator(self, update_forcings=None, export_func=None):
        if not self._initialized:
            self.initialize()
        self.options.use_limiter_for_tracers &= self.options.polynomial_degree > 0
        t_epsilon = 1.0e-5
        cputimestamp = time_mod.perf_counter()
        next_export_t = self.simulation_time + self.options.simulation_export_time
        initial_simulation_time = self.simulation_time
        internal_iteration = 0
        assert self.options.simulation_end_time is not None, "simulation_end_time must be set"
        self.print_state(0.0, print_header=True)
        if self.export_initial_state:
            self.export(time=self.simulation_time)
            if export_func is not None:
                self.sync_to_host()
                export_func()
        while self.simulation_time <= self.options.simulation_end_time - t_epsilon:
            self.timestepper.advance(self.simulation_time, update_forcings)
            yield self.simulation_time
            self.iteration += 1
            internal_iteration += 1
            self.simulation_time = initial_simulation_time + internal_iteration * self.dt
            if self.callbacks["timestep"]:
                self._run_callbacks("timestep")
            if self.simulation_time >= next_export_t - t_epsilon:
                self.i_export += 1
                next_export_t += self.options.simulation_export_time
                cputime = time_mod.perf_counter() - cputimestamp
                cputimestamp = time_mod.perf_counter()
                self.print_state(cputime)
                self.export(time=self.simulation_time)
                if export_func is not None:
                    self.sync_to_host()            # a user export function reads the host Functions
                    export_func()
        self._drain_exports(block=True)
        self.sync_to_host()
        return self.simulation_time
