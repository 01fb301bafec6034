"""
Bench / test drivers (NOT part of the product package): the BASELINE.json workloads run through the reference-shaped
surface (FlowSolver2d mirror -> SSPRK33 -> C-ABI) on one GPU or on a distributed mesh.  Shared by bench.py, the GPU
tests and __graft_entry__.smoke() so that they all run the same thing.
"""
from __future__ import annotations

import numpy as np

from thetis_b200.mesh import FACET_NODES
from thetis_b200.parallel import distribute_mesh

# ---------------------------------------------------------------------- bench drivers
def _make_solver(mesh, setup, wd, n_owned=None):
    """FlowSolver2d mirror configured for the North Sea workload (harness/workloads.py)."""
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, Constant, ShimMesh, as_shim_mesh
    sm = mesh if isinstance(mesh, ShimMesh) else as_shim_mesh(mesh)
    P1 = FunctionSpace(sm, "CG", 1)
    bath = Function(P1, name="Bathymetry")
    bath.dat.data[:] = setup["bath"]
    s = solver2d.FlowSolver2d(sm, bath)
    o = s.options
    o.swe_timestepper_type = "SSPRK33"
    o.swe_timestepper_options.use_automatic_timestep = False
    o.timestep = setup["dt"]
    o.simulation_end_time = 1e30
    o.simulation_export_time = 1e30
    o.use_wetting_and_drying = bool(wd)
    o.wetting_and_drying_alpha = Constant(setup["wd_alpha"])
    man = Function(P1, name="Manning coefficient")
    man.dat.data[:] = setup["manning"]
    cor = Function(P1, name="Coriolis forcing")
    cor.dat.data[:] = setup["coriolis"]
    o.manning_drag_coefficient = man
    o.coriolis_frequency = cor
    o.horizontal_velocity_scale = Constant(1.5)
    tide = Function(P1, name="Tidal elevation")
    s.bnd_functions["shallow_water"] = {100: {"elev": tide, "uv": Constant((0.0, 0.0))}}
    return s, tide


class SingleSWE:
    """North Sea workload on one GPU through the reference-shaped surface."""

    def __init__(self, mesh, setup, wd=True):
        import torch
        from .workloads import M2_PERIOD
        self.torch = torch
        self.solver, self.tide = _make_solver(mesh, setup, wd)
        mesh = self.solver.mesh2d.topology_mesh          # local Mesh2D (whole mesh on one GPU)
        self.mesh, self.setup = mesh, setup
        s = self.solver
        s.create_function_spaces()
        s.create_equations()
        uv0 = setup["uv0"]
        eta0 = setup["eta0"]
        s.initialize()
        s.fields.uv_2d.dat.data[:] = uv0.reshape(-1, 2)
        s.fields.elev_2d.dat.data[:] = eta0.reshape(-1)
        s.timestepper.initialize(s.fields.solution_2d)
        self.ts = s.timestepper
        self.eng = self.ts.engine
        self.t = 0.0
        self.dt = setup["dt"]
        # open-boundary vertices of the P1 tide Function and their phase (host-side forcing, like TPXO in the demo)
        m = mesh
        open_f = m.bf_marker == 100
        nodes = m.cells[m.bf_cell[open_f][:, None], FACET_NODES[m.bf_lf[open_f]]]      # geometric vertices (nb_open, 2)
        self._tide_nodes = m.topo[nodes].reshape(-1)
        self._tide_phase = setup["tide_phase"][open_f].reshape(-1)
        self._omega = 2 * np.pi / M2_PERIOD
        self._n_open = int(open_f.sum())
        self._norms = torch.zeros(4, dtype=torch.float64, device=self.eng.device)
        self._norms_host = torch.zeros(4, dtype=torch.float64).pin_memory()
        self.update_forcings(0.0)
        self.ts._push_dynamic()

    def update_forcings(self, t):
        """user callback of iterate(update_forcings=...): set the tidal elevation Function at time t"""
        self.tide.dat.data[self._tide_nodes] = np.sin(self._omega * t + self._tide_phase)

    def n_owned(self):
        return self.mesh.n_cells

    def stage_launches_per_step(self):
        return 3

    def launches_per_step(self):
        return 3

    def enable_graph(self):
        """Capture one resident step (3 fused stage launches + halo traffic) in a CUDA graph."""
        torch = self.torch
        self.ts.advance_device()                 # warm-up outside capture (lazy uploads)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.ts.advance_device()
        self._graph = g

    def use_fused_norms(self, on=True):
        """e2e path: the print_state norms are reduced in the epilogue of the last RK stage (tb_stage_integrals) instead
        of by a separate pass over the state.  Call before `enable_stage_graphs`."""
        self.ts.fused_norms = self._norms if on else None

    def enable_stage_graphs(self):
        """One CUDA graph per RK stage for the e2e path: the host-side forcing refresh stays between the launches."""
        torch, ts = self.torch, self.ts
        ts.advance_device()
        torch.cuda.synchronize()
        self._stage_graphs = []
        for i in range(ts.n_stages):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                ts._launch_stage(i)
            self._stage_graphs.append(g)
        ts.stage_graphs = self._stage_graphs     # SSPRK33.solve_stage replays them instead of re-launching

    def enable_step_graph(self):
        """e2e path with ONE graph per step: the tidal elevation of every stage goes into its own bank of the device
        boundary arrays (tb_set_bc_bank), then all stages replay from one graph."""
        self.ts.stage_graphs = None
        self.ts.enable_step_graph()

    def launches(self):
        return self.eng.launch_count() + getattr(self, "_replays", 0) * self.launches_per_step()

    def step_resident(self):
        g = getattr(self, "_graph", None)
        if g is not None:
            g.replay()
            self._replays = getattr(self, "_replays", 0) + 1
        else:
            self.ts.advance_device()

    def step_e2e(self):
        self.ts.advance(self.t, self.update_forcings)
        if getattr(self.ts, "stage_graphs", None) or getattr(self.ts, "step_graph", None) is not None:
            self._replays = getattr(self, "_replays", 0) + 1
        self.t += self.dt
        if self.ts.fused_norms is None:
            self.eng.swe_integrals(self.ts.device_state(), self._norms)
        self._norms_host.copy_(self._norms, non_blocking=True)   # else: reduced by the last stage (tb_stage_integrals)

    def e2e_path(self):
        how = ("one CUDA graph per step, the boundary data of stage i in bank i" if getattr(self.ts, "step_graph", None)
               is not None else "one CUDA graph per RK stage" if getattr(self.ts, "stage_graphs", None) else "direct launches")
        return ("FlowSolver2d mirror -> SSPRK33.advance(t, update_forcings) -> C-ABI: tidal elevation Function updated "
                "on the host for every stage (H2D from pinned memory), print_state norms reduced on the device (fused "
                "into the last stage kernel) and read back every step; " + how)

    def h2d_bytes_per_step(self):
        return 3 * self._n_open * 2 * 8

    def d2h_bytes_per_step(self):
        return 4 * 8

    def state_nodal(self):
        self.ts._host_stale = True
        self.ts.sync_to_host()
        s = self.solver
        return (s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2).copy(), s.fields.elev_2d.dat.data_ro.reshape(-1, 3).copy())


def localize_setup(setup, lm):
    """Restrict the global workload arrays (thetis_b200.workloads.north_sea_setup) to a rank's local mesh."""
    gv, gc, gb = lm.meta["global_vertices"], lm.meta["global_cells"], lm.meta["global_bfacets"]
    out = dict(setup)
    for k in ("bath", "coriolis", "manning"):
        out[k] = setup[k][gv]
    for k in ("eta0", "uv0"):
        out[k] = setup[k][gc]
    out["tide_phase"] = setup["tide_phase"][gb]
    return out


class PartitionedSWE(SingleSWE):
    """
    North Sea workload on `world` GPUs through the SAME reference-shaped surface as `SingleSWE` (FlowSolver2d mirror
    -> SSPRK33): the mesh is distributed with `distribute_mesh`, the integrator picks the rank's HaloPlan up from it
    and exchanges the one-deep halo once per RK stage.
    """

    def __init__(self, mesh, setup, rank, world, wd=True, transport="auto", overlap=True, halo="facet", fused=True):
        self.rank, self.world = rank, world
        sm = distribute_mesh(mesh, rank, world, halo=halo, transport=transport, overlap=overlap, fused=fused)
        self.part = sm.halo_plan.part
        super().__init__(sm, localize_setup(setup, self.part.mesh), wd=wd)
        self.plan = sm.halo_plan
        self.transport = self.plan.transport
        self.overlap = self.plan.overlap

    def n_owned(self):
        return self.part.n_owned

    def stage_launches_per_step(self):
        return 3 * (1 if self.plan.fused else (2 if self.plan.overlap else 1))

    def launches_per_step(self):
        return 3 * self.plan.kernels_per_swe_stage()

    def e2e_path(self):
        return ("distribute_mesh -> " + SingleSWE.e2e_path(self) + "; one halo exchange per RK stage ("
                + self.plan.transport + ")")

    def owned_nodal(self):
        uv, eta = self.state_nodal()
        n = self.part.n_owned
        return uv[:n], eta[:n]


class ConfigRun:
    """
    BASELINE configs 1-4 through the FlowSolver2d mirror on one GPU or on a distributed mesh (same public surface as
    config 5's `SingleSWE` / `PartitionedSWE`).  One step = SSPRK33 of the SWE (+ SSPRK33 of the tracer + limiter for
    config 4).
    """

    def __init__(self, cfg, rank=0, world=1, scale=1.0, transport="auto", fused=True):
        import torch
        from .workloads import config_mesh, config_solver
        self.torch = torch
        self.cfg, self.rank, self.world = cfg, rank, world
        gm = config_mesh(cfg, scale)
        self.n_global = gm.n_cells
        if world > 1:
            sm = distribute_mesh(gm, rank, world, halo="vertex" if cfg == 4 else "facet", transport=transport, fused=fused)
            self.plan = sm.halo_plan
            self.part = self.plan.part
            self.solver = config_solver(cfg, sm, global_mesh=gm)
            self._n_owned = self.part.n_owned
        else:
            self.plan = None
            self.solver = config_solver(cfg, gm)
            self._n_owned = gm.n_cells
        s = self.solver
        self.dt = s.dt
        self.t = 0.0
        ts = s.timestepper
        self.coupled = hasattr(ts, "timesteppers")
        self.swe = ts.timesteppers["swe2d"] if self.coupled else ts
        self.tracers = [ts.timesteppers[k] for k in s.options.tracer_fields] if self.coupled else []
        self.eng = self.swe.engine
        self.dofs_per_cell = 9 + 3 * len(self.tracers)
        self.transport = self.plan.transport if self.plan is not None else None
        self.overlap = self.plan.overlap if self.plan is not None else None
        self._norms = torch.zeros(4, dtype=torch.float64, device=self.eng.device)
        self._norms_host = torch.zeros(4, dtype=torch.float64).pin_memory()

    def n_owned(self):
        return self._n_owned

    def _limit(self):
        s = self.solver
        for system in s.options.tracer_fields:
            if s.options.use_limiter_for_tracers:
                s.tracer_limiter.apply(s.fields[system])

    def _swaps_buffers(self):
        # the limiter works out of place and swaps the tracer integrator's solution / scratch buffers every step
        return bool(self.tracers) and bool(self.solver.options.use_limiter_for_tracers)

    def step_resident(self):
        gs = getattr(self, "_graphs", None)
        if gs is not None:
            gs[self._gi].replay()
            if len(gs) == 2:
                self._gi ^= 1
                for tr in self.tracers:          # keep the integrators' view in step with what the graph did
                    tr.buf[0], tr.buf[1] = tr.buf[1], tr.buf[0]
            self._replays = getattr(self, "_replays", 0) + 1
            return
        self.swe.advance_device()
        for tr in self.tracers:
            tr.advance_device()
        self._limit()

    def enable_graph(self):
        """Capture the resident step in a CUDA graph -- two graphs replayed alternately when the step swaps buffers."""
        torch = self.torch
        self.step_resident()
        torch.cuda.synchronize()
        graphs = []
        for _ in range(2 if self._swaps_buffers() else 1):
            l0 = self.eng.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.step_resident()
            self._launches_per_step = self.eng.launch_count() - l0
            graphs.append(g)
        self._graphs, self._gi = graphs, 0

    def launches(self):
        return self.eng.launch_count() + getattr(self, "_replays", 0) * getattr(self, "_launches_per_step", 0)

    def use_fused_norms(self, on=True):
        self.swe.fused_norms = self._norms if on else None

    def step_e2e(self):
        """the call a user makes: timestepper.advance(t) of the (coupled) integrator + print_state norms read back"""
        self.solver.timestepper.advance(self.t, None)
        self.t += self.dt
        if self.swe.fused_norms is None:
            self.eng.swe_integrals(self.swe.device_state(), self._norms)
        self._norms_host.copy_(self._norms, non_blocking=True)

    def e2e_path(self):
        return ("FlowSolver2d mirror -> (Coupled)TimeIntegrator.advance(t) -> C-ABI; no time-dependent host forcing in "
                "this configuration; print_state norms reduced on the device and read back every step")

    def h2d_bytes_per_step(self):
        return 0

    def d2h_bytes_per_step(self):
        return 4 * 8

    def owned_nodal(self):
        for st in [self.swe] + self.tracers:
            st._host_stale = True             # graph replays bypass the steppers' bookkeeping
        self.solver.sync_to_host()
        s = self.solver
        n = self._n_owned
        out = [s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2)[:n].copy(), s.fields.elev_2d.dat.data_ro.reshape(-1, 3)[:n].copy()]
        for system in s.options.tracer_fields:
            out.append(s.fields[system].dat.data_ro.reshape(-1, 3)[:n].copy())
        return out
