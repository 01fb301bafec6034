"""
Domain decomposition for the explicit P1DG path.

The reference distributes the mesh inside Firedrake/PETSc (DMPlex overlap 1, PyOP2
halo exchanges in every assemble).  Here: the SFC-ordered cell range is cut into
`world` contiguous chunks, each rank owns one chunk plus a one-deep halo of ghost
cells (facet neighbours owned by other ranks, appended after the owned cells and
grouped by owner).  Once per RK stage every rank packs the records of the owned
cells its peers need (tb_gather_cells) and one all-to-all over NCCL/NVLink drops
them straight into the peers' ghost regions (the ghost block of a peer is
contiguous, so the receive side needs no unpack).

Cells are evaluated from their own side only, so the result of an owned cell does
not depend on who owns its neighbours: an N-GPU run is bit-identical to the 1-GPU run.

Also holds the two bench drivers (`SingleSWE`, `PartitionedSWE`) that bench.py,
the GPU tests and __graft_entry__.smoke() share.
"""
from __future__ import annotations

import numpy as np

from .mesh import Mesh2D, FACET_NODES

__all__ = ["LocalPart", "partition_mesh", "SingleSWE", "PartitionedSWE", "exchange_halo"]

INT32_MIN = np.iinfo(np.int32).min


class LocalPart:
    """One rank's share of a partitioned mesh (numpy only)."""

    def __init__(self, rank, world, mesh, owned_global, ghost_global, ghost_owner, send_lists):
        self.rank, self.world = rank, world
        self.owned_global = owned_global            # (n_owned,) global cell ids, ascending
        self.ghost_global = ghost_global            # (n_ghost,) grouped by owner rank, ascending inside a group
        self.ghost_owner = ghost_owner              # (n_ghost,)
        self.send_lists = send_lists                # {peer: local owned cell ids the peer needs, in the peer's ghost order}
        self.n_owned = owned_global.shape[0]
        self.n_ghost = ghost_global.shape[0]
        self.recv_counts = np.array([(ghost_owner == p).sum() for p in range(world)], dtype=np.int64)
        self.send_counts = np.array([send_lists[p].shape[0] if p in send_lists else 0 for p in range(world)], dtype=np.int64)
        self.mesh = mesh                            # local Mesh2D: owned cells first, then ghosts


def partition_mesh(mesh: Mesh2D, world: int, halo: str = "facet"):
    """
    Cut the (SFC-ordered) mesh into `world` contiguous chunks with a one-deep halo of ghost cells:
    ``halo='facet'`` = facet neighbours (enough for the SWE / tracer stage kernels),
    ``halo='vertex'`` = every cell sharing a vertex with an owned cell (needed by the vertex-based limiter,
    thetis/limiter.py:100-145: bounds are over all cells around a vertex).
    Deterministic: every rank computes the same partition from the same global mesh.
    Returns a list of `LocalPart`.
    """
    nt = mesh.n_cells
    bounds = np.linspace(0, nt, world + 1).astype(np.int64)
    owner = np.zeros(nt, dtype=np.int32)
    for r in range(world):
        owner[bounds[r]:bounds[r + 1]] = r
    if halo == "vertex":
        v2c_ptr, v2c_idx = mesh.vertex_to_cell_csr()
        tv = mesh.topo[mesh.cells]
    parts = []
    ghosts_of = []
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        if halo == "vertex":
            verts = np.unique(tv[lo:hi].reshape(-1))
            cnt = v2c_ptr[verts + 1] - v2c_ptr[verts]
            starts = np.repeat(v2c_ptr[verts], cnt)
            offs = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            cand = v2c_idx[starts + offs]
            g = np.unique(cand[(cand < lo) | (cand >= hi)])
        elif halo == "facet":
            nb = mesh.nbr[lo:hi]
            g = np.unique(nb[(nb >= 0) & ((nb < lo) | (nb >= hi))])
        else:
            raise ValueError(halo)
        go = owner[g]
        order = np.lexsort((g, go))
        ghosts_of.append((g[order].astype(np.int64), go[order].astype(np.int32)))
    blen = mesh.boundary_length()
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        owned = np.arange(lo, hi, dtype=np.int64)
        gg, go = ghosts_of[r]
        # what I must send to peer p = peer p's ghosts that I own, in p's ghost order
        send = {}
        for p in range(world):
            if p == r:
                continue
            pg, po = ghosts_of[p]
            mine = pg[po == r]
            if mine.size:
                send[p] = (mine - lo).astype(np.int64)
        # local mesh: owned cells first, then ghosts (grouped by owner)
        glob = np.concatenate([owned, gg])
        loc_of = np.full(nt, -1, dtype=np.int64)
        loc_of[glob] = np.arange(glob.shape[0])
        cells_g = mesh.cells[glob]
        vused, vinv = np.unique(cells_g.reshape(-1), return_inverse=True)
        cells_l = vinv.reshape(-1, 3).astype(np.int32)
        coords_l = mesh.coords[vused]
        _, topo_l = np.unique(mesh.topo[vused], return_inverse=True)
        nbr_g = mesh.nbr[glob].astype(np.int64)
        nbr_l = np.full(nbr_g.shape, INT32_MIN, dtype=np.int64)       # unknown: neighbour not on this rank
        pos = nbr_g >= 0
        present = np.zeros_like(pos)
        present[pos] = loc_of[nbr_g[pos]] >= 0
        nbr_l[present] = loc_of[nbr_g[present]]
        n_own = owned.shape[0]
        assert np.all(present[:n_own][pos[:n_own]]), "facet neighbours of owned cells must be local"
        # exterior facets of every local cell (ghosts included: the limiter needs their facet means), renumbered
        bsel = nbr_g < 0
        gb = -(nbr_g[bsel] + 1)
        ub, binv = np.unique(gb, return_inverse=True)
        nbr_l[bsel] = -(1 + binv)
        m = Mesh2D(coords=coords_l, cells=cells_l, topo=topo_l.astype(np.int32), periodic=mesh.periodic)
        m.nbr = nbr_l.astype(np.int32)
        m.nbr_lf = mesh.nbr_lf[glob].copy()
        m.bf_cell = loc_of[mesh.bf_cell[ub]].astype(np.int32)
        m.bf_lf = mesh.bf_lf[ub].copy()
        m.bf_marker = mesh.bf_marker[ub].copy()
        m.meta = dict(mesh.meta)
        m.meta.update(global_bfacets=ub, global_vertices=vused, global_boundary_len=blen, global_cells=glob,
                      n_owned=int(n_own), halo=halo, sfc=True)
        parts.append(LocalPart(r, world, m, owned, gg, go, send))
    return parts


def exchange_halo(part: LocalPart, sendbuf, recv_view, group=None):
    """
    One halo exchange with torch.distributed: sendbuf = packed records grouped by peer,
    recv_view = the ghost block (grouped by owner).  Works with NCCL (GPU) and gloo (CPU).
    """
    import torch.distributed as dist
    rec = sendbuf.shape[-1] if sendbuf.dim() > 1 else 1
    ins = [int(c) * rec for c in part.send_counts]
    outs = [int(c) * rec for c in part.recv_counts]
    try:
        dist.all_to_all_single(recv_view.view(-1), sendbuf.view(-1), output_split_sizes=outs, input_split_sizes=ins,
                               group=group)
    except (RuntimeError, NotImplementedError):
        # backends without all-to-all (older gloo): pairwise non-blocking send/recv
        reqs = []
        so = np.concatenate([[0], np.cumsum(ins)])
        ro = np.concatenate([[0], np.cumsum(outs)])
        sflat, rflat = sendbuf.view(-1), recv_view.view(-1)
        for p in range(part.world):
            if p == part.rank:
                continue
            if outs[p]:
                reqs.append(dist.irecv(rflat[ro[p]:ro[p + 1]], src=p, group=group))
            if ins[p]:
                reqs.append(dist.isend(sflat[so[p]:so[p + 1]].contiguous(), dst=p, group=group))
        for q in reqs:
            q.wait()


class HaloPlan:
    """
    Per-rank halo machinery shared by the integrators and the limiter of a distributed run:
    buffer allocation (symmetric memory when available), the per-stage exchange, and the
    boundary-first / interior-overlapped launch of the SWE stage kernel.
    """

    def __init__(self, parts, rank, transport="auto", overlap=True):
        self.parts = parts
        self.part = parts[rank]
        self.rank, self.world = rank, len(parts)
        self.requested_transport = transport
        self.want_overlap = overlap
        self.transport = None
        self.engine = None
        self.overlap = False
        self._groups = {}          # rec_len -> dict(L, tensor, handle, bufs, dst_ptrs)
        self._buf_info = {}        # data_ptr -> (rec_len, buffer index)
        self._sendbuf = {}

    # ------------------------------------------------------------ set-up (needs the engine: patch size, device)
    def attach(self, engine):
        if self.engine is not None:
            return
        import torch
        self.torch = torch
        self.engine = engine
        p = self.part
        send_idx = np.concatenate([p.send_lists[q] for q in range(self.world) if q in p.send_lists]) \
            if p.send_lists else np.zeros(0, np.int64)
        self.n_send = int(send_idx.shape[0])
        self.send_idx = torch.as_tensor(send_idx.astype(np.int32)).to(engine.device)
        self._send_idx_np = send_idx
        P = engine.patch_size
        self._pads = [((q.n_owned + P - 1) // P) * P for q in self.parts]
        self.transport = "nccl"
        if self.requested_transport in ("auto", "symm"):
            try:
                import torch.distributed._symmetric_memory as symm_mem   # noqa: F401
                self._symm_mem = symm_mem
                self._probe = self._alloc_symmetric(1, 1)       # fails early if symmetric memory is unusable
                self.transport = "symm"
            except Exception as exc:                            # noqa: BLE001
                if self.requested_transport == "symm":
                    raise
                self.symm_error = repr(exc)
        if self.want_overlap and self.n_send:
            bp = np.unique(send_idx // P)
            mask = np.zeros(engine.n_patches, dtype=bool)
            mask[bp] = True
            self._plist_b = torch.as_tensor(np.nonzero(mask)[0].astype(np.int32)).to(engine.device)
            self._plist_i = torch.as_tensor(np.nonzero(~mask)[0].astype(np.int32)).to(engine.device)
            self._comm_stream = torch.cuda.Stream(device=engine.device, priority=-1)   # boundary work goes first
            self._ev_b = torch.cuda.Event()
            self._ev_x = torch.cuda.Event()
            self.overlap = self._plist_b.numel() > 0 and self._plist_i.numel() > 0

    def _alloc_symmetric(self, rec, nbuf):
        import torch.distributed as dist
        torch, eng = self.torch, self.engine
        lens = [(self._pads[r] + self.parts[r].n_ghost) * rec for r in range(self.world)]
        L = (max(lens) + 31) // 32 * 32                 # same size on every rank; 256-B aligned sub-buffers (TMA)
        t = self._symm_mem.empty(nbuf * L, dtype=torch.float64, device=eng.device)
        t.zero_()
        hdl = self._symm_mem.rendezvous(t, dist.group.WORLD)
        return L, t, hdl

    def alloc(self, rec, nbuf=3):
        """`nbuf` state arrays of record length `rec` (9 = SWE, 3 = tracer) whose ghost blocks peers can write."""
        torch, eng, p = self.torch, self.engine, self.part
        n_local = (eng.n_owned_pad + p.n_ghost) * rec
        if self.transport != "symm":
            bufs = [torch.zeros(n_local, dtype=torch.float64, device=eng.device) for _ in range(nbuf)]
            for b, t in enumerate(bufs):
                self._buf_info[t.data_ptr()] = (rec, None, b)
            if rec not in self._sendbuf:
                self._sendbuf[rec] = torch.zeros((max(self.n_send, 1), rec), dtype=torch.float64, device=eng.device)
            return bufs
        L, t, hdl = self._alloc_symmetric(rec, nbuf)
        bufs = [t[b * L:b * L + n_local] for b in range(nbuf)]
        ptrs = [int(x) for x in hdl.buffer_ptrs]
        dst = np.zeros((nbuf, max(self.n_send, 1)), dtype=np.uint64)
        e = 0
        for q in range(self.world):
            if q not in p.send_lists:
                continue
            n = p.send_lists[q].shape[0]
            # my cells sit in q's ghost block after the ghosts owned by lower ranks, in q's ghost order
            first = int((self.parts[q].ghost_owner < self.rank).sum())
            slot = self._pads[q] + first + np.arange(n, dtype=np.int64)
            for b in range(nbuf):
                dst[b, e:e + n] = np.uint64(ptrs[q]) + ((b * L + slot * rec) * 8).astype(np.uint64)
            e += n
        grp = dict(L=L, tensor=t, handle=hdl,
                   dst_ptrs=[torch.as_tensor(dst[b].view(np.int64)).to(eng.device) for b in range(nbuf)])
        gid = len(self._groups)
        self._groups[gid] = grp
        for b, bt in enumerate(bufs):
            self._buf_info[bt.data_ptr()] = (rec, gid, b)
        return bufs

    # ------------------------------------------------------------ per-stage exchange
    def exchange(self, state):
        """Make the ghost block of `state` current on every rank (stream ordered on the current stream)."""
        eng, p = self.engine, self.part
        rec, gid, b = self._buf_info[state.data_ptr()]
        if self.transport == "symm":
            grp = self._groups[gid]
            if self.n_send:
                eng.push_cells(state, self.send_idx, grp["dst_ptrs"][b], rec)
            grp["handle"].barrier(channel=0)        # every rank's stores have landed before anyone reads its ghosts
            return
        sb = self._sendbuf[rec]
        if self.n_send:
            eng.gather_cells(state, self.send_idx, rec, sb)
        g0 = eng.n_owned_pad * rec
        ghost = state[g0:g0 + p.n_ghost * rec].view(-1, rec)
        exchange_halo(p, sb[:self.n_send], ghost)

    def swe_stage(self, a0, a1, bdt, src, u0, dst):
        """Stage kernel + halo exchange of its output; boundary patches first so the exchange overlaps the rest."""
        eng, torch = self.engine, self.torch
        if not self.overlap:
            eng.swe_stage(a0, a1, bdt, src, u0, dst)
            self.exchange(dst)
            return
        # The partition-boundary patches and the interior patches write disjoint parts of dst and read the same src:
        # they run CONCURRENTLY, the (few) boundary patches on a high-priority side stream followed by the push and
        # the cross-rank barrier, the interior patches on the main stream filling the SMs the boundary launch leaves
        # idle.  The next stage waits for both.
        main = torch.cuda.current_stream(eng.device)
        self._ev_b.record(main)                                  # everything src / u0 depend on
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(self._ev_b)
            eng.set_patch_list(self._plist_b)
            eng.swe_stage(a0, a1, bdt, src, u0, dst)             # partition-boundary patches
            self.exchange(dst)                                    # push + barrier while the interior computes
            self._ev_x.record(self._comm_stream)
        eng.set_patch_list(self._plist_i)
        eng.swe_stage(a0, a1, bdt, src, u0, dst)                 # interior patches
        eng.set_patch_list(None)
        main.wait_event(self._ev_x)                              # boundary patches + ghosts of dst are complete

    def kernels_per_swe_stage(self):
        return (2 if self.overlap else 1) + (1 if self.n_send else 0)

    def allreduce_sum(self, t):
        import torch.distributed as dist
        dist.all_reduce(t)
        return t

    def allreduce(self, t, ops):
        """Element-wise global reduction of a small device vector: ops[i] in 's' (sum), 'm' (min), 'M' (max)."""
        import torch.distributed as dist
        for kind, op in (("s", dist.ReduceOp.SUM), ("m", dist.ReduceOp.MIN), ("M", dist.ReduceOp.MAX)):
            idx = [i for i, c in enumerate(ops) if c == kind]
            if idx:
                sub = t[idx].contiguous()
                dist.all_reduce(sub, op=op)
                t[idx] = sub
        return t


def distribute_mesh(mesh: Mesh2D, rank=None, world=None, halo="vertex", transport="auto", overlap=True):
    """
    This rank's share of `mesh` as a shim mesh (owned cells first, then ghosts) carrying a `HaloPlan`; the
    integrators, the limiter and FlowSolver2d pick the plan up from the mesh.  Analogue of Firedrake distributing a
    mesh over COMM_WORLD; rank / world default to torch.distributed's.
    """
    import torch.distributed as dist
    from .shim import ShimMesh
    if rank is None:
        rank = dist.get_rank()
    if world is None:
        world = dist.get_world_size()
    parts = partition_mesh(mesh, world, halo=halo)
    lm = parts[rank].mesh
    sm = ShimMesh(lm)
    sm.boundary_len = dict(lm.meta["global_boundary_len"])     # boundary lengths are global sums (utility.py:821-832)
    sm.halo_plan = HaloPlan(parts, rank, transport=transport, overlap=overlap)
    sm.global_mesh = mesh
    return sm


# ---------------------------------------------------------------------- bench drivers
def _make_solver(mesh, setup, wd, n_owned=None):
    """FlowSolver2d mirror configured for the North Sea workload (thetis_b200/workloads.py)."""
    from . import solver2d
    from .shim import Function, FunctionSpace, Constant, ShimMesh, as_shim_mesh
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
        if getattr(self.ts, "stage_graphs", None):
            self._replays = getattr(self, "_replays", 0) + 1
        self.t += self.dt
        if self.ts.fused_norms is None:
            self.eng.swe_integrals(self.ts.device_state(), self._norms)
        self._norms_host.copy_(self._norms, non_blocking=True)   # else: reduced by the last stage (tb_stage_integrals)

    def e2e_path(self):
        return ("FlowSolver2d mirror -> SSPRK33.advance(t, update_forcings) -> C-ABI: tidal elevation Function updated "
                "on the host every stage (H2D from pinned memory), print_state norms reduced on the device (fused "
                "into the last stage kernel) and read back every step")

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

    def __init__(self, mesh, setup, rank, world, wd=True, transport="auto", overlap=True, halo="facet"):
        self.rank, self.world = rank, world
        sm = distribute_mesh(mesh, rank, world, halo=halo, transport=transport, overlap=overlap)
        self.part = sm.halo_plan.part
        super().__init__(sm, localize_setup(setup, self.part.mesh), wd=wd)
        self.plan = sm.halo_plan
        self.transport = self.plan.transport
        self.overlap = self.plan.overlap

    def n_owned(self):
        return self.part.n_owned

    def stage_launches_per_step(self):
        return 3 * (2 if self.plan.overlap else 1)

    def launches_per_step(self):
        return 3 * self.plan.kernels_per_swe_stage()

    def e2e_path(self):
        return ("distribute_mesh -> " + SingleSWE.e2e_path(self) + "; one halo exchange per RK stage ("
                + self.plan.transport + ")")

    def owned_nodal(self):
        uv, eta = self.state_nodal()
        n = self.part.n_owned
        return uv[:n], eta[:n]
