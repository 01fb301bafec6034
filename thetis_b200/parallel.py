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


def partition_mesh(mesh: Mesh2D, world: int):
    """
    Cut the (SFC-ordered) mesh into `world` contiguous chunks with a one-deep facet halo.
    Deterministic: every rank computes the same partition from the same global mesh.
    Returns a list of `LocalPart`.
    """
    nt = mesh.n_cells
    bounds = np.linspace(0, nt, world + 1).astype(np.int64)
    # align chunk boundaries to the patch size so that patches never straddle ranks
    owner = np.zeros(nt, dtype=np.int32)
    for r in range(world):
        owner[bounds[r]:bounds[r + 1]] = r
    parts = []
    ghosts_of = []
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        nb = mesh.nbr[lo:hi]
        ext = nb[(nb >= 0) & ((nb < lo) | (nb >= hi))]
        g = np.unique(ext)
        go = owner[g]
        order = np.lexsort((g, go))
        ghosts_of.append((g[order].astype(np.int64), go[order].astype(np.int32)))
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
        # local mesh
        glob = np.concatenate([owned, gg])
        g2l = {}
        loc_of = np.full(nt, -1, dtype=np.int64)
        loc_of[glob] = np.arange(glob.shape[0])
        cells_g = mesh.cells[glob]
        vused, vinv = np.unique(cells_g.reshape(-1), return_inverse=True)
        cells_l = vinv.reshape(-1, 3).astype(np.int32)
        coords_l = mesh.coords[vused]
        topo_l = mesh.topo[vused]
        _, topo_l = np.unique(topo_l, return_inverse=True)
        nbr_g = mesh.nbr[glob].astype(np.int64)
        nbr_l = np.full(nbr_g.shape, INT32_MIN, dtype=np.int64)
        n_own = owned.shape[0]
        own_rows = np.arange(glob.shape[0]) < n_own
        pos = (nbr_g >= 0) & own_rows[:, None]
        nbr_l[pos] = loc_of[nbr_g[pos]]
        assert np.all(nbr_l[pos] >= 0)
        # exterior facets of owned cells, renumbered locally
        bsel = (nbr_g < 0) & own_rows[:, None]
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
        m.meta.update(global_bfacets=ub, global_vertices=vused, global_boundary_len=mesh.boundary_length())
        del g2l
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


# ---------------------------------------------------------------------- bench drivers
def _make_solver(mesh, setup, wd, n_owned=None):
    """FlowSolver2d mirror configured for the North Sea workload (thetis_b200/workloads.py)."""
    from . import solver2d
    from .shim import Function, FunctionSpace, Constant, as_shim_mesh
    sm = as_shim_mesh(mesh)
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
        self.mesh, self.setup = mesh, setup
        self.solver, self.tide = _make_solver(mesh, setup, wd)
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

    def launches(self):
        return self.eng.launch_count()

    def stage_launches_per_step(self):
        return 3

    def step_resident(self):
        self.ts.advance_device()

    def step_e2e(self):
        self.ts.advance(self.t, self.update_forcings)
        self.t += self.dt
        self.eng.swe_integrals(self.ts.device_state(), self._norms)
        self._norms_host.copy_(self._norms, non_blocking=True)

    def e2e_path(self):
        return ("FlowSolver2d mirror -> SSPRK33.advance(t, update_forcings) -> C-ABI: tidal elevation Function updated "
                "on the host every stage (H2D from pinned memory), print_state norms reduced on the device and read "
                "back every step")

    def h2d_bytes_per_step(self):
        return 3 * self._n_open * 2 * 8

    def d2h_bytes_per_step(self):
        return 4 * 8

    def state_nodal(self):
        self.ts._host_stale = True
        self.ts.sync_to_host()
        s = self.solver
        return (s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2).copy(), s.fields.elev_2d.dat.data_ro.reshape(-1, 3).copy())


class PartitionedSWE:
    """North Sea workload on `world` GPUs: one process per GPU, one halo exchange per RK stage over NCCL."""

    def __init__(self, mesh, setup, rank, world, wd=True, transport="auto", overlap=True):
        """
        transport: 'nccl'  pack + NCCL all-to-all into the ghost block;
                   'symm'  state buffers in torch symmetric memory, send cells stored straight into the peers'
                           ghost blocks over NVLink (tb_push_cells) + one device-side barrier per stage;
                   'auto'  'symm' when symmetric memory can be set up, else 'nccl'.
        """
        import torch
        from . import _lib as L
        from .engine import Engine
        from .workloads import M2_PERIOD
        self.torch = torch
        self.rank, self.world = rank, world
        parts = partition_mesh(mesh, world)
        self.part = parts[rank]
        p = self.part
        lm = p.mesh
        self.eng = eng = Engine(lm, n_owned=p.n_owned)
        gv = lm.meta["global_vertices"]
        eng.set_option(L.OPT_NONLINEAR, 1)
        eng.set_option(L.OPT_LAX_FRIEDRICHS, 1)
        eng.set_option(L.OPT_WETTING_DRYING, bool(wd))
        eng.set_option(L.OPT_WD_ALPHA, setup["wd_alpha"])
        eng.set_field(L.F_BATHYMETRY, setup["bath"][gv])
        eng.set_field(L.F_MANNING, setup["manning"][gv])
        eng.set_field(L.F_CORIOLIS, setup["coriolis"][gv])
        for mk, ln in lm.meta["global_boundary_len"].items():
            eng.set_boundary_length(mk, ln)
        eng.set_bc(0, 100, L.BC_ELEV | L.BC_UV, [0, 0, 0, 0, 0, 0])
        gb = lm.meta["global_bfacets"]
        self._phase = setup["tide_phase"][gb]
        self._has_open = bool((lm.bf_marker == 100).any())
        self._omega = 2 * np.pi / M2_PERIOD
        self._n_open = int((lm.bf_marker == 100).sum())
        if self._has_open:
            eng.set_bc_array(0, 100, L.BC_ELEV, np.sin(self._phase))
        glob = np.concatenate([p.owned_global, p.ghost_global])
        uv = setup["uv0"][glob]
        eta = setup["eta0"][glob]
        # state: owned (padded) + ghosts
        dev = eng.device
        send_idx = np.concatenate([p.send_lists[q] for q in range(world) if q in p.send_lists]) if p.send_lists else np.zeros(0, np.int64)
        self.transport = "nccl"
        self._symm = None
        if transport in ("auto", "symm"):
            try:
                self._setup_symmetric(parts, send_idx)
                self.transport = "symm"
            except Exception as exc:            # noqa: BLE001  (fall back to NCCL all-to-all, same results)
                if transport == "symm":
                    raise
                self._symm_error = repr(exc)
        if self.transport == "nccl":
            self.buf = [eng.new_state(), eng.new_state(), eng.new_state()]
        A = self.buf[0]
        own = eng.upload_nodal(uv[:p.n_owned], eta[:p.n_owned])
        A[:own.numel()].copy_(own)
        if p.n_ghost:
            grec = np.concatenate([uv[p.n_owned:].reshape(-1, 6), eta[p.n_owned:]], axis=1)
            A[eng.n_owned_pad * 9:eng.n_owned_pad * 9 + grec.size] = torch.as_tensor(grec.reshape(-1)).to(dev)
        self.send_idx = torch.as_tensor(send_idx.astype(np.int32)).to(dev)
        self.sendbuf = torch.zeros((max(int(send_idx.shape[0]), 1), 9), dtype=torch.float64, device=dev)
        self.n_send = int(send_idx.shape[0])
        self.dt = setup["dt"]
        self.t = 0.0
        self._norms = torch.zeros(4, dtype=torch.float64, device=dev)
        self._norms_host = torch.zeros(4, dtype=torch.float64).pin_memory()
        self._L = L
        from .rungekutta import SSPRK33, butcher_to_shuosher_form
        self._alpha, self._beta = butcher_to_shuosher_form(SSPRK33.a, SSPRK33.b)
        self._c = [float(v) for v in SSPRK33.c]
        self.overlap = False
        if overlap:
            self._setup_overlap()

    def _setup_symmetric(self, parts, send_idx):
        """State buffers in symmetric memory; per buffer the peer addresses every send cell must be stored to."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        eng, p, world, rank = self.eng, self.part, self.world, self.rank
        P = eng.patch_size
        pads = [((q.n_owned + P - 1) // P) * P for q in parts]
        lens = [(pads[r] + parts[r].n_ghost) * 9 for r in range(world)]
        L = (max(lens) + 31) // 32 * 32                 # same size on every rank; 256-B aligned sub-buffers (TMA)
        t = symm_mem.empty(3 * L, dtype=torch.float64, device=eng.device)
        t.zero_()
        hdl = symm_mem.rendezvous(t, dist.group.WORLD)
        self._symm = (t, hdl)
        self.buf = [t[i * L:i * L + eng.state_len] for i in range(3)]
        ptrs = [int(x) for x in hdl.buffer_ptrs]
        dst = np.zeros((3, max(send_idx.shape[0], 1)), dtype=np.uint64)
        e = 0
        for q in range(world):
            if q not in p.send_lists:
                continue
            n = p.send_lists[q].shape[0]
            # my cells sit in q's ghost block after the ghosts owned by lower ranks, in q's ghost order
            first = int((parts[q].ghost_owner < rank).sum())
            slot = pads[q] + first + np.arange(n, dtype=np.int64)
            for b in range(3):
                dst[b, e:e + n] = np.uint64(ptrs[q]) + ((b * L + slot * 9) * 8).astype(np.uint64)
            e += n
        self._dst_ptrs = [torch.as_tensor(dst[b].view(np.int64)).to(eng.device) for b in range(3)]
        self._buf_index = {self.buf[b].data_ptr(): b for b in range(3)}

    def _exchange(self, state):
        eng, p = self.eng, self.part
        if self.transport == "symm":
            b = self._buf_index[state.data_ptr()]
            if self.n_send:
                eng.push_cells(state, self.send_idx, self._dst_ptrs[b], 9)
            self._symm[1].barrier(channel=0)        # every rank's stores have landed before anyone reads its ghosts
            return
        if self.n_send:
            eng.gather_cells(state, self.send_idx, 9, self.sendbuf)
        ghost = state[eng.n_owned_pad * 9:eng.n_owned_pad * 9 + p.n_ghost * 9].view(-1, 9)
        exchange_halo(p, self.sendbuf[:self.n_send], ghost)

    def _setup_overlap(self):
        """Patches holding cells a peer needs run first; their halo push overlaps the remaining patches."""
        torch = self.torch
        eng = self.eng
        P = eng.patch_size
        bp = np.unique(self.send_idx.cpu().numpy().astype(np.int64) // P) if self.n_send else np.zeros(0, np.int64)
        mask = np.zeros(eng.n_patches, dtype=bool)
        mask[bp] = True
        self._plist_b = torch.as_tensor(np.nonzero(mask)[0].astype(np.int32)).to(eng.device)
        self._plist_i = torch.as_tensor(np.nonzero(~mask)[0].astype(np.int32)).to(eng.device)
        self._comm_stream = torch.cuda.Stream(device=eng.device)
        self._ev_b = torch.cuda.Event()
        self._ev_x = torch.cuda.Event()
        self.overlap = self._plist_b.numel() > 0 and self._plist_i.numel() > 0

    def _stage(self, a0, a1, bdt, src, u0, dst):
        eng, torch = self.eng, self.torch
        if not getattr(self, "overlap", False):
            eng.swe_stage(a0, a1, bdt, src, u0, dst)
            self._exchange(dst)
            return
        main = torch.cuda.current_stream(eng.device)
        eng.set_patch_list(self._plist_b)
        eng.swe_stage(a0, a1, bdt, src, u0, dst)                 # partition-boundary patches
        self._ev_b.record(main)
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(self._ev_b)
            self._exchange(dst)                                   # push + barrier while the interior computes
            self._ev_x.record(self._comm_stream)
        eng.set_patch_list(self._plist_i)
        eng.swe_stage(a0, a1, bdt, src, u0, dst)                 # interior patches
        eng.set_patch_list(None)
        main.wait_event(self._ev_x)                              # ghosts of dst are complete before the next stage

    def _step(self, forcing=None):
        A, B, C = self.buf
        dt = self.dt
        al, be, c = self._alpha, self._beta, self._c      # the reference's own Shu-Osher coefficients
        if forcing:
            forcing(self.t + c[0] * dt)
        self._stage(0.0, float(al[1][0]), float(be[1][0]) * dt, A, None, B)
        if forcing:
            forcing(self.t + c[1] * dt)
        self._stage(float(al[2][0]), float(al[2][1]), float(be[2][1]) * dt, B, A, C)
        if forcing:
            forcing(self.t + c[2] * dt)
        self._stage(float(al[3][0]), float(al[3][2]), float(be[3][2]) * dt, C, A, A)

    def enable_graph(self):
        """Capture one whole resident step (3 stages incl. halo traffic) in a CUDA graph: at 8 GPUs a stage is ~45 us
        of GPU work, less than what the Python launch path costs."""
        torch = self.torch
        self._step()                                             # warm-up outside capture (lazy uploads, allocations)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step()
        self._graph = g
        self._graph_launches = self.launches_per_step()

    def launches_per_step(self):
        per_stage = (2 if getattr(self, "overlap", False) else 1) + (1 if self.n_send else 0) + \
            (0 if self.transport == "symm" else 0)
        return 3 * per_stage

    def update_forcings(self, t):
        if self._has_open:
            if not hasattr(self, "_tide_buf"):
                self._open_rows = np.nonzero(self.part.mesh.bf_marker == 100)[0]
                self._open_phase = self._phase[self._open_rows]
                self._tide_buf = np.zeros_like(self._phase)
            self._tide_buf[self._open_rows] = np.sin(self._omega * t + self._open_phase)
            self.eng.set_bc_array(0, 100, self._L.BC_ELEV, self._tide_buf)

    def n_owned(self):
        return self.part.n_owned

    def launches(self):
        return self.eng.launch_count() + getattr(self, "_replays", 0) * getattr(self, "_graph_launches", 0)

    def stage_launches_per_step(self):
        return 6 if getattr(self, "overlap", False) else 3

    def step_resident(self):
        g = getattr(self, "_graph", None)
        if g is not None:
            g.replay()
            self._replays = getattr(self, "_replays", 0) + 1
        else:
            self._step()

    def enable_stage_graphs(self):
        """One CUDA graph per RK stage (kernels + halo traffic); the host-side forcing upload stays between them."""
        torch = self.torch
        A, B, C = self.buf
        dt = self.dt
        al, be = self._alpha, self._beta
        args = [(0.0, float(al[1][0]), float(be[1][0]) * dt, A, None, B),
                (float(al[2][0]), float(al[2][1]), float(be[2][1]) * dt, B, A, C),
                (float(al[3][0]), float(al[3][2]), float(be[3][2]) * dt, C, A, A)]
        self._step()
        torch.cuda.synchronize()
        self._stage_graphs = []
        for a in args:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._stage(*a)
            self._stage_graphs.append(g)
        # graph capture ran each stage once more on real data: harmless for the bench (state stays finite), and
        # tests that need exact step counts use _step() directly

    def step_e2e(self):
        sg = getattr(self, "_stage_graphs", None)
        if sg is None:
            self._step(self.update_forcings)
        else:
            for i in range(3):
                self.update_forcings(self.t + self._c[i] * self.dt)     # host forcing -> H2D, stream ordered
                sg[i].replay()
            self._replays_stage = getattr(self, "_replays_stage", 0) + 1
        self.t += self.dt
        self.eng.swe_integrals(self.buf[0], self._norms)
        self._norms_host.copy_(self._norms, non_blocking=True)

    def e2e_path(self):
        return ("PartitionedSWE driver -> C-ABI (tb_swe_stage + tb_push_cells, one CUDA graph per RK stage): tidal "
                "elevation computed on the host every stage and copied H2D from pinned memory (tb_set_bc_array), "
                "print_state integrals reduced on the device and read back every step")

    def h2d_bytes_per_step(self):
        return 3 * self._n_open * 2 * 8

    def d2h_bytes_per_step(self):
        return 4 * 8

    def owned_nodal(self):
        return self.eng.download_nodal(self.buf[0])
