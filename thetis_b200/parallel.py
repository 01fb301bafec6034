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

The bench / test drivers built on top of this (`SingleSWE`, `PartitionedSWE`,
`ConfigRun`) live in the top-level `harness/` package, outside the product.
"""
from __future__ import annotations

import numpy as np

from .mesh import Mesh2D, FACET_NODES

__all__ = ["LocalPart", "PeerInfo", "partition_mesh", "HaloPlan", "distribute_mesh", "exchange_halo",
           "mark_unknown_facets", "build_overlap_connectivity", "local_contribution", "part_from_gathered",
           "plan_from_local_mesh", "init_torch_distributed_from_comm"]

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


def fused_push_tables(send_idx, n_patches, P):
    """
    Tables of the fused compute + halo-push launches (tb_halo_fused_setup) from this rank's send list (owned cell
    ids, peer by peer, in the order `HaloPlan.alloc` computes the destination addresses):
    `order`  launch order of the patches, those holding a cell some peer needs first;
    `push_ptr` / `push_cell`  CSR over those leading patches: for CTA b the cells (index inside the patch) it pushes,
    one entry per (cell, peer);  `perm`  maps the entries back to positions in `send_idx` (so that entry e goes to
    destination address dst[perm[e]]).
    """
    send_idx = np.asarray(send_idx, dtype=np.int64)
    bp = np.unique(send_idx // P).astype(np.int64)
    mask = np.zeros(n_patches, dtype=bool)
    mask[bp] = True
    order = np.concatenate([bp, np.nonzero(~mask)[0]]).astype(np.int32)
    pos = np.full(n_patches, -1, dtype=np.int64)
    pos[bp] = np.arange(bp.shape[0])
    key = pos[send_idx // P]
    perm = np.argsort(key, kind="stable")
    cnt = np.bincount(key, minlength=bp.shape[0]) if send_idx.size else np.zeros(0, np.int64)
    push_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    push_cell = (send_idx % P)[perm].astype(np.int32)
    return order, push_ptr, push_cell, perm


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

    def __init__(self, parts, rank, transport="auto", overlap=True, fused=True):
        self.parts = parts
        self.part = parts[rank]
        self.rank, self.world = rank, len(parts)
        self.requested_transport = transport
        self.want_overlap = overlap
        self.requested_fused = fused
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
        self.fused = False
        if self.transport == "symm" and self.requested_fused:
            self._setup_fused(send_idx, P)
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

    def _setup_fused(self, send_idx, P):
        """Tables of the fused compute + halo-push launch (tb_swe_stage_fused): launch order with the partition-
        boundary patches first, per-patch push entries, and the per-peer epoch flags in symmetric memory."""
        import torch.distributed as dist
        torch, eng, p = self.torch, self.engine, self.part
        order, push_ptr, push_cell, self._push_perm = fused_push_tables(send_idx, eng.n_patches, P)
        nflag = max(self.world, 4)
        self._flags = self._symm_mem.empty(nflag, dtype=torch.int64, device=eng.device)
        self._flags.zero_()
        self._flags_hdl = self._symm_mem.rendezvous(self._flags, dist.group.WORLD)
        ptrs = [int(x) for x in self._flags_hdl.buffer_ptrs]
        send_peers = [q for q in range(self.world) if q in p.send_lists and p.send_lists[q].shape[0]]
        recv_peers = [q for q in range(self.world) if p.recv_counts[q] > 0]
        eng.halo_fused_setup(order, push_ptr, push_cell, recv_peers, [ptrs[q] + 8 * self.rank for q in send_peers],
                             ptrs[self.rank])
        torch.cuda.synchronize(eng.device)
        dist.barrier()            # every rank's flags are zeroed and registered before anybody publishes into them
        self.fused = True

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
        if self.fused:
            # the same addresses grouped by patch, for the fused launches' epilogue push (SWE: 9 doubles per cell,
            # tracer stage / limiter: 3)
            grp["push_dst"] = [torch.as_tensor(dst[b][:self.n_send][self._push_perm].view(np.int64).copy()).to(eng.device)
                               if self.n_send else torch.zeros(1, dtype=torch.int64, device=eng.device)
                               for b in range(nbuf)]
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
        import torch.distributed as dist
        if dist.get_backend() == "gloo":
            # host-staged transport: gloo moves CPU tensors only.  Slow and synchronous -- it exists so that N ranks
            # can share ONE GPU (debugging, and the partition / ghost logic in a single-GPU test run)
            recv_h = self.torch.empty(ghost.shape, dtype=ghost.dtype)
            exchange_halo(p, sb[:self.n_send].cpu(), recv_h)
            ghost.copy_(recv_h)
            return
        exchange_halo(p, sb[:self.n_send], ghost)

    def swe_stage(self, a0, a1, bdt, src, u0, dst, fused=True):
        """Stage kernel + halo exchange of its output; boundary patches first so the exchange overlaps the rest.
        ``fused``: one launch that pushes from its epilogue and signals per-peer flags (safe when ghost blocks are
        only ever read by stage launches, i.e. the Shu-Osher integrators); otherwise boundary launch + push kernel +
        cross-rank barrier."""
        eng, torch = self.engine, self.torch
        if fused and self.fused:
            rec, gid, b = self._buf_info[dst.data_ptr()]
            eng.swe_stage_fused(a0, a1, bdt, src, u0, dst, self._groups[gid]["push_dst"][b])
            return
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

    def tracer_stage(self, a0, a1, bdt, src, u0, dst, swe_state):
        """Tracer stage + halo exchange of its output.  Fused: one launch (the SWE, tracer and limiter launches share
        one epoch sequence, so its boundary patches also see the SWE ghosts of the last SWE stage)."""
        eng = self.engine
        if self.fused:
            rec, gid, b = self._buf_info[dst.data_ptr()]
            eng.tracer_stage_fused(a0, a1, bdt, src, u0, dst, swe_state, self._groups[gid]["push_dst"][b])
            return
        eng.tracer_stage(a0, a1, bdt, src, u0, dst, swe_state)
        self.exchange(dst)

    def limiter_apply_to(self, src, dst):
        """Limiter (out of place) + halo exchange of the limited values."""
        eng = self.engine
        if self.fused:
            rec, gid, b = self._buf_info[dst.data_ptr()]
            eng.limiter_apply_to_fused(src, dst, self._groups[gid]["push_dst"][b])
            return
        eng.limiter_apply_to(src, dst)
        self.exchange(dst)

    def wait_ghosts(self):
        """Stream-ordered: the ghost records written by the peers' last fused stage launch have arrived."""
        if self.fused:
            self.engine.halo_fused_wait()

    def kernels_per_swe_stage(self):
        if self.fused:
            return 1
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


def distribute_mesh(mesh: Mesh2D, rank=None, world=None, halo="vertex", transport="auto", overlap=True, fused=True):
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
    sm.halo_plan = HaloPlan(parts, rank, transport=transport, overlap=overlap, fused=fused)
    sm.global_mesh = mesh
    return sm


# ---------------------------------------------------------------------------------------------------------------
# A mesh that is ALREADY distributed (Firedrake under mpiexec: DMPlex has cut the mesh, every rank holds its owned
# cells followed by an overlap of ghost cells, `firedrake/mesh.py` distribution_parameters["overlap_type"]).  The
# plan is built from what a rank knows locally -- its cells, which of them it owns and a global cell numbering --
# plus ONE all-gather of the id lists.  No rank ever sees the global mesh.
# ---------------------------------------------------------------------------------------------------------------
class PeerInfo:
    """What `HaloPlan` needs to know about another rank: sizes and the owner-grouping of its ghost block."""

    def __init__(self, rank, n_owned, n_ghost, ghost_owner):
        self.rank, self.n_owned, self.n_ghost = rank, int(n_owned), int(n_ghost)
        self.ghost_owner = np.asarray(ghost_owner, dtype=np.int32)


def mark_unknown_facets(mesh: Mesh2D, unknown):
    """
    Turn the boundary facets flagged in `unknown` (bool over the mesh's boundary facets) into facets whose neighbour
    is not on this rank (INT32_MIN in `nbr`, the convention of `partition_mesh`): the outer edge of the overlap of a
    distributed mesh has one cell locally but is not part of the domain boundary.  In place; returns the mesh.
    """
    unknown = np.asarray(unknown, dtype=bool)
    if not unknown.any():
        return mesh
    nbr = mesh.nbr.astype(np.int64)
    nbr[mesh.bf_cell[unknown], mesh.bf_lf[unknown]] = INT32_MIN
    keep = np.nonzero(~unknown)[0]
    nbr[mesh.bf_cell[keep], mesh.bf_lf[keep]] = -(1 + np.arange(keep.shape[0], dtype=np.int64))
    mesh.nbr = nbr.astype(np.int32)
    mesh.bf_cell, mesh.bf_lf, mesh.bf_marker = mesh.bf_cell[keep], mesh.bf_lf[keep], mesh.bf_marker[keep]
    return mesh


def build_overlap_connectivity(mesh: Mesh2D, exterior_edges):
    """
    Connectivity of a rank-local mesh with an overlap: `exterior_edges` = {(topological vertex a, b): marker} of the
    facets that lie on the DOMAIN boundary (Firedrake: `mesh.exterior_facets`, which carries the overlap cells' too);
    every other one-sided facet is the outer edge of the overlap and gets an unknown neighbour.
    """
    sentinel = INT32_MIN + 1
    mesh.build_connectivity(edge_markers={(min(a, b), max(a, b)): mk for (a, b), mk in exterior_edges.items()},
                            default_marker=sentinel)
    return mark_unknown_facets(mesh, mesh.bf_marker == sentinel)


def _owned_boundary_length(mesh: Mesh2D, n_owned):
    """Length of the exterior facets of the OWNED cells per marker (each facet counted on exactly one rank)."""
    sel = mesh.bf_cell < n_owned
    x = mesh.cell_coords()
    p = x[mesh.bf_cell[sel], FACET_NODES[mesh.bf_lf[sel], 0]]
    q = x[mesh.bf_cell[sel], FACET_NODES[mesh.bf_lf[sel], 1]]
    ln = np.hypot(*(q - p).T)
    return {int(mk): float(ln[mesh.bf_marker[sel] == mk].sum()) for mk in np.unique(mesh.bf_marker[sel])}


def local_contribution(mesh: Mesh2D, n_owned, global_ids):
    """This rank's entry of the all-gather behind `part_from_gathered` (small: two id lists and a dict)."""
    gids = np.asarray(global_ids, dtype=np.int64)
    if gids.shape[0] != mesh.n_cells:
        raise ValueError("one global id per local cell (owned cells first, then the overlap)")
    return dict(owned=gids[:n_owned].copy(), ghost=gids[n_owned:].copy(),
                boundary_len=_owned_boundary_length(mesh, n_owned))


def _owner_table(gathered):
    """(sorted global ids of all owned cells, their owner ranks): the lookup table behind `_owners_of`."""
    allo = np.concatenate([g["owned"] for g in gathered])
    allr = np.concatenate([np.full(g["owned"].shape[0], r, dtype=np.int32) for r, g in enumerate(gathered)])
    so = np.argsort(allo, kind="stable")
    allo, allr = allo[so], allr[so]
    if np.any(allo[1:] == allo[:-1]):
        raise ValueError("a cell is owned by two ranks")
    return allo, allr


def _owners_of(ids, table):
    """Owner rank of every global cell id in `ids` (-1: owned by nobody)."""
    allo, allr = table
    if ids.size == 0:
        return np.zeros(0, dtype=np.int32)
    if allo.size == 0:
        return np.full(ids.shape[0], -1, dtype=np.int32)
    pos = np.minimum(np.searchsorted(allo, ids), allo.shape[0] - 1)
    return np.where(allo[pos] == ids, allr[pos], -1).astype(np.int32)


def part_from_gathered(mesh: Mesh2D, n_owned, global_ids, gathered, rank, halo="facet", renumber=True):
    """
    `LocalPart` of `rank` + the `PeerInfo` of the others from a local mesh (connectivity built, overlap edge marked
    with `mark_unknown_facets`), the global cell ids and the gathered `local_contribution`s.

    The device layout wants owned cells along a space-filling curve and the ghost block grouped by owner (a peer
    writes ONE contiguous run of it, in ascending global id): the local cells are permuted accordingly;
    `part.mesh.cell_perm[new] = old` local index.  Returns (parts, part) with parts[rank] is part.
    """
    from .mesh import hilbert_index, sfc_renumber
    world = len(gathered)
    gids = np.asarray(global_ids, dtype=np.int64)
    n_owned = int(n_owned)
    ghost_g = gids[n_owned:]
    table = _owner_table(gathered)
    go = _owners_of(ghost_g, table)
    if np.any(go < 0) or np.any(go == rank):
        raise ValueError("overlap cell owned by no other rank: global ids and ownership are inconsistent")
    own_nbr = mesh.nbr[:n_owned]
    if np.any(own_nbr == INT32_MIN):
        raise ValueError("a facet neighbour of an owned cell is not on this rank: the mesh needs an overlap of at "
                         "least one cell (Firedrake's default)")
    gorder = np.lexsort((ghost_g, go))
    if renumber and n_owned:
        oorder = np.argsort(hilbert_index(mesh.cell_centroids()[:n_owned]), kind="stable")
    else:
        oorder = np.arange(n_owned)
    perm = np.concatenate([oorder, n_owned + gorder]).astype(np.int64)
    lm = sfc_renumber(mesh, perm)
    new_gids = gids[perm]
    # what I send to peer q = q's ghosts that I own, in q's ghost order (= ascending global id inside my run)
    so = np.argsort(new_gids[:n_owned], kind="stable")
    sorted_owned = new_gids[:n_owned][so]
    send, peers = {}, []
    for q, g in enumerate(gathered):
        qo = _owners_of(g["ghost"], table) if q != rank else go
        peers.append(PeerInfo(q, g["owned"].shape[0], g["ghost"].shape[0], np.sort(qo)))
        if q == rank:
            continue
        mine = np.sort(g["ghost"][qo == rank])
        if mine.size:
            pos = np.searchsorted(sorted_owned, mine)
            assert np.array_equal(sorted_owned[pos], mine)
            send[q] = so[pos].astype(np.int64)
    blen = {}
    for g in gathered:
        for mk, ln in g["boundary_len"].items():
            blen[mk] = blen.get(mk, 0.0) + ln
    lm.meta.update(global_cells=new_gids, global_boundary_len=blen, n_owned=n_owned, halo=halo, sfc=True)
    part = LocalPart(rank, world, lm, new_gids[:n_owned], new_gids[n_owned:], go[gorder], send)
    peers[rank] = part
    return peers, part


def torch_allgather(obj):
    """All-gather of a picklable object over torch.distributed's default group (gloo or NCCL)."""
    import torch.distributed as dist
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def plan_from_local_mesh(mesh: Mesh2D, n_owned, global_ids, rank=None, allgather=None, halo="facet",
                         transport="auto", overlap=True, fused=True, renumber=True):
    """
    HaloPlan of a mesh that arrives already distributed: `mesh` = this rank's cells, the `n_owned` owned ones first
    (connectivity built; facets on the outer edge of the overlap marked with `mark_unknown_facets`), `global_ids` =
    a global number per local cell.  `allgather(obj) -> [obj of rank 0, ...]` defaults to torch.distributed's; an
    mpi4py communicator's `comm.allgather` (Firedrake's `mesh.comm`) works as well, provided both number the ranks
    alike.  `halo` says what the overlap holds ('facet' neighbours, or 'vertex': every cell around an owned cell's
    vertices -- the limiter needs that, DistributedMeshOverlapType.VERTEX on the Firedrake side).
    Returns (plan, part); part.mesh is the device-ordered local mesh, part.mesh.cell_perm[new] = old local cell.
    """
    if allgather is None:
        allgather = torch_allgather
    gathered = allgather(local_contribution(mesh, n_owned, global_ids))
    if rank is None:
        import torch.distributed as dist
        rank = dist.get_rank()
    parts, part = part_from_gathered(mesh, n_owned, global_ids, gathered, rank, halo=halo, renumber=renumber)
    return HaloPlan(parts, rank, transport=transport, overlap=overlap, fused=fused), part


def init_torch_distributed_from_comm(comm, backend=None):
    """
    Bring torch.distributed up with the rank numbering of an MPI communicator (mpi4py's interface: `rank`, `size`,
    `bcast`, `allgather`) -- the situation of a Thetis script started with `mpiexec -n N`, where nobody set
    RANK / WORLD_SIZE / MASTER_ADDR.  Rank 0 picks a free TCP port and broadcasts its address; every rank binds the
    GPU with the index of its position among the ranks of its host (one process per GPU).  A group that is already
    initialised is only checked for the same numbering.  Returns (rank, world).
    """
    import socket
    import torch
    import torch.distributed as dist
    rank, world = int(comm.rank), int(comm.size)
    if dist.is_initialized():
        if dist.get_rank() != rank or dist.get_world_size() != world:
            raise RuntimeError(f"torch.distributed is rank {dist.get_rank()} of {dist.get_world_size()} but the mesh's "
                               f"communicator is rank {rank} of {world}: both must number the processes alike")
        return rank, world
    host = socket.gethostname()
    hosts = comm.allgather(host)
    local_rank = sum(1 for h in hosts[:rank] if h == host)
    one_node = all(h == hosts[0] for h in hosts)
    addr_port = None
    if rank == 0:
        s = socket.socket()
        s.bind(("127.0.0.1" if one_node else "", 0))
        port = s.getsockname()[1]
        s.close()
        addr = "127.0.0.1"
        if not one_node:
            try:
                addr = socket.gethostbyname(host)
            except OSError:
                addr = host
        addr_port = (addr, port)
    addr, port = comm.bcast(addr_port, root=0)
    use_cuda = torch.cuda.is_available()
    kw = {}
    if use_cuda:
        dev = local_rank % torch.cuda.device_count()
        torch.cuda.set_device(dev)
        if (backend or "nccl") == "nccl":
            kw["device_id"] = torch.device("cuda", dev)      # binds the communicator to this GPU (as bench.py does)
    dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), init_method=f"tcp://{addr}:{port}",
                            rank=rank, world_size=world, **kw)
    return rank, world
