"""
Multi-rank host logic on CPU (gloo, world_size 2 and 3): mesh partitioning with a one-deep halo and the
per-stage halo exchange.  Each rank evaluates the numpy oracle on its LOCAL sub-mesh (owned + ghost cells) with
ghost records received over the exchange; the owned part must equal the global oracle.  No GPU involved.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from thetis_b200.mesh import rectangle_mesh, sfc_renumber, read_gmsh
from thetis_b200.parallel import partition_mesh, exchange_halo
from oracle import swe_oracle as O

HERE = os.path.dirname(__file__)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mesh():
    return sfc_renumber(read_gmsh(os.path.join(HERE, "golden", "mini_tagged.msh")))


def _global_state(mesh):
    x = mesh.coords[mesh.cells]
    uv = np.stack([0.3 * np.sin(x[..., 0] / 9.0) + 0.1, 0.2 * np.cos(x[..., 1] / 7.0)], -1)
    eta = 0.4 * np.cos(x[..., 0] / 11.0) * np.sin(x[..., 1] / 8.0)
    return uv, eta


def test_partition_covers_mesh_and_halo_is_one_deep():
    mesh = _mesh()
    for world in (2, 3, 4):
        parts = partition_mesh(mesh, world)
        owned = np.concatenate([p.owned_global for p in parts])
        assert np.array_equal(np.sort(owned), np.arange(mesh.n_cells))
        for p in parts:
            # every facet neighbour of an owned cell is owned or a ghost; ghosts are exactly those neighbours
            nb = mesh.nbr[p.owned_global]
            ext = np.unique(nb[(nb >= 0) & ~np.isin(nb, p.owned_global)])
            assert np.array_equal(np.sort(p.ghost_global), ext)
            # send lists mirror the peers' ghost lists
            for q, lst in p.send_lists.items():
                peer = parts[q]
                assert np.array_equal(p.owned_global[lst], peer.ghost_global[peer.ghost_owner == p.rank])
            # local boundary facets keep their markers
            gb = p.mesh.meta["global_bfacets"]
            assert np.array_equal(p.mesh.bf_marker, mesh.bf_marker[gb])
            assert p.mesh.boundary_length() != {} or p.mesh.n_bfacets == 0


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = _mesh()
        part = partition_mesh(mesh, world)[rank]
        uv, eta = _global_state(mesh)
        glob = np.concatenate([part.owned_global, part.ghost_global])
        # local records: owned from the global state, ghosts zero until exchanged
        rec = np.concatenate([uv.reshape(-1, 6), eta], axis=1)
        local = np.zeros((glob.shape[0], 9))
        local[:part.n_owned] = rec[part.owned_global]
        send_idx = np.concatenate([part.send_lists[q] for q in range(world) if q in part.send_lists]) \
            if part.send_lists else np.zeros(0, np.int64)
        sendbuf = torch.as_tensor(local[send_idx].copy())
        ghost = torch.zeros((part.n_ghost, 9), dtype=torch.float64)
        exchange_halo(part, sendbuf, ghost)
        local[part.n_owned:] = ghost.numpy()
        assert np.array_equal(local[part.n_owned:], rec[part.ghost_global])      # bit-exact transport
        # oracle on the local sub-mesh: ghost cells' unknown facets are irrelevant for owned cells
        lm = part.mesh
        lm2 = type(lm)(coords=lm.coords, cells=lm.cells, topo=lm.topo)
        nbr = lm.nbr.copy().astype(np.int64)
        unknown = nbr == np.iinfo(np.int32).min
        # treat unknown facets of ghost cells as (fake) closed boundary facets appended after the real ones
        nfake = int(unknown.sum())
        nbr[unknown] = -(1 + lm.n_bfacets + np.arange(nfake))
        lm2.nbr = nbr.astype(np.int32)
        lm2.nbr_lf = lm.nbr_lf
        cu, fu = np.nonzero(unknown)
        lm2.bf_cell = np.concatenate([lm.bf_cell, cu]).astype(np.int32)
        lm2.bf_lf = np.concatenate([lm.bf_lf, fu]).astype(np.int8)
        lm2.bf_marker = np.concatenate([lm.bf_marker, np.full(nfake, 999)]).astype(np.int32)
        gv = lm.meta["global_vertices"]
        bath_v = 30.0 + 5 * np.sin(mesh.coords[:, 0] / 7.0)
        orc = O.SWEOracle(lm2, bath_v[gv][lm.cells], bnd_conditions={100: {"elev": 0.3, "uv": (0.0, 0.0)}})
        orc.boundary_len = dict(mesh.boundary_length())
        orc.boundary_len[999] = 1.0
        luv = local[:, :6].reshape(-1, 3, 2)
        leta = local[:, 6:]
        ku, ke = orc.tendency(luv, leta)
        gorc = O.SWEOracle(mesh, bath_v[mesh.cells], bnd_conditions={100: {"elev": 0.3, "uv": (0.0, 0.0)}})
        gu, ge = gorc.tendency(uv, eta)
        eu = np.abs(ku[:part.n_owned] - gu[part.owned_global]).max() / np.abs(gu).max()
        ee = np.abs(ke[:part.n_owned] - ge[part.owned_global]).max() / np.abs(ge).max()
        out[rank] = (eu, ee)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        eu, ee = out[r]
        assert eu < 1e-13 and ee < 1e-13, (r, eu, ee)


# ---------------------------------------------------------------- SURVEY 8f rows on a partitioned mesh (CPU, gloo)
def _local_oracle(mesh, part, bath_v, **kw):
    """numpy oracle on a rank's local sub-mesh (owned + ghost cells); unknown facets of ghost cells become fake
    closed facets -- they only affect the ghost cells' own (discarded) residuals"""
    lm = part.mesh
    lm2 = type(lm)(coords=lm.coords, cells=lm.cells, topo=lm.topo)
    nbr = lm.nbr.copy().astype(np.int64)
    unknown = nbr == np.iinfo(np.int32).min
    nfake = int(unknown.sum())
    nbr[unknown] = -(1 + lm.n_bfacets + np.arange(nfake))
    lm2.nbr = nbr.astype(np.int32)
    lm2.nbr_lf = lm.nbr_lf
    cu, fu = np.nonzero(unknown)
    lm2.bf_cell = np.concatenate([lm.bf_cell, cu]).astype(np.int32)
    lm2.bf_lf = np.concatenate([lm.bf_lf, fu]).astype(np.int8)
    lm2.bf_marker = np.concatenate([lm.bf_marker, np.full(nfake, 999)]).astype(np.int32)
    gv = lm.meta["global_vertices"]
    orc = O.SWEOracle(lm2, bath_v[gv][lm.cells], **kw)
    orc.boundary_len = dict(mesh.boundary_length())
    orc.boundary_len[999] = 1.0
    return orc


def _exchange(part, world, local):
    """ghost rows of `local` (n_owned + n_ghost, 9) <- owners' rows, like HaloPlan.exchange"""
    send_idx = np.concatenate([part.send_lists[q] for q in range(world) if q in part.send_lists]) \
        if part.send_lists else np.zeros(0, np.int64)
    sendbuf = torch.as_tensor(local[send_idx].copy())
    ghost = torch.zeros((part.n_ghost, 9), dtype=torch.float64)
    exchange_halo(part, sendbuf, ghost)
    local[part.n_owned:] = ghost.numpy()


def _worker_erk_visc(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = _mesh()
        part = partition_mesh(mesh, world)[rank]
        uv, eta = _global_state(mesh)
        bath_v = 30.0 + 5 * np.sin(mesh.coords[:, 0] / 7.0)
        kw = dict(options=dict(use_grad_div_viscosity_term=True), fields={"viscosity_h": 0.8},
                  bnd_conditions={100: {"elev": 0.3, "uv": (0.0, 0.0)}})
        orc = _local_oracle(mesh, part, bath_v, **kw)
        glob = np.concatenate([part.owned_global, part.ghost_global])
        rec = np.concatenate([uv.reshape(-1, 6), eta], axis=1)
        n_own = part.n_owned
        U = rec[glob].copy()                       # owned + ghost records (ghosts start current, like upload())
        a, b, c, _ = O.ERK_TABLEAUX["ERKLSPUM2"]
        dt, nsteps = 0.02, 3

        def tend(S):
            ku, ke = orc.tendency(S[:, :6].reshape(-1, 3, 2), S[:, 6:], dt=dt)
            K = np.zeros_like(S)
            K[:n_own] = np.concatenate([ku.reshape(-1, 6), ke], axis=1)[:n_own]
            _exchange(part, world, K)              # the tendency buffers are exchanged like state arrays
            return K

        for _ in range(nsteps):
            Ks = []
            for i in range(len(b)):
                S = U + sum(a[i][j] * Ks[j] for j in range(i)) if i else U      # owned AND ghost rows (tb_lincomb)
                if i < len(b) - 1:
                    Ks.append(tend(S))
                else:
                    # last stage: new solution into a separate buffer, ghosts by exchange (rungekutta.ERKGeneric)
                    ku, ke = orc.tendency(S[:, :6].reshape(-1, 3, 2), S[:, 6:], dt=dt)
                    Kl = np.concatenate([ku.reshape(-1, 6), ke], axis=1)
                    Un = np.zeros_like(U)
                    Un[:n_own] = (U + sum(b[j] * Ks[j] for j in range(i)))[:n_own] + b[i] * Kl[:n_own]
                    _exchange(part, world, Un)
                    U = Un
        out[rank] = (part.owned_global.copy(), U[:n_own].copy())
    finally:
        dist.destroy_process_group()


def test_distributed_butcher_erk_with_viscosity_gloo():
    """Butcher-form ERK + SIPG viscosity on 2 ranks with a one-deep facet halo == the global run: the neighbour's
    gradient needs the ghost cell's three nodal values and its geometry only, and exchanged tendencies make the
    stage inputs of ghost cells current without an extra exchange"""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_erk_visc, args=(world, _free_port(), out), nprocs=world, join=True)
    mesh = _mesh()
    uv, eta = _global_state(mesh)
    bath_v = 30.0 + 5 * np.sin(mesh.coords[:, 0] / 7.0)
    gorc = O.SWEOracle(mesh, bath_v[mesh.cells], options=dict(use_grad_div_viscosity_term=True),
                       fields={"viscosity_h": 0.8}, bnd_conditions={100: {"elev": 0.3, "uv": (0.0, 0.0)}})
    a, b, c, _ = O.ERK_TABLEAUX["ERKLSPUM2"]
    st = O.ButcherStepper(gorc, [uv, eta], 0.02, a, b, c)
    for i in range(3):
        st.advance(i * 0.02)
    ref = np.concatenate([uv.reshape(-1, 6), eta], axis=1)
    for r in range(world):
        owned, U = out[r]
        assert np.abs(U - ref[owned]).max() < 1e-13 * np.abs(ref).max()


def _worker_allreduce(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from thetis_b200.parallel import HaloPlan
        plan = HaloPlan(partition_mesh(_mesh(), world), rank)
        t = torch.tensor([1.0 + rank, 10.0 * (rank + 1), 5.0 - rank, -3.0 + 2 * rank], dtype=torch.float64)
        plan.allreduce(t, "ssmM")
        out[rank] = t.numpy().copy()
    finally:
        dist.destroy_process_group()


def test_callback_allreduce_sum_min_max_gloo():
    """device callbacks on a distributed mesh: integrals are summed, extrema are min / max-reduced (callback.py:477-481)"""
    world = 3
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_allreduce, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        assert np.array_equal(out[r], np.array([1.0 + 2.0 + 3.0, 10.0 + 20.0 + 30.0, 3.0, 1.0]))


def _worker_auto_dt(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from thetis_b200 import solver2d
        from thetis_b200.parallel import distribute_mesh
        from thetis_b200.shim import Function, FunctionSpace
        sm = distribute_mesh(_mesh(), rank, world, halo="vertex", transport="nccl")
        b = Function(FunctionSpace(sm, "CG", 1)).interpolate(lambda x, y: 5.0 + 0.4 * x)    # deepest at large x
        s = solver2d.FlowSolver2d(sm, b)
        assert s.options.swe_timestepper_options.use_automatic_timestep      # the reference's default (options.py:26)
        s.create_function_spaces()
        s.create_fields()
        s.set_time_step()
        out[rank] = float(s.dt)
    finally:
        dist.destroy_process_group()


def test_automatic_timestep_is_min_reduced_over_ranks_gloo():
    """solver2d.py:241: dt = comm.allreduce(dt, op=MPI.MIN) -- every rank of a distributed run must advance with the same
    time step, and it is the most restrictive one (close to the serial value: the local P1 projections differ from
    the global one only near the cuts)."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_auto_dt, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] == out[1]
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, as_shim_mesh
    sm = as_shim_mesh(_mesh())
    b = Function(FunctionSpace(sm, "CG", 1)).interpolate(lambda x, y: 5.0 + 0.4 * x)
    s = solver2d.FlowSolver2d(sm, b)
    s.create_function_spaces()
    s.create_fields()
    s.set_time_step()
    assert abs(out[0] - s.dt) / s.dt < 0.05


@pytest.mark.parametrize("halo", ["facet", "vertex"])
def test_fused_push_tables_deliver_every_ghost(halo):
    """Host logic of the fused compute + halo-push launches, emulated in numpy for 3 ranks: with the launch order,
    the per-patch push entries (`fused_push_tables`) and the destination slots `HaloPlan.alloc` computes, every ghost
    slot of every rank receives exactly the record of the owned cell it mirrors, and only partition-boundary patches
    push (they come first in the launch order, which is a permutation of all patches)."""
    from thetis_b200.parallel import fused_push_tables
    mesh = _mesh()
    world, P = 3, 16                                     # small patches so that a rank has several boundary patches
    parts = partition_mesh(mesh, world, halo=halo)
    pads = [((q.n_owned + P - 1) // P) * P for q in parts]
    # "device" arrays: owned cells (padded) then ghosts; the value of a cell is its global id
    arrays = [np.full(pads[r] + parts[r].n_ghost, -1.0) for r in range(world)]
    for r, p in enumerate(parts):
        arrays[r][: p.n_owned] = p.owned_global
    for r, p in enumerate(parts):
        send_idx = np.concatenate([p.send_lists[q] for q in range(world) if q in p.send_lists])
        n_patches = pads[r] // P
        order, push_ptr, push_cell, perm = fused_push_tables(send_idx, n_patches, P)
        assert np.array_equal(np.sort(order), np.arange(n_patches))
        n_b = push_ptr.shape[0] - 1
        assert n_b == np.unique(send_idx // P).shape[0] and push_ptr[-1] == send_idx.shape[0]
        # destination (rank, slot) of every send entry, as in HaloPlan.alloc
        dst = []
        for q in range(world):
            if q not in p.send_lists:
                continue
            n = p.send_lists[q].shape[0]
            first = int((parts[q].ghost_owner < r).sum())
            dst += [(q, pads[q] + first + k) for k in range(n)]
        for b in range(n_b):                              # CTA b of the launch = patch order[b]
            for e in range(push_ptr[b], push_ptr[b + 1]):
                cell = order[b] * P + push_cell[e]
                assert cell == send_idx[perm[e]]
                q, slot = dst[perm[e]]
                arrays[q][slot] = arrays[r][cell]
    for r, p in enumerate(parts):
        assert np.array_equal(arrays[r][pads[r]:], p.ghost_global.astype(float))
