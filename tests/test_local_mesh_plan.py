"""
Halo plan of a mesh that arrives ALREADY distributed (the situation under `mpiexec -n N` with Firedrake: DMPlex cut
the mesh, each rank sees its owned cells + an overlap and a global numbering, nobody sees the global mesh;
reference analogue: every Thetis test that runs `@pytest.mark.parallel(nprocs=2)`,
test/swe2d/test_steady_state_channel.py:6).

Each rank's view is made the way a DMPlex distribution looks from inside: an arbitrary (non-contiguous) ownership,
owned cells and overlap cells in arbitrary order, private vertex numbering, no neighbour table, exterior-facet
markers only for facets of the domain boundary.  `plan_from_local_mesh` must turn that into the device layout
(owned cells on a space-filling curve, ghost block grouped by owner) with send lists that mirror the peers' ghost
blocks, and the numpy oracle evaluated on the local mesh with exchanged ghosts must reproduce the global tendency on
the owned cells.  CPU only (in-process all-gather, and gloo with world_size 2).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from thetis_b200.mesh import Mesh2D, rectangle_mesh, sfc_renumber, read_gmsh, FACET_NODES
from thetis_b200 import parallel as PA
from oracle import swe_oracle as O

HERE = os.path.dirname(__file__)
INT32_MIN = np.iinfo(np.int32).min


def _mesh(kind):
    if kind == "gmsh":
        return sfc_renumber(read_gmsh(os.path.join(HERE, "golden", "mini_tagged.msh")))
    return rectangle_mesh(9, 7, 90.0, 70.0)


def _ownership(mesh, world, seed):
    """Non-contiguous, DMPlex-like ownership: Voronoi cells of `world` random seeds."""
    rng = np.random.default_rng(seed)
    c = mesh.cell_centroids()
    pts = c[rng.choice(mesh.n_cells, world, replace=False)]
    return np.argmin(((c[:, None, :] - pts[None]) ** 2).sum(-1), axis=1).astype(np.int32)


def _local_view(mesh, owner, rank, halo, seed):
    """(Mesh2D without global knowledge, n_owned, global ids, global vertex of each local vertex)."""
    rng = np.random.default_rng(1000 * seed + rank)
    owned = np.nonzero(owner == rank)[0]
    if halo == "facet":
        nb = mesh.nbr[owned]
        cand = np.unique(nb[nb >= 0])
    else:
        ptr, idx = mesh.vertex_to_cell_csr()
        tv = np.unique(mesh.topo[mesh.cells[owned]])
        cand = np.unique(np.concatenate([idx[ptr[v]:ptr[v + 1]] for v in tv]))
    ghost = cand[owner[cand] != rank]
    gids = np.concatenate([rng.permutation(owned), rng.permutation(ghost)]).astype(np.int64)
    cells_g = mesh.cells[gids]
    vused = rng.permutation(np.unique(cells_g))
    vloc = np.full(mesh.n_vertices, -1, dtype=np.int64)
    vloc[vused] = np.arange(vused.shape[0])
    lm = Mesh2D(coords=mesh.coords[vused], cells=vloc[cells_g].astype(np.int32),
                topo=np.unique(mesh.topo[vused], return_inverse=True)[1].astype(np.int32), periodic=mesh.periodic)
    ext = {}
    for c_loc, c in enumerate(gids):
        for f in range(3):
            if mesh.nbr[c, f] < 0:
                a, b = lm.topo[lm.cells[c_loc, FACET_NODES[f]]]
                ext[(int(a), int(b))] = int(mesh.bf_marker[-(mesh.nbr[c, f] + 1)])
    PA.build_overlap_connectivity(lm, ext)
    return lm, owned.shape[0], gids, vused


def _records(mesh):
    x = mesh.coords[mesh.cells]
    uv = np.stack([0.3 * np.sin(x[..., 0] / 9.0) + 0.1, 0.2 * np.cos(x[..., 1] / 7.0)], -1)
    eta = 0.4 * np.cos(x[..., 0] / 11.0) * np.sin(x[..., 1] / 8.0)
    return uv, eta, np.concatenate([uv.reshape(-1, 6), eta], axis=1)


def _closed_copy(lm):
    """The local mesh with the unknown facets of its ghost cells closed by a fake marker (oracle input)."""
    nbr = lm.nbr.astype(np.int64)
    unknown = nbr == INT32_MIN
    nfake = int(unknown.sum())
    nbr[unknown] = -(1 + lm.n_bfacets + np.arange(nfake))
    cu, fu = np.nonzero(unknown)
    m2 = Mesh2D(coords=lm.coords, cells=lm.cells, topo=lm.topo)
    m2.nbr, m2.nbr_lf = nbr.astype(np.int32), lm.nbr_lf
    m2.bf_cell = np.concatenate([lm.bf_cell, cu]).astype(np.int32)
    m2.bf_lf = np.concatenate([lm.bf_lf, fu]).astype(np.int8)
    m2.bf_marker = np.concatenate([lm.bf_marker, np.full(nfake, 999)]).astype(np.int32)
    return m2


def _check_owned_tendency(mesh, part, local_rec, vused):
    lm = part.mesh
    bath_v = 30.0 + 5 * np.sin(mesh.coords[:, 0] / 7.0)
    bc = {100: {"elev": 0.3, "uv": (0.0, 0.0)}, 1: {"un": 0.05}}
    gv = vused[lm.meta["vertex_perm"]]                       # global vertex of each device-ordered local vertex
    assert np.array_equal(mesh.coords[gv], lm.coords)
    orc = O.SWEOracle(_closed_copy(lm), bath_v[gv][lm.cells], bnd_conditions=bc)
    orc.boundary_len = dict(lm.meta["global_boundary_len"])
    orc.boundary_len[999] = 1.0
    ku, ke = orc.tendency(local_rec[:, :6].reshape(-1, 3, 2), local_rec[:, 6:])
    uv, eta, _ = _records(mesh)
    gorc = O.SWEOracle(mesh, bath_v[mesh.cells], bnd_conditions=bc)
    gu, ge = gorc.tendency(uv, eta)
    # the local cell's vertices may be rotated against the global cell's: compare per cell through the vertex ids
    own = part.owned_global
    for k_loc, k_glob in ((ku, gu), (ke, ge)):
        for a in range(3):
            match = mesh.cells[own] == gv[lm.cells[:part.n_owned, a]][:, None]       # (n_owned, 3)
            assert np.all(match.sum(1) == 1)
            ref = k_glob[own][match]
            err = np.abs(k_loc[:part.n_owned, a] - ref).max() / np.abs(k_glob).max()
            assert err < 1e-12, err


@pytest.mark.parametrize("kind", ["gmsh", "rect"])
@pytest.mark.parametrize("halo", ["facet", "vertex"])
@pytest.mark.parametrize("world", [2, 3, 5])
def test_plan_from_local_views_in_process(kind, halo, world):
    mesh = _mesh(kind)
    owner = _ownership(mesh, world, seed=world)
    views = [_local_view(mesh, owner, r, halo, seed=7) for r in range(world)]
    gathered = [PA.local_contribution(lm, n, g) for lm, n, g, _ in views]
    built = [PA.part_from_gathered(lm, n, g, gathered, r, halo=halo) for r, (lm, n, g, _) in enumerate(views)]
    _, _, rec = _records(mesh)
    blen = mesh.boundary_length()
    for r, (peers, part) in enumerate(built):
        lm, n_owned, gids, vused = views[r]
        assert peers[r] is part and part.n_owned == n_owned
        assert np.array_equal(np.sort(part.owned_global), np.nonzero(owner == r)[0])
        assert np.array_equal(part.mesh.meta["global_cells"][part.mesh.cell_perm.argsort()], gids)
        # ghost block: grouped by owner, ascending global id inside a group, owners correct
        assert np.array_equal(part.ghost_owner, owner[part.ghost_global])
        key = part.ghost_owner.astype(np.int64) * mesh.n_cells + part.ghost_global
        assert np.all(np.diff(key) > 0)
        # every rank derives the same description of every peer
        for q in range(world):
            other = built[q][1]
            assert peers[q].n_owned == other.n_owned and peers[q].n_ghost == other.n_ghost
            assert np.array_equal(np.asarray(peers[q].ghost_owner), other.ghost_owner)
        # send lists mirror the peers' ghost runs
        for q in range(world):
            if q == r:
                continue
            want = built[q][1].ghost_global[built[q][1].ghost_owner == r]
            got = part.owned_global[part.send_lists[q]] if q in part.send_lists else np.zeros(0, np.int64)
            assert np.array_equal(got, want)
        assert np.array_equal(part.send_counts, [built[q][1].recv_counts[r] for q in range(world)])
        # boundary lengths are the global ones although no rank saw the global mesh
        gl = part.mesh.meta["global_boundary_len"]
        assert set(gl) == set(blen)
        for mk in blen:
            assert abs(gl[mk] - blen[mk]) < 1e-9 * max(1.0, blen[mk])
        # no unknown facet on an owned cell; ghosts may have them
        assert not np.any(part.mesh.nbr[:n_owned] == INT32_MIN)
    # in-process exchange through the send lists, then the oracle on every rank's local mesh
    for r, (peers, part) in enumerate(built):
        local = np.zeros((part.n_owned + part.n_ghost, 9))
        local[:part.n_owned] = rec[part.owned_global]
        off = part.n_owned
        for q in range(world):
            if q == r or part.recv_counts[q] == 0:
                continue
            src = built[q][1]
            payload = rec[src.owned_global][src.send_lists[r]]
            local[off:off + payload.shape[0]] = payload
            off += payload.shape[0]
        assert off == local.shape[0]
        assert np.array_equal(local[part.n_owned:], rec[part.ghost_global])
        # records are per global cell with the GLOBAL vertex order: rotate them into the local cells' vertex order
        lmesh = part.mesh
        gv = views[r][3][lmesh.meta["vertex_perm"]]
        glob_cells = np.concatenate([part.owned_global, part.ghost_global])
        rot = np.argmax(mesh.cells[glob_cells][:, None, :] == gv[lmesh.cells][:, :, None], axis=2)   # local a -> global a
        uv = local[:, :6].reshape(-1, 3, 2)
        eta = local[:, 6:]
        idx = np.arange(local.shape[0])[:, None]
        local_rot = np.concatenate([uv[idx, rot].reshape(-1, 6), eta[idx, rot]], axis=1)
        _check_owned_tendency(mesh, part, local_rot, views[r][3])


def test_inconsistent_inputs_are_rejected():
    mesh = _mesh("rect")
    owner = _ownership(mesh, 2, seed=3)
    views = [_local_view(mesh, owner, r, "facet", seed=1) for r in range(2)]
    gathered = [PA.local_contribution(lm, n, g) for lm, n, g, _ in views]
    lm, n, g, _ = views[0]
    with pytest.raises(ValueError, match="one global id per local cell"):
        PA.local_contribution(lm, n, g[:-1])
    bad = [dict(gathered[0]), dict(gathered[1])]
    bad[1]["owned"] = np.concatenate([bad[1]["owned"], gathered[0]["owned"][:1]])
    with pytest.raises(ValueError, match="owned by two ranks"):
        PA.part_from_gathered(lm, n, g, bad, 0)
    g2 = g.copy()
    g2[-1] = 10 ** 9
    with pytest.raises(ValueError, match="owned by no other rank"):
        PA.part_from_gathered(lm, n, g2, gathered, 0)
    # no overlap at all: an owned cell's neighbour is missing
    own_only = Mesh2D(coords=lm.coords, cells=lm.cells[:n], topo=lm.topo)
    PA.build_overlap_connectivity(own_only, {})
    with pytest.raises(ValueError, match="overlap of at least one cell"):
        PA.part_from_gathered(own_only, n, g[:n], [dict(owned=g[:n], ghost=g[:0], boundary_len={}),
                                                   gathered[1]], 0)


# ---------------------------------------------------------------- the same over gloo, world_size 2
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = _mesh("gmsh")
        owner = _ownership(mesh, world, seed=11)
        lm, n_owned, gids, vused = _local_view(mesh, owner, rank, "vertex", seed=5)
        plan, part = PA.plan_from_local_mesh(lm, n_owned, gids, halo="vertex")       # torch.distributed all-gather
        assert plan.part is part and plan.world == world and part.mesh.meta["halo"] == "vertex"
        _, _, rec = _records(mesh)
        local = np.zeros((part.n_owned + part.n_ghost, 9))
        local[:part.n_owned] = rec[part.owned_global]
        send_idx = np.concatenate([part.send_lists[q] for q in range(world) if q in part.send_lists])
        ghost = torch.zeros((part.n_ghost, 9), dtype=torch.float64)
        PA.exchange_halo(part, torch.as_tensor(local[send_idx].copy()), ghost)
        out[rank] = bool(np.array_equal(ghost.numpy(), rec[part.ghost_global]))
    finally:
        dist.destroy_process_group()


def test_plan_from_local_mesh_gloo():
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


# ---------------------------------------------------------------- torch.distributed from an MPI-shaped communicator
class _BoardComm:
    """mpi4py-shaped communicator (rank, size, bcast, allgather) over a multiprocessing manager: what a Thetis script
    under `mpiexec` hands over through `mesh.comm`."""

    def __init__(self, rank, size, board, barrier):
        self.rank, self.size, self._board, self._barrier, self._n = rank, size, board, barrier, 0

    def allgather(self, obj):
        self._n += 1
        self._board[(self._n, self.rank)] = obj
        self._barrier.wait()
        return [self._board[(self._n, r)] for r in range(self.size)]

    def bcast(self, obj, root=0):
        return self.allgather(obj)[root]


def _worker_from_comm(rank, world, board, barrier, out):
    for k in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):          # an mpiexec launch sets none of these
        os.environ.pop(k, None)
    comm = _BoardComm(rank, world, board, barrier)
    try:
        assert PA.init_torch_distributed_from_comm(comm) == (rank, world)
        assert dist.get_rank() == rank and dist.get_world_size() == world and dist.get_backend() == "gloo"
        assert PA.init_torch_distributed_from_comm(comm) == (rank, world)     # second call: only the check
        mesh = _mesh("rect")
        owner = _ownership(mesh, world, seed=4)
        lm, n_owned, gids, _ = _local_view(mesh, owner, rank, "facet", seed=9)
        plan, part = PA.plan_from_local_mesh(lm, n_owned, gids, rank=comm.rank, allgather=comm.allgather)
        _, _, rec = _records(mesh)
        send_idx = np.concatenate([part.send_lists[q] for q in range(world) if q in part.send_lists])
        ghost = torch.zeros((part.n_ghost, 9), dtype=torch.float64)
        PA.exchange_halo(part, torch.as_tensor(rec[part.owned_global][send_idx].copy()), ghost)
        ok = bool(np.array_equal(ghost.numpy(), rec[part.ghost_global]))
        wrong = _BoardComm((rank + 1) % world, world, board, barrier)
        try:
            PA.init_torch_distributed_from_comm(wrong)
            ok = False
        except RuntimeError as exc:
            ok = ok and "number the processes alike" in str(exc)
        out[rank] = ok
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_process_group_from_an_mpi_shaped_communicator():
    mgr = mp.Manager()
    out, board, barrier = mgr.dict(), mgr.dict(), mgr.Barrier(2)
    mp.spawn(_worker_from_comm, args=(2, board, barrier, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


# ---------------------------------------------------------------- the two routes agree
@pytest.mark.parametrize("halo", ["facet", "vertex"])
@pytest.mark.parametrize("world", [2, 4])
def test_local_route_reproduces_partition_mesh(halo, world):
    """Fed with the parts `partition_mesh` cuts (the route every multi-GPU measurement went through), the local route
    must arrive at the SAME device layout: cell order, ghost block, send lists, neighbour table, exterior facets and
    boundary lengths -- so what the GPU tests established for one holds for the other."""
    mesh = _mesh("gmsh")
    ref = PA.partition_mesh(mesh, world, halo=halo)
    gathered, inputs = [], []
    for p in ref:
        lm = p.mesh
        bare = Mesh2D(coords=lm.coords, cells=lm.cells, topo=lm.topo, periodic=lm.periodic)
        tv = lm.topo[lm.cells]
        ext = {(int(tv[c, FACET_NODES[f, 0]]), int(tv[c, FACET_NODES[f, 1]])): int(mk)
               for c, f, mk in zip(lm.bf_cell, lm.bf_lf, lm.bf_marker)}
        PA.build_overlap_connectivity(bare, ext)
        gids = np.concatenate([p.owned_global, p.ghost_global])
        inputs.append((bare, p.n_owned, gids))
        gathered.append(PA.local_contribution(bare, p.n_owned, gids))
    for r, p in enumerate(ref):
        bare, n_owned, gids = inputs[r]
        peers, part = PA.part_from_gathered(bare, n_owned, gids, gathered, r, halo=halo, renumber=False)
        assert np.array_equal(part.owned_global, p.owned_global)
        assert np.array_equal(part.ghost_global, p.ghost_global)
        assert np.array_equal(part.ghost_owner, p.ghost_owner)
        assert set(part.send_lists) == set(p.send_lists)
        for q in p.send_lists:
            assert np.array_equal(part.send_lists[q], p.send_lists[q])
        assert np.array_equal(part.send_counts, p.send_counts) and np.array_equal(part.recv_counts, p.recv_counts)
        a, b = part.mesh, p.mesh
        assert np.array_equal(a.coords[a.cells], b.coords[b.cells])              # same cells, same local vertex order
        assert np.array_equal(a.nbr >= 0, b.nbr >= 0) and np.array_equal(a.nbr[a.nbr >= 0], b.nbr[b.nbr >= 0])
        assert np.array_equal(a.nbr == INT32_MIN, b.nbr == INT32_MIN)
        assert np.array_equal(a.nbr_lf[a.nbr >= 0], b.nbr_lf[b.nbr >= 0])
        # exterior facets: same (cell, local facet, marker) set, and nbr points at the right row of each list
        fa = sorted(zip(a.bf_cell.tolist(), a.bf_lf.tolist(), a.bf_marker.tolist()))
        fb = sorted(zip(b.bf_cell.tolist(), b.bf_lf.tolist(), b.bf_marker.tolist()))
        assert fa == fb
        for m in (a, b):
            i = np.arange(m.n_bfacets)
            assert np.array_equal(m.nbr[m.bf_cell, m.bf_lf], -(1 + i))
        gl, bl = a.meta["global_boundary_len"], b.meta["global_boundary_len"]
        assert set(gl) == set(bl) and all(abs(gl[k] - bl[k]) < 1e-9 * max(1.0, bl[k]) for k in bl)
        assert a.meta["n_owned"] == b.meta["n_owned"] and a.meta["halo"] == b.meta["halo"] == halo
        for q in range(world):
            assert (peers[q].n_owned, peers[q].n_ghost) == (ref[q].n_owned, ref[q].n_ghost)
            assert np.array_equal(np.asarray(peers[q].ghost_owner), ref[q].ghost_owner)
