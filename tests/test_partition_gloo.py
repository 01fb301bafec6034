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
