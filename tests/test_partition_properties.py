"""
Property tests (hypothesis) of the host side of the multi-GPU path: `partition_mesh` and
`fused_push_tables` over random meshes (structured, periodic, Delaunay), world sizes 1 - 8 (also more ranks than is
sensible for the mesh) and both halo kinds.  Invariants: every cell is owned exactly once; the ghosts of a rank are
exactly the facet (or vertex) neighbours of its owned cells that it does not own; send lists mirror the peers' ghost
lists entry by entry, grouped by owner in rank order (the layout HaloPlan.alloc and the fused push rely on); the
sub-mesh keeps geometry, markers and boundary lengths; emulating the fused push delivers every ghost record.
"""
import numpy as np
from hypothesis import given, settings, strategies as st, HealthCheck

from thetis_b200.mesh import rectangle_mesh, periodic_rectangle_mesh, delaunay_mesh, sfc_renumber, FACET_NODES
from thetis_b200.parallel import partition_mesh, fused_push_tables


def _mesh(kind, n, m, seed):
    if kind == "rect":
        return sfc_renumber(rectangle_mesh(n, m, 7.0 * n, 5.0 * m))
    if kind == "periodic":
        return sfc_renumber(periodic_rectangle_mesh(max(n, 3), m, 7.0 * n, 5.0 * m))
    return sfc_renumber(delaunay_mesh(max(n * m * 2, 12), 900.0, 700.0, seed=seed))


def _vertex_neighbours(mesh, owned):
    """cells sharing a (topological) vertex with an owned cell, not owned themselves"""
    tv = mesh.topo[mesh.cells]
    mark = np.zeros(int(mesh.topo.max()) + 1, dtype=bool)
    mark[tv[owned].reshape(-1)] = True
    touching = mark[tv].any(axis=1)
    touching[owned] = False
    return np.nonzero(touching)[0]


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(kind=st.sampled_from(["rect", "periodic", "delaunay"]), n=st.integers(2, 9), m=st.integers(2, 7),
       seed=st.integers(0, 50), world=st.integers(1, 8), halo=st.sampled_from(["facet", "vertex"]),
       P=st.sampled_from([4, 16, 128]))
def test_partition_and_push_tables(kind, n, m, seed, world, halo, P):
    mesh = _mesh(kind, n, m, seed)
    world = min(world, mesh.n_cells)                     # at least one cell per rank
    parts = partition_mesh(mesh, world, halo=halo)
    assert len(parts) == world
    owned = np.concatenate([p.owned_global for p in parts])
    assert np.array_equal(np.sort(owned), np.arange(mesh.n_cells))            # a partition: no overlap, no gap
    owner = np.empty(mesh.n_cells, dtype=np.int64)
    for p in parts:
        owner[p.owned_global] = p.rank
    for p in parts:
        og = p.owned_global
        if halo == "facet":
            nb = mesh.nbr[og]
            want = np.unique(nb[(nb >= 0) & (owner[np.maximum(nb, 0)] != p.rank)])
        else:
            want = _vertex_neighbours(mesh, og)
        assert np.array_equal(np.sort(p.ghost_global), want)
        assert np.array_equal(p.ghost_owner, owner[p.ghost_global])
        assert np.all(np.diff(p.ghost_owner) >= 0)                             # grouped by owner, rank order
        assert p.n_owned == og.shape[0] and p.n_ghost == p.ghost_global.shape[0]
        # send lists mirror the peers' ghost lists, entry by entry
        for q, lst in p.send_lists.items():
            peer = parts[q]
            assert q != p.rank and lst.shape[0] > 0
            assert np.array_equal(og[lst], peer.ghost_global[peer.ghost_owner == p.rank])
        for q in range(world):
            if q != p.rank and (parts[q].ghost_owner == p.rank).any():
                assert q in p.send_lists
        # the sub-mesh: owned cells first, then ghosts; geometry and markers of the global mesh
        lm = p.mesh
        glob = np.concatenate([og, p.ghost_global])
        gv = lm.meta["global_vertices"]
        assert np.array_equal(gv[lm.cells], mesh.cells[glob])
        assert np.array_equal(lm.coords, mesh.coords[gv])
        gb = lm.meta["global_bfacets"]
        assert np.array_equal(lm.bf_marker, mesh.bf_marker[gb])
        assert np.array_equal(glob[lm.bf_cell], mesh.bf_cell[gb]) and np.array_equal(lm.bf_lf, mesh.bf_lf[gb])
        # facet neighbours of owned cells are all present locally, with the same local facet on the other side
        nb_l = lm.nbr[: p.n_owned]
        nb_g = mesh.nbr[og]
        inner = nb_g >= 0
        assert np.all(nb_l[inner] >= 0) and np.array_equal(glob[nb_l[inner]], nb_g[inner])
        assert np.array_equal(lm.nbr_lf[: p.n_owned][inner], mesh.nbr_lf[og][inner])
    # fused push emulation (value of a cell = its global id)
    pads = [((q.n_owned + P - 1) // P) * P for q in parts]
    arrays = [np.full(pads[r] + parts[r].n_ghost, -1.0) for r in range(world)]
    for r, p in enumerate(parts):
        arrays[r][: p.n_owned] = p.owned_global
    for r, p in enumerate(parts):
        if not p.send_lists:
            continue
        send_idx = np.concatenate([p.send_lists[q] for q in range(world) if q in p.send_lists])
        n_patches = pads[r] // P
        order, push_ptr, push_cell, perm = fused_push_tables(send_idx, n_patches, P)
        assert np.array_equal(np.sort(order), np.arange(n_patches))
        n_b = push_ptr.shape[0] - 1
        assert push_ptr[-1] == send_idx.shape[0] and n_b == np.unique(send_idx // P).shape[0]
        dst = []
        for q in range(world):
            if q in p.send_lists:
                first = int((parts[q].ghost_owner < r).sum())
                dst += [(q, pads[q] + first + k) for k in range(p.send_lists[q].shape[0])]
        for b in range(n_b):
            for e in range(push_ptr[b], push_ptr[b + 1]):
                cell = order[b] * P + push_cell[e]
                assert cell == send_idx[perm[e]] and 0 <= push_cell[e] < P
                q, slot = dst[perm[e]]
                arrays[q][slot] = arrays[r][cell]
    for r, p in enumerate(parts):
        assert np.array_equal(arrays[r][pads[r]:], p.ghost_global.astype(float))
