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


# ---------------------------------------------------------------- the route of a mesh that arrives distributed
def _scrambled_view(mesh, owner, rank, halo, rng):
    """Rank-local view with arbitrary cell / vertex numbering and no neighbour table (see tests/test_local_mesh_plan.py)."""
    from thetis_b200.mesh import Mesh2D
    from thetis_b200.parallel import build_overlap_connectivity
    owned = np.nonzero(owner == rank)[0]
    if halo == "facet":
        nb = mesh.nbr[owned]
        cand = np.unique(nb[nb >= 0])
        ghost = cand[owner[cand] != rank]
    else:
        ghost = _vertex_neighbours(mesh, owned)
    gids = np.concatenate([rng.permutation(owned), rng.permutation(ghost)]).astype(np.int64)
    cells_g = mesh.cells[gids]
    vused = rng.permutation(np.unique(cells_g))
    vloc = np.full(mesh.n_vertices, -1, dtype=np.int64)
    vloc[vused] = np.arange(vused.shape[0])
    lm = Mesh2D(coords=mesh.coords[vused], cells=vloc[cells_g].astype(np.int32).reshape(-1, 3),
                topo=np.unique(mesh.topo[vused], return_inverse=True)[1].astype(np.int32), periodic=mesh.periodic)
    ext = {}
    cl, fl = np.nonzero(mesh.nbr[gids] < 0)
    for c_loc, f in zip(cl, fl):
        a, b = lm.topo[lm.cells[c_loc, FACET_NODES[f]]]
        ext[(int(a), int(b))] = int(mesh.bf_marker[-(mesh.nbr[gids[c_loc], f] + 1)])
    build_overlap_connectivity(lm, ext)
    return lm, owned.shape[0], gids


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(kind=st.sampled_from(["rect", "delaunay"]), n=st.integers(3, 8), m=st.integers(3, 6), seed=st.integers(0, 50),
       world=st.integers(1, 6), halo=st.sampled_from(["facet", "vertex"]), P=st.sampled_from([4, 16, 128]))
def test_plan_from_scrambled_local_views(kind, n, m, seed, world, halo, P):
    """Random ownership (not contiguous, ranks may own scattered cells), random local numberings: the parts built from
    local knowledge + one all-gather own every cell once, hold exactly the overlap they were given grouped by owner,
    mirror each other's ghost runs, keep neighbours and markers, and the fused push built on them delivers every ghost."""
    from thetis_b200.parallel import local_contribution, part_from_gathered
    INT32_MIN = np.iinfo(np.int32).min
    mesh = _mesh(kind, n, m, seed)
    rng = np.random.default_rng(seed)
    world = min(world, mesh.n_cells)
    # blocks of the SFC order dealt to random ranks: scattered but not pathological; every rank owns something
    nblk = max(world, min(mesh.n_cells, 3 * world))
    blk = (np.arange(mesh.n_cells) * nblk) // mesh.n_cells
    deal = np.concatenate([rng.permutation(world), rng.integers(0, world, nblk - world)])
    owner = deal[blk].astype(np.int32)
    views = [_scrambled_view(mesh, owner, r, halo, rng) for r in range(world)]
    gathered = [local_contribution(lm, k, g) for lm, k, g in views]
    built = [part_from_gathered(lm, k, g, gathered, r, halo=halo)[1] for r, (lm, k, g) in enumerate(views)]
    assert np.array_equal(np.sort(np.concatenate([p.owned_global for p in built])), np.arange(mesh.n_cells))
    cent = mesh.cell_centroids()
    for r, p in enumerate(built):
        lm, k, g = views[r]
        assert np.array_equal(np.sort(p.ghost_global), np.sort(g[k:]))
        assert np.array_equal(p.ghost_owner, owner[p.ghost_global]) and np.all(np.diff(p.ghost_owner) >= 0)
        glob = np.concatenate([p.owned_global, p.ghost_global])
        assert np.allclose(p.mesh.cell_centroids(), cent[glob], atol=1e-9)
        # neighbours: local ids point at the right global cells; unknown only on ghosts; exterior markers kept
        nb_l, nb_g = p.mesh.nbr, mesh.nbr[glob]
        # the local cell may list its facets in the global cell's order only (vertex order is kept by the view)
        loc = nb_l >= 0
        assert np.array_equal(glob[nb_l[loc]], nb_g[loc])
        unk = nb_l == INT32_MIN
        assert not unk[:p.n_owned].any() and np.all(nb_g[unk] >= 0)
        ext = (nb_l < 0) & ~unk
        assert np.array_equal(ext, nb_g < 0)
        assert np.array_equal(p.mesh.bf_marker[-(nb_l[ext] + 1)], mesh.bf_marker[-(nb_g[ext] + 1)])
        for q in range(world):
            if q == r:
                continue
            want = built[q].ghost_global[built[q].ghost_owner == r]
            got = p.owned_global[p.send_lists[q]] if q in p.send_lists else np.zeros(0, np.int64)
            assert np.array_equal(got, want)
    # fused push emulation: every rank's ghost block is filled with the owners' records, nothing else is touched
    rec = np.arange(mesh.n_cells, dtype=np.float64) + 0.5
    pads = [((p.n_owned + P - 1) // P) * P for p in built]
    state = [np.full(pads[r] + p.n_ghost, -1.0) for r, p in enumerate(built)]
    for r, p in enumerate(built):
        state[r][:p.n_owned] = rec[p.owned_global]
    for r, p in enumerate(built):
        if not p.send_lists:
            continue
        send_idx = np.concatenate([p.send_lists[q] for q in range(world) if q in p.send_lists])
        n_patches = pads[r] // P
        order, push_ptr, push_cell, perm = fused_push_tables(send_idx, n_patches, P)
        dst = []
        for q in range(world):
            if q not in p.send_lists:
                continue
            first = int((built[q].ghost_owner < r).sum())
            dst += [(q, pads[q] + first + i) for i in range(p.send_lists[q].shape[0])]
        dst = [dst[e] for e in perm]
        for b in range(push_ptr.shape[0] - 1):
            for e in range(push_ptr[b], push_ptr[b + 1]):
                q, slot = dst[e]
                state[q][slot] = state[r][order[b] * P + push_cell[e]]
    for r, p in enumerate(built):
        assert np.array_equal(state[r][pads[r]:], rec[p.ghost_global])
        assert np.array_equal(state[r][:p.n_owned], rec[p.owned_global])
