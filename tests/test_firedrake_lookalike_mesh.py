"""
CPU test of `adaptor._mesh2d_from_firedrake` and of the node maps built on it, against a LOOK-ALIKE of a Firedrake
mesh (Firedrake itself cannot be installed here).  The look-alike exposes exactly the attributes the
extraction reads -- the ones the reference itself uses:

    mesh.coordinates.function_space().cell_node_map().values       mesh.coordinates.dat.data_ro_with_halos
    mesh.cell_set.size          firedrake.FunctionSpace(mesh, "CG", 1).cell_node_map().values
    mesh.exterior_facets.facet_cell / .local_facet_dat.data_ro / .markers     (limiter.py:140-144, equation.py:29)

with Firedrake's conventions: local facet i is opposite local vertex i (FIAT), cells may have either orientation,
the coordinate Dat carries halo rows that no owned cell references, exterior-facet arrays have shape (n, 1), and a
periodic mesh has DG coordinates while the P1 space identifies the periodic vertices.

What this pins: the extraction logic (vertex compaction, topological ids, orientation fix, exterior-facet markers by
topological edge) and the node maps handed to the device (`dg_node_map`, `nodal_values`, `bfacet_values`).  What it
cannot pin: that real Firedrake objects behave like the look-alike.
"""
import sys
import types

import numpy as np
import pytest

from thetis_b200.mesh import rectangle_mesh, periodic_rectangle_mesh, FACET_NODES
from thetis_b200 import adaptor as AD


class _Map:
    def __init__(self, values):
        self.values = values
        self.arity = values.shape[1]


class _Element:
    def __init__(self, family, degree):
        self._f, self._d = family, degree

    def family(self):
        return self._f

    def degree(self):
        return self._d


class _Space:
    def __init__(self, mesh, family, values):
        self._mesh, self._el, self._map = mesh, _Element(family, 1), _Map(values)

    def mesh(self):
        return self._mesh

    def ufl_element(self):
        return self._el

    def cell_node_map(self):
        return self._map


class _Dat:
    def __init__(self, data):
        self.data = data
        self.dat_version = 0

    @property
    def data_ro(self):
        return self.data

    @property
    def data_ro_with_halos(self):
        return self.data


class _Function:
    def __init__(self, space, data):
        self._fs, self.dat = space, _Dat(data)

    def function_space(self):
        return self._fs


class _Facets:
    pass


class _LookalikeMesh:
    """A Firedrake-shaped view of a Mesh2D `m`: vertices renumbered by a random permutation, `n_flip` cells stored
    clockwise, `n_halo_rows` unreferenced rows appended to the coordinate Dat, `n_halo_cells` extra (non-owned) rows
    appended to every cell map."""

    def __init__(self, m, seed=0, n_flip=7, n_halo_rows=5, n_halo_cells=3):
        rng = np.random.default_rng(seed)
        nt = m.n_cells
        self.src = m
        self.flipped = np.zeros(nt, dtype=bool)
        self.flipped[rng.choice(nt, size=min(n_flip, nt), replace=False)] = True
        cells = m.cells.astype(np.int64).copy()
        cells[self.flipped, 1], cells[self.flipped, 2] = m.cells[self.flipped, 2], m.cells[self.flipped, 1]
        if m.periodic:
            # DG coordinates: every cell owns its three coordinate rows
            cmap = np.arange(3 * nt, dtype=np.int64).reshape(nt, 3)
            xy = m.coords[cells].reshape(-1, 2)
            tmap = m.topo[cells].astype(np.int64)            # the P1 space identifies the periodic vertices
            fam = "Discontinuous Lagrange"
        else:
            nv = m.n_vertices
            perm = rng.permutation(nv + n_halo_rows)[:nv]     # vertex v lives in row perm[v]; the other rows are halo
            xy = rng.standard_normal((nv + n_halo_rows, 2)) * 1e9   # garbage in the rows nobody may read
            xy[perm] = m.coords
            cmap = perm[cells]
            tmap = cmap.copy()
            fam = "Lagrange"
        halo_cells = np.zeros((n_halo_cells, 3), dtype=np.int64)
        self._cmap = np.vstack([cmap, halo_cells])
        self._tmap = np.vstack([tmap, halo_cells])
        self.cell_set = types.SimpleNamespace(size=nt)
        self.coordinates = _Function(_Space(self, fam, self._cmap), xy)
        # exterior facets in Firedrake's form; a flipped cell has its local vertices 1 and 2 exchanged, so its local
        # facets 1 and 2 are exchanged too (facet i is opposite vertex i)
        lf = m.bf_lf.astype(np.int64).copy()
        fl = self.flipped[m.bf_cell]
        lf[fl & (m.bf_lf == 1)] = 2
        lf[fl & (m.bf_lf == 2)] = 1
        order = rng.permutation(m.n_bfacets)
        ef = _Facets()
        ef.facet_cell = m.bf_cell.astype(np.int64)[order].reshape(-1, 1)
        ef.local_facet_dat = _Dat(lf[order].reshape(-1, 1))
        ef.markers = m.bf_marker.astype(np.int64)[order]
        ef.unique_markers = np.unique(ef.markers)
        self.exterior_facets = ef

    def p1_space(self):
        return _Space(self, "Lagrange", self._tmap)

    def p1dg_space(self):
        nt = self.cell_set.size
        vals = np.vstack([np.arange(3 * nt, dtype=np.int64).reshape(nt, 3), np.zeros((3, 3), dtype=np.int64)])
        return _Space(self, "Discontinuous Lagrange", vals)


@pytest.fixture
def fake_firedrake(monkeypatch):
    mod = types.ModuleType("firedrake")

    def FunctionSpace(mesh, family, degree):
        if (family, degree) == ("DG", 0):
            return mesh.dg0_space()
        assert family == "CG" and degree == 1
        return mesh.p1_space()
    mod.FunctionSpace = FunctionSpace
    monkeypatch.setitem(sys.modules, "firedrake", mod)
    return mod


def _edge_markers(m):
    """{sorted physical end points of an exterior facet: marker} -- independent of every numbering"""
    x = m.coords[m.cells]
    p = x[m.bf_cell, FACET_NODES[m.bf_lf, 0]]
    q = x[m.bf_cell, FACET_NODES[m.bf_lf, 1]]
    out = {}
    for a, b, mk in zip(np.round(p, 6), np.round(q, 6), m.bf_marker):
        out[tuple(sorted([tuple(a), tuple(b)]))] = int(mk)
    return out


@pytest.mark.parametrize("periodic", [False, True])
def test_extraction_from_a_firedrake_shaped_mesh(fake_firedrake, periodic):
    src = periodic_rectangle_mesh(6, 4, 6.0, 4.0) if periodic else rectangle_mesh(7, 5, 7.0, 5.0)
    fm = _LookalikeMesh(src, seed=3)
    m, swap = AD._mesh2d_from_firedrake(fm)
    assert m.n_cells == src.n_cells and m.n_bfacets == src.n_bfacets
    assert np.array_equal(swap, fm.flipped)                       # exactly the clockwise cells were re-oriented
    assert np.all(m.cell_area() > 0)
    # same cells at the same places: compare centroids and areas cell by cell (owned order is preserved)
    assert np.allclose(m.cell_centroids(), src.cell_centroids(), atol=1e-12)
    assert np.allclose(m.cell_area(), src.cell_area(), rtol=1e-12)
    assert m.periodic == periodic
    # same boundary: every exterior facet found, with the marker of the reference-side facet arrays
    assert _edge_markers(m) == _edge_markers(src)
    # same interior connectivity: as many interior facets, and the two cells of each really share its two vertices
    tv = m.topo[m.cells]
    assert int((m.nbr >= 0).sum()) == int((src.nbr >= 0).sum())
    for c in range(m.n_cells):
        for lf in range(3):
            n = m.nbr[c, lf]
            if n < 0:
                continue
            nl = m.nbr_lf[c, lf]
            mine = {int(tv[c, FACET_NODES[lf, 0]]), int(tv[c, FACET_NODES[lf, 1]])}
            theirs = {int(tv[n, FACET_NODES[nl, 0]]), int(tv[n, FACET_NODES[nl, 1]])}
            assert mine == theirs and m.nbr[n, nl] == c
    # topological vertices: as many as the source mesh has
    assert np.unique(m.topo).size == np.unique(src.topo).size


@pytest.mark.parametrize("periodic", [False, True])
def test_node_maps_of_the_adaptor_on_a_firedrake_shaped_mesh(fake_firedrake, periodic):
    """`MeshAdaptor` on the look-alike: the device cell / CCW node (c, a) must read the dof of the caller's P1DG
    Function that sits at the device mesh's node (c, a) -- through the orientation fix AND the SFC renumbering."""
    src = periodic_rectangle_mesh(6, 4, 6.0, 4.0) if periodic else rectangle_mesh(7, 5, 7.0, 5.0)
    fm = _LookalikeMesh(src, seed=11, n_flip=9)
    ad = AD.MeshAdaptor(fm)                                         # not a Mesh2D: goes through _mesh2d_from_firedrake
    fs = fm.p1dg_space()
    # a P1DG Function in the CALLER's numbering holding the node coordinates (cell-local order of the look-alike)
    cells = src.cells.astype(np.int64).copy()
    cells[fm.flipped, 1], cells[fm.flipped, 2] = src.cells[fm.flipped, 2], src.cells[fm.flipped, 1]
    xdg = src.coords[cells].reshape(-1, 2)
    fx = _Function(fs, xdg[:, 0].copy())
    fy = _Function(fs, xdg[:, 1].copy())
    dev = ad.mesh.coords[ad.mesh.cells]                              # (nt, 3, 2) device cells, CCW
    assert np.allclose(ad.nodal_values(fx), dev[..., 0], atol=1e-12)
    assert np.allclose(ad.nodal_values(fy), dev[..., 1], atol=1e-12)
    nm = ad.dg_node_map(fs)
    assert nm.dtype == np.int32 and nm.shape == (src.n_cells, 3)
    assert np.array_equal(np.sort(nm.reshape(-1)), np.arange(3 * src.n_cells))      # a permutation of the dofs
    # boundary data: values of a Function at the two nodes of every exterior facet of the device mesh
    m = ad.mesh
    if m.n_bfacets:
        for marker in np.unique(m.bf_marker):
            vals = ad.bfacet_values(fx, int(marker))
            sel = m.bf_marker == marker
            want = dev[m.bf_cell, :, 0][np.arange(m.n_bfacets)[:, None], FACET_NODES[m.bf_lf]]
            assert np.allclose(np.asarray(vals)[sel], want[sel], atol=1e-12)


# ---------------------------------------------------------------- a mesh Firedrake has distributed itself (mpiexec -n N)
class _HaloMap:
    """PyOP2 map of a distributed set: `values` stops at the owned entities, `values_with_halo` has them all."""

    def __init__(self, values, n_owned):
        self.values_with_halo = values
        self.values = values[:n_owned]
        self.arity = values.shape[1]


class _HaloDat:
    def __init__(self, data, n_owned_rows):
        self._all, self._n = data, n_owned_rows
        self.dat_version = 0

    @property
    def data_ro(self):
        return self._all[:self._n]

    data = data_ro

    @property
    def data_ro_with_halos(self):
        return self._all

    data_with_halos = data_ro_with_halos


class _HaloSpace(_Space):
    def __init__(self, mesh, family, values, n_owned, degree=1):
        self._mesh, self._el, self._map = mesh, _Element(family, degree), _HaloMap(values, n_owned)


class _DistributedLookalike:
    """Rank `rank`'s view of `src` under the ownership `owner`, the way Firedrake presents a distributed mesh: owned
    cells first, then a VERTEX overlap; private vertex numbering; DG0 dofs numbered globally through an lgmap;
    exterior facets of owned cells first (`facet_cell`), those of overlap cells only behind the `_with_halo`
    accessors; facets on the outer edge of the overlap in neither list."""

    def __init__(self, src, owner, rank, world, seed=0):
        rng = np.random.default_rng(100 * seed + rank)
        owned = np.nonzero(owner == rank)[0]
        ptr, idx = src.vertex_to_cell_csr()
        tv = np.unique(src.topo[src.cells[owned]])
        cand = np.unique(np.concatenate([idx[ptr[v]:ptr[v + 1]] for v in tv]))
        ghost = cand[owner[cand] != rank]
        self.gids = np.concatenate([rng.permutation(owned), rng.permutation(ghost)]).astype(np.int64)
        self.n_owned, n_local = owned.shape[0], self.gids.shape[0]
        cells_g = src.cells[self.gids].astype(np.int64)
        self.flipped = rng.random(n_local) < 0.2
        cells_g[self.flipped, 1], cells_g[self.flipped, 2] = cells_g[self.flipped, 2].copy(), cells_g[self.flipped, 1].copy()
        vused = rng.permutation(np.unique(cells_g))
        self.vused = vused
        vloc = np.full(src.n_vertices, -1, dtype=np.int64)
        vloc[vused] = np.arange(vused.shape[0])
        self._cmap = vloc[cells_g]
        self.local_cells_global_vertices = cells_g
        self.cell_set = types.SimpleNamespace(size=self.n_owned, total_size=n_local)
        self.coordinates = _Function(_HaloSpace(self, "Lagrange", self._cmap, self.n_owned), src.coords[vused].copy())
        self.comm = types.SimpleNamespace(size=world, rank=rank, allgather=None)
        self._distribution_parameters = {"overlap_type": (types.SimpleNamespace(name="VERTEX"), 1)}
        # DG0: local dof = a permutation of the local cells (owned dofs first), global number through the lgmap
        dof_of_cell = np.concatenate([rng.permutation(self.n_owned), self.n_owned + rng.permutation(n_local - self.n_owned)])
        lg = np.empty(n_local, dtype=np.int64)
        lg[dof_of_cell] = 7 * self.gids + 3                     # any injective global numbering will do
        self._dg0 = _HaloSpace(self, "Discontinuous Lagrange", dof_of_cell.reshape(-1, 1), self.n_owned, degree=0)
        self._dg0.dof_dset = types.SimpleNamespace(lgmap=types.SimpleNamespace(indices=lg))
        # exterior facets (of the DOMAIN boundary): owned cells' first
        rows = []
        for c_loc, c in enumerate(self.gids):
            for f in range(3):
                if src.nbr[c, f] < 0:
                    lf = f
                    if self.flipped[c_loc] and f in (1, 2):
                        lf = 3 - f
                    rows.append((c_loc, lf, int(src.bf_marker[-(src.nbr[c, f] + 1)])))
        rows = np.array(rows, dtype=np.int64).reshape(-1, 3)
        own = rows[rows[:, 0] < self.n_owned]
        gh = rows[rows[:, 0] >= self.n_owned]
        rows = np.vstack([own[rng.permutation(own.shape[0])], gh[rng.permutation(gh.shape[0])]])
        ef = _Facets()
        ef.facet_cell_map = _HaloMap(rows[:, :1].copy(), own.shape[0])
        ef.facet_cell = ef.facet_cell_map.values
        ef.local_facet_dat = _HaloDat(rows[:, 1:2].copy(), own.shape[0])
        ef.markers = rows[:, 2].copy()
        self.exterior_facets = ef

    def p1_space(self):
        return _HaloSpace(self, "Lagrange", self._cmap, self.n_owned)

    def dg0_space(self):
        return self._dg0

    def p1dg_function(self, nodal_of_global_vertex):
        """P1DG Function (dofs 3c..3c+2 of local cell c, overlap rows after the owned ones) holding a vertex field."""
        n_local = self.cell_set.total_size
        fs = _HaloSpace(self, "Discontinuous Lagrange", np.arange(3 * n_local, dtype=np.int64).reshape(n_local, 3),
                        self.n_owned)
        f = _Function(fs, None)
        f.dat = _HaloDat(nodal_of_global_vertex[self.local_cells_global_vertices].reshape(-1).copy(), 3 * self.n_owned)
        return f


@pytest.fixture
def no_process_group(monkeypatch):
    """The adaptor brings torch.distributed up from the mesh's communicator; one test process cannot be N ranks, so
    the call is recorded instead (the real thing runs in tests/test_local_mesh_plan.py over two processes)."""
    from thetis_b200 import parallel as PA
    calls = []
    monkeypatch.setattr(PA, "init_torch_distributed_from_comm", lambda comm, backend=None: calls.append(comm))
    return calls


@pytest.mark.parametrize("world", [2, 3])
def test_adaptor_on_a_distributed_firedrake_shaped_mesh(fake_firedrake, no_process_group, world):
    """`MeshAdaptor` on a mesh that arrives distributed: it must take the overlap along, build the halo plan from the
    owned / overlap split and the global DG0 numbering (one all-gather over the mesh's communicator), and its node
    maps must reach the overlap rows of the caller's Functions (`data_ro_with_halos`)."""
    from thetis_b200 import parallel as PA
    src = rectangle_mesh(8, 6, 8.0, 6.0)
    c = src.cell_centroids()
    owner = (np.floor(c[:, 0] / 8.0 * world).astype(np.int32) + (c[:, 1] > 3.0) * 1) % world
    fms = [_DistributedLookalike(src, owner, r, world, seed=2) for r in range(world)]
    # the all-gather: every rank's contribution, computed the way the adaptor computes it
    gathered = []
    for fm in fms:
        m, _ = AD._mesh2d_from_firedrake(fm, with_overlap=True)
        gids = AD._global_cell_ids(fm, m.n_cells)
        assert np.array_equal(gids, 7 * fm.gids + 3)
        gathered.append(PA.local_contribution(m, fm.n_owned, gids))
    blen = src.boundary_length()
    ads = []
    for r, fm in enumerate(fms):
        fm.comm.allgather = lambda obj, _g=gathered: _g
        ad = AD.MeshAdaptor(fm)
        ads.append(ad)
        assert ad.with_halos and ad.halo is not None and ad.halo.rank == r and ad.halo.world == world
        part = ad.halo.part
        assert ad.n_owned == fm.n_owned == part.n_owned and ad.mesh.n_cells == fm.cell_set.total_size
        assert ad.mesh.meta["halo"] == "vertex"
        assert np.all(ad.mesh.cell_area() > 0)
        # owned cells first, ghosts grouped by owner; global ids survive the renumbering
        glob = (ad.mesh.meta["global_cells"] - 3) // 7
        assert np.array_equal(np.sort(glob[:ad.n_owned]), np.nonzero(owner == r)[0])
        assert np.all(np.diff(owner[glob[ad.n_owned:]]) >= 0) and np.all(owner[glob[ad.n_owned:]] != r)
        assert np.allclose(ad.mesh.cell_centroids(), src.cell_centroids()[glob], atol=1e-12)
        # boundary lengths are global; the exterior facets of overlap cells were found, the overlap edge is unknown
        assert set(ad.boundary_len) == set(blen)
        assert all(abs(ad.boundary_len[k] - blen[k]) < 1e-12 for k in blen)
        assert not np.any(ad.mesh.nbr[:ad.n_owned] == np.iinfo(np.int32).min)
        nb_glob = src.nbr[glob]
        assert np.array_equal(np.sort((ad.mesh.nbr < 0) & (ad.mesh.nbr != np.iinfo(np.int32).min), axis=1).sum(1),
                              (nb_glob < 0).sum(1))
        # node maps reach the overlap rows: a P1DG Function holding x is read back at the device mesh's nodes
        fx = fm.p1dg_function(src.coords[:, 0])
        dev = ad.mesh.coords[ad.mesh.cells]
        assert np.allclose(ad.nodal_values(fx), dev[..., 0], atol=1e-12)
        nm = ad.dg_node_map(fx.function_space())
        assert np.array_equal(np.sort(nm.reshape(-1)), np.arange(3 * ad.mesh.n_cells))
        assert nm[:ad.n_owned].max() < 3 * ad.n_owned <= nm[ad.n_owned:].min()
        assert ad.dat_ro(fx).shape[0] == 3 * ad.mesh.n_cells and fx.dat.data_ro.shape[0] == 3 * ad.n_owned
    assert [c.rank for c in no_process_group] == list(range(world))      # every adaptor asked for the process group
    # send lists mirror the peers' ghost runs across the adaptors
    for r, ad in enumerate(ads):
        for q, lst in ad.halo.part.send_lists.items():
            peer = ads[q].halo.part
            assert np.array_equal(ad.halo.part.owned_global[lst], peer.ghost_global[peer.ghost_owner == r])


def test_distributed_mesh_without_overlap_is_rejected(fake_firedrake, no_process_group):
    src = rectangle_mesh(4, 4, 4.0, 4.0)
    owner = (src.cell_centroids()[:, 0] > 2.0).astype(np.int32)
    fm = _DistributedLookalike(src, owner, 0, 2)
    fm._distribution_parameters = {"overlap_type": (types.SimpleNamespace(name="NONE"), 0)}
    fm.comm.allgather = lambda obj: [obj, obj]
    with pytest.raises(NotImplementedError, match="overlap of at least one cell"):
        AD.MeshAdaptor(fm)
