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
