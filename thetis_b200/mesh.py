"""
Host-side triangular mesh toolkit (numpy only).

The reference never stores a mesh itself: every mesh comes from Firedrake
(`RectangleMesh`, `PeriodicRectangleMesh`, `UnitSquareMesh`, `Mesh("*.msh")`,
see e.g. /root/reference demos/demo_2d_channel.py:19-21,
test/swe2d/test_rossby_wave.py:146, demos/demo_2d_north_sea.py).  This module
provides the same meshes as plain arrays plus exactly the connectivity the
explicit P1DG path needs:

* ``cells``       (nt, 3)  geometric vertex ids, every cell counter-clockwise
* ``coords``      (nv, 2)  geometric vertex coordinates
* ``topo``        (nv,)    topological vertex id of each geometric vertex
                            (identity unless the mesh is periodic)
* ``nbr``         (nt, 3)  cell across local facet i (facet i is opposite local
                            vertex i, FIAT convention), or -(1+k) where k is the
                            index of the exterior facet in ``bf_*``
* ``nbr_lf``      (nt, 3)  local facet number of that facet in the neighbour
* ``bf_cell``, ``bf_lf``, ``bf_marker``  exterior facets

Because all cells are CCW, the two nodes of a shared facet match crosswise:
node (i+1)%3 of cell K is node (j+2)%3 of neighbour N and vice versa.
"""
from __future__ import annotations

import numpy as np
from dataclasses import dataclass, field

__all__ = [
    "Mesh2D", "rectangle_mesh", "unit_square_mesh", "periodic_rectangle_mesh",
    "read_gmsh", "refine_uniform", "delaunay_mesh", "hilbert_index",
    "sfc_renumber", "FACET_NODES", "load_npz_mesh",
]

# local facet i = edge opposite local vertex i; its two nodes in CCW order
FACET_NODES = np.array([[1, 2], [2, 0], [0, 1]], dtype=np.int32)


@dataclass
class Mesh2D:
    coords: np.ndarray          # (nv, 2) float64
    cells: np.ndarray           # (nt, 3) int32, CCW
    topo: np.ndarray            # (nv,) int32
    nbr: np.ndarray = None      # (nt, 3) int32
    nbr_lf: np.ndarray = None   # (nt, 3) int8
    bf_cell: np.ndarray = None  # (nb,) int32
    bf_lf: np.ndarray = None    # (nb,) int8
    bf_marker: np.ndarray = None  # (nb,) int32
    periodic: bool = False
    # permutation history: cell_perm[new] = old cell id of the mesh as first built
    cell_perm: np.ndarray = None
    meta: dict = field(default_factory=dict)

    # ------------------------------------------------------------------ sizes
    @property
    def n_cells(self):
        return self.cells.shape[0]

    @property
    def n_vertices(self):
        return self.coords.shape[0]

    @property
    def n_topo_vertices(self):
        return int(self.topo.max()) + 1

    @property
    def n_bfacets(self):
        return self.bf_cell.shape[0]

    # --------------------------------------------------------------- geometry
    def cell_coords(self):
        """(nt, 3, 2) vertex coordinates of every cell."""
        return self.coords[self.cells]

    def cell_area(self):
        x = self.cell_coords()
        d1 = x[:, 1] - x[:, 0]
        d2 = x[:, 2] - x[:, 0]
        return 0.5 * (d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0])

    def cell_centroids(self):
        return self.cell_coords().mean(axis=1)

    def facet_scaled_normals(self):
        """(nt, 3, 2): outward normal of local facet i times its length."""
        x = self.cell_coords()
        p = x[:, FACET_NODES[:, 0]]
        q = x[:, FACET_NODES[:, 1]]
        e = q - p
        return np.stack([e[..., 1], -e[..., 0]], axis=-1)

    def unique_markers(self):
        return sorted(int(m) for m in np.unique(self.bf_marker))

    def boundary_length(self):
        """
        Length of every boundary segment, keyed by marker
        (restates `compute_boundary_length`, thetis/utility.py:821-832:
        assemble(1*ds(marker))).
        """
        x = self.cell_coords()
        p = x[self.bf_cell, FACET_NODES[self.bf_lf, 0]]
        q = x[self.bf_cell, FACET_NODES[self.bf_lf, 1]]
        ln = np.hypot(*(q - p).T)
        return {m: float(ln[self.bf_marker == m].sum()) for m in self.unique_markers()}

    def interior_facets(self):
        """
        Each interior facet once: arrays (cell_p, lf_p, cell_m, lf_m), the
        '+' side being the lower cell id (an arbitrary but fixed choice, like
        UFL's '+'/'-' restriction).
        """
        c, f = np.nonzero(self.nbr >= 0)
        n = self.nbr[c, f]
        keep = c < n
        # periodic meshes one cell wide could pair a cell with itself; excluded
        c, f, n = c[keep], f[keep], n[keep]
        return (c.astype(np.int32), f.astype(np.int8), n.astype(np.int32),
                self.nbr_lf[c, f].astype(np.int8))

    # ---------------------------------------------------------- connectivity
    def build_connectivity(self, edge_markers=None, default_marker=0, marker_fn=None):
        """
        Fill nbr/nbr_lf/bf_* by hashing facets on topological vertex ids.

        :kwarg edge_markers: dict {(tv_min, tv_max): marker} (gmsh tagged lines)
        :kwarg marker_fn: callable(midpoints (nb,2), p (nb,2), q (nb,2)) -> markers
        """
        nt = self.n_cells
        tv = self.topo[self.cells]                       # (nt, 3)
        a = tv[:, FACET_NODES[:, 0]].reshape(-1)         # (nt*3,)
        b = tv[:, FACET_NODES[:, 1]].reshape(-1)
        lo = np.minimum(a, b).astype(np.int64)
        hi = np.maximum(a, b).astype(np.int64)
        key = lo * (int(self.topo.max()) + 2) + hi
        order = np.argsort(key, kind="stable")
        ks = key[order]
        same_next = np.zeros(ks.shape[0], dtype=bool)
        same_next[:-1] = ks[1:] == ks[:-1]
        same_prev = np.zeros_like(same_next)
        same_prev[1:] = same_next[:-1]
        # sanity: no facet shared by three cells
        if np.any(same_next[:-1] & same_next[1:]):
            raise ValueError("non-manifold mesh: a facet is shared by more than two cells")
        first = order[same_next]            # first of each interior pair
        second = order[np.roll(same_next, 1) & same_prev]
        nbr = np.full(nt * 3, -1, dtype=np.int64)
        nlf = np.zeros(nt * 3, dtype=np.int8)
        nbr[first] = second // 3
        nlf[first] = second % 3
        nbr[second] = first // 3
        nlf[second] = first % 3
        bnd = order[~same_next & ~same_prev]
        bnd.sort()
        self.bf_cell = (bnd // 3).astype(np.int32)
        self.bf_lf = (bnd % 3).astype(np.int8)
        nbr[bnd] = -(1 + np.arange(bnd.shape[0], dtype=np.int64))
        self.nbr = nbr.reshape(nt, 3).astype(np.int32)
        self.nbr_lf = nlf.reshape(nt, 3)
        # markers
        nb = bnd.shape[0]
        x = self.cell_coords()
        p = x[self.bf_cell, FACET_NODES[self.bf_lf, 0]]
        q = x[self.bf_cell, FACET_NODES[self.bf_lf, 1]]
        if marker_fn is not None:
            self.bf_marker = np.asarray(marker_fn(0.5 * (p + q), p, q), dtype=np.int32)
        elif edge_markers is not None:
            mk = np.full(nb, default_marker, dtype=np.int32)
            kk = key[bnd]
            base = int(self.topo.max()) + 2
            ek = np.array([min(e) * base + max(e) for e in edge_markers.keys()], dtype=np.int64)
            ev = np.array(list(edge_markers.values()), dtype=np.int32)
            if ek.size:
                so = np.argsort(ek)
                ek, ev = ek[so], ev[so]
                pos = np.searchsorted(ek, kk)
                pos[pos >= ek.size] = ek.size - 1
                hit = ek[pos] == kk
                mk[hit] = ev[pos[hit]]
            self.bf_marker = mk
        else:
            self.bf_marker = np.full(nb, default_marker, dtype=np.int32)
        return self

    def make_ccw(self):
        """Flip cells with negative area (swap local vertices 1 and 2)."""
        a = self.cell_area()
        neg = a < 0
        if neg.any():
            c = self.cells.copy()
            c[neg, 1], c[neg, 2] = self.cells[neg, 2], self.cells[neg, 1]
            self.cells = c
        if np.any(self.cell_area() <= 0):
            raise ValueError("degenerate cell (zero area)")
        return self

    def vertex_to_cell_csr(self):
        """CSR (ptr, idx) of cells around each *topological* vertex."""
        tv = self.topo[self.cells].reshape(-1)
        cell = np.repeat(np.arange(self.n_cells, dtype=np.int32), 3)
        order = np.argsort(tv, kind="stable")
        counts = np.bincount(tv, minlength=self.n_topo_vertices)
        ptr = np.zeros(self.n_topo_vertices + 1, dtype=np.int64)
        np.cumsum(counts, out=ptr[1:])
        return ptr, cell[order]


# --------------------------------------------------------------------------
# structured meshes with Firedrake's RectangleMesh conventions
# --------------------------------------------------------------------------
def rectangle_mesh(nx, ny, lx, ly, diagonal="left", origin=(0.0, 0.0)):
    """
    Triangulated rectangle, following Firedrake's `RectangleMesh` conventions
    (recalled, Firedrake is not available here): vertex (i, j) at
    (i*lx/nx, j*ly/ny) with id i*(ny+1)+j; ``diagonal='left'`` (the default in
    Firedrake) splits each quad along (i, j+1)-(i+1, j), ``'right'`` along
    (i, j)-(i+1, j+1), ``'crossed'`` into four triangles.  Boundary markers:
    1: x=0, 2: x=lx, 3: y=0, 4: y=ly.
    """
    xs = np.linspace(0.0, lx, nx + 1) + origin[0]
    ys = np.linspace(0.0, ly, ny + 1) + origin[1]
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    coords = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    i = i.reshape(-1)
    j = j.reshape(-1)
    c0 = i * (ny + 1) + j
    c1 = i * (ny + 1) + j + 1
    c2 = (i + 1) * (ny + 1) + j + 1
    c3 = (i + 1) * (ny + 1) + j
    if diagonal == "left":
        t = np.stack([np.stack([c0, c1, c3], 1), np.stack([c1, c2, c3], 1)], 1).reshape(-1, 3)
    elif diagonal == "right":
        t = np.stack([np.stack([c0, c1, c2], 1), np.stack([c0, c2, c3], 1)], 1).reshape(-1, 3)
    elif diagonal == "crossed":
        nq = nx * ny
        cc = coords.shape[0] + np.arange(nq)
        ctr = 0.25 * (coords[c0] + coords[c1] + coords[c2] + coords[c3])
        coords = np.vstack([coords, ctr])
        t = np.stack([np.stack([c0, c1, cc], 1), np.stack([c1, c2, cc], 1),
                      np.stack([c2, c3, cc], 1), np.stack([c3, c0, cc], 1)], 1).reshape(-1, 3)
    else:
        raise ValueError(f"unknown diagonal {diagonal!r}")
    m = Mesh2D(coords=coords, cells=t.astype(np.int32),
               topo=np.arange(coords.shape[0], dtype=np.int32))
    m.make_ccw()
    x0, y0 = origin
    tol = 1e-9 * max(lx, ly)

    def marker_fn(mid, p, q):
        mk = np.zeros(mid.shape[0], dtype=np.int32)
        mk[np.abs(mid[:, 0] - x0) < tol] = 1
        mk[np.abs(mid[:, 0] - (x0 + lx)) < tol] = 2
        mk[np.abs(mid[:, 1] - y0) < tol] = 3
        mk[np.abs(mid[:, 1] - (y0 + ly)) < tol] = 4
        return mk

    m.build_connectivity(marker_fn=marker_fn)
    m.cell_perm = np.arange(m.n_cells, dtype=np.int64)
    m.meta.update(kind="rectangle", nx=nx, ny=ny, lx=lx, ly=ly, diagonal=diagonal)
    return m


def unit_square_mesh(nx, ny, diagonal="left"):
    """`UnitSquareMesh(nx, ny)` (test/slopelimiter/test_slopelimiter.py:17)."""
    return rectangle_mesh(nx, ny, 1.0, 1.0, diagonal=diagonal)


def periodic_rectangle_mesh(nx, ny, lx, ly, direction="x", diagonal="left", origin=(0.0, 0.0)):
    """
    `PeriodicRectangleMesh(nx, ny, lx, ly, direction='x')`
    (test/swe2d/test_rossby_wave.py:146).  Geometric vertices keep their own
    coordinates (like Firedrake's DG coordinate field); the column x=lx is
    identified topologically with x=0.  Remaining boundary markers follow
    Firedrake: 1: y=0, 2: y=ly for direction 'x'.
    """
    if direction != "x":
        raise NotImplementedError("only x-periodic meshes are needed on this path")
    if nx < 3:
        raise ValueError("periodic meshes need at least 3 cells in the periodic direction")
    m = rectangle_mesh(nx, ny, lx, ly, diagonal=diagonal, origin=origin)
    topo = np.arange(m.n_vertices, dtype=np.int32)
    last = nx * (ny + 1) + np.arange(ny + 1)
    topo[last] = np.arange(ny + 1)
    # compress ids
    _, topo = np.unique(topo, return_inverse=True)
    m.topo = topo.astype(np.int32)
    m.periodic = True
    y0 = origin[1]
    tol = 1e-9 * max(lx, ly)

    def marker_fn(mid, p, q):
        mk = np.zeros(mid.shape[0], dtype=np.int32)
        mk[np.abs(mid[:, 1] - y0) < tol] = 1
        mk[np.abs(mid[:, 1] - (y0 + ly)) < tol] = 2
        return mk

    m.build_connectivity(marker_fn=marker_fn)
    m.meta.update(kind="periodic_rectangle", direction=direction)
    return m


# --------------------------------------------------------------------------
# Gmsh 2.2 ASCII reader (demos/north_sea.msh)
# --------------------------------------------------------------------------
def read_gmsh(path):
    """
    Read a Gmsh 2.2 ASCII mesh: 3-node triangles (type 2) and tagged 2-node
    lines (type 1, first tag = physical id = boundary marker).  This is what
    Firedrake's `Mesh("north_sea.msh")` consumes (examples/north_sea,
    demos/demo_2d_north_sea.py: markers 100 = open ocean, 200 = coast).
    """
    with open(path, "r") as fh:
        lines = fh.read().split("\n")
    nodes = None
    tris, edges, etags = [], [], []
    i = 0
    while i < len(lines):
        ln = lines[i].strip()
        if ln == "$MeshFormat":
            ver = lines[i + 1].split()[0]
            if not ver.startswith("2"):
                raise ValueError(f"unsupported gmsh format version {ver}")
            i += 2
        elif ln == "$Nodes":
            n = int(lines[i + 1])
            arr = np.array([l.split() for l in lines[i + 2:i + 2 + n]], dtype=np.float64)
            ids = arr[:, 0].astype(np.int64)
            nodes = (ids, arr[:, 1:3].copy())
            i += 2 + n
        elif ln == "$Elements":
            n = int(lines[i + 1])
            for l in lines[i + 2:i + 2 + n]:
                w = l.split()
                et = int(w[1])
                ntag = int(w[2])
                v = w[3 + ntag:]
                if et == 2:
                    tris.append((int(v[0]), int(v[1]), int(v[2])))
                elif et == 1:
                    edges.append((int(v[0]), int(v[1])))
                    etags.append(int(w[3]) if ntag > 0 else 0)
            i += 2 + n
        else:
            i += 1
    ids, xy = nodes
    remap = np.full(int(ids.max()) + 1, -1, dtype=np.int64)
    remap[ids] = np.arange(ids.shape[0])
    cells = remap[np.array(tris, dtype=np.int64)]
    used = np.zeros(ids.shape[0], dtype=bool)
    used[cells.reshape(-1)] = True
    comp = np.cumsum(used) - 1
    coords = xy[used]
    cells = comp[cells].astype(np.int32)
    m = Mesh2D(coords=coords, cells=cells, topo=np.arange(coords.shape[0], dtype=np.int32))
    m.make_ccw()
    em = {}
    for (a, b), t in zip(edges, etags):
        a2, b2 = int(comp[remap[a]]), int(comp[remap[b]])
        em[(min(a2, b2), max(a2, b2))] = t
    m.build_connectivity(edge_markers=em)
    m.cell_perm = np.arange(m.n_cells, dtype=np.int64)
    m.meta.update(kind="gmsh", path=str(path))
    return m


# --------------------------------------------------------------------------
# uniform k-section refinement (4 M-triangle North Sea: k = 19)
# --------------------------------------------------------------------------
def refine_uniform(mesh, k):
    """
    Split every triangle into k*k congruent children (k-section).  Boundary
    markers are inherited from the parent facet.  Lattice points on shared
    parent edges are computed so that both parents produce bit-identical
    coordinates (fp addition is commutative), then deduplicated exactly.
    """
    if k == 1:
        return mesh
    if mesh.periodic:
        raise NotImplementedError("refinement of periodic meshes is not needed on this path")
    nt = mesh.n_cells
    x = mesh.cell_coords()                   # (nt, 3, 2)
    # lattice (i, j) with l = k-i-j; point = (l*v0 + i*v1 + j*v2)/k
    ij = [(i, j) for i in range(k + 1) for j in range(k + 1 - i)]
    lat = np.array(ij, dtype=np.int64)
    npt = lat.shape[0]
    idx_of = -np.ones((k + 1, k + 1), dtype=np.int64)
    idx_of[lat[:, 0], lat[:, 1]] = np.arange(npt)
    wi = lat[:, 0].astype(np.float64)
    wj = lat[:, 1].astype(np.float64)
    wl = (k - lat[:, 0] - lat[:, 1]).astype(np.float64)
    # evaluate as sum of products with exact zero handling: order-independent on edges
    t0 = wl[None, :, None] * x[:, None, 0, :]
    t1 = wi[None, :, None] * x[:, None, 1, :]
    t2 = wj[None, :, None] * x[:, None, 2, :]
    # on an edge one of the terms is exactly 0 and a+b is commutative; at parent
    # vertices return the vertex coordinate itself
    pts = (t0 + t1 + t2) / float(k)          # (nt, npt, 2)
    for vloc, (ii, jj) in enumerate([(0, 0), (k, 0), (0, k)]):
        pts[:, idx_of[ii, jj]] = x[:, vloc]
    # edge points must be evaluated with the two non-zero terms only and in a
    # canonical (sorted global vertex id) order to be bit-identical
    gv = mesh.cells.astype(np.int64)
    for (va, vb, sel) in [
        (0, 1, [(i, 0) for i in range(1, k)]),      # j = 0: l*v0 + i*v1
        (0, 2, [(0, j) for j in range(1, k)]),      # i = 0: l*v0 + j*v2
        (1, 2, [(i, k - i) for i in range(1, k)]),  # l = 0: i*v1 + j*v2
    ]:
        for (ii, jj) in sel:
            w = {0: k - ii - jj, 1: ii, 2: jj}
            pa = float(w[va]) * x[:, va]
            pb = float(w[vb]) * x[:, vb]
            pts[:, idx_of[ii, jj]] = (pa + pb) / float(k)   # commutative => same from both parents
    # children
    up = [(idx_of[i, j], idx_of[i + 1, j], idx_of[i, j + 1])
          for i in range(k) for j in range(k - i)]
    dn = [(idx_of[i + 1, j], idx_of[i + 1, j + 1], idx_of[i, j + 1])
          for i in range(k) for j in range(k - i - 1)]
    child = np.array(up + dn, dtype=np.int64)              # (k*k, 3) local lattice ids
    glob = (np.arange(nt, dtype=np.int64)[:, None] * npt)  # (nt, 1)
    cells = (glob[:, :, None] + child[None, :, :]).reshape(-1, 3)
    allpts = pts.reshape(-1, 2)
    # exact dedup on the bit pattern
    view = np.ascontiguousarray(allpts).view(np.dtype((np.void, 16))).reshape(-1)
    _, first_idx, inv = np.unique(view, return_index=True, return_inverse=True)
    coords = allpts[first_idx]
    cells = inv.reshape(-1)[cells].astype(np.int32)
    m = Mesh2D(coords=coords, cells=cells, topo=np.arange(coords.shape[0], dtype=np.int32))
    m.make_ccw()
    # boundary markers: child boundary facet inherits marker of the parent boundary
    # facet it lies on.  Identify by parent cell + collinearity with parent facet.
    parent = np.repeat(np.arange(nt, dtype=np.int64), k * k)
    m.build_connectivity()
    pc = parent[m.bf_cell]
    xc = m.cell_coords()
    p = xc[m.bf_cell, FACET_NODES[m.bf_lf, 0]]
    q = xc[m.bf_cell, FACET_NODES[m.bf_lf, 1]]
    mid = 0.5 * (p + q)
    mk = np.zeros(m.n_bfacets, dtype=np.int32)
    found = np.zeros(m.n_bfacets, dtype=bool)
    px = mesh.cell_coords()[pc]                            # (nb, 3, 2)
    pn = mesh.nbr[pc]                                      # (nb, 3)
    best = np.full(m.n_bfacets, np.inf)
    for f in range(3):
        a = px[:, FACET_NODES[f, 0]]
        b = px[:, FACET_NODES[f, 1]]
        e = b - a
        d = mid - a
        cross = np.abs(e[:, 0] * d[:, 1] - e[:, 1] * d[:, 0]) / np.maximum(np.hypot(e[:, 0], e[:, 1]), 1e-300)
        isb = pn[:, f] < 0
        better = isb & (cross < best)
        best[better] = cross[better]
        kidx = -(pn[:, f] + 1)
        mk[better] = mesh.bf_marker[np.clip(kidx, 0, mesh.n_bfacets - 1)][better]
        found |= better
    if not found.all():
        raise RuntimeError("refinement produced a boundary facet without a parent boundary facet")
    m.bf_marker = mk
    m.cell_perm = np.arange(m.n_cells, dtype=np.int64)
    m.meta.update(kind="refined", k=k, parent=dict(mesh.meta))
    return m


def delaunay_mesh(n_points, lx=1.0, ly=1.0, seed=0, jitter=0.35):
    """
    'Unstructured' triangulation for the stommel2d-style configuration
    (BASELINE.json config 3): Delaunay triangulation of a jittered lattice on
    [0,lx]x[0,ly] whose boundary points stay on the boundary.  Markers as
    `rectangle_mesh`.
    """
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    nx = max(2, int(round(np.sqrt(n_points * lx / ly))))
    ny = max(2, int(round(n_points / nx)))
    xs = np.linspace(0, lx, nx + 1)
    ys = np.linspace(0, ly, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    dx, dy = lx / nx, ly / ny
    JX = rng.uniform(-jitter, jitter, X.shape) * dx
    JY = rng.uniform(-jitter, jitter, Y.shape) * dy
    JX[0, :] = JX[-1, :] = 0.0
    JY[:, 0] = JY[:, -1] = 0.0
    pts = np.stack([(X + JX).reshape(-1), (Y + JY).reshape(-1)], 1)
    tri = Delaunay(pts)
    cells = tri.simplices.astype(np.int32)
    m = Mesh2D(coords=pts, cells=cells, topo=np.arange(pts.shape[0], dtype=np.int32))
    m.make_ccw()
    # drop slivers on the hull (area ~ 0) that Qhull may emit for collinear boundary points
    a = m.cell_area()
    keep = a > 1e-12 * lx * ly
    m.cells = m.cells[keep]
    tol = 1e-9 * max(lx, ly)

    def marker_fn(mid, p, q):
        mk = np.zeros(mid.shape[0], dtype=np.int32)
        mk[np.abs(mid[:, 0]) < tol] = 1
        mk[np.abs(mid[:, 0] - lx) < tol] = 2
        mk[np.abs(mid[:, 1]) < tol] = 3
        mk[np.abs(mid[:, 1] - ly) < tol] = 4
        return mk

    m.build_connectivity(marker_fn=marker_fn)
    m.cell_perm = np.arange(m.n_cells, dtype=np.int64)
    m.meta.update(kind="delaunay", n_points=int(pts.shape[0]), seed=seed)
    return m


# --------------------------------------------------------------------------
# space-filling-curve renumbering
# --------------------------------------------------------------------------
def hilbert_index(xy, bits=16):
    """Hilbert curve index of points scaled into a 2^bits x 2^bits grid."""
    xy = np.asarray(xy, dtype=np.float64)
    lo = xy.min(axis=0)
    span = np.maximum(xy.max(axis=0) - lo, 1e-300)
    n = 1 << bits
    x = np.minimum(((xy[:, 0] - lo[0]) / span.max() * (n - 1)).astype(np.int64), n - 1)
    y = np.minimum(((xy[:, 1] - lo[1]) / span.max() * (n - 1)).astype(np.int64), n - 1)
    d = np.zeros(x.shape[0], dtype=np.int64)
    s = n >> 1
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        # rotate
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        x &= (s - 1)
        y &= (s - 1)
        s >>= 1
    return d


def sfc_renumber(mesh, perm=None):
    """
    Renumber cells along a Hilbert curve through the centroids (or by a given
    permutation ``perm[new] = old``), and vertices by first use, so that
    facet neighbours and vertex data of consecutive cells are close in memory.
    Returns a new mesh; ``cell_perm`` composes with previous renumberings.
    """
    if perm is None:
        perm = np.argsort(hilbert_index(mesh.cell_centroids()), kind="stable")
    perm = np.asarray(perm, dtype=np.int64)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.shape[0])
    cells = mesh.cells[perm]
    # vertices by first use
    flat = cells.reshape(-1)
    _, first = np.unique(flat, return_index=True)
    vorder = flat[np.sort(first)]
    vinv = np.full(mesh.n_vertices, -1, dtype=np.int64)
    vinv[vorder] = np.arange(vorder.shape[0])
    new = Mesh2D(coords=mesh.coords[vorder], cells=vinv[cells].astype(np.int32),
                 topo=None, periodic=mesh.periodic)
    # topological ids renumbered by first occurrence: identity for non-periodic meshes
    told = mesh.topo[vorder]
    _, first_t, t = np.unique(told, return_index=True, return_inverse=True)
    rank = np.empty(first_t.shape[0], dtype=np.int64)
    rank[np.argsort(first_t, kind="stable")] = np.arange(first_t.shape[0])
    new.topo = rank[t].astype(np.int32)
    nbr = mesh.nbr[perm].astype(np.int64)
    pos = nbr >= 0
    nbr[pos] = inv[nbr[pos]]
    new.nbr = nbr.astype(np.int32)
    new.nbr_lf = mesh.nbr_lf[perm]
    new.bf_cell = inv[mesh.bf_cell].astype(np.int32)
    new.bf_lf = mesh.bf_lf.copy()
    new.bf_marker = mesh.bf_marker.copy()
    base = mesh.cell_perm if mesh.cell_perm is not None else np.arange(mesh.n_cells, dtype=np.int64)
    new.cell_perm = base[perm]
    new.meta = dict(mesh.meta)
    new.meta["sfc"] = True
    new.meta["vertex_perm"] = vorder
    return new


def load_npz_mesh(path):
    """Mesh stored as arrays (coords, cells, bnd_edges, bnd_tags); see tests/golden/make_north_sea_fixture.py."""
    d = np.load(path)
    m = Mesh2D(coords=d["coords"].astype(np.float64), cells=d["cells"].astype(np.int32),
               topo=np.arange(d["coords"].shape[0], dtype=np.int32))
    m.make_ccw()
    em = {(int(min(a, b)), int(max(a, b))): int(t) for (a, b), t in zip(d["bnd_edges"], d["bnd_tags"])}
    m.build_connectivity(edge_markers=em)
    m.cell_perm = np.arange(m.n_cells, dtype=np.int64)
    m.meta.update(kind="npz", path=str(path))
    return m
