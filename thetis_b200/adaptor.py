"""
Adaptor between Firedrake-shaped objects (real Firedrake or thetis_b200.shim)
and the arrays the CUDA library consumes.  Read once at set-up; never on the
hot path.

Everything here relies only on attribute names that the reference itself uses
on Firedrake objects (SURVEY.md 8b lists the evidence for each one):
`Function.dat.data(_ro)`, `Function.function_space()`,
`FunctionSpace.cell_node_map().values`, `FunctionSpace.ufl_element().family()/
.degree()`, `mesh.coordinates`, `mesh.exterior_facets.unique_markers`,
`Constant.values()`, `float(Constant)`.

The Firedrake branch of `extract_mesh` cannot be exercised in this environment
(Firedrake is not installed); the shim branch is what the tests
run.  Both produce the same `Mesh2D`.
"""
from __future__ import annotations

import weakref

import numpy as np

from .mesh import Mesh2D, FACET_NODES, sfc_renumber

__all__ = ["MeshAdaptor", "is_constant", "is_function", "is_expression", "constant_value", "get_adaptor",
           "expression_leaves", "expression_degree"]


def _space_of(x):
    """`x.function_space()` or None.  A real firedrake.Constant has the method too -- it returns None."""
    fs = getattr(x, "function_space", None)
    if fs is None or not callable(fs):
        return None
    try:
        return fs()
    except Exception:
        return None


def is_function(x):
    """A Function: lives in a function space and carries data (`dat`, or `subfunctions` for a mixed one)."""
    return _space_of(x) is not None and (hasattr(x, "dat") or hasattr(x, "subfunctions"))


def is_constant(x):
    """Plain numbers, or a Constant: has `values()` and no function space (classified by capability, because a real
    firedrake.Constant also has `.dat` and `.function_space()`)."""
    if isinstance(x, (int, float, np.integer, np.floating)):
        return True
    if isinstance(x, (tuple, list)) and all(isinstance(v, (int, float, np.integer, np.floating)) for v in x):
        return True
    if isinstance(x, np.ndarray) and x.ndim <= 1 and x.size <= 3:
        return True
    return hasattr(x, "values") and callable(x.values) and _space_of(x) is None


def is_expression(x):
    """A UFL expression tree that is neither a Function nor a Constant (e.g. `elev_ramp * elev_tide_2d`)."""
    return hasattr(x, "ufl_operands") and not is_function(x) and not is_constant(x)


def expression_leaves(expr):
    """The Functions and Constants an expression reads (its change stamp is the tuple of theirs)."""
    out = []

    def walk(e):
        if is_function(e) or (is_constant(e) and hasattr(e, "values")):
            if not any(e is o for o in out):
                out.append(e)
            return
        for o in getattr(e, "ufl_operands", ()):
            walk(o)
    walk(expr)
    return out


def _condition_reads_a_function(expr):
    """True when the condition of some `conditional(...)` inside a (real-UFL) expression depends on a Function."""
    def has_function(e):
        return is_function(e) or any(has_function(o) for o in getattr(e, "ufl_operands", ()))

    def walk(e):
        ops = getattr(e, "ufl_operands", ())
        if type(e).__name__ == "Conditional" and ops and has_function(ops[0]):
            return True
        return any(walk(o) for o in ops)
    return walk(expr)


def expression_degree(expr):
    """Polynomial degree of an expression in its Function operands (P1 Functions count 1, Constants 0); None when it
    is not a polynomial (division by a Function, a condition that depends on a Function).  Degree <= 1 means nodal
    evaluation reproduces the expression exactly, which is what the accelerated path requires."""
    if is_constant(expr):
        return 0
    if is_function(expr):
        return 1
    kind = getattr(expr, "kind", None)
    if kind is None:           # real UFL
        from ufl.algorithms import estimate_total_polynomial_degree
        # UFL's estimate ignores the condition of a `conditional`: one that depends on a Function switches inside
        # cells and is not reproduced by nodal evaluation, whatever degree its two values have
        if _condition_reads_a_function(expr):
            return None
        return estimate_total_polynomial_degree(expr)
    ops = [expression_degree(o) for o in expr.ufl_operands]
    if any(d is None for d in ops):
        return None
    if kind == "mul":
        return sum(ops)
    if kind in ("add", "sub", "vector"):
        return max(ops)
    if kind == "div":
        return ops[0] if ops[1] == 0 else None
    if kind in ("lt", "gt", "le", "ge"):
        return 0 if max(ops) == 0 else None
    if kind == "conditional":
        return max(ops[1:]) if ops[0] == 0 else None
    return None


def constant_value(x):
    """Current value of a Constant-like object as a 1-D float array (read live, like the UFL forms do)."""
    if hasattr(x, "values") and callable(x.values):
        return np.atleast_1d(np.asarray(x.values(), dtype=np.float64))
    return np.atleast_1d(np.asarray(x, dtype=np.float64))


def _family(fs):
    el = fs.ufl_element()
    fam = el.family()
    # firedrake VectorElement wraps a scalar element
    sub = getattr(el, "sub_elements", None)
    if fam in ("Discontinuous Lagrange", "DG", "DP", "DQ"):
        return "DG"
    if fam in ("Lagrange", "CG", "P", "Q"):
        return "CG"
    if sub:
        return _family_of_name(sub[0].family())
    return fam


def _family_of_name(fam):
    if fam in ("Discontinuous Lagrange", "DG", "DP"):
        return "DG"
    if fam in ("Lagrange", "CG", "P"):
        return "CG"
    return fam


def _comm_size(mesh):
    """Number of ranks a Firedrake mesh is distributed over (1 for anything without an MPI communicator)."""
    comm = getattr(mesh, "comm", None)
    try:
        return int(comm.size) if comm is not None else 1
    except Exception:                                          # noqa: BLE001
        return 1


def _map_rows(pmap, n, with_overlap):
    """First `n` rows of a PyOP2 map; `values` stops at the owned entities, `values_with_halo` carries the overlap."""
    vals = getattr(pmap, "values_with_halo", None) if with_overlap else None
    if vals is None:
        vals = pmap.values
    return np.asarray(vals[:n], dtype=np.int64)


def _overlap_kind(mesh):
    """'vertex' / 'facet': what the overlap of a distributed Firedrake mesh holds (firedrake/mesh.py
    distribution_parameters['overlap_type'] = (DistributedMeshOverlapType.X, depth); Firedrake's default is FACET)."""
    dp = getattr(mesh, "_distribution_parameters", None) or {}
    ot = dp.get("overlap_type")
    if ot is None:
        return "facet"
    kind, depth = ot[0], (ot[1] if len(ot) > 1 else 1)
    name = str(getattr(kind, "name", kind)).lower()
    if "none" in name or depth < 1:
        raise NotImplementedError("a distributed mesh needs an overlap of at least one cell "
                                  "(distribution_parameters={'overlap_type': (DistributedMeshOverlapType.VERTEX, 1)})")
    return "vertex" if "vertex" in name else "facet"


def _global_cell_ids(mesh, n_local):
    """A global number per local cell (overlap included) of a distributed Firedrake mesh: the global DG0 dof numbers
    (PETSc local-to-global map of the DG0 space: one dof per cell)."""
    import firedrake as fd
    dg0 = fd.FunctionSpace(mesh, "DG", 0)
    cell_dof = _map_rows(dg0.cell_node_map(), n_local, True).reshape(-1)
    lg = np.asarray(dg0.dof_dset.lgmap.indices, dtype=np.int64)
    return lg[cell_dof]


def _mesh2d_from_firedrake(mesh, with_overlap=False):
    """
    Build a `Mesh2D` from a real Firedrake mesh.  Written against the attribute names listed in SURVEY.md 8b;
    exercised against a Firedrake-shaped look-alike (tests/test_firedrake_lookalike_mesh.py), never against
    Firedrake itself (not installable here).  ``with_overlap``: the mesh is distributed over MPI ranks -- the
    overlap cells are taken along (after the owned ones) and one-sided facets that are not in
    `mesh.exterior_facets` become facets with an unknown neighbour.
    """
    import firedrake as fd  # noqa: F401  (ImportError if absent, by design)
    coords_f = mesh.coordinates
    cfs = coords_f.function_space()
    ncell = mesh.cell_set.total_size if with_overlap else mesh.cell_set.size
    cmap = _map_rows(cfs.cell_node_map(), ncell, with_overlap)
    xy = np.asarray(coords_f.dat.data_ro_with_halos, dtype=np.float64)
    used = np.unique(cmap)
    remap = np.full(xy.shape[0], -1, dtype=np.int64)
    remap[used] = np.arange(used.shape[0])
    coords = xy[used][:, :2]
    cells = remap[cmap].astype(np.int32)
    p1 = fd.FunctionSpace(mesh, "CG", 1)
    tmap = _map_rows(p1.cell_node_map(), ncell, with_overlap)
    topo = np.zeros(coords.shape[0], dtype=np.int64)
    topo[cells.reshape(-1)] = tmap.reshape(-1)
    _, topo = np.unique(topo, return_inverse=True)
    m = Mesh2D(coords=coords, cells=cells, topo=topo.astype(np.int32),
               periodic=_family(cfs) == "DG")
    a = m.cell_area()
    swap = a < 0
    m.make_ccw()
    # exterior facet markers by topological edge
    ef = mesh.exterior_facets
    fcm = getattr(ef, "facet_cell_map", None) if with_overlap else None
    fcell = (_map_rows(fcm, None, True) if fcm is not None else np.asarray(ef.facet_cell)).reshape(-1)
    lfd = ef.local_facet_dat
    flocal = np.asarray(lfd.data_ro_with_halos if with_overlap and hasattr(lfd, "data_ro_with_halos")
                        else lfd.data_ro).reshape(-1)
    markers = np.asarray(ef.markers).reshape(-1)
    em = {}
    for c, lf, mk in zip(fcell, flocal, markers):
        if c >= ncell:
            continue
        ta, tb = int(topo[remap[cmap[c, FACET_NODES[lf, 0]]]]), int(topo[remap[cmap[c, FACET_NODES[lf, 1]]]])
        em[(min(ta, tb), max(ta, tb))] = int(mk)
    if with_overlap:
        from .parallel import build_overlap_connectivity
        build_overlap_connectivity(m, em)
    else:
        m.build_connectivity(edge_markers=em)
    m.cell_perm = np.arange(m.n_cells, dtype=np.int64)
    return m, swap


class MeshAdaptor:
    """
    One per mesh object: the SFC-renumbered `Mesh2D` the device uses plus the
    maps back to the caller's numbering.
    """

    def __init__(self, mesh_obj, renumber=True):
        tm = getattr(mesh_obj, "topology_mesh", None)
        dist_fd = False
        if isinstance(mesh_obj, Mesh2D):
            base, swap = mesh_obj, np.zeros(mesh_obj.n_cells, dtype=bool)
        elif isinstance(tm, Mesh2D):
            base, swap = tm, np.zeros(tm.n_cells, dtype=bool)
        else:
            dist_fd = _comm_size(mesh_obj) > 1
            base, swap = _mesh2d_from_firedrake(mesh_obj, with_overlap=dist_fd)
        self.mesh_obj_ref = weakref.ref(mesh_obj) if not isinstance(mesh_obj, Mesh2D) else (lambda: mesh_obj)
        self.base = base
        self.swap = swap
        # Functions on a mesh Firedrake distributed itself: overlap rows live behind the `_with_halos` accessors
        self.with_halos = dist_fd
        plan = None
        if dist_fd:
            # a mesh Firedrake has already distributed (mpiexec -n N): the halo plan comes from the local cells, the
            # owned / overlap split and the global DG0 numbering; Firedrake's communicator does the one all-gather.
            # torch.distributed (the transport of the stage exchanges) is brought up with the same rank numbering.
            from .parallel import plan_from_local_mesh, init_torch_distributed_from_comm
            comm = mesh_obj.comm
            init_torch_distributed_from_comm(comm)
            plan, part = plan_from_local_mesh(base, int(mesh_obj.cell_set.size), _global_cell_ids(mesh_obj, base.n_cells),
                                              rank=int(comm.rank), allgather=comm.allgather,
                                              halo=_overlap_kind(mesh_obj), renumber=renumber)
            self.mesh = part.mesh
            self.perm = np.asarray(part.mesh.cell_perm, dtype=np.int64)
        elif renumber and not base.meta.get("sfc"):
            self.mesh = sfc_renumber(base)
            # cell_perm composes with earlier renumberings of `base`; we need new -> base order
            base_perm = base.cell_perm if base.cell_perm is not None else np.arange(base.n_cells)
            inv = np.empty(base.n_cells, dtype=np.int64)
            inv[base_perm] = np.arange(base.n_cells)
            self.perm = inv[self.mesh.cell_perm]
        else:
            self.mesh = base
            self.perm = np.arange(base.n_cells, dtype=np.int64)
        self.engine = None
        # distributed run: the mesh object carries this rank's HaloPlan (thetis_b200.parallel.distribute_mesh)
        self.halo = plan if plan is not None else getattr(mesh_obj, "halo_plan", None)
        self.n_owned = self.halo.part.n_owned if self.halo is not None else self.mesh.n_cells
        bl = self.mesh.meta["global_boundary_len"] if plan is not None else getattr(mesh_obj, "boundary_len", None)
        self.boundary_len = dict(bl) if bl is not None else self.mesh.boundary_length()

    def get_engine(self):
        """The device context of this mesh (created on first use; shared by every integrator / limiter on it)."""
        if self.engine is None:
            from .engine import Engine
            self.engine = Engine(self.mesh, n_owned=self.n_owned)
            if self.halo is not None:
                self.halo.attach(self.engine)
                for mk, ln in self.boundary_len.items():
                    self.engine.set_boundary_length(mk, ln)
        return self.engine

    # ------------------------------------------------------------ node maps
    def dat_ro(self, func):
        """The Function's values as the node maps of this adaptor index them (overlap rows included)."""
        d = func.dat
        return np.asarray(d.data_ro_with_halos if self.with_halos else d.data_ro)

    def dat_rw(self, func):
        """Writable view of the same rows."""
        d = func.dat
        return d.data_with_halos if self.with_halos else d.data

    def _cell_nodes(self, fs):
        cm = _map_rows(fs.cell_node_map(), self.base.n_cells, self.with_halos)
        if cm.shape[1] != 3:
            raise NotImplementedError("only P1 / P1DG spaces are supported on the accelerated path")
        if self.swap.any():
            cm = cm.copy()
            cm[self.swap, 1], cm[self.swap, 2] = cm[self.swap, 2].copy(), cm[self.swap, 1].copy()
        return cm[self.perm]

    def dg_node_map(self, fs):
        """(nt, 3) int32: dof of device cell c / CCW local node a in a P1DG space."""
        if _family(fs) != "DG" or fs.ufl_element().degree() != 1:
            raise NotImplementedError("solution must live in a P1DG space (element_family 'dg-dg', degree 1)")
        return np.ascontiguousarray(self._cell_nodes(fs), dtype=np.int32)

    def nodal_values(self, func):
        """(nt, 3[,k]) values of a P1/P1DG Function at the device cells' nodes."""
        fs = func.function_space()
        return self.dat_ro(func)[self._cell_nodes(fs)]

    def evaluate(self, expr):
        """
        (nt, 3[,k]) nodal values, at the device cells' nodes, of a Function or of an expression over Functions and
        Constants that is affine in its Function operands (e.g. `elev_ramp * elev_tide_2d`,
        examples/north_sea/model_config.py:188-192): evaluating it at the nodes and interpolating linearly is then
        identical to evaluating the UFL expression at the quadrature points.  Anything of higher degree must be
        interpolated into a P1 Function by the caller (a modelling decision the library does not take silently).
        """
        if is_function(expr):
            return self.nodal_values(expr)
        if is_constant(expr):
            v = constant_value(expr)
            return np.broadcast_to(v if v.size > 1 else v[0], (self.mesh.n_cells, 3) + ((v.size,) if v.size > 1 else ()))
        deg = expression_degree(expr)
        if deg is None or deg > 1:
            raise NotImplementedError(
                "expression is not affine in its Function operands: interpolate it into a P1 / P1DG Function first")
        kind = getattr(expr, "kind", None)
        if kind is None:
            # real UFL: let Firedrake do the nodal evaluation (exact for degree <= 1).  Exercised on the numpy stand-in
            # for UFL only (tests/test_dropin_with_reference_objects.py), never against Firedrake itself.
            import firedrake as fd
            mesh_obj = self.mesh_obj_ref()
            shape = getattr(expr, "ufl_shape", ())
            fs = fd.VectorFunctionSpace(mesh_obj, "DG", 1) if shape else fd.FunctionSpace(mesh_obj, "DG", 1)
            return self.nodal_values(fd.Function(fs).interpolate(expr))
        return self._eval_tree(expr, self.evaluate)

    @staticmethod
    def _eval_tree(expr, leaf):
        """Apply the operators of a shim expression tree to arrays produced by ``leaf`` for Functions / Constants."""
        kind = getattr(expr, "kind", None)
        if kind is None:
            return leaf(expr)
        ops = [MeshAdaptor._eval_tree(o, leaf) if hasattr(o, "ufl_operands") else np.asarray(o, dtype=np.float64)
               for o in expr.ufl_operands]
        if kind == "mul":
            a, b = ops
            if a.ndim and b.ndim and a.ndim < b.ndim:
                a = a[..., None]
            elif a.ndim and b.ndim and b.ndim < a.ndim:
                b = b[..., None]
            return a * b
        if kind == "add":
            return ops[0] + ops[1]
        if kind == "sub":
            return ops[0] - ops[1]
        if kind == "div":
            return ops[0] / ops[1]
        if kind == "lt":
            return ops[0] < ops[1]
        if kind == "gt":
            return ops[0] > ops[1]
        if kind == "le":
            return ops[0] <= ops[1]
        if kind == "ge":
            return ops[0] >= ops[1]
        if kind == "conditional":
            return np.where(ops[0], ops[1], ops[2])
        if kind == "vector":
            return np.stack(np.broadcast_arrays(*ops), axis=-1)
        raise NotImplementedError(f"expression node {kind!r}")

    def coefficient_values(self, func):
        """
        ('vertex', (nv[,k])) for a continuous field (P1 CG, or a P1DG Function whose values agree at shared
        vertices), else ('cell', (nt, 3[,k])) for a genuinely discontinuous P1DG one (e.g. Coriolis / sources
        projected into H_2d, test/swe2d/test_steady_state_basin_mms.py:169-177).
        """
        nodal = np.asarray(self.evaluate(func), dtype=np.float64)
        vert = np.zeros((self.mesh.n_vertices,) + nodal.shape[2:])
        vert[self.mesh.cells] = nodal
        err = np.abs(vert[self.mesh.cells] - nodal).max() if nodal.size else 0.0
        scale = max(np.abs(nodal).max() if nodal.size else 0.0, 1e-300)
        if err > 1e-10 * scale:
            return "cell", np.ascontiguousarray(nodal)
        return "vertex", vert

    def vertex_values(self, func):
        """
        Values of a continuous P1 field (CG, or continuous DG) at the device mesh's geometric vertices; raises for
        discontinuous data (coefficients of facet terms -- bathymetry, viscosity, diffusivity, wetting-drying alpha --
        must be continuous on the accelerated path).
        """
        kind, vals = self.coefficient_values(func)
        if kind != "vertex":
            raise NotImplementedError("this coefficient enters facet terms and must be continuous: discontinuous "
                                      "fields are not supported on the accelerated path")
        return vals

    def bfacet_values(self, func, marker=None):
        """
        (nb, 2[,k]) values of a P1/P1DG Function at the two nodes of every exterior facet.  With ``marker`` only
        the rows of that marker are evaluated (the rest of the returned, reused, buffer is untouched).
        """
        if not is_function(func):
            # expression over Functions / Constants, affine in the Functions: evaluated at the facet nodes only (its
            # Function leaves through the cached index path below), every stage of a tidal run
            deg = expression_degree(func)
            if deg is None or deg > 1:
                raise NotImplementedError(
                    "expression is not affine in its Function operands: interpolate it into a P1 / P1DG Function first")
            if getattr(func, "kind", None) is None:       # real UFL: nodal interpolation by Firedrake
                nodal = np.asarray(self.evaluate(func), dtype=np.float64)
                m = self.mesh
                return np.stack([nodal[m.bf_cell, FACET_NODES[m.bf_lf, 0]], nodal[m.bf_cell, FACET_NODES[m.bf_lf, 1]]], axis=1)

            def leaf(x):
                if is_function(x):
                    return self.bfacet_values(x, marker)
                v = constant_value(x)
                return v if v.size > 1 else v[0]
            out = np.asarray(self._eval_tree(func, leaf), dtype=np.float64)
            nb = self.mesh.n_bfacets
            if out.ndim < 2 or out.shape[0] != nb:        # expression of Constants only: broadcast over the facets
                out = np.broadcast_to(out, (nb, 2) + ((out.shape[-1],) if out.ndim == 1 else ())).copy()
            return out
        fs = func.function_space()
        cache = self.__dict__.setdefault("_bf_nodes", {})
        idx = cache.get(id(fs))
        if idx is None:
            cn = self._cell_nodes(fs)
            m = self.mesh
            idx = np.stack([cn[m.bf_cell, FACET_NODES[m.bf_lf, 0]], cn[m.bf_cell, FACET_NODES[m.bf_lf, 1]]], axis=1)
            cache[id(fs)] = idx
            self.__dict__.setdefault("_bf_keep", []).append(fs)     # keep the key object alive
        data = self.dat_ro(func)
        if marker is None:
            return data[idx]
        key = (id(fs), int(marker))
        ent = cache.get(key)
        if ent is None:
            rows = np.nonzero(self.mesh.bf_marker == marker)[0]
            ent = (rows, idx[rows], np.zeros(idx.shape + data.shape[1:]))
            cache[key] = ent
        rows, ridx, out = ent
        out[rows] = data[ridx]
        return out


_ADAPTORS = weakref.WeakKeyDictionary()
_ADAPTORS_STRONG = {}


def get_adaptor(mesh_obj, renumber=True):
    try:
        ad = _ADAPTORS.get(mesh_obj)
        if ad is None:
            ad = MeshAdaptor(mesh_obj, renumber=renumber)
            _ADAPTORS[mesh_obj] = ad
        return ad
    except TypeError:      # unhashable / not weak-referenceable
        key = id(mesh_obj)
        if key not in _ADAPTORS_STRONG:
            _ADAPTORS_STRONG[key] = MeshAdaptor(mesh_obj, renumber=renumber)
        return _ADAPTORS_STRONG[key]
