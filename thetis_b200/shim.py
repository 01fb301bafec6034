"""
Firedrake-shaped data containers for running the stepper without Firedrake.

Firedrake is not installable in this environment, so the
stand-alone harness and the tests use these minimal look-alikes.  They expose
exactly the attribute names the adaptor reads from real Firedrake objects
(`Function.dat.data`, `function_space().cell_node_map().values`,
`ufl_element().family()/degree()`, `mesh.coordinates`,
`mesh.exterior_facets.unique_markers`, `Constant.values()/assign`,
`Function.subfunctions`), so the SAME adaptor code path serves both.
They hold data only: no forms, no assembly, no solves.
"""
from __future__ import annotations

import numpy as np

from .mesh import Mesh2D, FACET_NODES

__all__ = ["Constant", "Function", "FunctionSpace", "MixedFunctionSpace", "ShimMesh", "as_shim_mesh",
           "conditional", "lt", "gt", "le", "ge", "as_vector"]


class _Expr:
    """
    Minimal look-alike of a UFL expression tree over Constants and Functions (`ufl_operands`, operator overloads,
    `conditional`), enough to write boundary data the way the reference's examples do, e.g.
    ``elev_ramp * elev_tide_2d`` with ``elev_ramp = conditional(bnd_time < ramp_t, bnd_time / ramp_t, 1.0)``
    (examples/north_sea/model_config.py:181-192).  Data only: the adaptor evaluates it nodally.
    """
    ufl_operands = ()

    def __mul__(self, o): return _Op("mul", self, o)
    def __rmul__(self, o): return _Op("mul", o, self)
    def __add__(self, o): return _Op("add", self, o)
    def __radd__(self, o): return _Op("add", o, self)
    def __sub__(self, o): return _Op("sub", self, o)
    def __rsub__(self, o): return _Op("sub", o, self)
    def __truediv__(self, o): return _Op("div", self, o)
    def __rtruediv__(self, o): return _Op("div", o, self)
    def __neg__(self): return _Op("mul", -1.0, self)
    def __lt__(self, o): return _Op("lt", self, o)
    def __gt__(self, o): return _Op("gt", self, o)
    def __le__(self, o): return _Op("le", self, o)
    def __ge__(self, o): return _Op("ge", self, o)


class _Op(_Expr):
    def __init__(self, kind, *operands):
        self.kind = kind
        self.ufl_operands = tuple(operands)


def conditional(condition, true_value, false_value):
    return _Op("conditional", condition, true_value, false_value)


def lt(a, b): return _Op("lt", a, b)
def gt(a, b): return _Op("gt", a, b)
def le(a, b): return _Op("le", a, b)
def ge(a, b): return _Op("ge", a, b)


def as_vector(components):
    """`as_vector([..])` of numbers / scalar expressions."""
    if all(isinstance(c, (int, float, np.integer, np.floating)) for c in components):
        return np.asarray(components, dtype=np.float64)
    return _Op("vector", *components)


class _ConstDat:
    """`Constant.dat` of real Firedrake: a PyOP2 Global with a version counter."""

    def __init__(self, owner):
        self._owner = owner
        self.dat_version = 0

    @property
    def data(self):
        self.dat_version += 1
        return self._owner._v

    @property
    def data_ro(self):
        return self._owner._v


class Constant(_Expr):
    """Look-alike of firedrake.Constant (scalar or small vector): like the real one it carries `.dat`,
    `function_space()` (returning None) and `values()`."""

    def __init__(self, value):
        if isinstance(value, Constant):
            value = value._v if not value._scalar else float(value)
        self._v = np.atleast_1d(np.asarray(value, dtype=np.float64)).copy()
        self._scalar = np.ndim(value) == 0
        self.dat = _ConstDat(self)

    def function_space(self):
        return None

    def assign(self, value):
        if isinstance(value, Constant):
            value = value._v
        self.dat.data[...] = np.asarray(value, dtype=np.float64)
        return self

    def values(self):
        return self._v.copy()

    def __float__(self):
        if self._v.size != 1:
            raise TypeError("vector Constant cannot be converted to float")
        return float(self._v[0])

    @property
    def ufl_shape(self):
        return () if self._scalar else (self._v.size,)


class _Element:
    def __init__(self, family, degree, value_size):
        self._family = {"CG": "Lagrange", "DG": "Discontinuous Lagrange", "DP": "Discontinuous Lagrange"}[family]
        self._degree = degree
        self.value_size = value_size

    def family(self):
        return self._family

    def degree(self):
        return self._degree


class _NodeMap:
    def __init__(self, values):
        self.values = values
        self.arity = values.shape[1]


class _ExteriorFacets:
    def __init__(self, mesh: Mesh2D):
        self._m = mesh

    @property
    def unique_markers(self):
        return np.array(self._m.unique_markers(), dtype=np.int32)


class ShimMesh:
    """Look-alike of a Firedrake 2-D triangular mesh wrapping a `Mesh2D`."""

    def __init__(self, mesh: Mesh2D):
        self.topology_mesh = mesh
        self.exterior_facets = _ExteriorFacets(mesh)
        self.boundary_len = mesh.boundary_length()      # solver2d.py:97-98
        self.geometric_dimension = 2
        self.comm = None
        self._spaces = {}
        # coordinates: vector P1 on the geometric vertices (DG-like for periodic meshes)
        cfs = FunctionSpace(self, "CG", 1, value_size=2, _geometric=True)
        self.coordinates = Function(cfs, name="coordinates")
        self.coordinates.dat.data[:] = mesh.coords

    def cell_dimension(self):
        return 2

    def num_cells(self):
        return self.topology_mesh.n_cells


def as_shim_mesh(mesh):
    if isinstance(mesh, ShimMesh):
        return mesh
    if isinstance(mesh, Mesh2D):
        sm = getattr(mesh, "_shim", None)
        if sm is None:
            sm = ShimMesh(mesh)
            mesh._shim = sm
        return sm
    raise TypeError("expected Mesh2D or ShimMesh")


class FunctionSpace:
    """P1 CG / P1 DG / P0 DG scalar or vector spaces on a ShimMesh."""

    def __init__(self, mesh, family, degree, value_size=1, name=None, _geometric=False):
        mesh = as_shim_mesh(mesh) if not isinstance(mesh, ShimMesh) else mesh
        if family in ("DP", "Discontinuous Lagrange"):
            family = "DG"
        if family in ("Lagrange", "P"):
            family = "CG"
        if (family, degree) not in (("CG", 1), ("DG", 1), ("DG", 0)):
            raise NotImplementedError(f"shim supports P1/P1DG/P0 only, got {family}{degree}")
        self._mesh = mesh
        self.family = family
        self.degree = degree
        self.value_size = value_size
        self.name = name
        m = mesh.topology_mesh
        if family == "CG":
            if _geometric:
                self._map = m.cells.astype(np.int32)
                self._dim = m.n_vertices
            else:
                self._map = m.topo[m.cells].astype(np.int32)
                self._dim = m.n_topo_vertices
        elif degree == 1:
            self._map = np.arange(3 * m.n_cells, dtype=np.int32).reshape(-1, 3)
            self._dim = 3 * m.n_cells
        else:
            self._map = np.arange(m.n_cells, dtype=np.int32).reshape(-1, 1)
            self._dim = m.n_cells

    def mesh(self):
        return self._mesh

    def ufl_element(self):
        return _Element(self.family, self.degree, self.value_size)

    def cell_node_map(self):
        return _NodeMap(self._map)

    def dim(self):
        return self._dim * self.value_size

    def node_count(self):
        return self._dim

    @property
    def subspaces(self):
        return (self,)

    def __len__(self):
        return 1


class MixedFunctionSpace:
    def __init__(self, spaces, name=None):
        self._spaces = tuple(spaces)
        self.name = name

    def mesh(self):
        return self._spaces[0].mesh()

    @property
    def subspaces(self):
        return self._spaces

    def __iter__(self):
        return iter(self._spaces)

    def __len__(self):
        return len(self._spaces)

    def __getitem__(self, i):
        return self._spaces[i]


class _Dat:
    def __init__(self, data):
        self._data = data
        self.dat_version = 0

    @property
    def data(self):
        self.dat_version += 1
        return self._data

    @property
    def data_ro(self):
        return self._data

    @property
    def data_with_halos(self):
        return self.data

    @property
    def data_ro_with_halos(self):
        return self._data


class Function(_Expr):
    """Look-alike of firedrake.Function (data container only)."""

    def __init__(self, space, name=None, val=None):
        self._space = space
        self._name = name
        if isinstance(space, MixedFunctionSpace):
            self._subs = tuple(Function(s, name=f"{name}[{i}]") for i, s in enumerate(space.subspaces))
            self.dat = None
        else:
            shape = (space.node_count(),) if space.value_size == 1 else (space.node_count(), space.value_size)
            self.dat = _Dat(np.zeros(shape, dtype=np.float64))
            self._subs = (self,)
        if val is not None:
            self.assign(val)

    def name(self):
        return self._name

    def function_space(self):
        return self._space

    @property
    def subfunctions(self):
        return self._subs

    def sub(self, i):
        return self._subs[i]

    def assign(self, value):
        if self.dat is None:
            if isinstance(value, Function):
                for s, v in zip(self._subs, value._subs):
                    s.assign(v)
            else:
                for s in self._subs:
                    s.assign(value)
            return self
        if isinstance(value, Function):
            self.dat.data[...] = value.dat.data_ro
        elif isinstance(value, Constant):
            self.dat.data[...] = value.values() if self._space.value_size > 1 else float(value)
        else:
            self.dat.data[...] = np.asarray(value, dtype=np.float64)
        return self

    def node_coordinates(self):
        """(n_nodes, 2) coordinates of this space's nodes."""
        fs = self._space
        m = fs.mesh().topology_mesh
        x = m.coords[m.cells]                                    # (nt, 3, 2)
        out = np.zeros((fs.node_count(), 2))
        if fs.degree == 0:
            out[:] = x.mean(axis=1)
        else:
            out[fs.cell_node_map().values.reshape(-1)] = x.reshape(-1, 2)
        return out

    def interpolate(self, fn):
        """Nodal interpolation of ``fn(x, y)`` (returns scalar array or tuple/array of components)."""
        xy = self.node_coordinates()
        v = fn(xy[:, 0], xy[:, 1])
        if self._space.value_size > 1:
            v = np.stack([np.broadcast_to(np.asarray(c, dtype=float), xy.shape[:1]) for c in v], axis=-1)
        else:
            v = np.broadcast_to(np.asarray(v, dtype=float), xy.shape[:1])
        self.dat.data[...] = v
        return self

    def copy(self, deepcopy=True):
        f = Function(self._space, name=self._name)
        f.assign(self)
        return f
