"""
TEST INFRASTRUCTURE ONLY (used by tests/golden/make_reference_residual_golden.py, never by the product).

An eager numpy evaluator for the small part of the UFL / Firedrake surface that the reference's own term classes use
(thetis/shallowwater_eq.py, thetis/tracer_eq_2d.py, thetis/equation.py, thetis/utility.py DepthExpression,
thetis/rungekutta.py ERKGenericShuOsher).  Firedrake cannot be installed here; with this module standing in for
`firedrake` / `ufl`, the reference's Python source is IMPORTED FROM /root/reference AND EXECUTED UNMODIFIED: its
`residual()` methods build expression trees out of the operators below, and `assemble()` integrates them by
quadrature on P1 / P1DG triangles.  What comes out are numbers produced by the reference's own weak forms -- the
golden vectors the CPU oracle (and through it the CUDA path) is pinned against.

Scope: affine triangles, P1 (CG) and P1DG scalar / vector spaces and their mixed product, measures dx / dS / ds(marker),
first derivatives only (forward mode: every node evaluates to a value and, on request, its spatial gradient).

Conventions (Firedrake's): local facet i is opposite local vertex i; interior facets carry a '+' and a '-' cell with
n('-') = -n('+'); an unrestricted CONTINUOUS coefficient in a dS integral takes its '+' value, an unrestricted
discontinuous one is an error; `jump(v, n)` = v('+') n('+') + v('-') n('-') (contracting the last index for rank >= 1),
`jump(v)` = v('+') - v('-'), `avg(v)` = (v('+') + v('-')) / 2.
Quadrature: 2-point Gauss-Legendre on facets, the 6-point degree-3 rule in cells -- the rules of degree 2p + 1 = 3 that
the reference requests (shallowwater_eq.py:225-230); polynomial integrands up to degree 3 are integrated exactly.

Evaluated arrays have shape (E, Q, T, J) + value_shape: entities, quadrature points, test basis functions (1 when the
expression does not involve the test function), trial basis functions (1 likewise).
"""
from __future__ import annotations

import numbers

import numpy as np

FACET_NODES = np.array([[1, 2], [2, 0], [0, 1]])
_GAUSS = np.array([0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)])


def cell_rule():
    a, b, c = 0.659027622374092, 0.231933368553031, 0.109039009072877
    pts = np.array([[a, b], [a, c], [b, a], [b, c], [c, a], [c, b]])
    lam = np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], axis=1)
    return lam, np.full(6, 1.0 / 6.0)


# ------------------------------------------------------------------------------------------------ mesh
class ExteriorFacets:
    """mesh.exterior_facets: unique_markers, and Firedrake's per-facet arrays (cell, local facet number, marker)"""

    def __init__(self, m):
        self.unique_markers = np.unique(m.bf_marker)
        self.markers = np.asarray(m.bf_marker, dtype=np.int64)
        self.facet_cell = np.asarray(m.bf_cell, dtype=np.int64).reshape(-1, 1)
        self.local_facet_dat = _Dat()
        self.local_facet_dat.data = np.asarray(m.bf_lf, dtype=np.int64).reshape(-1, 1)


class Mesh:
    """Wraps a thetis_b200.mesh.Mesh2D (cells CCW, nbr / bf_* connectivity)."""

    def __init__(self, m):
        self.m = m
        self.exterior_facets = ExteriorFacets(m)
        self.cell_set = _Sized(m.n_cells)
        self.boundary_len = None
        self.geometric_dimension = 2
        x = m.coords[m.cells]                                   # (nt, 3, 2)
        self.x = x
        self.area = m.cell_area()
        p, q = x[:, FACET_NODES[:, 0]], x[:, FACET_NODES[:, 1]]
        e = q - p
        nrm = np.stack([e[..., 1], -e[..., 0]], axis=-1)        # (nt, 3, 2) outward, scaled by the facet length
        self.flen = np.hypot(e[..., 0], e[..., 1])
        self.unit_normal = nrm / self.flen[..., None]
        self.gradphi = -nrm / (2.0 * self.area[:, None, None])  # grad(phi_a) = -N_a / 2A
        self._ctx = {}

    def cell_dimension(self):
        return 2

    def ufl_cell(self):
        return triangle

    @property
    def coordinates(self):
        """the coordinate Function: P1 vector, DG on a periodic mesh (every cell carries its own vertex positions)"""
        if getattr(self, "_coords", None) is None:
            sp = Space(self, "DG" if self.m.periodic else "CG", 2, name="coordinates")
            f = Function(sp, name="coordinates")
            if self.m.periodic:
                f.dat.data[...] = self.x.reshape(-1, 2)
            else:
                f.dat.data[...] = self.m.coords
            self._coords = f
        return self._coords

    def context(self, kind, marker=None):
        key = (kind, marker)
        if key not in self._ctx:
            self._ctx[key] = Ctx(self, kind, marker)
        return self._ctx[key]


def _facet_lam(lf, reverse):
    """(E, 2, 3) barycentric coordinates of the two Gauss points on local facet lf (from its first node to its
    second; reversed for the '-' side, whose facet runs the other way)."""
    s = _GAUSS[::-1] if reverse else _GAUSS
    lam = np.zeros((lf.shape[0], 2, 3))
    e = np.arange(lf.shape[0])
    for k in range(2):
        lam[e, k, FACET_NODES[lf, 0]] = 1.0 - s[k]
        lam[e, k, FACET_NODES[lf, 1]] = s[k]
    return lam


class Ctx:
    """Integration context: cells / interior facets / exterior facets of one marker."""

    def __init__(self, mesh, kind, marker):
        m = mesh.m
        self.mesh, self.kind = mesh, kind
        if kind == "cell":
            lam, w = cell_rule()
            nt = m.n_cells
            self.cells = {None: np.arange(nt)}
            self.lam = {None: np.broadcast_to(lam, (nt,) + lam.shape)}
            self.wts = mesh.area[:, None] * w[None, :]
            self.normal = {}
            self.flen = None
        elif kind == "interior":
            cp, fp, cm, fm = m.interior_facets()
            self.cells = {"+": cp, "-": cm}
            self.lam = {"+": _facet_lam(fp, False), "-": _facet_lam(fm, True)}
            self.flen = mesh.flen[cp, fp]
            self.wts = 0.5 * self.flen[:, None] * np.ones((1, 2))
            self.normal = {"+": mesh.unit_normal[cp, fp], "-": -mesh.unit_normal[cp, fp]}
            # the two sides must describe the same physical points
            xp = np.einsum("eqa,eac->eqc", self.lam["+"], mesh.x[cp])
            xm = np.einsum("eqa,eac->eqc", self.lam["-"], mesh.x[cm])
            if not m.periodic:
                assert np.allclose(xp, xm, atol=1e-9 * max(1.0, np.abs(xp).max()))
        elif kind == "exterior":
            sel = np.nonzero(m.bf_marker == marker)[0] if marker is not None else np.arange(m.n_bfacets)
            c, f = m.bf_cell[sel], m.bf_lf[sel]
            self.cells = {None: c}
            self.lam = {None: _facet_lam(f, False)}
            self.flen = mesh.flen[c, f]
            self.wts = 0.5 * self.flen[:, None] * np.ones((1, 2))
            self.normal = {None: mesh.unit_normal[c, f]}
        else:
            raise ValueError(kind)
        self.n_ent = self.wts.shape[0]

    def side_key(self, side):
        if self.kind == "interior":
            if side is None:
                raise RestrictionError("unrestricted discontinuous quantity in an interior-facet integral")
            return side
        if side is not None:
            raise RestrictionError("restriction outside an interior-facet integral")
        return None


class RestrictionError(Exception):
    pass


# ------------------------------------------------------------------------------------------------ expressions
def _expand(a, rank_from, rank_to):
    """append value axes to an array of value rank `rank_from` so that it broadcasts against rank `rank_to`"""
    return a.reshape(a.shape + (1,) * (rank_to - rank_from))


def as_expr(x):
    if isinstance(x, Expr):
        return x
    if isinstance(x, numbers.Number):
        return Const(float(x))
    if isinstance(x, (list, tuple)):
        return ListTensor([as_expr(c) for c in x])
    if isinstance(x, np.ndarray):
        return Const(np.asarray(x, dtype=float))
    raise TypeError(f"cannot use {type(x).__name__} in an expression")


class Expr:
    shape = ()
    __array_ufunc__ = None          # numpy scalars defer to the reflected operators below

    def __bool__(self):
        return True

    # value-and-gradient evaluation: returns (v, g) with g = None unless need_grad
    def ev(self, ctx, side, need_grad):
        raise NotImplementedError

    @property
    def ufl_shape(self):
        return self.shape

    @property
    def ufl_operands(self):
        """child expressions, like ufl.core.Operator.ufl_operands (terminals: an empty tuple)"""
        if isinstance(self, (Function, Constant, Const, Argument, FacetNormal, _CellGeom, SpatialCoordinate)):
            return ()
        out = []
        for k in vars(self).values():
            if isinstance(k, (Expr, Condition)):
                out.append(k)
            elif isinstance(k, (list, tuple)):
                out.extend(c for c in k if isinstance(c, (Expr, Condition)))
        return tuple(out)

    @property
    def rank(self):
        return len(self.shape)

    # --- operators
    def __add__(self, o):
        if isinstance(o, Form):
            return NotImplemented
        if isinstance(o, numbers.Number) and o == 0:
            return self                      # UFL: expr + 0 is expr, whatever its shape (`sum(...)` starts at 0)
        return Sum(self, as_expr(o))

    def __radd__(self, o):
        if isinstance(o, numbers.Number) and o == 0:
            return self
        return Sum(as_expr(o), self)

    def __sub__(self, o):
        return Sum(self, Neg(as_expr(o)))

    def __rsub__(self, o):
        return Sum(as_expr(o), Neg(self))

    def __neg__(self):
        return Neg(self)

    def __pos__(self):
        return self

    def __mul__(self, o):
        if isinstance(o, (Form, Measure)):
            return o.__rmul__(self)
        return _product(self, as_expr(o))

    def __rmul__(self, o):
        return _product(as_expr(o), self)

    def __truediv__(self, o):
        return Division(self, as_expr(o))

    def __rtruediv__(self, o):
        return Division(as_expr(o), self)

    def __pow__(self, p):
        return Power(self, as_expr(p))

    def __abs__(self):
        return Abs(self)

    def __getitem__(self, i):
        return Indexed(self, i)

    def __call__(self, side):
        assert side in ("+", "-")
        return Restricted(self, side)

    def __gt__(self, o):
        return Condition(">", self, as_expr(o))

    def __lt__(self, o):
        return Condition("<", self, as_expr(o))

    def __ge__(self, o):
        return Condition(">=", self, as_expr(o))

    def __le__(self, o):
        return Condition("<=", self, as_expr(o))

    def __len__(self):
        if not self.shape:
            raise TypeError("scalar expression has no len()")
        return self.shape[0]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    @property
    def T(self):
        return Transposed(self)

    def dx(self, i):
        return Dx(self, i)


def _product(a, b):
    if a.rank and b.rank:
        # UFL: tensor * tensor is a contraction (matrix-vector, matrix-matrix); vector * vector is not allowed
        if a.rank == 2 and b.rank in (1, 2):
            return Dot(a, b)
        raise TypeError("product of two non-scalar expressions")
    return Product(a, b)


class Const(Expr):
    def __init__(self, value):
        self.value = np.asarray(value, dtype=float)
        self.shape = self.value.shape

    def ev(self, ctx, side, need_grad):
        v = self.value.reshape((1, 1, 1, 1) + self.shape)
        return v, (np.zeros(v.shape + (2,)) if need_grad else None)


class Constant(Expr):
    """firedrake.Constant: a mutable global value"""

    def __init__(self, value=0.0, domain=None, name=None, **kw):
        if isinstance(value, Expr):              # an expression of Constants, e.g. Constant(1.0 / rho0)
            value = value.ev(None, None, False)[0].reshape(value.shape)
        self._v = np.array(value, dtype=float)
        self.shape = self._v.shape
        self.name = name
        self.dat = _Dat()                        # a real firedrake.Constant carries a (versioned) PyOP2 Global
        self.dat.data = self._v

    def assign(self, value):
        v = value.values() if isinstance(value, Constant) else value
        self._v = np.array(v, dtype=float).reshape(self.shape)
        self.dat.data = self._v
        self.dat.dat_version += 1
        return self

    def values(self):
        return self._v.reshape(-1).copy()

    def __float__(self):
        return float(self._v)

    def ev(self, ctx, side, need_grad):
        v = self._v.reshape((1, 1, 1, 1) + self.shape)
        return v, (np.zeros(v.shape + (2,)) if need_grad else None)

    def function_space(self):
        return None


class Sum(Expr):
    def __init__(self, a, b):
        if a.shape != b.shape:
            raise TypeError(f"sum of shapes {a.shape} and {b.shape}")
        self.a, self.b, self.shape = a, b, a.shape

    def ev(self, ctx, side, need_grad):
        av, ag = self.a.ev(ctx, side, need_grad)
        bv, bg = self.b.ev(ctx, side, need_grad)
        return av + bv, (ag + bg if need_grad else None)

    def nodal(self):
        return _nodal_add(_nodal(self.a), _nodal(self.b))


class Neg(Expr):
    def __init__(self, a):
        self.a, self.shape = a, a.shape

    def ev(self, ctx, side, need_grad):
        v, g = self.a.ev(ctx, side, need_grad)
        return -v, (-g if need_grad else None)

    def nodal(self):
        return _nodal_scale(_nodal(self.a), -1.0)


class Product(Expr):
    """scalar * anything"""

    def __init__(self, a, b):
        if a.rank:
            a, b = b, a
        assert a.rank == 0
        self.a, self.b, self.shape = a, b, b.shape

    def ev(self, ctx, side, need_grad):
        av, ag = self.a.ev(ctx, side, need_grad)
        bv, bg = self.b.ev(ctx, side, need_grad)
        r = self.b.rank
        v = _expand(av, 0, r) * bv
        if not need_grad:
            return v, None
        g = ag.reshape(ag.shape[:-1] + (1,) * r + (2,)) * bv[..., None] + _expand(av, 0, r + 1) * bg
        return v, g

    def nodal(self):
        a, b = self.a, self.b
        for s, f in ((a, b), (b, a)):
            if isinstance(s, (Const, Constant)) and not s.shape:
                return _nodal_scale(_nodal(f), float(s.value if isinstance(s, Const) else s._v))
        raise NotImplementedError("nodal evaluation of a product of two fields")


class Division(Expr):
    def __init__(self, a, b):
        if b.rank:
            raise TypeError("division by a non-scalar")
        self.a, self.b, self.shape = a, b, a.shape

    def ev(self, ctx, side, need_grad):
        av, ag = self.a.ev(ctx, side, need_grad)
        bv, bg = self.b.ev(ctx, side, need_grad)
        r = self.a.rank
        v = av / _expand(bv, 0, r)
        if not need_grad:
            return v, None
        bg_e = bg.reshape(bg.shape[:-1] + (1,) * r + (2,))
        g = (ag - v[..., None] * bg_e) / _expand(bv, 0, r + 1)
        return v, g


class Power(Expr):
    def __init__(self, a, p):
        assert a.rank == 0 and p.rank == 0
        self.a, self.p = a, p

    def ev(self, ctx, side, need_grad):
        av, ag = self.a.ev(ctx, side, need_grad)
        pv, _ = self.p.ev(ctx, side, False)
        v = av ** pv
        if not need_grad:
            return v, None
        if not isinstance(self.p, (Const, Constant)):
            raise NotImplementedError("gradient of a power with a varying exponent")
        return v, (pv * av ** (pv - 1.0))[..., None] * ag


class _Unary(Expr):
    def __init__(self, a):
        assert a.rank == 0, "scalar function of a non-scalar"
        self.a = a

    def ev(self, ctx, side, need_grad):
        av, ag = self.a.ev(ctx, side, need_grad)
        v = self.f(av)
        return v, (self.df(av, v)[..., None] * ag if need_grad else None)


class Sqrt(_Unary):
    f = staticmethod(np.sqrt)
    df = staticmethod(lambda a, v: 0.5 / v)


class Abs(_Unary):
    f = staticmethod(np.abs)
    df = staticmethod(lambda a, v: np.sign(a))


class Ln(_Unary):
    f = staticmethod(np.log)
    df = staticmethod(lambda a, v: 1.0 / a)


class Sign(_Unary):
    f = staticmethod(np.sign)
    df = staticmethod(lambda a, v: np.zeros_like(a))


class Cos(_Unary):
    f = staticmethod(np.cos)
    df = staticmethod(lambda a, v: -np.sin(a))


class Sin(_Unary):
    f = staticmethod(np.sin)
    df = staticmethod(lambda a, v: np.cos(a))


class Exp(_Unary):
    f = staticmethod(np.exp)
    df = staticmethod(lambda a, v: v)


class Condition:
    def __init__(self, op, a, b):
        assert a.rank == 0 and b.rank == 0
        self.op, self.a, self.b = op, a, b

    @property
    def ufl_operands(self):
        return (self.a, self.b)

    def ev(self, ctx, side):
        av, _ = self.a.ev(ctx, side, False)
        bv, _ = self.b.ev(ctx, side, False)
        return {">": np.greater, "<": np.less, ">=": np.greater_equal, "<=": np.less_equal}[self.op](av, bv)


class Conditional(Expr):
    def __init__(self, c, a, b):
        assert a.shape == b.shape
        self.c, self.a, self.b, self.shape = c, a, b, a.shape

    def ev(self, ctx, side, need_grad):
        c = self.c.ev(ctx, side)
        av, ag = self.a.ev(ctx, side, need_grad)
        bv, bg = self.b.ev(ctx, side, need_grad)
        r = self.rank
        v = np.where(_expand(c, 0, r), av, bv)
        return v, (np.where(_expand(c, 0, r + 1), ag, bg) if need_grad else None)


class Indexed(Expr):
    def __init__(self, a, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        if len(idx) > a.rank:
            raise IndexError("too many indices")
        self.a, self.idx = a, idx
        shape = []
        for n, i in zip(a.shape, idx):
            if isinstance(i, slice):
                shape.append(len(range(*i.indices(n))))
            elif not -n <= i < n:
                raise IndexError(i)           # ends `for c in vector` loops
        self.shape = tuple(shape) + a.shape[len(idx):]

    def ev(self, ctx, side, need_grad):
        av, ag = self.a.ev(ctx, side, need_grad)
        r = self.a.rank
        lead = (slice(None),) * 4
        v = av[lead + self.idx]
        g = ag[lead + self.idx + (slice(None),) * (r - len(self.idx)) + (slice(None),)] if need_grad else None
        return v, g


class ListTensor(Expr):
    def __init__(self, comps):
        self.c = comps
        assert all(k.shape == comps[0].shape for k in comps)
        self.shape = (len(comps),) + comps[0].shape

    def ev(self, ctx, side, need_grad):
        ev = [k.ev(ctx, side, need_grad) for k in self.c]
        vs = np.broadcast_arrays(*[e[0] for e in ev])
        v = np.stack(vs, axis=4)
        if not need_grad:
            return v, None
        gs = np.broadcast_arrays(*[e[1] for e in ev])
        return v, np.stack(gs, axis=4)


class Transposed(Expr):
    def __init__(self, a):
        assert a.rank == 2
        self.a, self.shape = a, a.shape[::-1]

    def ev(self, ctx, side, need_grad):
        v, g = self.a.ev(ctx, side, need_grad)
        return np.swapaxes(v, 4, 5), (np.swapaxes(g, 4, 5) if need_grad else None)


class Sym(Expr):
    def __init__(self, a):
        assert a.rank == 2
        self.a, self.shape = a, a.shape

    def ev(self, ctx, side, need_grad):
        v, g = self.a.ev(ctx, side, need_grad)
        return 0.5 * (v + np.swapaxes(v, 4, 5)), (0.5 * (g + np.swapaxes(g, 4, 5)) if need_grad else None)


_LET = "abcdefgh"


def _contract(av, bv, ra, rb, n):
    """contract the last n value axes of a with the first n value axes of b (leading 4 axes broadcast)"""
    ia = _LET[:ra]
    ib = ia[ra - n:] + _LET[ra:ra + rb - n]
    out = ia[:ra - n] + ib[n:]
    return np.einsum(f"...{ia},...{ib}->...{out}", av, bv)


class Dot(Expr):
    """contraction of the last index of a with the first index of b; Inner contracts everything"""

    def __init__(self, a, b, n=1):
        self.a, self.b, self.n = a, b, n
        if a.rank < n or b.rank < n or a.shape[a.rank - n:] != b.shape[:n]:
            raise TypeError(f"cannot contract shapes {a.shape} and {b.shape} over {n} indices")
        self.shape = a.shape[:a.rank - n] + b.shape[n:]

    def ev(self, ctx, side, need_grad):
        av, ag = self.a.ev(ctx, side, need_grad)
        bv, bg = self.b.ev(ctx, side, need_grad)
        ra, rb, n = self.a.rank, self.b.rank, self.n
        v = _contract(av, bv, ra, rb, n)
        if not need_grad:
            return v, None
        # d(a.b) = da.b + a.db, the derivative index stays last
        ia = _LET[:ra]
        ib = ia[ra - n:] + _LET[ra:ra + rb - n]
        out = ia[:ra - n] + ib[n:]
        g = np.einsum(f"...{ia}z,...{ib}->...{out}z", ag, bv) + np.einsum(f"...{ia},...{ib}z->...{out}z", av, bg)
        return v, g


class Outer(Expr):
    def __init__(self, a, b):
        self.a, self.b, self.shape = a, b, a.shape + b.shape

    def ev(self, ctx, side, need_grad):
        av, ag = self.a.ev(ctx, side, need_grad)
        bv, bg = self.b.ev(ctx, side, need_grad)
        ra, rb = self.a.rank, self.b.rank
        ia, ib = _LET[:ra], _LET[ra:ra + rb]
        v = np.einsum(f"...{ia},...{ib}->...{ia}{ib}", av, bv)
        if not need_grad:
            return v, None
        g = np.einsum(f"...{ia}z,...{ib}->...{ia}{ib}z", ag, bv) + np.einsum(f"...{ia},...{ib}z->...{ia}{ib}z", av, bg)
        return v, g


class Grad(Expr):
    def __init__(self, a):
        self.a, self.shape = a, a.shape + (2,)

    def ev(self, ctx, side, need_grad):
        if need_grad:
            raise NotImplementedError("second derivatives")
        _, g = self.a.ev(ctx, side, True)
        return g, None


class NablaGrad(Expr):
    """(nabla_grad v)_{i...} = d_i v_{...}"""

    def __init__(self, a):
        self.a, self.shape = a, (2,) + a.shape

    def ev(self, ctx, side, need_grad):
        if need_grad:
            raise NotImplementedError("second derivatives")
        _, g = self.a.ev(ctx, side, True)
        return np.moveaxis(g, -1, 4), None


class Div(Expr):
    """div contracts the LAST index with the derivative; nabla_div the FIRST"""

    def __init__(self, a, first=False):
        assert a.rank >= 1
        self.a, self.first = a, first
        self.shape = a.shape[1:] if first else a.shape[:-1]

    def ev(self, ctx, side, need_grad):
        if need_grad:
            raise NotImplementedError("second derivatives")
        _, g = self.a.ev(ctx, side, True)
        r = self.a.rank
        ia = _LET[:r]
        k = ia[0] if self.first else ia[-1]
        out = ia[1:] if self.first else ia[:-1]
        return np.einsum(f"...{ia}{k}->...{out}", g), None


class Dx(Expr):
    def __init__(self, a, i):
        self.a, self.i, self.shape = a, i, a.shape

    def ev(self, ctx, side, need_grad):
        if need_grad:
            raise NotImplementedError("second derivatives")
        _, g = self.a.ev(ctx, side, True)
        return g[..., self.i], None


class Restricted(Expr):
    def __init__(self, a, side):
        self.a, self.side, self.shape = a, side, a.shape

    def ev(self, ctx, side, need_grad):
        if ctx.kind != "interior":
            raise RestrictionError("restricted expression outside dS")
        return self.a.ev(ctx, self.side, need_grad)


class FacetNormal(Expr):
    shape = (2,)

    def __init__(self, mesh):
        self.mesh = mesh

    def ev(self, ctx, side, need_grad):
        if ctx.kind == "cell":
            raise RestrictionError("FacetNormal in a cell integral")
        n = ctx.normal[ctx.side_key(side)]
        v = n[:, None, None, None, :]
        return v, (np.zeros(v.shape + (2,)) if need_grad else None)


class _CellGeom(Expr):
    def __init__(self, mesh):
        self.mesh = mesh

    def ev(self, ctx, side, need_grad):
        if ctx.kind == "interior" and side is None:
            side = "+"
        v = self.value(ctx, ctx.cells[ctx.side_key(side)])[:, None, None, None]
        return v, (np.zeros(v.shape + (2,)) if need_grad else None)


class CellVolume(_CellGeom):
    def value(self, ctx, cells):
        return self.mesh.area[cells]


class FacetArea(_CellGeom):
    def value(self, ctx, cells):
        if ctx.kind == "cell":
            raise RestrictionError("FacetArea in a cell integral")
        return ctx.flen


class CellSize(_CellGeom):
    """CellDiameter: the longest edge.  Constructed by the reference's term classes, read by none of the P1DG terms."""

    def value(self, ctx, cells):
        return self.mesh.flen[cells].max(axis=1)


CellDiameter = CellSize


class SpatialCoordinate(Expr):
    shape = (2,)

    def __init__(self, mesh):
        self.mesh = mesh

    def ev(self, ctx, side, need_grad):
        if ctx.kind == "interior" and side is None:
            side = "+"
        k = ctx.side_key(side)
        v = np.einsum("eqa,eac->eqc", ctx.lam[k], self.mesh.x[ctx.cells[k]])[:, :, None, None, :]
        g = np.broadcast_to(np.eye(2), v.shape + (2,)).copy() if need_grad else None
        return v, g


# ------------------------------------------------------------------------------------------------ spaces
class Element:
    def __init__(self, family, degree, value_shape=()):
        self._family, self._degree, self._vs = family, degree, value_shape

    def family(self):
        return self._family

    def degree(self):
        return self._degree

    def value_shape(self):
        return self._vs


class HDivElement:
    pass


class TensorProductElement:
    pass


class EnrichedElement:
    pass


class MixedElement(Element):
    def __init__(self, subs):
        super().__init__("Mixed", 1, (sum(max(1, int(np.prod(e.value_shape()))) for e in subs),))
        self.sub_elements = list(subs)


class VectorElement(Element):
    def __init__(self, sub, dim=2):
        super().__init__(sub.family(), sub.degree(), (dim,))
        self.sub_elements = [sub] * dim


def FiniteElement(family, cell=None, degree=1, variant=None, **kw):
    return Element({"CG": "Lagrange", "P": "Lagrange", "DG": "Discontinuous Lagrange",
                    "DP": "Discontinuous Lagrange"}.get(family, family), degree)


class _Sized:
    def __init__(self, size):
        self.size = size


class _Dat:
    """PyOP2 Dat look-alike: `data` is the nodal array (`data_ro`, `*_with_halos`: the same array, one process), and
    `dat_version` counts the assignments"""

    def __init__(self):
        self.dat_version = 0
        self.data = None

    data_ro = property(lambda self: self.data)
    data_with_halos = property(lambda self: self.data)
    data_ro_with_halos = property(lambda self: self.data)


class _NodeMap:
    def __init__(self, values):
        self.values = values
        self.arity = values.shape[1]


class Space:
    """P1 ('CG') or P1DG ('DG') space of scalars (vdim 0) or 2-vectors (vdim 2) on a Mesh"""

    def __init__(self, mesh, family, vdim=0, name=None):
        family = {"Lagrange": "CG", "Discontinuous Lagrange": "DG", "P": "CG", "DP": "DG"}.get(family, family)
        assert family in ("CG", "DG")
        self._mesh, self.family, self.vdim, self.name = mesh, family, vdim, name
        m = mesh.m
        if family == "DG":
            self.cell_nodes = np.arange(3 * m.n_cells).reshape(-1, 3)
            self.ndof = 3 * m.n_cells
        else:
            self.cell_nodes = m.topo[m.cells] if m.periodic else m.cells.astype(np.int64)
            self.ndof = int(self.cell_nodes.max()) + 1
        self.ncomp = max(vdim, 1)
        self.nb = 3 * self.ncomp
        self.value_shape = (vdim,) if vdim else ()
        self.size = self.ndof * self.ncomp
        self.subspaces = [self]

    def mesh(self):
        return self._mesh

    def ufl_domain(self):
        return self._mesh

    def ufl_element(self):
        el = Element("Discontinuous Lagrange" if self.family == "DG" else "Lagrange", 1)
        return VectorElement(el, self.vdim) if self.vdim else el

    def __mul__(self, other):
        return MixedSpace([self, other])

    def __getitem__(self, i):
        assert i == 0
        return self

    def __len__(self):
        return 1

    def cell_node_map(self):
        return _NodeMap(self.cell_nodes)

    # local basis function t of the space on one cell: node t // ncomp, component t % ncomp
    def basis_layout(self):
        node = np.repeat(np.arange(3), self.ncomp)
        comp = np.tile(np.arange(self.ncomp), 3)
        return node, comp

    def local_dofs(self, cells):
        """(E, nb) indices into the flattened global vector of the space"""
        node, comp = self.basis_layout()
        return self.cell_nodes[cells][:, node] * self.ncomp + comp[None, :]


class MixedSpace:
    def __init__(self, spaces):
        self.subspaces = list(spaces)
        self._mesh = spaces[0].mesh()
        self.value_shape = (sum(s.ncomp for s in spaces),)
        self.nb = sum(s.nb for s in spaces)
        self.size = sum(s.size for s in spaces)
        self.family = "mixed"

    def mesh(self):
        return self._mesh

    def ufl_domain(self):
        return self._mesh

    def ufl_element(self):
        return MixedElement([s.ufl_element() for s in self.subspaces])

    def sub(self, i):
        return self.subspaces[i]

    def __iter__(self):
        return iter(self.subspaces)

    def basis_layout(self):
        """(space index, node, component within the MIXED value) of every local basis function"""
        sp, node, comp = [], [], []
        c0 = 0
        for i, s in enumerate(self.subspaces):
            n, c = s.basis_layout()
            sp += [i] * s.nb
            node += list(n)
            comp += list(c + c0)
            c0 += s.ncomp
        return np.array(sp), np.array(node), np.array(comp)

    def local_dofs(self, cells):
        out, off = [], 0
        for s in self.subspaces:
            out.append(s.local_dofs(cells) + off)
            off += s.size
        return np.concatenate(out, axis=1)


def FunctionSpace(mesh, family, degree=1, name=None, **kw):
    if isinstance(family, Element):
        el = family
        assert el.degree() == 1, "P1 / P1DG only"
        return Space(mesh, el.family(), el.value_shape()[0] if el.value_shape() else 0, name)
    assert degree == 1, "P1 / P1DG only"
    return Space(mesh, family, 0, name)


def VectorFunctionSpace(mesh, family, degree=1, dim=2, name=None, **kw):
    if isinstance(family, Element):
        family, degree = family.family(), family.degree()
    assert degree == 1 and dim == 2
    return Space(mesh, family, 2, name)


def TensorFunctionSpace(*a, **kw):
    raise NotImplementedError("tensor spaces")


def MixedFunctionSpace(spaces, **kw):
    return MixedSpace(spaces)


def _basis_values(space, ctx, side, need_grad, axis):
    """Test (axis 2) or trial (axis 3) function of `space` in ctx: arrays (E, Q, T, J) + value_shape"""
    mesh = space.mesh()
    if ctx.kind == "interior":
        if side is None:
            raise RestrictionError("test / trial function must be restricted in dS")
        sides, n_slots = ("+", "-"), 2
    else:
        sides, n_slots = (None,), 1
        if side is not None:
            raise RestrictionError("restriction outside dS")
    if isinstance(space, MixedSpace):
        _, node, comp = space.basis_layout()
    else:
        node, comp = space.basis_layout()
    nb, nc = space.nb, (space.value_shape[0] if space.value_shape else 1)
    E, Q = ctx.n_ent, ctx.wts.shape[1]
    v = np.zeros((E, Q, n_slots * nb, nc))
    g = np.zeros((E, Q, n_slots * nb, nc, 2)) if need_grad else None
    slot = sides.index(side)
    k = ctx.side_key(side)
    lam, cells = ctx.lam[k], ctx.cells[k]
    t = slot * nb + np.arange(nb)
    v[:, :, t, comp] = lam[:, :, node]
    if need_grad:
        g[:, :, t, comp, :] = mesh.gradphi[cells][:, None, node, :]
    if not space.value_shape:
        v = v[..., 0]
        g = g[..., 0, :] if need_grad else None
    if axis == 2:
        v = v[:, :, :, None]
        g = g[:, :, :, None] if need_grad else None
    else:
        v = v[:, :, None, :]
        g = g[:, :, None, :] if need_grad else None
    return v, g


class Argument(Expr):
    def __init__(self, space, number):
        self.space, self.number, self.shape = space, number, tuple(space.value_shape)

    def function_space(self):
        return self.space

    def ev(self, ctx, side, need_grad):
        return _basis_values(self.space, ctx, side, need_grad, 2 if self.number == 0 else 3)


def TestFunction(space):
    return Argument(space, 0)


def TrialFunction(space):
    return Argument(space, 1)


def split(f):
    if isinstance(f, Function) and isinstance(f.space, MixedSpace):
        return tuple(f.subfunctions)
    sp = f.space if isinstance(f, Argument) else f.function_space()
    if not isinstance(sp, MixedSpace):
        return (f,)
    out, c0 = [], 0
    for s in sp.subspaces:
        out.append(Indexed(f, slice(c0, c0 + s.ncomp)) if s.vdim else Indexed(f, c0))
        c0 += s.ncomp
    return tuple(out)


def TestFunctions(space):
    return split(TestFunction(space))


def TrialFunctions(space):
    return split(TrialFunction(space))


class _DofDset:
    def __init__(self, dim):
        self.dim = dim


class Function(Expr):
    def __init__(self, space, val=None, name=None, **kw):
        if isinstance(space, Function):
            space, val = space.space, space
        self.space, self.name = space, name
        self.shape = tuple(space.value_shape)
        if isinstance(space, MixedSpace):
            self.subfunctions = [Function(s) for s in space.subspaces]
            self.dat = None
            self.dof_dset = _DofDset(tuple(s.ncomp for s in space.subspaces))
        else:
            self.subfunctions = [self]
            self.dat = _Dat()
            self.dat.data = np.zeros((space.ndof, 2) if space.vdim else space.ndof)
            self.dof_dset = _DofDset((space.ncomp,))
        if val is not None:
            self.assign(val)

    def function_space(self):
        return self.space

    def sub(self, i):
        return self.subfunctions[i]

    def split(self):
        return tuple(self.subfunctions)

    def copy(self, deepcopy=True):
        return Function(self.space).assign(self)

    def vector_data(self):
        """flattened global vector [space 0 (interleaved components), space 1, ...]"""
        if isinstance(self.space, MixedSpace):
            return np.concatenate([f.dat.data.reshape(-1) for f in self.subfunctions])
        return self.dat.data.reshape(-1).copy()

    def set_vector_data(self, vec):
        off = 0
        for f in self.subfunctions:
            n = f.space.size
            f.dat.data[...] = vec[off:off + n].reshape(f.dat.data.shape)
            f.dat.dat_version += 1
            off += n

    def nodal(self):
        return [f.dat.data.copy() for f in self.subfunctions]

    def assign(self, other):
        vals = other if isinstance(other, list) else _nodal(other)
        if isinstance(vals, float):
            for f in self.subfunctions:
                f.dat.data[...] = vals
                f.dat.dat_version += 1
        else:
            assert len(vals) == len(self.subfunctions)
            for f, v in zip(self.subfunctions, vals):
                f.dat.data[...] = v
                f.dat.dat_version += 1
        return self

    def __iadd__(self, other):
        if isinstance(other, numbers.Number) and other == 0:
            return self
        return self.assign(_nodal_add(self.nodal(), _nodal(other)))

    def interpolate(self, expr):
        """nodal interpolation (P1 / P1DG): the expression at the three vertices of every cell"""
        assert not isinstance(self.space, MixedSpace)
        mesh = self.space.mesh()
        ctx = _VertexCtx(mesh)
        v, _ = as_expr(expr).ev(ctx, None, False)
        v = np.broadcast_to(v, (ctx.n_ent, 3, 1, 1) + self.shape)[:, :, 0, 0]
        self.dat.data[self.space.cell_nodes.reshape(-1)] = v.reshape((-1,) + self.shape)
        self.dat.dat_version += 1
        return self

    def project(self, expr):
        return self.interpolate(expr)

    def ev(self, ctx, side, need_grad):
        if isinstance(self.space, MixedSpace):
            parts = [f.ev(ctx, side, need_grad) for f in self.subfunctions]
            vs = [p[0] if f.shape else p[0][..., None] for p, f in zip(parts, self.subfunctions)]
            v = np.concatenate(vs, axis=4)
            g = None
            if need_grad:
                gs = [p[1] if f.shape else p[1][..., None, :] for p, f in zip(parts, self.subfunctions)]
                g = np.concatenate(gs, axis=4)
            return v, g
        sp, mesh = self.space, self.space.mesh()
        if ctx.kind == "interior" and side is None:
            if sp.family == "DG":
                raise RestrictionError(f"discontinuous function {self.name!r} must be restricted in dS")
            side = "+"
        k = ctx.side_key(side)
        cells, lam = ctx.cells[k], ctx.lam[k]
        nod = self.dat.data[sp.cell_nodes[cells]]                 # (E, 3[, 2])
        v = np.einsum("eqa,ea...->eq...", lam, nod)[:, :, None, None]
        g = None
        if need_grad:
            gg = np.einsum("ea...,eaz->e...z", nod, mesh.gradphi[cells])
            g = np.broadcast_to(gg[:, None, None, None], v.shape + (2,))
        return v, g


class _VertexCtx:
    """evaluation at the three vertices of every cell (interpolation)"""
    kind = "cell"

    def __init__(self, mesh):
        nt = mesh.m.n_cells
        self.mesh = mesh
        self.cells = {None: np.arange(nt)}
        self.lam = {None: np.broadcast_to(np.eye(3), (nt, 3, 3))}
        self.wts = np.zeros((nt, 3))
        self.n_ent = nt
        self.normal = {}

    def side_key(self, side):
        return None


def _nodal(x):
    if isinstance(x, numbers.Number):
        return float(x)
    if isinstance(x, (Const, Constant)) and not x.shape:
        return float(x.value if isinstance(x, Const) else x._v)
    if hasattr(x, "nodal"):
        return x.nodal()
    raise NotImplementedError(f"nodal evaluation of {type(x).__name__}")


def _nodal_add(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return a + b
    if isinstance(a, float):
        a, b = b, a
    if isinstance(b, float):
        return [x + b for x in a]
    return [x + y for x, y in zip(a, b)]


def _nodal_scale(a, s):
    return a * s if isinstance(a, float) else [x * s for x in a]


# ------------------------------------------------------------------------------------------------ forms
class Measure:
    def __init__(self, kind, marker=None):
        self.kind, self.marker = kind, marker

    def __call__(self, subdomain_id=None, degree=None, domain=None, **kw):
        if isinstance(subdomain_id, Mesh):
            subdomain_id = None
        m = self.marker if subdomain_id is None else subdomain_id
        return Measure(self.kind, None if m is None else int(m))

    def __rmul__(self, integrand):
        return Form([(as_expr(integrand), self)])


dx = Measure("cell")
dS = Measure("interior")
ds = Measure("exterior")
ds_t = ds_b = ds_v = Measure("exterior")
dS_h = dS_v = Measure("interior")


class Form:
    def __init__(self, integrals):
        self.integrals = list(integrals)

    def __add__(self, o):
        if isinstance(o, numbers.Number) and o == 0:
            return self
        if not isinstance(o, Form):
            return NotImplemented
        return Form(self.integrals + o.integrals)

    __radd__ = __add__

    def __neg__(self):
        return Form([(Neg(e), m) for e, m in self.integrals])

    def __sub__(self, o):
        return self + (-o)

    def __rmul__(self, s):
        s = as_expr(s)
        return Form([(Product(s, e), m) for e, m in self.integrals])

    __mul__ = __rmul__

    def __eq__(self, o):
        if isinstance(o, Form):
            return FormEquation(self, o)          # `a == L`, the argument of solve()
        return self is o

    def __ne__(self, o):
        return not (self is o)

    __hash__ = object.__hash__

    def arguments(self):
        return _arguments(self)


def _walk(e, seen):
    if id(e) in seen:
        return
    seen[id(e)] = e
    for k in vars(e).values():
        if isinstance(k, (Expr, Condition)):
            _walk(k, seen)
        elif isinstance(k, (list, tuple)):
            for c in k:
                if isinstance(c, (Expr, Condition)):
                    _walk(c, seen)


def _arguments(form):
    seen = {}
    for e, _ in form.integrals:
        _walk(e, seen)
    return sorted({(a.number, id(a.space)): a for a in seen.values() if isinstance(a, Argument)}.values(),
                  key=lambda a: a.number)


def assemble(form, tensor=None, **kw):
    """Scalar (no arguments), vector (test function) -> Function, or cell-block matrix (test + trial, dx only)."""
    if isinstance(form, Expr):
        raise TypeError("assemble of a bare expression")
    args = _arguments(form)
    mesh = None
    if not args:
        total = 0.0
        for e, ms in form.integrals:
            mesh = _find_mesh(e)
            ctx = mesh.context(ms.kind, ms.marker)
            v, _ = e.ev(ctx, None, False)
            v = np.broadcast_to(v, (ctx.n_ent, ctx.wts.shape[1], 1, 1))[:, :, 0, 0]
            total += float((v * ctx.wts).sum())
        return total
    test = args[0].space
    mesh = test.mesh()
    if len(args) == 1:
        vec = np.zeros(test.size)
        for e, ms in form.integrals:
            if e.shape:
                raise TypeError("integrand is not a scalar")
            ctx = mesh.context(ms.kind, ms.marker)
            if ctx.n_ent == 0:
                continue
            v, _ = e.ev(ctx, None, False)
            nslot = 2 if ctx.kind == "interior" else 1
            if v.shape[2] != nslot * test.nb or v.shape[3] != 1:
                raise TypeError("integrand of a linear form must be linear in the test function")
            v = np.broadcast_to(v, (ctx.n_ent, ctx.wts.shape[1], nslot * test.nb, 1))[..., 0]
            loc = np.einsum("eqt,eq->et", v, ctx.wts)
            sides = ("+", "-") if ctx.kind == "interior" else (None,)
            dofs = np.concatenate([test.local_dofs(ctx.cells[s]) for s in sides], axis=1)
            np.add.at(vec, dofs.reshape(-1), loc.reshape(-1))
        out = tensor if tensor is not None else Function(test)
        out.set_vector_data(vec)
        return out
    trial = args[1].space
    blocks = np.zeros((mesh.m.n_cells, test.nb, trial.nb))
    for e, ms in form.integrals:
        if ms.kind != "cell":
            raise NotImplementedError("bilinear forms: cell integrals only (block-diagonal DG mass matrices)")
        ctx = mesh.context("cell")
        v, _ = e.ev(ctx, None, False)
        v = np.broadcast_to(v, (ctx.n_ent, ctx.wts.shape[1], test.nb, trial.nb))
        blocks += np.einsum("eqtj,eq->etj", v, ctx.wts)
    return BlockMatrix(test, trial, blocks)


def _find_mesh(e):
    seen = {}
    _walk(e, seen)
    for k in seen.values():
        m = getattr(k, "mesh", None)
        if isinstance(m, Mesh):
            return m
        sp = getattr(k, "space", None)
        if sp is not None:
            return sp.mesh()
    raise ValueError("no mesh in expression")


class BlockMatrix:
    def __init__(self, test, trial, blocks):
        self.test, self.trial, self.blocks = test, trial, blocks


class FormEquation:
    def __init__(self, lhs, rhs):
        self.lhs, self.rhs = lhs, rhs

    def __bool__(self):
        return False


def solve(equation, u, bcs=None, solver_parameters=None, **kw):
    """solve(a == L, u): the cell blocks of `a` scattered into a global sparse matrix (continuous spaces couple the
    cells), direct solve"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    A = assemble(equation.lhs)
    b = assemble(equation.rhs).vector_data()
    cells = np.arange(A.test.mesh().m.n_cells)
    rows = A.test.local_dofs(cells)
    cols = A.trial.local_dofs(cells)
    M = sp.coo_matrix((A.blocks.reshape(-1), (np.repeat(rows, cols.shape[1], axis=1).reshape(-1),
                                             np.tile(cols, (1, rows.shape[1])).reshape(-1))),
                      shape=(A.test.size, A.trial.size)).tocsc()
    u.set_vector_data(spla.spsolve(M, b))
    return u


class LinearVariationalProblem:
    def __init__(self, a, L, u, bcs=None, **kw):
        self.a, self.L, self.u = a, L, u


class LinearVariationalSolver:
    """Solves a(u, v) = L(v) for a cell-block-diagonal a (DG mass matrices): re-assembles both sides at every solve,
    like the reference's solver does for its (non-constant-Jacobian) default"""

    def __init__(self, problem, **kw):
        self.problem = problem

    def solve(self):
        p = self.problem
        A = assemble(p.a)
        b = assemble(p.L)
        mesh = A.test.mesh()
        cells = np.arange(mesh.m.n_cells)
        dofs = A.test.local_dofs(cells)
        rhs = b.vector_data()[dofs]
        x = np.linalg.solve(A.blocks, rhs[..., None])[..., 0]
        vec = np.zeros(A.trial.size)
        vec[A.trial.local_dofs(cells)] = x
        p.u.set_vector_data(vec)


# ------------------------------------------------------------------------------------------------ UFL functions
def inner(a, b):
    a, b = as_expr(a), as_expr(b)
    if a.rank == 0 or b.rank == 0:
        return _product(a, b)
    return Dot(a, b, n=a.rank)


def dot(a, b):
    a, b = as_expr(a), as_expr(b)
    if a.rank == 0 or b.rank == 0:
        return _product(a, b)
    return Dot(a, b, n=1)


def outer(a, b):
    return Outer(as_expr(a), as_expr(b))


def grad(a):
    return Grad(as_expr(a))


def nabla_grad(a):
    return NablaGrad(as_expr(a))


def div(a):
    return Div(as_expr(a))


def nabla_div(a):
    return Div(as_expr(a), first=True)


def sym(a):
    return Sym(a)


def transpose(a):
    return Transposed(a)


def sqrt(a):
    return Sqrt(as_expr(a))


def ln(a):
    return Ln(as_expr(a))


def sign(a):
    return Sign(as_expr(a))


def cos(a):
    return Cos(as_expr(a))


def sin(a):
    return Sin(as_expr(a))


def exp(a):
    return Exp(as_expr(a))


def conditional(c, a, b):
    return Conditional(c, as_expr(a), as_expr(b))


def gt(a, b):
    return as_expr(a) > b


def lt(a, b):
    return as_expr(a) < b


def ge(a, b):
    return as_expr(a) >= b


def le(a, b):
    return as_expr(a) <= b


def max_value(a, b):
    a, b = as_expr(a), as_expr(b)
    return Conditional(a > b, a, b)


def min_value(a, b):
    a, b = as_expr(a), as_expr(b)
    return Conditional(a < b, a, b)


def as_vector(c):
    return ListTensor([as_expr(k) for k in c])


def as_matrix(rows):
    return ListTensor([ListTensor([as_expr(k) for k in r]) for r in rows])


as_tensor = as_matrix


def avg(a):
    a = as_expr(a)
    return 0.5 * (a("+") + a("-"))


def jump(a, n=None):
    a = as_expr(a)
    if n is None:
        return a("+") - a("-")
    n = as_expr(n)
    if a.rank == 0:
        return a("+") * n("+") + a("-") * n("-")
    return dot(a("+"), n("+")) + dot(a("-"), n("-"))


def Dx_(a, i):
    return Dx(as_expr(a), i)


def unit_vectors(d):
    return tuple(Const(np.eye(d)[i]) for i in range(d))


def Identity(d):
    return Const(np.eye(d))


class VertexBasedLimiter:
    """Placeholder so that `thetis/limiter.py` imports: Firedrake's limiter kernels are C strings run by PyOP2 and are
    not restated here (the reference's limiter cannot be executed on this stand-in)."""

    def __init__(self, space):
        raise NotImplementedError("firedrake.VertexBasedLimiter is not available on the numpy stand-in")


def estimate_total_polynomial_degree(e, default_degree=1):
    """ufl.algorithms.estimate_total_polynomial_degree for the node types above, with UFL's own rules: coefficients
    count their element degree, sums take the maximum, products AND quotients add, a conditional takes the maximum of
    its two values (UFL ignores the condition), non-polynomial functions add 2"""
    d = estimate_total_polynomial_degree
    if isinstance(e, (Function, Argument)):
        return 1
    if isinstance(e, SpatialCoordinate):
        return 1
    if isinstance(e, (Const, Constant, FacetNormal, _CellGeom)):
        return 0
    if isinstance(e, (Sum, ListTensor)):
        return max(d(k) for k in e.ufl_operands)
    if isinstance(e, (Product, Division, Dot, Outer)):
        return sum(d(k) for k in e.ufl_operands)
    if isinstance(e, Power):
        p = e.p.ev(None, None, False)[0].reshape(()) if isinstance(e.p, (Const, Constant)) else 2
        return d(e.a) * int(p) if float(p) == int(p) and p >= 0 else d(e.a) + 2
    if isinstance(e, Conditional):
        return max(d(e.a), d(e.b))
    if isinstance(e, (Sqrt, Ln, Cos, Sin, Exp)):
        return d(e.a) + 2
    if isinstance(e, (Grad, NablaGrad, Div, Dx)):
        return max(d(e.a) - 1, 0)
    ops = e.ufl_operands
    return max(d(k) for k in ops) if ops else 0


triangle = "triangle"
quadrilateral = "quadrilateral"


def errornorm(a, b, **kw):
    d = as_expr(a) - as_expr(b)
    return float(np.sqrt(assemble(inner(d, d) * dx)))


def norm(a, **kw):
    a = as_expr(a)
    return float(np.sqrt(assemble(inner(a, a) * dx)))
