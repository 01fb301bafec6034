"""
Set-ups of the reference's known-answer tests for the explicit dg-dg path, restated in numpy (shared by the CPU
oracle KATs in tests/test_oracle_reference_kat.py and the GPU runs in tests/test_gpu_reference_kat.py):

* Rossby soliton           /root/reference/test/swe2d/test_rossby_wave.py:23-257
* steady-state basin MMS   /root/reference/test/swe2d/test_steady_state_basin_mms.py:16-245
* tracer h-advection       /root/reference/test/tracerEq/test_h-advection_mes_2d.py:9-122

The analytic fields are re-derived here (sympy for the MMS sources); tests/golden/reference_kat_fields.npz holds the
values the reference's own functions produce at sample points (tests/golden/make_reference_kat_golden.py) and the
tests pin these restatements to them.
"""
import functools

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# ------------------------------------------------------------------------------------------------ quadrature / projection
@functools.lru_cache(None)
def duffy_rule(n=6):
    """Gauss x Gauss rule on the triangle through the Duffy map, exact to degree 2n-2; barycentric points, weights sum 1."""
    g, w = np.polynomial.legendre.leggauss(n)
    g, w = 0.5 * (g + 1.0), 0.5 * w
    u, v = np.meshgrid(g, g, indexing="ij")
    wu, wv = np.meshgrid(w, w, indexing="ij")
    x, y = u.ravel(), (v * (1.0 - u)).ravel()
    wt = (wu * wv * (1.0 - u)).ravel()
    lam = np.stack([1.0 - x - y, x, y], 1)
    return lam, wt / wt.sum()


_MREF = (np.ones((3, 3)) + np.eye(3)) / 12.0


def project_dg1(mesh, fn, n=6):
    """L2 projection of fn(x, y) [scalar or (...,k)] onto P1DG: (nt, 3[,k])  (Function(H_2d).project(expr))."""
    lam, w = duffy_rule(n)
    xc = mesh.coords[mesh.cells]
    xq = np.einsum("qa,cai->cqi", lam, xc)
    f = np.asarray(fn(xq[..., 0], xq[..., 1]))
    rhs = np.einsum("q,qa,cq...->ca...", w, lam, f)
    return np.einsum("ab,cb...->ca...", np.linalg.inv(_MREF), rhs)


def project_dg1_nodal_product(mesh, *nodal, scale=1.0):
    """P1DG projection of the pointwise product of P1(DG) nodal fields (nt, 3) (e.g. uv_ana[0]*(bath+elev_ana)*ly)."""
    lam, w = duffy_rule(4)
    prod = np.ones((mesh.n_cells, lam.shape[0]))
    for f in nodal:
        prod = prod * np.einsum("qa,ca->cq", lam, f)
    rhs = np.einsum("q,qa,cq->ca", w, lam, prod) * scale
    return np.einsum("ab,cb->ca", np.linalg.inv(_MREF), rhs)


def project_cg1(mesh, fn, n=6):
    """L2 projection onto continuous P1 (one value per vertex): Function(P1_2d).project(expr)."""
    lam, w = duffy_rule(n)
    xc = mesh.coords[mesh.cells]
    area = mesh.cell_area()
    xq = np.einsum("qa,cai->cqi", lam, xc)
    f = np.asarray(fn(xq[..., 0], xq[..., 1]))
    rhs_c = np.einsum("c,q,qa,cq->ca", area, w, lam, f)
    nv = mesh.n_vertices
    rhs = np.bincount(mesh.cells.ravel(), rhs_c.ravel(), nv)
    rows = np.repeat(mesh.cells, 3, axis=1).ravel()
    cols = np.tile(mesh.cells, (1, 3)).ravel()
    vals = (area[:, None, None] * _MREF[None]).ravel()
    M = sp.csr_matrix((vals, (rows, cols)), shape=(nv, nv))
    return spla.spsolve(M.tocsc(), rhs)


# ------------------------------------------------------------------------------------------------ Rossby soliton
# Boyd's equatorial Rossby soliton, zeroth + first order asymptotic solution (test_rossby_wave.py:23-130).  The
# unnormalised Hermite-series coefficients are the published constants of the first-order correction.
_ROSSBY_U = {0: 1.7892760e+00, 2: 0.1164146e+00, 4: -0.3266961e-03, 6: -0.1274022e-02, 8: 0.4762876e-04,
             10: -0.1120652e-05, 12: 0.1996333e-07, 14: -0.2891698e-09, 16: 0.3543594e-11, 18: -0.3770130e-13,
             20: 0.3547600e-15, 22: -0.2994113e-17, 24: 0.2291658e-19, 26: -0.1178252e-21}
_ROSSBY_V = {3: -0.6697824e-01, 5: -0.2266569e-02, 7: 0.9228703e-04, 9: -0.1954691e-05, 11: 0.2925271e-07,
             13: -0.3332983e-09, 15: 0.2916586e-11, 17: -0.1824357e-13, 19: 0.4920951e-16, 21: 0.6302640e-18,
             23: -0.1289167e-19, 25: 0.1471189e-21}
_ROSSBY_E = {0: -3.0714300e+00, 2: -0.3508384e-01, 4: -0.1861060e-01, 6: -0.2496364e-03, 8: 0.1639537e-04,
             10: -0.4410177e-06, 12: 0.8354759e-09, 14: -0.1254222e-09, 16: 0.1573519e-11, 18: -0.1702300e-13,
             20: 0.1621976e-15, 22: -0.1382304e-17, 24: 0.1066277e-19, 26: -0.1178252e-21}


def _hermite_series(coef, y):
    h = [np.ones_like(y), 2.0 * y]
    for i in range(2, 28):
        h.append(2.0 * y * h[i - 1] - 2.0 * (i - 1) * h[i - 2])
    return sum(c * h[i] for i, c in coef.items())


def rossby_soliton(x, y, time=0.0, order=1, amplitude=0.395):
    """(u, v, eta) of the asymptotic solution at (x, y, time)."""
    x, y = np.asarray(x, float), np.asarray(y, float)
    B = amplitude
    c = -1.0 / 3.0 - (0.395 * B * B if order == 1 else 0.0)
    xi = x - c * time
    psi = np.exp(-0.5 * y * y)
    phi = 0.771 * (B / np.cosh(B * xi)) ** 2
    dphi = -2.0 * B * phi * np.tanh(B * xi)
    u = phi * 0.25 * (-9.0 + 6.0 * y * y) * psi
    v = 2.0 * y * dphi * psi
    eta = phi * 0.25 * (3.0 + 6.0 * y * y) * psi
    if order == 1:
        C = -0.395 * B * B
        u = u + C * phi * 0.5625 * (3.0 + 2.0 * y * y) * psi + phi * phi * psi * _hermite_series(_ROSSBY_U, y)
        v = v + dphi * phi * psi * _hermite_series(_ROSSBY_V, y)
        eta = eta + C * phi * 0.5625 * (-5.0 + 2.0 * y * y) * psi + phi * phi * psi * _hermite_series(_ROSSBY_E, y)
    return u, v, eta


def rossby_mesh(refinement):
    """PeriodicRectangleMesh(2r, r, 48, 24, direction='x') shifted to be centred on the origin (:143-148)."""
    from thetis_b200.mesh import periodic_rectangle_mesh
    return periodic_rectangle_mesh(2 * refinement, refinement, 48.0, 24.0, direction="x", origin=(-24.0, -12.0))


def rossby_metrics(mesh, eta):
    """Relative peak heights and phase speeds (:193-213) from a P1DG elevation (nt, 3)."""
    xc = mesh.coords[mesh.cells]
    s = np.sign(xc[..., 1]) * eta
    i_n, i_s = np.unravel_index(np.argmax(s), s.shape), np.unravel_index(np.argmin(s), s.shape)
    h_n, h_s = s[i_n] / 0.1567020, s[i_s] / -0.1567020
    c_n, c_s = (48.0 - xc[i_n][0]) / 47.18, (48.0 - xc[i_s][0]) / 47.18
    return h_n, h_s, c_n, c_s


def rossby_check_convergence(metrics_by_refinement):
    """run_convergence's criterion (:239-257): 1 - |1 - m| must not decrease by more than 2 % between refinements."""
    for k in range(4):
        for i in range(1, len(metrics_by_refinement)):
            slope = (1 - abs(1 - metrics_by_refinement[i][k])) / (1 - abs(1 - metrics_by_refinement[i - 1][k]))
            assert slope > 1.0 - 0.02, ("h+", "h-", "c+", "c-")[k] + f" diverges: {slope}"


# ------------------------------------------------------------------------------------------------ steady-state basin MMS
MMS = dict(lx=15e3, ly=10e3, h0=10.0, f0=5e-3, nu0=100.0, g=9.81, t_end=1000.0)


@functools.lru_cache(None)
def mms_setup(name):
    """
    Analytic fields of setup7/8/9 and the sources that make them a steady solution of the nonlinear equations:
      res_elev = div(H u),   res_uv = u.grad(u) + f e_z x u + g grad(eta) - div(stress) - diag(grad(H)/H) diag(stress)
    (stress = nu (grad u + grad u^T), set-up 9 only).  Returns numpy callables f(x, y) and the boundary tags.
    """
    import sympy as sy
    x, y = sy.symbols("x y", real=True)
    lx, ly, h0, f0, nu0, g = (MMS[k] for k in ("lx", "ly", "h0", "f0", "nu0", "g"))
    pi = sy.pi
    bath = h0 * sy.sqrt(0.3 * x ** 2 + 0.2 * y ** 2 + 0.1) / lx + 4.0
    elev = sy.cos(pi * (3.0 * x + 1.0 * y) / lx)
    cori = visc = None
    if name == "setup7":
        u = sy.sin(pi * (-2.0 * x + 1.0 * y) / lx) * sy.sin(pi * y / ly)
        v = 0.5 * sy.sin(pi * x / lx) * sy.sin(pi * (-3.0 * x + 1.0 * y) / lx)
        cori = f0 * sy.cos(pi * (x + y) / lx)
        bnd = {1: ("elev", "flux_left"), 2: ("flux_right",), 3: ("elev", "flux_lower"), 4: ("un_upper",)}
    else:
        u = sy.sin(pi * (-2.0 * x + 1.0 * y) / lx)
        v = 0.5 * sy.sin(pi * (-3.0 * x + 1.0 * y) / lx)
        if name == "setup8":
            cori = f0 * sy.cos(pi * (x + y) / lx)
            bnd = {m: ("elev", "uv") for m in (1, 2, 3, 4)}
        elif name == "setup9":
            visc = nu0 * (1.0 + x / lx)
            bnd = {m: ("uv",) for m in (1, 2, 3, 4)}
        else:
            raise ValueError(name)
    H = bath + elev
    res_e = sy.diff(H * u, x) + sy.diff(H * v, y)
    f = cori if cori is not None else 0
    res_u = u * sy.diff(u, x) + v * sy.diff(u, y) - f * v + g * sy.diff(elev, x)
    res_v = u * sy.diff(v, x) + v * sy.diff(v, y) + f * u + g * sy.diff(elev, y)
    if visc is not None:
        sxx, syy = 2 * visc * sy.diff(u, x), 2 * visc * sy.diff(v, y)
        sxy = visc * (sy.diff(u, y) + sy.diff(v, x))
        gx, gy = sy.diff(H, x) / H, sy.diff(H, y) / H
        # NB the expressions the reference ships (:85-89) carry only the diagonal stress entries in the grad(H)/H term
        # although HorizontalViscosityTerm uses the full product (shallowwater_eq.py:611-612); the known answer is the
        # reference's source, so that is what is restated (the omitted part is ~5e-4 of the source).
        res_u += -(sy.diff(sxx, x) + sy.diff(sxy, y)) - gx * sxx
        res_v += -(sy.diff(sxy, x) + sy.diff(syy, y)) - gy * syy
    num = lambda e: sy.lambdify((x, y), e, "numpy")
    bc = lambda fn: (lambda X, Y: fn(np.asarray(X, float), np.asarray(Y, float)) + 0.0 * np.asarray(X, float))
    out = dict(bath=bc(num(bath)), elev=bc(num(elev)), u=bc(num(u)), v=bc(num(v)), res_elev=bc(num(res_e)),
               res_u=bc(num(res_u)), res_v=bc(num(res_v)), bnd=bnd,
               cori=None if cori is None else bc(num(cori)), visc=None if visc is None else bc(num(visc)),
               options={"use_grad_div_viscosity_term": True, "use_grad_depth_viscosity_term": True}
               if name == "setup9" else {})
    return out


def mms_problem(name, refinement):
    """
    Mesh, projected fields and boundary data of run() (test_steady_state_basin_mms.py:114-245) for dg-dg / SSPRK33:
    RectangleMesh(5r, 5r), dt = 4/r, bathymetry projected to P1, sources / Coriolis / boundary data to P1DG.
    """
    from thetis_b200.mesh import rectangle_mesh
    s = mms_setup(name)
    lx, ly = MMS["lx"], MMS["ly"]
    mesh = rectangle_mesh(5 * refinement, 5 * refinement, lx, ly)
    bath_v = project_cg1(mesh, s["bath"])                                   # (nv,)
    bath = bath_v[mesh.cells]
    elev = project_dg1(mesh, s["elev"])
    uv = project_dg1(mesh, lambda X, Y: np.stack([s["u"](X, Y), s["v"](X, Y)], -1))
    p = dict(mesh=mesh, dt=4.0 / refinement, bath_vertex=bath_v, bath=bath, elev=elev, uv=uv, setup=s,
             momentum_source=project_dg1(mesh, lambda X, Y: np.stack([s["res_u"](X, Y), s["res_v"](X, Y)], -1)),
             volume_source=project_dg1(mesh, s["res_elev"]),
             coriolis=None if s["cori"] is None else project_dg1(mesh, s["cori"]),
             viscosity_vertex=None if s["visc"] is None else project_cg1(mesh, s["visc"]))
    # un / flux boundary data (:190-203): P1DG projections of products of the projected fields
    un_x, un_y = uv[..., 0], uv[..., 1]                                     # projecting a P1DG function is the identity
    H = bath + elev
    flux_x = project_dg1_nodal_product(mesh, uv[..., 0], H, scale=ly)
    flux_y = project_dg1_nodal_product(mesh, uv[..., 1], H, scale=lx)
    mapping = {"elev": elev, "uv": uv, "un_left": -un_x, "un_right": un_x, "un_lower": -un_y, "un_upper": un_y,
               "flux_left": -flux_x, "flux_right": flux_x, "flux_lower": -flux_y, "flux_upper": flux_y}
    p["bnd"] = {m: {t.split("_")[0]: mapping[t] for t in tags} for m, tags in s["bnd"].items()}
    return p


def mms_errors(p, uv, eta):
    """elev and uv L2 errors / sqrt(area) against the analytic fields (:247-248)."""
    from oracle import swe_oracle as O
    s, mesh = p["setup"], p["mesh"]
    area = MMS["lx"] * MMS["ly"]
    ee = O.l2_error(mesh, eta, s["elev"])
    eu = np.hypot(O.l2_error(mesh, uv[..., 0], s["u"]), O.l2_error(mesh, uv[..., 1], s["v"]))
    return ee / np.sqrt(area), eu / np.sqrt(area)


def convergence_slope(ref_list, errs):
    from scipy import stats
    return stats.linregress(np.log10(1.0 / np.asarray(ref_list, float)), np.log10(np.asarray(errs))).slope


# ------------------------------------------------------------------------------------------------ tracer h-advection
HADV = dict(lx=15.0e3, depth=40.0, u=1.0, t_end=3000.0, x0=0.3 * 15.0e3, sigma=1600.0)


def hadv_mesh(refinement):
    from thetis_b200.mesh import rectangle_mesh
    return rectangle_mesh(6 * refinement + 1, 1, HADV["lx"], 6.0e3 / refinement)


def hadv_exact(t):
    return lambda X, Y: np.exp(-(X - HADV["x0"] - HADV["u"] * t) ** 2 / HADV["sigma"] ** 2)


def hadv_timestep(mesh):
    """The automatic time step the reference would use (solver2d.py:150-177,214-241) with horizontal_velocity_scale = |u|:
    uniform mesh and constant depth make both P1 projections exact."""
    h = np.sqrt(mesh.cell_area().min())
    return 0.05 * h / (np.sqrt(9.81 * HADV["depth"]) + abs(HADV["u"]))


# ------------------------------------------------------------------------------------------------ Thacker basin
# test/swe2d/test_thacker.py:40-96 (wetting-drying; the reference runs it with its implicit integrators only)
THACKER = dict(l_mesh=951646.46, D0=50.0, L=430620.0, eta0=2.0, t_end=43200.0,
               # max_err of test_thacker.py:17-26 for (n, stepper family)
               max_err={(10, "BackwardEuler"): 0.33, (25, "BackwardEuler"): 0.19, (10, "other"): 0.26, (25, "other"): 0.15})


def thacker_problem(n, alpha_max=None):
    """Mesh (`SquareMesh(n, n, l_mesh)`), nodal bathymetry, projected initial elevation and the P1 wetting-drying alpha
    of `use_automatic_wetting_and_drying_alpha` (solver2d.py:279-287: cell widths . |grad b|, utility.py:716-739,
    interpolated into P1 -- where cells meet at a vertex Firedrake keeps the value of whichever cell it visits last; the
    maximum over the adjacent cells is taken here).  ``alpha_max``: `wetting_and_drying_alpha_max` (options.py:897-902;
    the reference's default is Constant(2.0), which the test leaves in place; None = uncapped, 5 - 44 m on the 10 x 10
    mesh)."""
    from thetis_b200.mesh import rectangle_mesh
    T = THACKER
    l, D0, L, eta0 = T["l_mesh"], T["D0"], T["L"], T["eta0"]
    A = ((D0 + eta0) ** 2 - D0 ** 2) / ((D0 + eta0) ** 2 + D0 ** 2)
    x0 = y0 = l / 2
    r2 = lambda x, y: (x - x0) ** 2 + (y - y0) ** 2            # noqa: E731
    bath = lambda x, y: D0 * (1 - r2(x, y) / L ** 2)            # noqa: E731
    elev = lambda x, y: D0 * (np.sqrt(1 - A * A) / (1 - A) - 1 - r2(x, y) * ((1 + A) / (1 - A) - 1) / L ** 2)   # noqa: E731
    mesh = rectangle_mesh(n, n, l, l)
    bv = bath(mesh.coords[:, 0], mesh.coords[:, 1])
    xc = mesh.coords[mesh.cells]
    widths = np.ptp(xc, axis=1)                                                # (nt, 2)
    area2 = ((xc[:, 1, 0] - xc[:, 0, 0]) * (xc[:, 2, 1] - xc[:, 0, 1])
             - (xc[:, 1, 1] - xc[:, 0, 1]) * (xc[:, 2, 0] - xc[:, 0, 0]))
    # gradient of the P1 bathymetry per cell
    b = bv[mesh.cells]
    gx = (b[:, 0] * (xc[:, 1, 1] - xc[:, 2, 1]) + b[:, 1] * (xc[:, 2, 1] - xc[:, 0, 1]) + b[:, 2] * (xc[:, 0, 1] - xc[:, 1, 1])) / area2
    gy = (b[:, 0] * (xc[:, 2, 0] - xc[:, 1, 0]) + b[:, 1] * (xc[:, 0, 0] - xc[:, 2, 0]) + b[:, 2] * (xc[:, 1, 0] - xc[:, 0, 0])) / area2
    al_cell = widths[:, 0] * np.abs(gx) + widths[:, 1] * np.abs(gy)
    al_v = np.zeros(mesh.n_vertices)
    np.maximum.at(al_v, mesh.cells.reshape(-1), np.repeat(al_cell, 3))
    if alpha_max is not None:
        al_v = np.minimum(al_v, alpha_max)
    return dict(mesh=mesh, bath=bv[mesh.cells], alpha=al_v[mesh.cells], elev_init=elev, eta0=project_dg1(mesh, elev),
                centre=(x0, y0))


def thacker_error(p, eta, k=8):
    """test_thacker.py:83-93: mask = (1 - tanh((r - 420 km) / 1 km)) / 2, eta <- project(mask * eta),
    errornorm(mask * elev_init, eta) / l_mesh.  Integrals by a degree-10 Duffy rule on a k x k sub-triangulation of
    every cell (the 1 km wide mask edge is far below the cell size of the test's meshes)."""
    mesh = p["mesh"]
    lam6, w6 = duffy_rule(6)
    pts, wts = [], []
    for i in range(k):
        for j in range(k - i):
            tris = [np.array([[i, j], [i + 1, j], [i, j + 1]], float) / k]
            if j < k - i - 1:
                tris.append(np.array([[i + 1, j], [i + 1, j + 1], [i, j + 1]], float) / k)
            for t in tris:
                xy = lam6 @ t
                pts.append(np.stack([1 - xy[:, 0] - xy[:, 1], xy[:, 0], xy[:, 1]], 1))
                wts.append(w6 / k ** 2)
    lam, w = np.concatenate(pts), np.concatenate(wts)
    xc = mesh.coords[mesh.cells]
    xq = np.einsum("qa,cai->cqi", lam, xc)
    r = np.hypot(xq[..., 0] - p["centre"][0], xq[..., 1] - p["centre"][1])
    mask = 0.5 * (1 - np.tanh((r - 420000.0) / 1000.0))
    rhs = np.einsum("q,cq,qa->ca", w, mask * np.einsum("qa,ca->cq", lam, eta), lam)
    proj = np.einsum("ab,cb->ca", np.linalg.inv(_MREF), rhs)
    diff = mask * p["elev_init"](xq[..., 0], xq[..., 1]) - np.einsum("qa,ca->cq", lam, proj)
    return float(np.sqrt((mesh.cell_area()[:, None] * w[None] * diff ** 2).sum())) / THACKER["l_mesh"]
