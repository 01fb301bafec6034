"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  NOT A PRODUCT PATH.

Plain numpy restatement of the reference's explicit P1DG shallow-water path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm
may import this module; `thetis_b200/` never does.

What is restated (all paths relative to /root/reference):

* thetis/shallowwater_eq.py:335-393   ExternalPressureGradientTerm (dg branch)
* thetis/shallowwater_eq.py:396-450   HUDivTerm (by-parts branch)
* thetis/shallowwater_eq.py:453-510   HorizontalAdvectionTerm (+ Lax-Friedrichs)
* thetis/shallowwater_eq.py:619-663   CoriolisTerm, WindStressTerm, AtmosphericPressureTerm
* thetis/shallowwater_eq.py:666-701   QuadraticDragTerm (constant / Manning)
* thetis/shallowwater_eq.py:728-740   LinearDragTerm
* thetis/shallowwater_eq.py:704-726   BoundaryDragTerm ('drag' boundary tag)
* thetis/shallowwater_eq.py:834-850,917-920  BathymetryDisplacementMassTerm / the wetting-drying mass functional
                                      (displaced_mass, solve_displaced_mass, DisplacedMassShuOsherStepper)
* thetis/shallowwater_eq.py:513-616   HorizontalViscosityTerm (SIPG; grad-div and grad-depth variants)
* thetis/shallowwater_eq.py:794-831   MomentumSourceTerm, ContinuitySourceTerm
* thetis/shallowwater_eq.py:232-296   get_bnd_functions / impose_dynamic_bnd
* thetis/utility.py:936-996           DepthExpression
* thetis/equation.py:99-105           mass term
* thetis/rungekutta.py:13-87,326-347,870-952  Shu-Osher SSPRK33
* thetis/tracer_eq_2d.py:78-193,281-298       tracer advection + source
* thetis/tracer_eq_2d.py:196-278              HorizontalDiffusionTerm (SIPG)
* thetis/tracer_eq_2d.py:300-437              conservative tracer advection + source
* thetis/rungekutta.py:762-867                ERKGeneric (Butcher form)
* thetis/limiter.py:48-198 + firedrake.VertexBasedLimiter (recalled); the exterior-facet kernel :123-145 is pinned
  by executing the reference's kernel text (limiter_boundary_bounds)

The arithmetic itself lives in Firedrake/TSFC/PyOP2/PETSc, which is NOT under
/root/reference and is unpinned upstream (CI image
firedrakeproject/firedrake-vanilla-default:dev-main).  It cannot be imported
here, so this oracle restates the *published weak forms* with the same
quadrature the reference requests (`degree=2p+1=3`, shallowwater_eq.py:225-230):
2-point Gauss-Legendre on facets and a 6-point degree-3 rule in cells.

PARITY STATUS: (1) pinned FIELD BY FIELD against numbers produced by executing the reference's own source:
tests/golden/make_reference_residual_golden.py imports thetis/shallowwater_eq.py, tracer_eq_2d.py, equation.py,
utility.py, rungekutta.py and timeintegrator.py from the reference tree and runs their residual() / mass_term() /
advance() on a numpy stand-in for the UFL operators they use (tests/golden/ufl_lite.py; Firedrake is not installable
here); tests/test_oracle_reference_residuals.py: every SWE term, all open-boundary combinations, wetting-drying
depth, the three drag laws, SIPG viscosity, ModeSplit2DEquations, the tracer terms in both forms, and whole steps of
SSPRK33 / ERKLSPUM2 / ERKLPUM2 / ERKMidpoint / ERKEuler / ForwardEuler are reproduced to <= 2e-15 (tolerance 1e-12).
What that cannot pin is Firedrake's own assembly of those forms (quadrature rule choice for non-polynomial
integrands) and Firedrake's VertexBasedLimiter (Thetis' own exterior-facet kernel of the limiter,
limiter.py:123-145, IS pinned: its C text is compiled and executed, tests/test_oracle_reference_limiter.py).
(2) pinned against the reference's own known-answer criteria
(tests/test_oracle_kat.py: Shu-Osher coefficients produced by executing
rungekutta.py:13-87 itself, ODE convergence slope, eta-norm 6251.2574, standing
wave thresholds, limiter invariants, tracer conservation, atmospheric-pressure
criteria; tests/test_oracle_reference_kat.py: Rossby-soliton peak/phase criteria,
steady-state basin MMS order 2 with flux/un/elev/uv boundary data, tracer
h-advection slope -- their analytic fields pinned to values produced by executing
the reference's own test functions, tests/golden/reference_kat_fields.npz).  There are no stored field dumps in the
reference for this path and Firedrake cannot run here, so parity vs a live Firedrake run rests on (1) + (2); the
explicit wetting-drying STEP is "parity unpinned" (not a reference code path) -- its residual is pinned
by (1).

Deliberately written differently from the CUDA kernels: every form is
evaluated by quadrature with tabulated basis functions, interior facets are
visited once with '+'/'-' restrictions exactly like UFL, and the mass system
is *solved* (batched LU) instead of using a closed-form inverse.
"""
from __future__ import annotations

import numpy as np

FACET_NODES = np.array([[1, 2], [2, 0], [0, 1]], dtype=np.int64)

# ---------------------------------------------------------------- quadrature
# facet: 2-point Gauss-Legendre on [0, 1] (degree 3)
_GS = np.array([0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)])
_GW = np.array([0.5, 0.5])


def cell_quadrature(name="strang_fix6"):
    """
    Degree-3 cell rules in barycentric coordinates; weights sum to 1 (multiply
    by the cell area).  FIAT's default degree-3 triangle rule differs between
    FIAT versions; it only matters for non-polynomial
    integrands (Manning drag, wind stress / H, wetting-drying depth).
    """
    if name == "strang_fix6":
        a, b, c = 0.659027622374092, 0.231933368553031, 0.109039009072877
        pts = np.array([[a, b, c], [a, c, b], [b, a, c], [b, c, a], [c, a, b], [c, b, a]])
        # FIAT lists reference coordinates (x, y); barycentric = (1-x-y, x, y)
        lam = np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], axis=1)
        w = np.full(6, 1.0 / 6.0)
        return lam, w
    if name == "dunavant6":      # degree 4, 6 points
        a1, w1 = 0.445948490915965, 0.223381589678011
        a2, w2 = 0.091576213509771, 0.109951743655322
        lam = np.array([[1 - 2 * a1, a1, a1], [a1, 1 - 2 * a1, a1], [a1, a1, 1 - 2 * a1],
                        [1 - 2 * a2, a2, a2], [a2, 1 - 2 * a2, a2], [a2, a2, 1 - 2 * a2]])
        w = np.array([w1, w1, w1, w2, w2, w2])
        return lam, w / w.sum()
    if name == "collapsed_gauss4":   # FIAT 'canonical' 2x2 collapsed Gauss-Jacobi
        # Gauss-Legendre (2 pts) in eta1, Gauss-Jacobi(1,0) (2 pts) in eta2
        g = np.array([-1.0, 1.0]) / np.sqrt(3.0)
        gw = np.array([1.0, 1.0])
        # Gauss-Jacobi alpha=1, beta=0 two-point rule on [-1, 1]
        j = np.array([-0.6898979485566356, 0.2898979485566356])
        # weights of GJ(1,0) n=2: solve moments  int (1-x) dx = 2, int (1-x) x dx = -2/3
        A = np.array([[1.0, 1.0], [j[0], j[1]]])
        jw = np.linalg.solve(A, np.array([2.0, -2.0 / 3.0]))
        lam, w = [], []
        for a_, wa in zip(g, gw):
            for b_, wb in zip(j, jw):
                x = 0.25 * (1 + a_) * (1 - b_)
                y = 0.5 * (1 + b_)
                lam.append([1 - x - y, x, y])
                w.append(wa * wb * 0.125)
        lam = np.array(lam)
        w = np.array(w)
        return lam, w / w.sum()
    raise ValueError(name)


# ------------------------------------------------------- Shu-Osher (restated)
def butcher_to_shuosher_form(a, b):
    """Restates thetis/rungekutta.py:13-87 (explicit branch only)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    butcher = np.vstack((a, b))
    if np.diag(a).any():
        raise NotImplementedError("implicit tableaux are out of scope")
    aa = butcher[1:, :]
    be_0 = np.diag(np.diag(aa))
    n = aa.shape[0]
    al_0 = np.eye(n) - be_0 @ np.linalg.inv(aa)
    alpha = np.zeros((n + 1, n + 1))
    alpha[1:, 1:] = al_0
    alpha[:, 0] = 1.0 - alpha.sum(axis=1)
    beta = np.zeros((n + 1, n + 1))
    beta[1:, :-1] = be_0
    alpha[np.abs(alpha) < 1e-13] = 0.0
    beta[np.abs(beta) < 1e-13] = 0.0
    return alpha, beta


SSPRK33_A = [[0, 0, 0], [1.0, 0, 0], [0.25, 0.25, 0]]      # rungekutta.py:342-344
SSPRK33_B = [1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0]              # :345
SSPRK33_C = [0, 1.0, 0.5]                                  # :346


class _Geom:
    """Per-cell affine geometry from vertex coordinates."""

    def __init__(self, coords, cells):
        x = coords[cells]                                   # (nt, 3, 2)
        self.x = x
        d1 = x[:, 1] - x[:, 0]
        d2 = x[:, 2] - x[:, 0]
        self.area = 0.5 * (d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0])
        assert np.all(self.area > 0)
        p = x[:, FACET_NODES[:, 0]]
        q = x[:, FACET_NODES[:, 1]]
        e = q - p
        self.flen = np.hypot(e[..., 0], e[..., 1])          # (nt, 3)
        self.fnormal = np.stack([e[..., 1], -e[..., 0]], -1) / self.flen[..., None]   # unit outward
        # grad phi_a = - N_a |e_a| / (2 A)
        self.grad = -self.fnormal * (self.flen / (2.0 * self.area[:, None]))[..., None]  # (nt, 3, 2)


def _nodal(val, nt, ncomp=None):
    """Broadcast a constant or nodal array to a P1DG nodal array."""
    if ncomp is None:
        out = np.empty((nt, 3))
        out[...] = np.asarray(val, dtype=float) if np.ndim(val) else float(val)
        return out
    out = np.empty((nt, 3, ncomp))
    out[...] = np.asarray(val, dtype=float)
    return out


class SWEOracle:
    """
    Explicit residual of `ShallowWaterEquations` on P1DG-P1DG triangles.

    State layout: uv (nt, 3, 2), eta (nt, 3) -- nodal values per cell in the
    cell's local vertex order.

    :arg mesh: object with coords, cells (CCW), nbr, nbr_lf, bf_cell, bf_lf, bf_marker
    :arg bathymetry: constant or nodal (nt, 3)
    :kwarg options: dict with use_nonlinear_equations, use_lax_friedrichs_velocity,
        use_wetting_and_drying, wetting_and_drying_alpha, norm_smoother
    :kwarg fields: dict like solver2d.py:546-558 (values: None, constants, nodal arrays)
    :kwarg bnd_conditions: {marker: {'elev'|'uv'|'un'|'flux'|'drag': value}}
    """

    def __init__(self, mesh, bathymetry, options=None, fields=None, bnd_conditions=None,
                 g_grav=9.81, rho0=1000.0, cell_rule="strang_fix6"):
        self.mesh = mesh
        self.nt = mesh.cells.shape[0]
        self.geom = _Geom(mesh.coords, mesh.cells)
        self.bath = _nodal(bathymetry, self.nt)
        o = dict(use_nonlinear_equations=True, use_lax_friedrichs_velocity=True,
                 use_wetting_and_drying=False, wetting_and_drying_alpha=0.5,
                 norm_smoother=0.0, use_grad_div_viscosity_term=False,
                 use_grad_depth_viscosity_term=True, sipg_factor=1.0,
                 # False: ModeSplit2DEquations (shallowwater_eq.py:931-966): the term list has no
                 # HorizontalAdvectionTerm (nor drag / wind / viscosity terms) although the depth is nonlinear
                 include_momentum_advection=True)
        o.update(options or {})
        self.options = o
        self.fields = dict(fields or {})
        self.bnd = dict(bnd_conditions or {})
        self.g = float(g_grav)
        self.rho0 = float(rho0)
        self.lam, self.qw = cell_quadrature(cell_rule)
        c, f, n, g = mesh.interior_facets()
        self.if_p, self.if_fp, self.if_m, self.if_fm = (np.asarray(c, np.int64), np.asarray(f, np.int64),
                                                        np.asarray(n, np.int64), np.asarray(g, np.int64))
        self.boundary_len = mesh.boundary_length()
        # mass matrix per cell (equation.py:99-105): M_ab = int phi_a phi_b
        lam, w = self.lam, self.qw
        mref = np.einsum("q,qa,qb->ab", w, lam, lam)
        self.mass = self.geom.area[:, None, None] * mref[None]

    # ------------------------------------------------------------ depth (utility.py:975-996)
    def wd_bathymetry_displacement(self, b, eta, alpha=None):
        """utility.py:975-985; ``alpha``: wetting_and_drying_alpha at the evaluation points when it is a P1 field
        (solver2d.py:279-287), else the constant option."""
        if self.options["use_wetting_and_drying"]:
            H = b + eta
            al = self.options["wetting_and_drying_alpha"] if alpha is None else alpha
            return 0.5 * (np.sqrt(H ** 2 + np.asarray(al) ** 2) - H)
        return 0.0

    def total_depth(self, b, eta, alpha=None):
        if self.options["use_nonlinear_equations"]:
            return b + eta + self.wd_bathymetry_displacement(b, eta, alpha)
        return b + 0.0 * eta

    def _alpha_nodal(self):
        """(nt, 3) nodal values when wetting_and_drying_alpha is a spatially varying P1 field, else None."""
        al = self.options["wetting_and_drying_alpha"]
        if isinstance(al, np.ndarray) and al.ndim == 2:
            return al
        return None

    # ------------------------------------------------------------ helpers
    def _field(self, name, ncomp=None):
        v = self.fields.get(name)
        if v is None:
            return None
        return _nodal(v, self.nt, ncomp)

    @staticmethod
    def _at_cell_q(nodal, lam):
        """nodal (nt,3[,k]) -> values at quadrature points (nt, nq[,k])"""
        return np.einsum("qa,ca...->cq...", lam, nodal)

    @staticmethod
    def _facet_trace(nodal, cells, lf, s, reverse=False):
        """
        Trace of a P1DG nodal field on local facet `lf` of `cells` at facet
        parameters s (ngp,) measured from the facet's first node (or from its
        second node if reverse).  Returns (nf, ngp[,k]).
        """
        n0 = FACET_NODES[lf, 0]
        n1 = FACET_NODES[lf, 1]
        if reverse:
            n0, n1 = n1, n0
        v0 = nodal[cells, n0]
        v1 = nodal[cells, n1]
        if v0.ndim == 1:
            return v0[:, None] * (1 - s)[None, :] + v1[:, None] * s[None, :]
        return v0[:, None, :] * (1 - s)[None, :, None] + v1[:, None, :] * s[None, :, None]

    def _bc_value(self, val, cells, lf, ncomp=None):
        """Evaluate a boundary datum (constant or nodal DG array) at the facet Gauss points."""
        if isinstance(val, np.ndarray) and val.ndim >= 2 and val.shape[0] == self.nt:
            return self._facet_trace(val, cells, lf, _GS)
        nf = cells.shape[0]
        if ncomp is None:
            return np.full((nf, _GS.shape[0]), float(val))
        out = np.empty((nf, _GS.shape[0], ncomp))
        out[...] = np.asarray(val, dtype=float)
        return out

    def get_bnd_functions(self, eta_in, uv_in, marker, funcs, b, normal):
        """shallowwater_eq.py:232-272; everything evaluated at facet Gauss points."""
        bnd_len = self.boundary_len[marker]
        cells, lf = self._bf_sel[marker]
        aln = self._alpha_nodal()
        al_b = None if aln is None else self._facet_trace(aln, cells, lf, _GS)
        eta_ext = uv_ext = None
        if 'elev' in funcs and 'uv' in funcs:
            eta_ext = self._bc_value(funcs['elev'], cells, lf)
            uv_ext = self._bc_value(funcs['uv'], cells, lf, 2)
        elif 'elev' in funcs and 'un' in funcs:
            eta_ext = self._bc_value(funcs['elev'], cells, lf)
            uv_ext = self._bc_value(funcs['un'], cells, lf)[..., None] * normal
        elif 'elev' in funcs and 'flux' in funcs:
            eta_ext = self._bc_value(funcs['elev'], cells, lf)
            h_ext = self.total_depth(b, eta_ext, al_b)
            area = h_ext * bnd_len
            uv_ext = (self._bc_value(funcs['flux'], cells, lf) / area)[..., None] * normal
        elif 'elev' in funcs:
            eta_ext = self._bc_value(funcs['elev'], cells, lf)
            uv_ext = uv_in
        elif 'uv' in funcs:
            eta_ext = eta_in
            uv_ext = self._bc_value(funcs['uv'], cells, lf, 2)
        elif 'un' in funcs:
            eta_ext = eta_in
            uv_ext = self._bc_value(funcs['un'], cells, lf)[..., None] * normal
        elif 'flux' in funcs:
            eta_ext = eta_in
            h_ext = self.total_depth(b, eta_ext, al_b)
            area = h_ext * bnd_len
            uv_ext = (self._bc_value(funcs['flux'], cells, lf) / area)[..., None] * normal
        if eta_ext is None or uv_ext is None:
            raise Exception('Unsupported bnd type, one of "elev", "uv", "un", or "flux" must be defined')
        return eta_ext, uv_ext

    @staticmethod
    def impose_dynamic_bnd(funcs, marker):
        """shallowwater_eq.py:274-296"""
        open_tags = ['elev', 'uv', 'un', 'flux']
        all_tags = open_tags + ['drag']
        if funcs is None:
            return False
        for k in funcs.keys():
            if k not in all_tags:
                raise Exception(f'Invalid boundary tag "{k}" specified on boundary {marker}')
            if k in open_tags:
                return True
        return False

    # ------------------------------------------------------------ residual
    def residual(self, uv, eta):
        """
        Sum over all terms of `-f` tested against every basis function:
        returns (Ru (nt,3,2), Re (nt,3)) such that  M d(u,eta)/dt = (Ru, Re).
        In the explicit integrator solution_old == solution and
        fields_old == fields (rungekutta.py:901-904).
        """
        nt, g, geo = self.nt, self.g, self.geom
        o = self.options
        lam, qw = self.lam, self.qw
        nonlin = o["use_nonlinear_equations"]
        adv = nonlin and o["include_momentum_advection"]
        # ModeSplit2DEquations (include_momentum_advection=False) has no BoundaryDragTerm (shallowwater_eq.py:953-957)
        boundary_drag = bool(o["include_momentum_advection"])
        Ru = np.zeros((nt, 3, 2))
        Re = np.zeros((nt, 3))
        A = geo.area
        grad = geo.grad                                   # (nt, 3, 2)
        wq = A[:, None] * qw[None, :]                     # (nt, nq)
        phi = lam                                         # (nq, 3)

        # ---- cell integrals
        u_q = self._at_cell_q(uv, lam)                    # (nt, nq, 2)
        eta_q = self._at_cell_q(eta, lam)
        b_q = self._at_cell_q(self.bath, lam)
        aln = self._alpha_nodal()
        if aln is not None and self.fields.get("viscosity_h") is not None and o["use_grad_depth_viscosity_term"]:
            raise NotImplementedError("grad-depth viscosity term with a spatially varying wetting-drying alpha")
        H_q = self.total_depth(b_q, eta_q, None if aln is None else self._at_cell_q(aln, lam))
        # ExternalPressureGradient: f = -g*eta*div(psi) dx ; R = -f
        Ru += g * np.einsum("cq,cq,cai->cai", wq, eta_q, grad)
        # HUDiv: f = -inner(grad(phi), H*uv) dx
        Re += np.einsum("cq,cqi,cai->ca", wq, H_q[..., None] * u_q, grad)
        if adv:
            # HorizontalAdvection: f = -inner(div(outer(psi, uv_old)), uv) dx
            divu = np.einsum("cai,cai->c", grad, uv)      # constant per cell
            gu = np.einsum("cai,cqi->cqa", grad, u_q)     # grad(phi_a).u at q
            coef = gu + phi[None, :, :] * divu[:, None, None]       # (nt, nq, 3)
            Ru += np.einsum("cq,cqa,cqi->cai", wq, coef, u_q)
        cor = self._field("coriolis")
        if cor is not None:
            f_q = self._at_cell_q(cor, lam)
            # f = coriolis*(-uv[1]*psi[0] + uv[0]*psi[1]) dx ; R = -f
            Ru[..., 0] += np.einsum("cq,cq,qa->ca", wq, f_q * u_q[..., 1], phi)
            Ru[..., 1] -= np.einsum("cq,cq,qa->ca", wq, f_q * u_q[..., 0], phi)
        wind = self._field("wind_stress", 2)
        if wind is not None:
            t_q = self._at_cell_q(wind, lam)
            Ru += np.einsum("cq,cqi,qa->cai", wq, t_q / H_q[..., None] / self.rho0, phi)
        pa = self._field("atmospheric_pressure")
        if pa is not None:
            gp = np.einsum("ca,cai->ci", pa, grad)
            Ru -= np.einsum("cq,ci,qa->cai", wq, gp / self.rho0, phi)
        mann = self._field("manning_drag_coefficient")
        cd = self._field("quadratic_drag_coefficient")
        nik = self._field("nikuradse_bed_roughness")
        cd_q = None
        if nik is not None:
            # shallowwater_eq.py:689-697
            if mann is not None:
                raise Exception('Cannot set both Nikuradse drag and Manning drag parameter')
            if cd is not None:
                raise Exception('Cannot set both dimensionless and Nikuradse drag parameter')
            kappa = float(self.fields.get("von_karman", 0.4))
            ks_q = self._at_cell_q(nik, lam)
            with np.errstate(divide="ignore", invalid="ignore"):
                cd_n = 2 * kappa ** 2 / np.log(11.036 * H_q / ks_q) ** 2
            cd_q = np.where(H_q > ks_q, cd_n, 0.0)
        if mann is not None:
            if cd is not None:
                raise Exception('Cannot set both dimensionless and Manning drag parameter')
            cd_q = g * self._at_cell_q(mann, lam) ** 2 / H_q ** (1. / 3.)
        elif cd is not None and nik is None:
            cd_q = self._at_cell_q(cd, lam)
        if cd_q is not None:
            eps = float(o["norm_smoother"])
            mag = np.sqrt(u_q[..., 0] ** 2 + u_q[..., 1] ** 2 + eps ** 2)
            Ru -= np.einsum("cq,cq,cqi,qa->cai", wq, cd_q * mag / H_q, u_q, phi)
        lin = self._field("linear_drag_coefficient")
        if lin is not None:
            Ru -= np.einsum("cq,cq,cqi,qa->cai", wq, self._at_cell_q(lin, lam), u_q, phi)
        ms = self._field("momentum_source", 2)
        if ms is not None:
            Ru += np.einsum("cq,cqi,qa->cai", wq, self._at_cell_q(ms, lam), phi)
        vs = self._field("volume_source")
        if vs is not None:
            Re += np.einsum("cq,cq,qa->ca", wq, self._at_cell_q(vs, lam), phi)
        nu = self._field("viscosity_h")
        graddiv = bool(o["use_grad_div_viscosity_term"])
        if nu is not None:
            # HorizontalViscosityTerm cell part (shallowwater_eq.py:554-573, 611-614)
            G = np.einsum("cai,caj->cij", uv, grad)              # grad(uv)[i, j] = d u_i / d x_j, constant per cell
            S = G + np.swapaxes(G, 1, 2) if graddiv else G       # stress / nu
            nu_q = self._at_cell_q(nu, lam)
            # f = inner(grad(psi), stress) dx ; R = -f
            Ru -= np.einsum("cq,cq,caj,cij->cai", wq, nu_q, grad, S)
            if o["use_grad_depth_viscosity_term"]:
                # f += -dot(psi, dot(grad(total_h)/total_h, stress)) dx
                if nonlin:
                    ghl = np.einsum("ca,caj->cj", self.bath + eta, grad)          # grad(b + eta)
                    gH = np.broadcast_to(ghl[:, None, :], (nt, lam.shape[0], 2))
                    if o["use_wetting_and_drying"]:
                        hl_q = b_q + eta_q
                        al = np.asarray(o["wetting_and_drying_alpha"])
                        gH = gH * (0.5 * (1.0 + hl_q / np.sqrt(hl_q ** 2 + al ** 2)))[..., None]
                else:
                    gb = np.einsum("ca,caj->cj", self.bath, grad)
                    gH = np.broadcast_to(gb[:, None, :], (nt, lam.shape[0], 2))
                w_q = np.einsum("cqi,cq,cij->cqj", gH / H_q[..., None], nu_q, S)
                Ru += np.einsum("cq,cqj,qa->caj", wq, w_q, phi)

        # ---- interior facets ('+' = if_p, '-' = if_m), 2-point Gauss
        cp, fp, cm, fm = self.if_p, self.if_fp, self.if_m, self.if_fm
        if cp.size:
            s = _GS
            n_p = geo.fnormal[cp, fp][:, None, :]         # (nf, 1, 2)
            n_m = -n_p
            flen = geo.flen[cp, fp]
            wf = flen[:, None] * _GW[None, :]             # (nf, ngp)
            up = self._facet_trace(uv, cp, fp, s)
            um = self._facet_trace(uv, cm, fm, s, reverse=True)
            ep = self._facet_trace(eta, cp, fp, s)
            em = self._facet_trace(eta, cm, fm, s, reverse=True)
            bp = self._facet_trace(self.bath, cp, fp, s)
            bm = self._facet_trace(self.bath, cm, fm, s, reverse=True)
            al_f = None if aln is None else self._facet_trace(aln, cp, fp, s)       # alpha is continuous (P1)
            Hp = self.total_depth(bp, ep, al_f)
            Hm = self.total_depth(bm, em, al_f)
            # test functions: '+' nodes (n0: 1-s, n1: s); '-' nodes reversed
            php = np.stack([1 - s, s], -1)                # (ngp, 2) for nodes FACET_NODES[fp]
            nodes_p = FACET_NODES[fp]                     # (nf, 2)
            nodes_m = FACET_NODES[fm][:, ::-1]            # matching order
            h_av = 0.5 * (Hp + Hm)
            # --- PG: head_star = avg(head) + sqrt(avg(H)/g)*jump(uv, n)
            jump_un = np.einsum("fqi,fqi->fq", up, np.broadcast_to(n_p, up.shape)) \
                + np.einsum("fqi,fqi->fq", um, np.broadcast_to(n_m, um.shape))
            head_star = 0.5 * (ep + em) + np.sqrt(h_av / g) * jump_un
            # f += g*head_star*jump(psi, n) dS
            fu_p = g * head_star[..., None] * n_p         # tested with phi on '+'
            fu_m = g * head_star[..., None] * n_m
            # --- HUDiv: uv_rie = avg(uv) + sqrt(g/h)*jump(eta, n); hu_star = h*uv_rie
            jump_eta_n = ep[..., None] * n_p + em[..., None] * n_m
            uv_rie = 0.5 * (up + um) + np.sqrt(g / h_av)[..., None] * jump_eta_n
            hu_star = h_av[..., None] * uv_rie
            fe_p = np.einsum("fqi,fqi->fq", hu_star, np.broadcast_to(n_p, hu_star.shape))
            fe_m = np.einsum("fqi,fqi->fq", hu_star, np.broadcast_to(n_m, hu_star.shape))
            if adv:
                uv_avg = 0.5 * (up + um)
                un_av = np.einsum("fqi,fqi->fq", uv_avg, np.broadcast_to(n_m, uv_avg.shape))
                # inner(uv_avg, jump(outer(psi, uv_old), n))
                fu_p = fu_p + uv_avg * np.einsum("fqi,fqi->fq", up, np.broadcast_to(n_p, up.shape))[..., None]
                fu_m = fu_m + uv_avg * np.einsum("fqi,fqi->fq", um, np.broadcast_to(n_m, um.shape))[..., None]
                if o["use_lax_friedrichs_velocity"]:
                    sig = float(self.fields.get("lax_friedrichs_velocity_scaling_factor", 1.0))
                    gamma = 0.5 * np.abs(un_av) * sig
                    ju = up - um
                    fu_p = fu_p + gamma[..., None] * ju
                    fu_m = fu_m - gamma[..., None] * ju
            if nu is not None:
                # SIPG interior-facet terms (shallowwater_eq.py:575-590)
                sipg = float(o["sipg_factor"])
                cp_el = 3.0                                       # (p+1)(p+2)/2 for p = 1 triangles
                sig_p = sipg * cp_el * flen / A[cp]               # sigma = sipg*cp / (CellVolume/FacetArea)
                sig_m = sipg * cp_el * flen / A[cm]
                sig_max = np.maximum(sig_p, sig_m)[:, None]
                nu_p = self._facet_trace(nu, cp, fp, s)
                nu_m = self._facet_trace(nu, cm, fm, s, reverse=True)
                nu_av = 0.5 * (nu_p + nu_m)
                npb = np.broadcast_to(n_p, up.shape)
                J = np.einsum("fqi,fqj->fqij", up - um, npb)      # tensor_jump(uv, n)
                SJ = nu_av[..., None, None] * (J + np.swapaxes(J, 2, 3) if graddiv else J)
                avS = 0.5 * (nu_p[..., None, None] * S[cp][:, None] + nu_m[..., None, None] * S[cm][:, None])
                # + sigma_max*inner(tensor_jump(psi, n), stress_jump) - inner(tensor_jump(psi, n), avg(stress))
                t13 = np.einsum("fqij,fqj->fqi", sig_max[..., None, None] * SJ - avS, npb)
                fu_p = fu_p + t13
                fu_m = fu_m - t13
                # - inner(avg(grad(psi)), stress_jump): touches all three nodes of each side
                SJint = np.einsum("fq,fqij->fij", wf, SJ)
                np.add.at(Ru, cp, 0.5 * np.einsum("faj,fij->fai", grad[cp], SJint))
                np.add.at(Ru, cm, 0.5 * np.einsum("faj,fij->fai", grad[cm], SJint))
            # scatter: R -= int f * phi
            for k in range(2):
                np.subtract.at(Ru, (cp, nodes_p[:, k]), np.einsum("fq,fqi,q->fi", wf, fu_p, php[:, k]))
                np.subtract.at(Ru, (cm, nodes_m[:, k]), np.einsum("fq,fqi,q->fi", wf, fu_m, php[:, k]))
                np.subtract.at(Re, (cp, nodes_p[:, k]), np.einsum("fq,fq,q->f", wf, fe_p, php[:, k]))
                np.subtract.at(Re, (cm, nodes_m[:, k]), np.einsum("fq,fq,q->f", wf, fe_m, php[:, k]))

        # ---- exterior facets, one integral per marker
        mesh = self.mesh
        self._bf_sel = {}
        for marker in sorted(set(int(m) for m in np.unique(mesh.bf_marker))):
            sel = mesh.bf_marker == marker
            cells = mesh.bf_cell[sel].astype(np.int64)
            lf = mesh.bf_lf[sel].astype(np.int64)
            self._bf_sel[marker] = (cells, lf)
            funcs = self.bnd.get(marker)
            s = _GS
            n = geo.fnormal[cells, lf][:, None, :]
            nb = np.broadcast_to(n, (cells.shape[0], s.shape[0], 2))
            wf = geo.flen[cells, lf][:, None] * _GW[None, :]
            u = self._facet_trace(uv, cells, lf, s)
            e = self._facet_trace(eta, cells, lf, s)
            b = self._facet_trace(self.bath, cells, lf, s)
            al_b = None if aln is None else self._facet_trace(aln, cells, lf, s)
            H = self.total_depth(b, e, al_b)
            nodes = FACET_NODES[lf]
            php = np.stack([1 - s, s], -1)
            fu = np.zeros_like(u)
            fe = np.zeros_like(e)
            if self.impose_dynamic_bnd(funcs, marker):
                eta_ext, uv_ext = self.get_bnd_functions(e, u, marker, funcs, b, nb)
                # PG (shallowwater_eq.py:370-375)
                un_jump = np.einsum("fqi,fqi->fq", u - uv_ext, nb)
                eta_rie = 0.5 * (e + eta_ext) + np.sqrt(H / g) * un_jump
                fu = fu + g * eta_rie[..., None] * nb
                # HUDiv (:431-442)
                H_ext = self.total_depth(b, eta_ext, al_b)
                h_av = 0.5 * (H + H_ext)
                eta_jump = e - eta_ext
                un_rie = 0.5 * np.einsum("fqi,fqi->fq", u + uv_ext, nb) + np.sqrt(g / h_av) * eta_jump
                eta_rie2 = 0.5 * (e + eta_ext) + np.sqrt(h_av / g) * un_jump
                h_rie = self.total_depth(b, eta_rie2, al_b)
                fe = fe + h_rie * un_rie
                if adv:
                    # advection (:498-509)
                    un_rie_a = 0.5 * np.einsum("fqi,fqi->fq", u + uv_ext, nb) + np.sqrt(g / H) * eta_jump
                    uv_av = 0.5 * (uv_ext + u)
                    fu = fu + un_rie_a[..., None] * uv_av
            else:
                # land boundary (:376-381)
                un_jump = np.einsum("fqi,fqi->fq", u, nb)
                head_rie = e + np.sqrt(H / g) * un_jump
                fu = fu + g * head_rie[..., None] * nb
                if adv and o["use_lax_friedrichs_velocity"]:
                    # mirror velocity (:489-497)
                    sig = float(self.fields.get("lax_friedrichs_velocity_scaling_factor", 1.0))
                    uv_ext = u - 2 * un_jump[..., None] * nb
                    gamma = 0.5 * np.abs(un_jump) * sig
                    fu = fu + gamma[..., None] * (u - uv_ext)
            if nu is not None and self.impose_dynamic_bnd(funcs, marker):
                # Dirichlet bcs of the viscosity term (shallowwater_eq.py:592-609)
                delta = None
                if 'un' in funcs:
                    un_ext = self._bc_value(funcs['un'], cells, lf)
                    delta = (np.einsum("fqi,fqi->fq", u, nb) - un_ext)[..., None] * nb
                else:
                    eta_ext, uv_ext = self.get_bnd_functions(e, u, marker, funcs, b, nb)
                    if uv_ext is not u:
                        delta = u - uv_ext
                if delta is not None:
                    sipg = float(o["sipg_factor"])
                    sig = (sipg * 3.0 * geo.flen[cells, lf] / A[cells])[:, None]
                    nu_b = self._facet_trace(nu, cells, lf, s)
                    Jb = np.einsum("fqi,fqj->fqij", delta, nb)
                    SJb = nu_b[..., None, None] * (Jb + np.swapaxes(Jb, 2, 3) if graddiv else Jb)
                    Sb = nu_b[..., None, None] * S[cells][:, None]
                    fu = fu + np.einsum("fqij,fqj->fqi", sig[..., None, None] * SJb - Sb, nb)
                    SJint = np.einsum("fq,fqij->fij", wf, SJb)
                    np.add.at(Ru, cells, np.einsum("faj,fij->fai", grad[cells], SJint))
            if funcs is not None and 'drag' in funcs and boundary_drag:
                # BoundaryDragTerm (shallowwater_eq.py:704-726): C_D |u_t| u_t, u_t the tangential velocity
                cd = self._bc_value(funcs['drag'], cells, lf)
                ut = u - np.einsum("fqi,fqi->fq", u, nb)[..., None] * nb
                ut_mag = np.sqrt(np.einsum("fqi,fqi->fq", ut, ut))
                fu = fu + (cd * ut_mag)[..., None] * ut
            for k in range(2):
                np.subtract.at(Ru, (cells, nodes[:, k]), np.einsum("fq,fqi,q->fi", wf, fu, php[:, k]))
                np.subtract.at(Re, (cells, nodes[:, k]), np.einsum("fq,fq,q->f", wf, fe, php[:, k]))
        return Ru, Re

    # ------------------------------------------------------------ mass solve
    def solve_mass(self, Ru, Re):
        """`M k = R` (rungekutta.py:921-924; PETSc cg/bjacobi/ilu is an exact block solve)."""
        ku = np.linalg.solve(self.mass, Ru)
        ke = np.linalg.solve(self.mass, Re[..., None])[..., 0]
        return ku, ke

    def tendency(self, uv, eta, dt=1.0):
        Ru, Re = self.residual(uv, eta)
        return self.solve_mass(dt * Ru, dt * Re)

    # ------------------------------------------------------------ the reference's wetting-drying mass functional
    def displaced_mass(self, eta):
        """`ShallowWaterEquations.mass_term` of the elevation with wetting-drying (shallowwater_eq.py:917-920 =
        `Equation.mass_term` + `BathymetryDisplacementMassTerm`, :834-850): F(eta)_a = int (eta + f(b + eta)) phi_a dx,
        f = wd_bathymetry_displacement (utility.py:975-985), cell rule of degree 3 like every `self.dx` integral."""
        lam, qw = self.lam, self.qw
        aln = self._alpha_nodal()
        f = self.wd_bathymetry_displacement(self._at_cell_q(self.bath, lam), self._at_cell_q(eta, lam),
                                            None if aln is None else self._at_cell_q(aln, lam))
        F = np.einsum("cab,cb->ca", self.mass, eta)
        if self.options["use_wetting_and_drying"]:
            F = F + self.geom.area[:, None] * np.einsum("q,cq,qa->ca", qw, f, lam)
        return F

    def solve_displaced_mass(self, target, guess, tol=1e-12, max_it=60):
        """eta with displaced_mass(eta) = target (the functional is the gradient of a strictly convex potential:
        d/d eta (eta + f) = (1 + H / sqrt(H^2 + alpha^2)) / 2 in (0, 1), so the root is unique)."""
        shift = np.einsum("cab,cb->ca", self.mass, self.bath)       # eta + f = Ht - b:  F = depth_mass - M b
        return self.solve_depth_mass(target + shift, guess, tol, max_it)

    def _depth_and_slope(self, H, al):
        """total depth Ht = (sqrt(H^2 + alpha^2) + H) / 2 (utility.py:987-996) and dHt/dH, without cancellation on the
        dry side (H < 0: Ht = alpha^2 / (2 (r - H)))"""
        if not self.options["use_wetting_and_drying"]:
            return H, np.ones_like(H)
        aa = np.asarray(al, dtype=float) ** 2 + 0.0 * H
        r = np.sqrt(H * H + aa)
        wet = H >= 0.0
        d = np.where(wet, 1.0, r - H)
        with np.errstate(divide="ignore", invalid="ignore"):
            ht = np.where(wet, 0.5 * (r + H), 0.5 * aa / d)
            dht = np.where(wet, np.where(r > 0, 0.5 * (1.0 + H / np.where(r > 0, r, 1.0)), 0.5), 0.5 * aa / (d * np.where(r > 0, r, 1.0)))
        return ht, dht

    def depth_mass(self, eta):
        """S(eta)_a = int Ht(b + eta) phi_a dx = displaced_mass(eta) + int b phi_a: the same functional up to a constant,
        evaluated without the cancellation eta + f suffers from in almost dry cells"""
        lam, qw = self.lam, self.qw
        aln = self._alpha_nodal()
        al = self.options["wetting_and_drying_alpha"] if aln is None else self._at_cell_q(aln, lam)
        ht, _ = self._depth_and_slope(self._at_cell_q(self.bath, lam) + self._at_cell_q(eta, lam), al)
        return self.geom.area[:, None] * np.einsum("q,cq,qa->ca", qw, ht, lam)

    def solve_depth_mass(self, target, guess, tol=1e-12, max_it=60):
        """eta with depth_mass(eta) = target: cell-local Newton iteration with step halving (a step that increases the
        residual of its cell -- a jump across the kink of Ht from the flat, dry side -- is halved)."""
        lam, qw = self.lam, self.qw
        aln = self._alpha_nodal()
        al = self.options["wetting_and_drying_alpha"] if aln is None else self._at_cell_q(aln, lam)
        b_q = self._at_cell_q(self.bath, lam)
        area = self.geom.area
        e = np.array(guess, dtype=float, copy=True)
        base, step = e.copy(), np.zeros_like(e)
        gprev = np.full(e.shape[0], np.inf)
        prev = np.inf
        for _ in range(max_it):
            ht, dht = self._depth_and_slope(b_q + self._at_cell_q(e, lam), al)
            G = area[:, None] * np.einsum("q,cq,qa->ca", qw, ht, lam) - target
            gn = np.abs(G).max(axis=1)
            noise = 4e-15 * (np.abs(G + target).sum(axis=1) + np.abs(target).sum(axis=1))     # rounding level of S - T
            if _ > 0 and (gn <= noise).all():
                return e
            rej = (gn > gprev) & (gn > noise)
            if rej.any():
                step[rej] *= 0.5
                e[rej] = base[rej] - step[rej]
                if rej.all():
                    continue
            acc = ~rej
            J = area[:, None, None] * np.einsum("q,cq,qa,qb->cab", qw, dht, lam, lam)
            try:
                d = np.linalg.solve(J[acc], G[acc][..., None])[..., 0]
            except np.linalg.LinAlgError as err:      # an iterate ran off (target no elevation can meet, non-finite state)
                raise RuntimeError("displaced-mass Newton iteration: " + str(err))
            gprev[acc] = gn[acc]
            base[acc] = e[acc]
            step[acc] = d
            e[acc] = base[acc] - d
            dm, scale = np.abs(d).max() if d.size else 0.0, max(1.0, np.abs(e).max())
            if not np.isfinite(dm):
                break
            if not rej.any() and (dm <= tol * scale or (dm >= 0.5 * prev and dm <= 1e-8 * scale)):
                return e
            if not rej.any():
                prev = dm
        raise RuntimeError("displaced-mass Newton iteration did not converge")


class ShuOsherStepper:
    """
    `ERKGenericShuOsher` (rungekutta.py:870-952) over any object exposing
    ``tendency(*state, dt) -> tuple of arrays``.  ``state`` is a list of arrays
    updated in place (like `solution.assign`).
    """

    def __init__(self, rhs, state, dt, a=SSPRK33_A, b=SSPRK33_B, c=SSPRK33_C):
        self.rhs = rhs
        self.state = state
        self.dt = float(dt)
        self.a = np.array(a, dtype=float)
        self.b = np.array(b, dtype=float)
        self.c = np.array(c, dtype=float)
        self.n_stages = len(self.b)
        self.alpha, self.beta = butcher_to_shuosher_form(self.a, self.b)
        self.stage_sol = [[np.zeros_like(s) for s in state] for _ in range(self.n_stages)]
        self.cfl_coeff = 1.0

    def set_dt(self, dt):
        self.dt = float(dt)

    def solve_stage(self, i, t, update_forcings=None):
        if update_forcings is not None:
            update_forcings(t + self.c[i] * self.dt)
        if i == 0:
            for d, s in zip(self.stage_sol[0], self.state):
                d[...] = s
        k = self.rhs.tendency(*self.state, dt=self.dt)
        for j, s in enumerate(self.state):
            new = self.beta[i + 1][i] * k[j]
            for jj in range(i + 1):
                if self.alpha[i + 1][jj] != 0.0:
                    new = new + self.alpha[i + 1][jj] * self.stage_sol[jj][j]
            s[...] = new
        if i < self.n_stages - 1:
            for d, s in zip(self.stage_sol[i + 1], self.state):
                d[...] = s

    def advance(self, t, update_forcings=None):
        for i in range(self.n_stages):
            self.solve_stage(i, t, update_forcings)


class DisplacedMassShuOsherStepper(ShuOsherStepper):
    """
    Shu-Osher stepper that advances the reference's OWN wetting-drying mass functional (shallowwater_eq.py:917-920):
        F(eta^(i+1)) = sum_j alpha_ij F(eta^(j)) + beta_i dt R_eta(u^(i)),   F = SWEOracle.displaced_mass,
    solved cell by cell for eta^(i+1); the velocity keeps the plain P1DG mass.  This is what `d/dt mass_term(u) = R(u)`
    (the equation every implicit integrator of the reference solves with wetting-drying on, e.g. test_thacker.py)
    becomes under an explicit Shu-Osher scheme.  `ERKGenericShuOsher` itself cannot be run with wetting-drying in the
    reference -- its `LinearVariationalProblem` puts the TrialFunction into this nonlinear term -- so
    neither this stepper nor `ShuOsherStepper` with wetting-drying (plain mass: the extension the CUDA path
    implements) has a reference counterpart; tests/test_oracle_reference_kat.py compares the two
    on the Thacker basin of test/swe2d/test_thacker.py.
    """

    def solve_stage(self, i, t, update_forcings=None):
        if update_forcings is not None:
            update_forcings(t + self.c[i] * self.dt)
        orc = self.rhs
        uv, eta = self.state
        if i == 0:
            self.stage_sol[0][0][...] = uv
            self.stage_sol[0][1][...] = eta
            self._F = [orc.depth_mass(eta)]       # = displaced_mass + const: the constant cancels (sum alpha = 1)
        Ru, Re = orc.residual(uv, eta)
        ku = np.linalg.solve(orc.mass, self.dt * Ru)
        new_uv = self.beta[i + 1][i] * ku
        target = self.beta[i + 1][i] * self.dt * Re
        for j in range(i + 1):
            if self.alpha[i + 1][j] != 0.0:
                new_uv = new_uv + self.alpha[i + 1][j] * self.stage_sol[j][0]
                target = target + self.alpha[i + 1][j] * self._F[j]
        eta[...] = orc.solve_depth_mass(target, eta)
        uv[...] = new_uv
        if i < self.n_stages - 1:
            self.stage_sol[i + 1][0][...] = uv
            self.stage_sol[i + 1][1][...] = eta
            self._F.append(orc.depth_mass(eta))


class ButcherStepper:
    """
    `ERKGeneric` (rungekutta.py:762-867): explicit Runge-Kutta in Butcher form.
    Stage i: solution = solution_old + sum_j a[i][j]*k_j (update_solution, :816-828), forcings at t + c_i*dt and
    k_i = dt*M^-1 R(solution) (solve_tendency, :831-838); after the last stage
    solution = solution_old + sum_j b[j]*k_j (get_final_solution, :841-852).
    """

    def __init__(self, rhs, state, dt, a, b, c, cfl_coeff=1.0):
        self.rhs = rhs
        self.state = state
        self.dt = float(dt)
        self.a = [list(map(float, row)) for row in a]
        self.b = list(map(float, b))
        self.c = list(map(float, c))
        self.n_stages = len(self.b)
        self.cfl_coeff = cfl_coeff
        self.solution_old = [np.array(s, copy=True) for s in state]
        self.tendency = [None] * self.n_stages

    def set_dt(self, dt):
        self.dt = float(dt)

    def solve_stage(self, i, t, update_forcings=None):
        for j, s in enumerate(self.state):
            new = self.solution_old[j].copy()
            for jj in range(i):
                if self.a[i][jj] != 0.0:
                    new = new + self.a[i][jj] * self.tendency[jj][j]
            s[...] = new
        if update_forcings is not None:
            update_forcings(t + self.c[i] * self.dt)
        self.tendency[i] = self.rhs.tendency(*self.state, dt=self.dt)

    def advance(self, t, update_forcings=None):
        for i in range(self.n_stages):
            self.solve_stage(i, t, update_forcings)
        for j, s in enumerate(self.state):
            new = self.solution_old[j].copy()
            for jj in range(self.n_stages):
                if self.b[jj] != 0.0:
                    new = new + self.b[jj] * self.tendency[jj][j]
            s[...] = new
            self.solution_old[j][...] = new


ERK_TABLEAUX = {
    # name: (a, b, c, cfl_coeff)   rungekutta.py:142-149, 350-392
    "ERKEuler": ([[0]], [1.0], [0], 1.0),
    "ERKLSPUM2": ([[0, 0, 0], [5.0 / 6.0, 0, 0], [11.0 / 24.0, 11.0 / 24.0, 0]],
                  [24.0 / 55.0, 1.0 / 5.0, 4.0 / 11.0], [0, 5.0 / 6.0, 11.0 / 12.0], 1.2),
    "ERKLPUM2": ([[0, 0, 0], [0.5, 0, 0], [0.5, 0.5, 0]], [1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0], [0, 0.5, 1.0], 2.0),
    "ERKMidpoint": ([[0.0, 0.0], [0.5, 0.0]], [0.0, 1.0], [0.0, 0.5], 1.0),
}


class TracerOracle:
    """
    2-D tracer advection + SIPG diffusion + source on P1DG, non-conservative
    (tracer_eq_2d.py:124-298) or conservative form (``use_conservative_form``,
    tracer_eq_2d.py:323-437: the unknown is the depth-integrated tracer).
    ``uv`` (nt,3,2) and ``elev`` (nt,3) are the live SWE fields (frozen during
    the tracer stages, coupled_timeintegrator_2d.py:99-101).

    fields: 'source', 'diffusivity_h' (constant or nodal (nt,3)),
    'tracer_advective_velocity_factor', 'lax_friedrichs_tracer_scaling_factor'.
    bnd tags: 'value', 'uv', 'un', 'flux', 'elev', 'diff_flux'.
    """

    def __init__(self, swe: SWEOracle, bnd_conditions=None, fields=None, options=None):
        self.swe = swe
        self.mesh = swe.mesh
        self.nt = swe.nt
        self.geom = swe.geom
        self.bnd = dict(bnd_conditions or {})
        self.fields = dict(fields or {})
        o = dict(use_lax_friedrichs_tracer=False, use_conservative_form=False, sipg_factor_tracer=1.0)
        o.update(options or {})
        self.options = o
        self.uv = None
        self.elev = None
        self.mass = swe.mass

    def set_velocity(self, uv, elev):
        self.uv, self.elev = uv, elev

    def residual(self, c):
        swe, geo, nt = self.swe, self.geom, self.nt
        lam, qw = swe.lam, swe.qw
        corr = float(self.fields.get("tracer_advective_velocity_factor", 1.0))
        uv = corr * self.uv
        R = np.zeros((nt, 3))
        wq = geo.area[:, None] * qw[None, :]
        grad = geo.grad
        cons = bool(self.options["use_conservative_form"])
        u_q = swe._at_cell_q(uv, lam)
        c_q = swe._at_cell_q(c, lam)
        gu = np.einsum("cai,cqi->cqa", grad, u_q)
        if cons:
            # f = -(Dx(test,0)*uv[0]*c + Dx(test,1)*uv[1]*c) dx (tracer_eq_2d.py:356-357) ; R = -f
            R += np.einsum("cq,cqa,cq->ca", wq, gu, c_q)
        else:
            divu = np.einsum("cai,cai->c", grad, uv)
            # f = -(Dx(uv[0]*test,0)*c + Dx(uv[1]*test,1)*c) dx ; R = -f
            coef = gu + lam[None] * divu[:, None, None]
            R += np.einsum("cq,cqa,cq->ca", wq, coef, c_q)
        src = self.fields.get("source")
        if src is not None:
            s_q = swe._at_cell_q(_nodal(src, nt), lam)
            if cons:
                # ConservativeSourceTerm (:429-437): H*source
                s_q = s_q * swe.total_depth(swe._at_cell_q(swe.bath, lam), swe._at_cell_q(self.elev, lam))
            R += np.einsum("cq,cq,qa->ca", wq, s_q, lam)
        mu = self.fields.get("diffusivity_h")
        if mu is not None:
            mu = _nodal(mu, nt)
            gc = np.einsum("ca,caj->cj", c, grad)                 # grad(c), constant per cell
            # HorizontalDiffusionTerm cell part (:238): f = inner(grad(test), mu*grad(c)) dx
            R -= np.einsum("cq,cq,caj,cj->ca", wq, swe._at_cell_q(mu, lam), grad, gc)
        # interior facets
        cp, fp, cm, fm = swe.if_p, swe.if_fp, swe.if_m, swe.if_fm
        s = _GS
        php = np.stack([1 - s, s], -1)
        if cp.size:
            n_p = np.broadcast_to(geo.fnormal[cp, fp][:, None, :], (cp.shape[0], 2, 2))
            n_m = -n_p
            wf = geo.flen[cp, fp][:, None] * _GW[None, :]
            up = swe._facet_trace(uv, cp, fp, s)
            um = swe._facet_trace(uv, cm, fm, s, reverse=True)
            c_p = swe._facet_trace(c, cp, fp, s)
            c_m = swe._facet_trace(c, cm, fm, s, reverse=True)
            uv_av = 0.5 * (up + um)
            un_av = np.einsum("fqi,fqi->fq", uv_av, n_m)
            sg = 0.5 * (np.sign(un_av) + 1.0)
            if cons:
                # flux_up = c('-')*uv('-')*s + c('+')*uv('+')*(1-s)   (:364-367)
                flux_up = (c_m * sg)[..., None] * um + (c_p * (1 - sg))[..., None] * up
                f_p = np.einsum("fqi,fqi->fq", flux_up, n_p)
                f_m = np.einsum("fqi,fqi->fq", flux_up, n_m)
            else:
                c_up = c_m * sg + c_p * (1 - sg)
                f_p = c_up * np.einsum("fqi,fqi->fq", up, n_p)
                f_m = c_up * np.einsum("fqi,fqi->fq", um, n_m)
            if self.options["use_lax_friedrichs_tracer"]:
                lf = float(self.fields.get("lax_friedrichs_tracer_scaling_factor", 1.0))
                gamma = 0.5 * np.abs(un_av) * lf
                f_p = f_p + gamma * (c_p - c_m)
                f_m = f_m - gamma * (c_p - c_m)
            if mu is not None:
                # SIPG interior-facet terms (:240-258)
                sipg = float(self.options["sipg_factor_tracer"])
                flen = geo.flen[cp, fp]
                sig_max = np.maximum(sipg * 3.0 * flen / geo.area[cp], sipg * 3.0 * flen / geo.area[cm])[:, None]
                mu_p = swe._facet_trace(mu, cp, fp, s)
                mu_m = swe._facet_trace(mu, cm, fm, s, reverse=True)
                dc = c_p - c_m                                    # jump(c, n) = dc * n('+')
                av_flux = 0.5 * (mu_p[..., None] * gc[cp][:, None, :] + mu_m[..., None] * gc[cm][:, None, :])
                t = sig_max * 0.5 * (mu_p + mu_m) * dc - np.einsum("fqi,fqi->fq", av_flux, n_p)
                f_p = f_p + t
                f_m = f_m - t
                # -inner(avg(mu*grad(test)), jump(c, n)): all three nodes of each side
                gn_p = np.einsum("faj,fj->fa", grad[cp], n_p[:, 0, :])
                gn_m = np.einsum("faj,fj->fa", grad[cm], n_p[:, 0, :])
                np.add.at(R, cp, 0.5 * gn_p * np.einsum("fq,fq->f", wf, mu_p * dc)[:, None])
                np.add.at(R, cm, 0.5 * gn_m * np.einsum("fq,fq->f", wf, mu_m * dc)[:, None])
            nodes_p = FACET_NODES[fp]
            nodes_m = FACET_NODES[fm][:, ::-1]
            for k in range(2):
                np.subtract.at(R, (cp, nodes_p[:, k]), np.einsum("fq,fq,q->f", wf, f_p, php[:, k]))
                np.subtract.at(R, (cm, nodes_m[:, k]), np.einsum("fq,fq,q->f", wf, f_m, php[:, k]))
        # exterior facets
        mesh = self.mesh
        for marker in sorted(set(int(m) for m in np.unique(mesh.bf_marker))):
            sel = mesh.bf_marker == marker
            cells = mesh.bf_cell[sel].astype(np.int64)
            lf = mesh.bf_lf[sel].astype(np.int64)
            funcs = self.bnd.get(marker)
            n = np.broadcast_to(geo.fnormal[cells, lf][:, None, :], (cells.shape[0], 2, 2))
            wf = geo.flen[cells, lf][:, None] * _GW[None, :]
            u = swe._facet_trace(uv, cells, lf, s)
            cin = swe._facet_trace(c, cells, lf, s)
            if funcs is not None:
                swe._bf_sel = getattr(swe, "_bf_sel", {})
                swe._bf_sel[marker] = (cells, lf)
                c_ext = swe._bc_value(funcs['value'], cells, lf) if 'value' in funcs else cin
                if 'uv' in funcs:
                    uv_ext = corr * swe._bc_value(funcs['uv'], cells, lf, 2)
                elif 'flux' in funcs:
                    e_in = swe._facet_trace(self.elev, cells, lf, s)
                    e_ext = swe._bc_value(funcs['elev'], cells, lf) if 'elev' in funcs else e_in
                    b = swe._facet_trace(swe.bath, cells, lf, s)
                    h_ext = swe.total_depth(b, e_ext)
                    area = h_ext * swe.boundary_len[marker]
                    uv_ext = (corr * swe._bc_value(funcs['flux'], cells, lf) / area)[..., None] * n
                elif 'un' in funcs:
                    uv_ext = swe._bc_value(funcs['un'], cells, lf)[..., None] * n
                else:
                    uv_ext = u
                uv_av = 0.5 * (u + uv_ext)
                un_av = np.einsum("fqi,fqi->fq", uv_av, n)
                sg = 0.5 * (np.sign(un_av) + 1.0)
                if cons:
                    # flux_up = c_in*uv*s + c_ext*uv_ext*(1-s)   (:391-394)
                    f = np.einsum("fqi,fqi->fq", (cin * sg)[..., None] * u + (c_ext * (1 - sg))[..., None] * uv_ext, n)
                else:
                    c_up = cin * sg + c_ext * (1 - sg)
                    f = c_up * un_av
                if mu is not None:
                    mu_b = swe._facet_trace(mu, cells, lf, s)
                    if 'diff_flux' in funcs:
                        # f += -test*diff_flux ds   (:264-265)
                        f = f - swe._bc_value(funcs['diff_flux'], cells, lf)
                    else:
                        # f += -test*dot(mu*grad(c_up), n) ds  (:267-272); grad(c_ext) = 0 for a Constant 'value',
                        # d(sign)/dx = 0 in UFL
                        if 'value' in funcs:
                            if isinstance(funcs['value'], np.ndarray) and np.ndim(funcs['value']) >= 2:
                                raise NotImplementedError("diffusive boundary flux with a spatially varying 'value'")
                            wgt = sg
                        else:
                            wgt = np.ones_like(sg)
                        f = f - wgt * mu_b * np.einsum("fj,fqj->fq", gc[cells], n)
            else:
                f = cin * np.einsum("fqi,fqi->fq", u, n)
            nodes = FACET_NODES[lf]
            for k in range(2):
                np.subtract.at(R, (cells, nodes[:, k]), np.einsum("fq,fq,q->f", wf, f, php[:, k]))
        return R

    def tendency(self, c, dt=1.0):
        R = self.residual(c)
        return (np.linalg.solve(self.mass, (dt * R)[..., None])[..., 0],)


def limiter_boundary_bounds(mesh, q, qmax, qmin):
    """Thetis' own addition to the vertex bounds (limiter.py:109-145, the `my_kernel` par_loop over the exterior
    facets): the arithmetic mean of the two nodal values on every exterior facet enters the max / min of the facet's two
    vertices.  In place on qmax / qmin (one value per topological vertex).  Pinned to the reference's kernel text,
    compiled and executed: tests/golden/make_reference_limiter_golden.py."""
    tv = mesh.topo[mesh.cells]
    bc = mesh.bf_cell.astype(np.int64)
    bl = mesh.bf_lf.astype(np.int64)
    n0 = FACET_NODES[bl, 0]
    n1 = FACET_NODES[bl, 1]
    face_mean = (q[bc, n0] + q[bc, n1]) / 2
    for nn in (n0, n1):
        np.maximum.at(qmax, tv[bc, nn], face_mean)
        np.minimum.at(qmin, tv[bc, nn], face_mean)


def vertex_based_limiter(mesh, q):
    """
    `VertexBasedP1DGLimiter.apply` for a scalar P1DG field on a 2-D mesh
    (limiter.py:100-145,182-198 + Firedrake's VertexBasedLimiter kernels,
    recalled): returns the limited nodal array (nt, 3).
    """
    q = np.array(q, dtype=float)
    nt = q.shape[0]
    tv = mesh.topo[mesh.cells]                             # topological vertex of each node
    nv = int(mesh.topo.max()) + 1
    # centroids: P0 projection = mean of the nodal values (limiter.py:90-97)
    qbar = q.mean(axis=1)
    qmax = np.full(nv, -1.0e10)
    qmin = np.full(nv, 1.0e10)
    for a in range(3):
        np.maximum.at(qmax, tv[:, a], qbar)
        np.minimum.at(qmin, tv[:, a], qbar)
    limiter_boundary_bounds(mesh, q, qmax, qmin)
    # limit
    alpha = np.ones(nt)
    for a in range(3):
        qa = q[:, a]
        with np.errstate(divide="ignore", invalid="ignore"):
            a1 = np.minimum(alpha, np.minimum(1.0, (qmax[tv[:, a]] - qbar) / (qa - qbar)))
            a2 = np.minimum(alpha, np.minimum(1.0, (qbar - qmin[tv[:, a]]) / (qbar - qa)))
        alpha = np.where(qa > qbar, a1, np.where(qa < qbar, a2, alpha))
    return qbar[:, None] + alpha[:, None] * (q - qbar[:, None])


# ------------------------------------------------------------------ utilities
def l2_norm(mesh, nodal, lam_w=None):
    """sqrt(int f^2 dx) of a P1DG nodal field (scalar (nt,3) or vector (nt,3,k))."""
    geo = _Geom(mesh.coords, mesh.cells)
    mref = (np.ones((3, 3)) + np.eye(3)) / 12.0
    if nodal.ndim == 2:
        v = np.einsum("ca,ab,cb->c", nodal, mref, nodal)
    else:
        v = np.einsum("cai,ab,cbi->c", nodal, mref, nodal)
    return float(np.sqrt((geo.area * v).sum()))


def interpolate(mesh, fn):
    """P1DG nodal interpolation of fn(x, y) -> (nt, 3[,k])."""
    x = mesh.coords[mesh.cells]
    return np.asarray(fn(x[..., 0], x[..., 1]))


def l2_error(mesh, nodal, fn, degree_rule="dunavant6"):
    """L2 error of a P1DG field against an analytic function (quadrature)."""
    lam, w = cell_quadrature(degree_rule)
    geo = _Geom(mesh.coords, mesh.cells)
    xq = np.einsum("qa,cai->cqi", lam, geo.x)
    vq = np.einsum("qa,ca->cq", lam, nodal)
    ex = fn(xq[..., 0], xq[..., 1])
    return float(np.sqrt((geo.area[:, None] * w[None] * (vq - ex) ** 2).sum()))
