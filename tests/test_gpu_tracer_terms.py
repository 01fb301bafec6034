"""
GPU parity of the tracer stage kernel's SURVEY 8f terms against the numpy oracle, through the C-ABI:
SIPG `HorizontalDiffusionTerm` (tracer_eq_2d.py:226-278), the conservative form
(tracer_eq_2d.py:323-437), the Butcher-form stage combination (tb_lincomb) and the
device diagnostics (tb_tracer_integrals / tb_swe_integrals).  fp64, tolerance relative
to the max-norm of the tendency: 1e-12 (polynomial integrands), 1e-11 with sqrt terms.
"""
import numpy as np
import pytest

from thetis_b200.mesh import rectangle_mesh, periodic_rectangle_mesh, delaunay_mesh, sfc_renumber
from oracle import swe_oracle as O

pytestmark = pytest.mark.gpu


def _fields(mesh, seed):
    rng = np.random.default_rng(seed)
    x = mesh.coords[mesh.cells]
    L = np.ptp(mesh.coords, axis=0).max()
    k = 2 * np.pi / L
    u = 0.4 * np.sin(k * x[..., 0]) * np.cos(k * x[..., 1]) + 0.03 * rng.standard_normal(x.shape[:2])
    v = -0.3 * np.cos(k * x[..., 0]) * np.sin(2 * k * x[..., 1]) + 0.03 * rng.standard_normal(x.shape[:2])
    e = 0.2 * np.cos(k * x[..., 0]) + 0.01 * rng.standard_normal(x.shape[:2])
    c = 2.0 + np.sin(2 * k * x[..., 0]) * np.cos(k * x[..., 1]) + 0.1 * rng.standard_normal(x.shape[:2])
    return np.stack([u, v], -1), e, c


def _run(mesh, bath_v, swe_opts, tr_opts, tr_fields_v, bnd, tol=1e-12, seed=0, corr=1.0):
    import thetis_b200._lib as L
    from thetis_b200.engine import Engine
    uv, eta, c = _fields(mesh, seed)
    cells = mesh.cells
    is_dg = lambda a: isinstance(a, np.ndarray) and a.ndim == 2 and a.shape == (mesh.n_cells, 3)
    to_nodal = lambda a: a[cells] if (isinstance(a, np.ndarray) and not is_dg(a) and a.shape[0] == mesh.n_vertices) else a
    swe = O.SWEOracle(mesh, to_nodal(bath_v), options=swe_opts)
    of = {k: to_nodal(v) for k, v in tr_fields_v.items()}
    of["tracer_advective_velocity_factor"] = corr
    trc = O.TracerOracle(swe, bnd_conditions=bnd, fields=of, options=tr_opts)
    trc.set_velocity(uv, eta)
    (kc,) = trc.tendency(c)

    eng = Engine(mesh)
    eng.set_option(L.OPT_NONLINEAR, swe_opts.get("use_nonlinear_equations", True))
    eng.set_option(L.OPT_WETTING_DRYING, swe_opts.get("use_wetting_and_drying", False))
    eng.set_option(L.OPT_WD_ALPHA, swe_opts.get("wetting_and_drying_alpha", 0.5))
    eng.set_option(L.OPT_LF_TRACER, tr_opts.get("use_lax_friedrichs_tracer", False))
    eng.set_option(L.OPT_TRACER_CONSERVATIVE, tr_opts.get("use_conservative_form", False))
    eng.set_option(L.OPT_SIPG_FACTOR_TRACER, tr_opts.get("sipg_factor_tracer", 1.0))
    eng.set_option(L.OPT_TRACER_VEL_FACTOR, corr)
    eng.set_field(L.F_BATHYMETRY, bath_v)
    if tr_fields_v.get("diffusivity_h") is not None:
        eng.set_field(L.F_DIFFUSIVITY, tr_fields_v["diffusivity_h"])
    if tr_fields_v.get("source") is not None:
        eng.set_field(L.F_TRACER_SOURCE, tr_fields_v["source"])
    tags = {"elev": L.BC_ELEV, "uv": L.BC_UV, "un": L.BC_UN, "flux": L.BC_FLUX, "value": L.BC_VALUE,
            "diff_flux": L.BC_DIFF_FLUX}
    slot = {"elev": 0, "uv": 1, "un": 3, "flux": 4, "value": 5, "diff_flux": 6}
    for m, d in bnd.items():
        op, consts = 0, np.zeros(8)
        for tag, val in d.items():
            op |= tags[tag]
            v = np.atleast_1d(np.asarray(val, dtype=float))
            consts[slot[tag]:slot[tag] + v.size] = v
        eng.set_bc(1, m, op, consts)
    st = eng.upload_nodal(uv, eta)
    cd = eng.upload_tracer(c)
    k = eng.new_tracer()
    eng.tracer_stage(0.0, 0.0, 1.0, cd, None, k, st)
    gk = eng.download_tracer(k)
    err = np.abs(gk - kc).max() / np.abs(kc).max()
    assert err < tol, err
    # the specialised (plain advection) kernel and the generic one must agree to rounding
    eng.set_option(L.OPT_FORCE_GENERIC_KERNEL, 1)
    k2 = eng.new_tracer()
    eng.tracer_stage(0.0, 0.0, 1.0, cd, None, k2, st)
    assert np.abs(eng.download_tracer(k2) - gk).max() <= 1e-13 * np.abs(gk).max()
    return err


def test_plain_advection_specialised_kernel():
    """BASELINE config 4 physics (non-conservative advection only): TSPEC 1 kernel, multi-patch unstructured mesh"""
    mesh = sfc_renumber(delaunay_mesh(2500, 1.0, 1.0, seed=9))
    _run(mesh, 1.0, {}, {}, {}, {1: {"value": 0.5}}, tol=1e-12)


def test_diffusion_constant_closed():
    mesh = sfc_renumber(rectangle_mesh(18, 12, 900.0, 600.0))
    _run(mesh, 10.0, {}, dict(sipg_factor_tracer=1.5), {"diffusivity_h": 4.0}, {})


def test_diffusion_variable_unstructured_multi_patch():
    mesh = sfc_renumber(delaunay_mesh(3000, 4e4, 3e4, seed=4))
    X, Y = mesh.coords[:, 0], mesh.coords[:, 1]
    mu = 30.0 * (1.0 + 0.4 * np.cos(X / 6e3) * np.sin(Y / 5e3))
    _run(mesh, 25.0 + 5.0 * np.sin(X / 8e3), {}, {}, {"diffusivity_h": mu}, {}, tol=1e-11)


def test_diffusion_periodic_with_lax_friedrichs():
    mesh = sfc_renumber(periodic_rectangle_mesh(20, 10, 40.0, 20.0))
    _run(mesh, 2.0, {}, dict(use_lax_friedrichs_tracer=True), {"diffusivity_h": 0.02}, {})


@pytest.mark.parametrize("bc", [
    {"value": 3.0},
    {"diff_flux": 0.02},
    {"value": 1.5, "uv": (0.3, -0.1)},
    {"value": 2.5, "diff_flux": -0.01},
    {"un": 0.2},
    {"value": 2.0, "flux": 400.0, "elev": 0.1},
    {},
])
def test_diffusion_boundary_terms(bc):
    mesh = sfc_renumber(rectangle_mesh(12, 10, 600.0, 500.0))
    b = 12 + 0.004 * mesh.coords[:, 0]
    _run(mesh, b, {}, {}, {"diffusivity_h": 1.5 + 0.001 * mesh.coords[:, 1]}, {1: bc, 3: bc, 2: {"value": 1.0}},
         tol=1e-11)


@pytest.mark.parametrize("wd", [False, True])
def test_conservative_form_with_source_and_diffusion(wd):
    mesh = sfc_renumber(rectangle_mesh(14, 8, 14e3, 2e3))
    X = mesh.coords[:, 0]
    b = 4.0 - 4.5 * X / 14e3 if wd else 6.0 + 0.0001 * X
    src = 1e-4 * (1.0 + np.sin(X / 2e3))
    _run(mesh, b, dict(use_wetting_and_drying=wd, wetting_and_drying_alpha=0.4),
         dict(use_conservative_form=True, use_lax_friedrichs_tracer=True),
         {"source": src, "diffusivity_h": 2.0}, {1: {"value": 3.0}, 2: {"value": 1.0, "uv": (0.2, 0.0)}},
         tol=1e-11, corr=0.9)


def test_conservative_form_linear_equations():
    mesh = sfc_renumber(rectangle_mesh(10, 10, 100.0, 100.0, diagonal="right"))
    _run(mesh, 3.0, dict(use_nonlinear_equations=False), dict(use_conservative_form=True), {"source": 0.01}, {})


def test_lincomb_and_integrals():
    """tb_lincomb (Butcher stage combinations) and the device diagnostics against numpy"""
    import torch
    import thetis_b200._lib as L
    from thetis_b200.engine import Engine
    mesh = sfc_renumber(delaunay_mesh(1200, 3e3, 2e3, seed=7))
    uv, eta, c = _fields(mesh, 3)
    eng = Engine(mesh)
    X = mesh.coords[:, 0]
    bath_v = 8.0 + 0.001 * X
    eng.set_field(L.F_BATHYMETRY, bath_v)
    st = eng.upload_nodal(uv, eta)
    a = torch.randn(eng.state_len, dtype=torch.float64, device=eng.device)
    b = torch.randn_like(a)
    out = torch.empty_like(a)
    eng.lincomb([(1.0, st), (0.25, a), (-1.5, b)], out)
    assert torch.allclose(out, st + 0.25 * a - 1.5 * b, rtol=0, atol=1e-14)
    eng.lincomb([(2.0, out), (1.0, a)], out)                      # in place
    assert torch.allclose(out, 2.0 * (st + 0.25 * a - 1.5 * b) + a, rtol=0, atol=1e-13)
    # diagnostics
    area = mesh.cell_area()
    bn = bath_v[mesh.cells]
    o4 = torch.zeros(4, dtype=torch.float64, device=eng.device)
    eng.swe_integrals(st, o4)
    vol = (area * (eta + bn).mean(1)).sum()                          # comp_volume_2d: int (eta + bath)
    assert abs(o4[3].item() - vol) / vol < 1e-13
    assert abs(o4[0].item() - O.l2_norm(mesh, eta) ** 2) / O.l2_norm(mesh, eta) ** 2 < 1e-12
    cd = eng.upload_tracer(c)
    t4 = torch.zeros(4, dtype=torch.float64, device=eng.device)
    eng.tracer_integrals(cd, st, t4)
    mref = (np.ones((3, 3)) + np.eye(3)) / 12.0
    H = bn + eta
    mass = (area * np.einsum("ca,ab,cb->c", H, mref, c)).sum()      # comp_tracer_mass_2d: int H c
    t = t4.cpu().numpy()
    assert abs(t[0] - (area * c.mean(1)).sum()) / abs(t[0]) < 1e-13
    assert abs(t[1] - mass) / mass < 1e-13
    assert t[2] == c.min() and t[3] == c.max()
    # wetting-drying depth in the tracer mass: cell rule
    eng.set_option(L.OPT_WETTING_DRYING, 1)
    eng.set_option(L.OPT_WD_ALPHA, 0.7)
    eng.tracer_integrals(cd, st, t4)
    lam, w = O.cell_quadrature()
    hq = np.einsum("qa,ca->cq", lam, H)
    Hq = 0.5 * (hq + np.sqrt(hq ** 2 + 0.49))
    mass_wd = (area[:, None] * w[None] * Hq * np.einsum("qa,ca->cq", lam, c)).sum()
    assert abs(t4[1].item() - mass_wd) / mass_wd < 1e-13


@pytest.mark.parametrize("cons", [False, True])
def test_discontinuous_p1dg_tracer_source(cons):
    """`source` of the tracer equation given as a genuinely discontinuous P1DG Function (the usual space of tracer
    sources, Q_2d): stored per cell node (tb_set_field_cell); non-conservative and depth-integrated forms"""
    mesh = sfc_renumber(delaunay_mesh(800, 900.0, 700.0, seed=6))
    rng = np.random.default_rng(8)
    src = 1e-3 * (1.0 + rng.standard_normal((mesh.n_cells, 3)))
    b = 20.0 + 4.0 * np.sin(mesh.coords[:, 0] / 150.0)
    _run(mesh, b, {}, dict(use_conservative_form=cons), {"source": src}, {}, tol=1e-11)
