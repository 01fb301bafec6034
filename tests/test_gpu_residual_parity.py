"""
GPU parity: M^-1 R(u) from the CUDA stage kernel (through the C-ABI) against the
numpy oracle on the same seeded inputs.  fp64; tolerance: 1e-12 relative to the
max-norm of the tendency for polynomial integrands, 1e-11 where sqrt/cbrt/division
order differs (nonlinear depth, Manning).
"""
import numpy as np
import pytest

from thetis_b200.mesh import (rectangle_mesh, periodic_rectangle_mesh, delaunay_mesh, read_gmsh, sfc_renumber,
                              FACET_NODES)
from oracle.swe_oracle import SWEOracle

pytestmark = pytest.mark.gpu


def _state(mesh, seed, amp_u=0.5, amp_e=0.3):
    rng = np.random.default_rng(seed)
    x = mesh.coords[mesh.cells]
    L = np.ptp(mesh.coords, axis=0).max()
    k = 2 * np.pi / L
    u = amp_u * np.sin(k * x[..., 0] + 0.3) * np.cos(k * x[..., 1]) + 0.05 * rng.standard_normal(x.shape[:2])
    v = -amp_u * np.cos(2 * k * x[..., 0]) * np.sin(k * x[..., 1] + 0.1) + 0.05 * rng.standard_normal(x.shape[:2])
    e = amp_e * np.cos(k * x[..., 0]) * np.sin(k * x[..., 1]) + 0.02 * rng.standard_normal(x.shape[:2])
    return np.stack([u, v], -1), e


def _vertex_field(mesh, fn):
    return fn(mesh.coords[:, 0], mesh.coords[:, 1])


def _run(mesh, bath_v, options, fields_v, bnd, g=9.81, tol=1e-12, seed=0, bc_arrays=None, return_engine=False,
         cell_rule=None):
    """fields_v: dict name -> None | const | per-vertex array; bnd: {marker: {tag: const}};
    cell_rule: name of another degree-3 cell rule (oracle.swe_oracle.cell_quadrature) for oracle and library"""
    import thetis_b200._lib as L
    from thetis_b200.engine import Engine
    uv, eta = _state(mesh, seed)
    cells = mesh.cells
    # per-vertex arrays (P1) -> cell-nodal; (nt, 3[,2]) arrays are genuinely discontinuous P1DG fields: as they are
    is_dg = lambda a: isinstance(a, np.ndarray) and a.ndim >= 2 and a.shape[:2] == (mesh.n_cells, 3)
    to_nodal = lambda a: a[cells] if (isinstance(a, np.ndarray) and not is_dg(a) and a.shape[0] == mesh.n_vertices) else a
    ofields = {k: to_nodal(v) for k, v in fields_v.items() if v is not None}
    options = dict(options)
    wd_alpha = options.get("wetting_and_drying_alpha", 0.5)
    if isinstance(wd_alpha, np.ndarray):
        options["wetting_and_drying_alpha"] = wd_alpha[cells]          # P1 field: nodal for the oracle
    obnd = {}
    for m, d in bnd.items():
        obnd[m] = dict(d)
    if bc_arrays:
        # per-exterior-facet nodal arrays -> oracle wants DG nodal arrays (nt,3[,2])
        for (m, tag), arr in bc_arrays.items():
            full = np.zeros((mesh.n_cells, 3) + arr.shape[2:])
            for side in range(2):
                full[mesh.bf_cell, FACET_NODES[mesh.bf_lf, side]] = arr[:, side]
            obnd.setdefault(m, {})[tag] = full
    orc = SWEOracle(mesh, to_nodal(bath_v), options=options, fields=ofields, bnd_conditions=obnd, g_grav=g,
                    **({"cell_rule": cell_rule} if cell_rule else {}))
    ku, ke = orc.tendency(uv, eta)

    eng = Engine(mesh)
    if cell_rule:
        eng.set_cell_quadrature(orc.lam, orc.qw)
    eng.set_option(L.OPT_G_GRAV, g)
    eng.set_option(L.OPT_NONLINEAR, options.get("use_nonlinear_equations", True))
    eng.set_option(L.OPT_LAX_FRIEDRICHS, options.get("use_lax_friedrichs_velocity", True))
    eng.set_option(L.OPT_NORM_SMOOTHER, options.get("norm_smoother", 0.0))
    eng.set_option(L.OPT_WETTING_DRYING, options.get("use_wetting_and_drying", False))
    if isinstance(wd_alpha, np.ndarray):
        eng.set_field(L.F_WD_ALPHA, wd_alpha)
    else:
        eng.set_option(L.OPT_WD_ALPHA, wd_alpha)
    eng.set_option(L.OPT_LF_SCALING, fields_v.get("lax_friedrichs_velocity_scaling_factor", 1.0))
    eng.set_option(L.OPT_SIPG_FACTOR, options.get("sipg_factor", 1.0))
    eng.set_option(L.OPT_GRAD_DIV_VISCOSITY, options.get("use_grad_div_viscosity_term", False))
    eng.set_option(L.OPT_GRAD_DEPTH_VISCOSITY, options.get("use_grad_depth_viscosity_term", True))
    eng.set_field(L.F_BATHYMETRY, bath_v)
    names = {"coriolis": L.F_CORIOLIS, "manning_drag_coefficient": L.F_MANNING,
             "quadratic_drag_coefficient": L.F_QUAD_DRAG, "linear_drag_coefficient": L.F_LINEAR_DRAG,
             "wind_stress": L.F_WIND_STRESS, "atmospheric_pressure": L.F_ATM_PRESSURE,
             "momentum_source": L.F_MOMENTUM_SOURCE, "volume_source": L.F_VOLUME_SOURCE,
             "viscosity_h": L.F_VISCOSITY, "nikuradse_bed_roughness": L.F_NIKURADSE}
    for k, fid in names.items():
        if fields_v.get(k) is not None:
            eng.set_field(fid, fields_v[k])
    tags = {"elev": L.BC_ELEV, "uv": L.BC_UV, "un": L.BC_UN, "flux": L.BC_FLUX, "drag": L.BC_DRAG}
    for m, d in obnd.items():
        op = 0
        consts = np.zeros(8)
        for tag, val in d.items():
            op |= tags[tag]
            if not (isinstance(val, np.ndarray) and val.ndim >= 2):
                if tag == "elev": consts[0] = val
                if tag == "uv": consts[1:3] = val
                if tag == "un": consts[3] = val
                if tag == "flux": consts[4] = val
                if tag == "drag": consts[7] = val
        eng.set_bc(0, m, op, consts)
    if bc_arrays:
        for (m, tag), arr in bc_arrays.items():
            eng.set_bc_array(0, m, tags[tag], arr)
    st = eng.upload_nodal(uv, eta)
    k = eng.new_state()
    eng.swe_tendency(st, k)
    gu, ge = eng.download_nodal(k)
    # the specialised stage kernels and the generic one must agree to rounding (the specialised ones interpolate to the
    # cell-rule points through the symmetric form of the default rule: same values, another summation order)
    eng.set_option(L.OPT_FORCE_GENERIC_KERNEL, 1)
    k2 = eng.new_state()
    eng.swe_tendency(st, k2)
    gu2, ge2 = eng.download_nodal(k2)
    assert np.abs(gu2 - gu).max() <= 1e-12 * np.abs(gu).max() and np.abs(ge2 - ge).max() <= 1e-12 * np.abs(ge).max()
    su = np.abs(ku).max()
    se = np.abs(ke).max()
    eu = np.abs(gu - ku).max() / su
    ee = np.abs(ge - ke).max() / se
    assert eu < tol and ee < tol, (eu, ee)
    if return_engine:
        return eng, orc, (uv, eta)
    return eu, ee


def test_linear_constant_depth_closed():
    mesh = sfc_renumber(rectangle_mesh(24, 20, 1000.0, 800.0))
    _run(mesh, 50.0, dict(use_nonlinear_equations=False), {}, {})


def test_linear_variable_depth_small_ragged():
    # 2 x 3 cells x 2 = 12 triangles: a single ragged patch
    mesh = rectangle_mesh(2, 3, 10.0, 10.0)
    b = _vertex_field(mesh, lambda x, y: 5.0 + 0.1 * x + 0.05 * y)
    _run(mesh, b, dict(use_nonlinear_equations=False), {}, {})


def test_nonlinear_lf_closed_channel():
    mesh = sfc_renumber(rectangle_mesh(40, 25, 40e3, 2e3))
    _run(mesh, 20.0, {}, {}, {}, tol=1e-11)


def test_nonlinear_no_lf():
    mesh = sfc_renumber(rectangle_mesh(17, 9, 300.0, 200.0, diagonal="right"))
    b = _vertex_field(mesh, lambda x, y: 10 + 2 * np.sin(x / 50.0))
    _run(mesh, b, dict(use_lax_friedrichs_velocity=False), {}, {}, tol=1e-11)


def test_periodic_coriolis_uv_bc():
    # Rossby-soliton style set-up (test/swe2d/test_rossby_wave.py): g=1, h=1, f=y, 'uv' BCs
    mesh = sfc_renumber(periodic_rectangle_mesh(48, 24, 48.0, 24.0, origin=(-24.0, -12.0)))
    f = _vertex_field(mesh, lambda x, y: y)
    bnd = {m: {"uv": (0.0, 0.0)} for m in mesh.unique_markers()}
    _run(mesh, 1.0, {}, {"coriolis": f}, bnd, g=1.0, tol=1e-11)


def test_unstructured_stommel_terms():
    # stommel2d style: linear, coriolis beta-plane, wind stress, linear drag
    mesh = sfc_renumber(delaunay_mesh(3000, 1e6, 1e6, seed=0))
    f = _vertex_field(mesh, lambda x, y: 1e-4 + 2e-11 * y)
    tau = np.stack([0.1 * np.sin(np.pi * (mesh.coords[:, 1] / 1e6 - 0.5)), 0 * mesh.coords[:, 1]], -1)
    _run(mesh, 1000.0, dict(use_nonlinear_equations=False),
         {"coriolis": f, "wind_stress": tau, "linear_drag_coefficient": 1e-6}, {})


def test_manning_atm_pressure_sources_nonlinear():
    mesh = sfc_renumber(rectangle_mesh(16, 16, 10e3, 10e3))
    X, Y = mesh.coords[:, 0], mesh.coords[:, 1]
    b = 5.0 + 1.0 * np.cos(np.pi * X / 10e3)
    pa = 2.0 * 9.81 * 1000.0 * np.cos(np.pi * X / 10e3) * np.cos(np.pi * Y / 10e3)
    ms = np.stack([1e-4 * np.sin(X / 3e3), 2e-4 * np.cos(Y / 2e3)], -1)
    vs = 1e-4 * np.sin((X + Y) / 4e3)
    _run(mesh, b, dict(norm_smoother=0.01),
         {"manning_drag_coefficient": 0.03 + 0 * X, "atmospheric_pressure": pa, "momentum_source": ms,
          "volume_source": vs, "coriolis": 1.2e-4, "wind_stress": (0.05, -0.02)}, {}, tol=1e-11)


def test_quadratic_drag_const_linear_drag_field():
    mesh = sfc_renumber(rectangle_mesh(10, 12, 1e3, 1e3))
    X = mesh.coords[:, 0]
    _run(mesh, 8.0, {}, {"quadratic_drag_coefficient": 2.5e-3, "linear_drag_coefficient": 1e-3 * (1 + X / 1e3)},
         {}, tol=1e-11)


@pytest.mark.parametrize("bc", [
    {"elev": 0.3, "uv": (0.2, -0.1)},
    {"elev": 0.3, "un": 0.15},
    {"elev": -0.2, "flux": 500.0},
    {"elev": 0.25},
    {"uv": (0.1, 0.3)},
    {"un": -0.2},
    {"flux": -300.0},
])
@pytest.mark.parametrize("nonlinear", [True, False])
def test_open_boundary_opcodes(bc, nonlinear):
    mesh = sfc_renumber(rectangle_mesh(12, 10, 600.0, 500.0))
    b = _vertex_field(mesh, lambda x, y: 12 + 0.004 * x)
    bnd = {1: bc, 2: {"elev": 0.1}, 3: bc}
    _run(mesh, b, dict(use_nonlinear_equations=nonlinear), {}, bnd, tol=1e-11)


@pytest.mark.parametrize("nonlinear", [True, False])
def test_boundary_drag_term(nonlinear):
    """BoundaryDragTerm (shallowwater_eq.py:704-726): quadratic friction of the tangential velocity on markers that carry
    a 'drag' tag -- alone (closed boundary) and combined with open tags; Manning + Coriolis are on so that a
    specialised kernel would be chosen without the tag (the term lives in the generic kernel).  The oracle side of this
    comparison reproduces the reference's own BoundaryDragTerm (tests/test_oracle_reference_residuals.py)."""
    import thetis_b200._lib as L
    mesh = sfc_renumber(delaunay_mesh(900, 5.0e3, 4.0e3, seed=5))
    X, Y = mesh.coords[:, 0], mesh.coords[:, 1]
    b = 20.0 + 3.0 * np.sin(X / 900.0) * np.cos(Y / 700.0)
    bnd = {1: {"drag": 0.05}, 2: {"elev": -0.2, "un": 0.15, "drag": 0.02}, 3: {"uv": (0.1, 0.05), "drag": 0.1}, 4: {"drag": 0.2}}
    fields = {"manning_drag_coefficient": 0.03, "coriolis": 1e-4 + 2e-8 * Y} if nonlinear else {}
    eng, orc, (uv, eta) = _run(mesh, b, dict(use_nonlinear_equations=nonlinear), fields, bnd, tol=1e-11, return_engine=True)
    # the term is really there: without the tags the tendency moves far beyond the tolerance
    ku, _ = orc.tendency(uv, eta)
    orc.bnd = {m: {t: v for t, v in d.items() if t != "drag"} for m, d in bnd.items()}
    ku0, _ = orc.tendency(uv, eta)
    assert np.abs(ku - ku0).max() > 1e-4 * np.abs(ku).max()
    # fluid at rest: |u_t| = 0 exactly on every boundary facet, the root must not produce NaN
    eng.set_option(L.OPT_FORCE_GENERIC_KERNEL, 0)
    st = eng.upload_nodal(np.zeros_like(uv), eta)
    k = eng.new_state()
    eng.swe_tendency(st, k)
    gu, ge = eng.download_nodal(k)
    orc.bnd = bnd
    ku, ke = orc.tendency(np.zeros_like(uv), eta)
    assert np.isfinite(gu).all() and np.isfinite(ge).all()
    assert np.abs(gu - ku).max() <= 1e-11 * np.abs(ku).max() and np.abs(ge - ke).max() <= 1e-11 * np.abs(ke).max()


def test_boundary_drag_is_a_shallow_water_tag():
    import thetis_b200._lib as L
    from thetis_b200.engine import Engine
    eng = Engine(rectangle_mesh(4, 4, 10.0, 10.0))
    c = np.zeros(8)
    c[7] = 0.1
    with pytest.raises(L.TbError):
        eng.set_bc(1, 1, L.BC_DRAG, c)          # the tracer equation has no boundary drag


def test_open_boundary_arrays_north_sea():
    # tidal elevation Function on marker 100 + uv Constant (demos/demo_2d_north_sea.py), Manning + Coriolis
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "mini_tagged.msh")
    mesh = sfc_renumber(read_gmsh(path))
    rng = np.random.default_rng(3)
    elev = 0.5 + 0.1 * rng.standard_normal((mesh.n_bfacets, 2))
    X, Y = mesh.coords[:, 0], mesh.coords[:, 1]
    b = 30.0 + 5 * np.sin(X / 7.0) * np.cos(Y / 5.0)
    _run(mesh, b, {}, {"manning_drag_coefficient": 0.03, "coriolis": 1e-4 + 1e-6 * Y},
         {100: {"elev": 0.0, "uv": (0.0, 0.0)}}, tol=1e-11, bc_arrays={(100, "elev"): elev})


def test_wetting_drying_residual():
    mesh = sfc_renumber(rectangle_mesh(14, 6, 14e3, 1.2e3))
    X = mesh.coords[:, 0]
    b = 3.0 - 5.0 * X / 14e3          # dries out: negative bathymetry on the right
    _run(mesh, b, dict(use_wetting_and_drying=True, wetting_and_drying_alpha=0.4),
         {"manning_drag_coefficient": 0.02}, {1: {"elev": 0.5}}, tol=1e-10)


@pytest.mark.parametrize("wd", [False, True])
def test_config5_specialised_kernels(wd):
    """North-Sea physics (Manning + Coriolis + LF + tidal elevation array [+ wetting-drying]): SPEC 2 / SPEC 3 kernels"""
    import os
    from harness.workloads import north_sea_mesh, north_sea_setup, tide_values
    mesh = north_sea_mesh(k=1)
    setup = north_sea_setup(mesh, wetting_drying=wd)
    tv = tide_values(setup, 1234.0)
    _run(mesh, setup["bath"], dict(use_wetting_and_drying=wd, wetting_and_drying_alpha=0.5),
         {"manning_drag_coefficient": setup["manning"], "coriolis": setup["coriolis"]},
         {100: {"elev": 0.0, "uv": (0.0, 0.0)}}, tol=1e-10, bc_arrays={(100, "elev"): tv})


def test_other_cell_rule_is_served_by_the_generic_kernel():
    """tb_set_cell_quadrature with a rule that is not one symmetric orbit (Dunavant's 6-point rule: two orbits, two
    weights): the specialised kernels are written for the default rule's structure, so the library must route this
    to the generic kernel -- same North-Sea physics as above, oracle with the same rule."""
    from harness.workloads import north_sea_mesh, north_sea_setup, tide_values
    from oracle.swe_oracle import cell_quadrature
    from thetis_b200.engine import Engine
    mesh = north_sea_mesh(k=1)
    setup = north_sea_setup(mesh, wetting_drying=True)
    tv = tide_values(setup, 1234.0)
    try:
        _run(mesh, setup["bath"], dict(use_wetting_and_drying=True, wetting_and_drying_alpha=0.5),
             {"manning_drag_coefficient": setup["manning"], "coriolis": setup["coriolis"]},
             {100: {"elev": 0.0, "uv": (0.0, 0.0)}}, tol=1e-10, bc_arrays={(100, "elev"): tv}, cell_rule="dunavant6")
    finally:
        # the rule lives in constant memory of the device (one per process): back to the default for the other tests
        Engine(mesh).set_cell_quadrature(*cell_quadrature())


# ---------------------------------------------------------------- HorizontalViscosityTerm (SIPG), SURVEY 8f rank 1
@pytest.mark.parametrize("graddiv", [False, True])
@pytest.mark.parametrize("graddepth", [False, True])
def test_viscosity_structured_closed(graddiv, graddepth):
    mesh = sfc_renumber(rectangle_mesh(20, 14, 2000.0, 1500.0))
    b = _vertex_field(mesh, lambda x, y: 12.0 + 3.0 * np.sin(x / 400.0) * np.cos(y / 300.0))
    _run(mesh, b, dict(use_grad_div_viscosity_term=graddiv, use_grad_depth_viscosity_term=graddepth, sipg_factor=1.5),
         {"viscosity_h": 35.0}, {}, tol=1e-11)


def test_viscosity_linear_equations_variable_nu():
    mesh = sfc_renumber(rectangle_mesh(16, 16, 1000.0, 1000.0, diagonal="right"))
    nu = _vertex_field(mesh, lambda x, y: 5.0 + 0.01 * x + 0.004 * y)
    b = _vertex_field(mesh, lambda x, y: 20.0 + 0.002 * x)
    _run(mesh, b, dict(use_nonlinear_equations=False), {"viscosity_h": nu}, {}, tol=1e-11)


def test_viscosity_unstructured_multi_patch():
    # 6000 triangles = 47 patches: neighbour gradients come from halo cells whose vertices are outside the patch
    mesh = sfc_renumber(delaunay_mesh(3000, 5e4, 4e4, seed=2))
    X, Y = mesh.coords[:, 0], mesh.coords[:, 1]
    nu = 50.0 * (1.0 + 0.5 * np.sin(X / 7e3) * np.cos(Y / 9e3))
    b = 30.0 + 10.0 * np.cos(X / 1e4)
    _run(mesh, b, dict(use_grad_div_viscosity_term=True, sipg_factor=2.0),
         {"viscosity_h": nu, "manning_drag_coefficient": 0.025, "coriolis": 1e-4}, {}, tol=1e-10)


def test_viscosity_periodic():
    mesh = sfc_renumber(periodic_rectangle_mesh(24, 12, 48.0, 24.0, origin=(-24.0, -12.0)))
    bnd = {m: {"uv": (0.0, 0.0)} for m in mesh.unique_markers()}
    _run(mesh, 1.0, dict(use_grad_depth_viscosity_term=False), {"viscosity_h": 0.05}, bnd, g=1.0, tol=1e-11)


@pytest.mark.parametrize("bc", [
    {"elev": 0.3, "uv": (0.2, -0.1)},
    {"elev": 0.3, "un": 0.15},
    {"elev": -0.2, "flux": 500.0},
    {"elev": 0.25},
    {"uv": (0.1, 0.3)},
    {"un": -0.2},
    {"flux": -300.0},
])
@pytest.mark.parametrize("graddiv", [False, True])
def test_viscosity_open_boundary_dirichlet_terms(bc, graddiv):
    mesh = sfc_renumber(rectangle_mesh(12, 10, 600.0, 500.0))
    b = _vertex_field(mesh, lambda x, y: 12 + 0.004 * x)
    nu = _vertex_field(mesh, lambda x, y: 2.0 + 0.002 * y)
    bnd = {1: bc, 2: {"elev": 0.1}, 3: bc}
    _run(mesh, b, dict(use_grad_div_viscosity_term=graddiv), {"viscosity_h": nu}, bnd, tol=1e-11)


def test_viscosity_wetting_drying_grad_depth():
    mesh = sfc_renumber(rectangle_mesh(14, 6, 14e3, 1.2e3))
    X = mesh.coords[:, 0]
    b = 3.0 - 5.0 * X / 14e3
    _run(mesh, b, dict(use_wetting_and_drying=True, wetting_and_drying_alpha=0.4, use_grad_div_viscosity_term=True),
         {"viscosity_h": 10.0, "manning_drag_coefficient": 0.02}, {1: {"elev": 0.5}}, tol=1e-10)


def test_viscosity_toggle_rebuilds_patch_tables():
    """the halo-geometry tables are built when viscosity is switched on and dropped when it is cleared"""
    import thetis_b200._lib as L
    from thetis_b200.engine import Engine
    mesh = sfc_renumber(delaunay_mesh(800, 1e4, 1e4, seed=5))
    uv, eta = _state(mesh, 1)
    eng = Engine(mesh)
    eng.set_field(L.F_BATHYMETRY, 20.0)
    st = eng.upload_nodal(uv, eta)
    k0, k1, k2 = eng.new_state(), eng.new_state(), eng.new_state()
    eng.swe_tendency(st, k0)
    eng.set_field(L.F_VISCOSITY, 3.0)
    eng.swe_tendency(st, k1)
    eng.set_field(L.F_VISCOSITY, None)
    eng.swe_tendency(st, k2)
    import torch
    assert torch.equal(k0, k2)
    assert not torch.equal(k0, k1)
    orc = SWEOracle(mesh, 20.0, fields={"viscosity_h": 3.0})
    ku, ke = orc.tendency(uv, eta)
    gu, ge = eng.download_nodal(k1)
    assert np.abs(gu - ku).max() / np.abs(ku).max() < 1e-11


@pytest.mark.parametrize("wd", [False, True])
@pytest.mark.parametrize("graddiv", [False, True])
def test_config5_with_viscosity_specialised_kernels(wd, graddiv):
    """North-Sea physics + horizontal viscosity (examples/north_sea prescribes a viscosity sponge): SPEC 5 / SPEC 6"""
    from harness.workloads import north_sea_mesh, north_sea_setup, tide_values
    mesh = north_sea_mesh(k=1)
    setup = north_sea_setup(mesh, wetting_drying=wd)
    tv = tide_values(setup, 4321.0)
    X = mesh.coords[:, 0]
    nu = 50.0 + 200.0 * np.exp(-((X - X.min()) / 5e4) ** 2)          # sponge towards the open boundary
    _run(mesh, setup["bath"], dict(use_wetting_and_drying=wd, wetting_and_drying_alpha=0.5,
                                   use_grad_div_viscosity_term=graddiv),
         {"manning_drag_coefficient": setup["manning"], "coriolis": setup["coriolis"], "viscosity_h": nu},
         {100: {"elev": 0.0, "uv": (0.0, 0.0)}}, tol=1e-10, bc_arrays={(100, "elev"): tv})


@pytest.mark.parametrize("bc", [{"elev": 0.3, "uv": (0.2, -0.1)}, {"un": -0.2}, {}])
def test_modesplit_equations_no_momentum_advection(bc):
    """ModeSplit2DEquations (shallowwater_eq.py:931-966): nonlinear depth in HUDiv, external pressure
    gradient, Coriolis, momentum source, atmospheric pressure -- and no HorizontalAdvectionTerm"""
    import thetis_b200._lib as L
    from thetis_b200.engine import Engine
    mesh = sfc_renumber(delaunay_mesh(1500, 2e4, 1.5e4, seed=6))
    X, Y = mesh.coords[:, 0], mesh.coords[:, 1]
    b = 15.0 + 4.0 * np.sin(X / 4e3)
    f = 1e-4 + 1e-9 * Y
    ms = np.stack([1e-4 * np.sin(X / 3e3), 2e-4 * np.cos(Y / 2e3)], -1)
    pa = 500.0 * np.cos(X / 5e3)
    uv, eta = _state(mesh, 4)
    cells = mesh.cells
    bnd = {1: bc} if bc else {}
    orc = SWEOracle(mesh, b[cells], options=dict(include_momentum_advection=False),
                    fields={"coriolis": f[cells], "momentum_source": ms[cells], "atmospheric_pressure": pa[cells]},
                    bnd_conditions=bnd)
    ku, ke = orc.tendency(uv, eta)
    full = SWEOracle(mesh, b[cells], fields={"coriolis": f[cells], "momentum_source": ms[cells],
                                             "atmospheric_pressure": pa[cells]}, bnd_conditions=bnd).tendency(uv, eta)
    assert np.abs(full[0] - ku).max() > 1e-3 * np.abs(ku).max()        # the advection term is not negligible here
    eng = Engine(mesh)
    eng.set_option(L.OPT_MOMENTUM_ADVECTION, 0)
    eng.set_field(L.F_BATHYMETRY, b)
    eng.set_field(L.F_CORIOLIS, f)
    eng.set_field(L.F_MOMENTUM_SOURCE, ms)
    eng.set_field(L.F_ATM_PRESSURE, pa)
    if bc:
        tags = {"elev": L.BC_ELEV, "uv": L.BC_UV, "un": L.BC_UN}
        op, consts = 0, np.zeros(8)
        for tag, val in bc.items():
            op |= tags[tag]
            if tag == "elev": consts[0] = val
            if tag == "uv": consts[1:3] = val
            if tag == "un": consts[3] = val
        eng.set_bc(0, 1, op, consts)
    st = eng.upload_nodal(uv, eta)
    k = eng.new_state()
    eng.swe_tendency(st, k)
    gu, ge = eng.download_nodal(k)
    assert np.abs(gu - ku).max() / np.abs(ku).max() < 1e-11 and np.abs(ge - ke).max() / np.abs(ke).max() < 1e-11


# ---------------------------------------------------------------- round 2: coefficient forms the reference accepts
def _dg_field(mesh, seed, base, amp, ncomp=None):
    """genuinely discontinuous P1DG nodal data (nt, 3[,k]): smooth part + per-node noise"""
    rng = np.random.default_rng(seed)
    shape = (mesh.n_cells, 3) + (() if ncomp is None else (ncomp,))
    return base + amp * rng.standard_normal(shape)


def test_p1dg_coefficient_fields_linear():
    """Coriolis / sources / drag / wind / pressure given as discontinuous P1DG Functions (the reference projects
    Coriolis and the MMS sources into H_2d / U_2d, test_steady_state_basin_mms.py:169-177, and passes P1DG atmospheric
    pressure, test_atmospheric_pressure.py:57-63): stored per cell node (tb_set_field_cell), generic kernel"""
    mesh = sfc_renumber(delaunay_mesh(900, 1000.0, 800.0, seed=3))
    fields = {"coriolis": _dg_field(mesh, 1, 1e-2, 3e-3), "linear_drag_coefficient": _dg_field(mesh, 2, 2e-3, 5e-4),
              "wind_stress": _dg_field(mesh, 3, 0.1, 0.05, 2), "atmospheric_pressure": _dg_field(mesh, 4, 1e5, 300.0),
              "momentum_source": _dg_field(mesh, 5, 0.0, 1e-3, 2), "volume_source": _dg_field(mesh, 6, 0.0, 1e-3)}
    _run(mesh, 30.0, dict(use_nonlinear_equations=False), fields, {})


def test_p1dg_coefficient_fields_nonlinear_manning():
    mesh = sfc_renumber(delaunay_mesh(700, 1000.0, 800.0, seed=5))
    b = _vertex_field(mesh, lambda x, y: 12.0 + 3 * np.sin(x / 200.0) * np.cos(y / 150.0))
    fields = {"coriolis": _dg_field(mesh, 1, 1e-2, 3e-3), "manning_drag_coefficient": _dg_field(mesh, 2, 0.03, 0.004),
              "momentum_source": _dg_field(mesh, 5, 0.0, 1e-3, 2), "volume_source": _dg_field(mesh, 6, 0.0, 1e-3)}
    bnd = {m: {"elev": 0.2, "uv": (0.1, -0.05)} for m in mesh.unique_markers()[:1]}
    _run(mesh, b, dict(norm_smoother=1e-3), fields, bnd, tol=1e-11)


def test_discontinuous_coefficient_of_a_facet_term_is_rejected():
    import thetis_b200._lib as L
    from thetis_b200.engine import Engine
    mesh = rectangle_mesh(4, 4, 10.0, 10.0)
    eng = Engine(mesh)
    for fid in (L.F_BATHYMETRY, L.F_VISCOSITY, L.F_WD_ALPHA):
        with pytest.raises(L.TbError):
            eng.set_field(fid, _dg_field(mesh, 0, 10.0, 1.0))


@pytest.mark.parametrize("as_dg", [False, True])
def test_nikuradse_bed_roughness(as_dg):
    """QuadraticDragTerm with the Nikuradse law (shallowwater_eq.py:689-697), incl. cells where H < k_s (C_D = 0)"""
    mesh = sfc_renumber(rectangle_mesh(14, 11, 700.0, 500.0))
    b = _vertex_field(mesh, lambda x, y: 0.6 + 0.5 * np.sin(x / 90.0) * np.cos(y / 70.0))       # H in ~[0.1, 1.4]
    ks = _vertex_field(mesh, lambda x, y: 0.25 + 0.2 * np.cos(x / 130.0))
    if as_dg:
        ks = ks[mesh.cells] * (1.0 + 0.1 * np.random.default_rng(0).standard_normal((mesh.n_cells, 3)))
    _run(mesh, b, dict(norm_smoother=1e-2), {"nikuradse_bed_roughness": ks}, {}, tol=1e-11)


def test_nikuradse_and_manning_together_raise():
    import thetis_b200._lib as L
    from thetis_b200.engine import Engine
    mesh = rectangle_mesh(4, 4, 10.0, 10.0)
    eng = Engine(mesh)
    eng.set_field(L.F_BATHYMETRY, 5.0)
    eng.set_field(L.F_MANNING, 0.02)
    eng.set_field(L.F_NIKURADSE, 0.1)
    st = eng.new_state()
    with pytest.raises(L.TbError, match="Nikuradse"):
        eng.swe_tendency(st, eng.new_state())


def test_wetting_drying_alpha_as_p1_field():
    """wetting_and_drying_alpha as a P1 Function (use_automatic_wetting_and_drying_alpha, solver2d.py:279-287;
    utility.py:981-983): interior facets, cell rule (Manning), closed and open boundaries"""
    mesh = sfc_renumber(read_gmsh(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "mini_tagged.msh")))
    b = _vertex_field(mesh, lambda x, y: 1.5 + 2.0 * np.sin(x / 9.0) * np.cos(y / 7.0))          # dries in places
    alpha = _vertex_field(mesh, lambda x, y: 0.4 + 0.3 * np.cos(x / 11.0) ** 2 + 0.1 * np.sin(y / 5.0))
    bnd = {100: {"elev": 0.3, "flux": -40.0}}
    _run(mesh, b, dict(use_wetting_and_drying=True, wetting_and_drying_alpha=alpha, norm_smoother=1e-3),
         {"manning_drag_coefficient": 0.03}, bnd, tol=1e-11)


def test_time_dependent_p1_field_updates_only_its_columns():
    """new VALUES of an existing P1 coefficient (wind stress / pressure assigned in update_forcings) are refreshed
    stream-ordered column by column (tb_sync_fields) and give the same tendency as a fresh context"""
    mesh = sfc_renumber(delaunay_mesh(600, 1000.0, 800.0, seed=9))
    w0 = np.stack([_vertex_field(mesh, lambda x, y: 0.1 * np.sin(x / 300.0)), _vertex_field(mesh, lambda x, y: 0.05 + 0 * x)], -1)
    pa0 = _vertex_field(mesh, lambda x, y: 1e5 + 200.0 * np.cos(y / 200.0))
    opts = dict(use_nonlinear_equations=False)
    eng, orc, (uv, eta) = _run(mesh, 25.0, opts, {"wind_stress": w0, "atmospheric_pressure": pa0}, {}, return_engine=True)
    import thetis_b200._lib as L
    eng.set_option(L.OPT_FORCE_GENERIC_KERNEL, 0)
    n0 = eng.launch_count()
    for step in range(1, 4):
        w = w0 * (1.0 + 0.3 * step)
        pa = pa0 + 50.0 * step * np.sin(mesh.coords[:, 0] / 150.0)
        eng.set_field(L.F_WIND_STRESS, w)
        eng.set_field(L.F_ATM_PRESSURE, pa)
        st = eng.upload_nodal(uv, eta)
        k = eng.new_state()
        eng.swe_tendency(st, k)
        gu, ge = eng.download_nodal(k)
        orc.fields["wind_stress"] = w[mesh.cells]
        orc.fields["atmospheric_pressure"] = pa[mesh.cells]
        ku, ke = orc.tendency(uv, eta)
        assert np.abs(gu - ku).max() < 1e-12 * np.abs(ku).max() and np.abs(ge - ke).max() < 1e-12 * np.abs(ke).max()
    # per refresh: one column-scatter launch per field (2), not a rebuild; plus layout conversion and the stage itself
    assert eng.launch_count() - n0 <= 3 * (2 + 3)
