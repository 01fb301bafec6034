"""
The reference's own known-answer tests for the explicit dg-dg path, run on the GPU through the reference-shaped
surface (FlowSolver2d mirror -> SSPRK33 -> C-ABI -> CUDA) with the reference's set-ups, time steps and criteria:

* Rossby soliton           test/swe2d/test_rossby_wave.py:133-257   (refinements 24, 48, T = 30)
* steady-state basin MMS   test/swe2d/test_steady_state_basin_mms.py:114-311 (set-ups 7, 8, 9; refinements 1, 2, 4, 6)
* tracer h-advection       test/tracerEq/test_h-advection_mes_2d.py:9-180    (refinements 1, 2, 3; automatic dt)

and, on the coarsest level of each, against the CPU oracle stepping the same problem.
"""
import numpy as np
import pytest

import kat_setups as K
from oracle import swe_oracle as O
from oracle import c_oracle as CO

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _solver(mesh, bath_vertex, **opts):
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, as_shim_mesh
    sm = as_shim_mesh(mesh)
    P1 = FunctionSpace(sm, "CG", 1, _geometric=mesh.periodic)
    b = Function(P1, name="Bathymetry")
    b.assign(bath_vertex)
    s = solver2d.FlowSolver2d(sm, b)
    s.options.swe_timestepper_type = "SSPRK33"
    s.options.no_exports = True
    s.options.update(opts)
    return s


def _dg(space, nodal):
    from thetis_b200.shim import Function
    f = Function(space)
    f.dat.data[...] = np.asarray(nodal).reshape(f.dat.data_ro.shape)
    return f


def _nodal(solver, mesh):
    uv = solver.fields.uv_2d.dat.data_ro.reshape(mesh.n_cells, 3, 2).copy()
    eta = solver.fields.elev_2d.dat.data_ro.reshape(mesh.n_cells, 3).copy()
    return uv, eta


# ---------------------------------------------------------------- Rossby soliton
def _rossby_gpu(refinement, t_end=30.0):
    from thetis_b200.shim import Function, FunctionSpace, Constant
    from thetis_b200.solver2d import physical_constants
    mesh = K.rossby_mesh(refinement)
    g_old = float(physical_constants["g_grav"])
    physical_constants["g_grav"].assign(1.0)                                    # test_rossby_wave.py:151-152
    try:
        s = _solver(mesh, 1.0, timestep=0.96 / refinement, simulation_end_time=t_end, simulation_export_time=5.0,
                    use_grad_div_viscosity_term=False, use_grad_depth_viscosity_term=False, horizontal_viscosity=None)
        s.options.swe_timestepper_options.use_automatic_timestep = False
        s.create_function_spaces()
        P1 = FunctionSpace(s.mesh2d, "CG", 1, _geometric=True)
        s.options.coriolis_frequency = Function(P1).interpolate(lambda x, y: y)
        for tag in s.mesh2d.exterior_facets.unique_markers:
            s.bnd_functions["shallow_water"][int(tag)] = {"uv": Constant(np.array([0.0, 0.0]))}
        xc = mesh.coords[mesh.cells]
        u, v, e = K.rossby_soliton(xc[..., 0], xc[..., 1])
        s.create_equations()
        s.assign_initial_conditions(uv=_dg(s.function_spaces.U_2d, np.stack([u, v], -1)),
                                    elev=_dg(s.function_spaces.H_2d, e))
        s.iterate()
    finally:
        physical_constants["g_grav"].assign(g_old)
    return mesh, s


def test_rossby_soliton_reference_criteria_on_gpu():
    metrics = []
    for r in (24, 48):
        mesh, s = _rossby_gpu(r)
        assert s.iteration == int(round(30.0 / (0.96 / r)))
        uv, eta = _nodal(s, mesh)
        metrics.append(K.rossby_metrics(mesh, eta))
        if r == 24:
            # same run on the CPU: C oracle, 750 SSPRK33 steps
            xc = mesh.coords[mesh.cells]
            u0, v0, e0 = K.rossby_soliton(xc[..., 0], xc[..., 1])
            state = CO.records_from_nodal(np.stack([u0, v0], -1), e0)
            CO.COracle(mesh, 1.0, g=1.0, coriolis=mesh.coords[:, 1],
                       bnd={m: {"uv": (0.0, 0.0)} for m in mesh.unique_markers()}).ssprk33(state, 0.96 / r, s.iteration)
            uv_c, e_c = CO.nodal_from_records(state)
            assert _rel(eta, e_c) < 1e-9 and _rel(uv, uv_c) < 1e-9
    K.rossby_check_convergence(metrics)


# ---------------------------------------------------------------- steady-state basin MMS
def _mms_gpu(name, refinement, nsteps=None):
    from thetis_b200.shim import Function, FunctionSpace
    p = K.mms_problem(name, refinement)
    mesh = p["mesh"]
    t_end = K.MMS["t_end"] if nsteps is None else nsteps * p["dt"]
    opts = dict(timestep=p["dt"], simulation_end_time=t_end, simulation_export_time=K.MMS["t_end"] / 10.0,
                horizontal_velocity_scale=1.0)
    opts.update(p["setup"]["options"])
    s = _solver(mesh, p["bath_vertex"], **opts)
    s.options.swe_timestepper_options.use_automatic_timestep = False
    s.create_function_spaces()
    fs = s.function_spaces
    s.options.momentum_source_2d = _dg(fs.U_2d, p["momentum_source"])
    s.options.volume_source_2d = _dg(fs.H_2d, p["volume_source"])
    if p["coriolis"] is not None:
        s.options.coriolis_frequency = _dg(fs.H_2d, p["coriolis"])             # projected into H_2d: discontinuous
    if p["viscosity_vertex"] is not None:
        nu = Function(FunctionSpace(s.mesh2d, "CG", 1))
        nu.assign(p["viscosity_vertex"])
        s.options.horizontal_viscosity = nu
    for m, d in p["bnd"].items():
        s.bnd_functions["shallow_water"][m] = {
            tag: _dg(fs.U_2d if tag == "uv" else fs.H_2d, val) for tag, val in d.items()}
    s.create_equations()
    s.assign_initial_conditions(elev=_dg(fs.H_2d, p["elev"]), uv=_dg(fs.U_2d, p["uv"]))
    s.iterate()
    return p, s


@pytest.mark.parametrize("name", ["setup7", "setup8", "setup9"])
def test_mms_convergence_reference_refinements_on_gpu(name):
    """run_convergence(setup, [1, 2, 4, 6], 1) (:306-311): slopes within 20 % of 2 for elev and uv"""
    refs = [1, 2, 4, 6]
    errs = []
    for r in refs:
        p, s = _mms_gpu(name, r)
        uv, eta = _nodal(s, p["mesh"])
        errs.append(K.mms_errors(p, uv, eta))
    se = K.convergence_slope(refs, [e[0] for e in errs])
    su = K.convergence_slope(refs, [e[1] for e in errs])
    assert abs(se - 2.0) / 2.0 < 0.2, (name, "elev", se, errs)
    assert abs(su - 2.0) / 2.0 < 0.2, (name, "uv", su, errs)


@pytest.mark.parametrize("name", ["setup7", "setup8", "setup9"])
def test_mms_steps_match_numpy_oracle(name):
    """40 steps of refinement 2 against the UFL-literal oracle: P1DG Coriolis / sources, flux / un / elev / uv
    boundary Functions, SIPG viscosity with grad-div and grad-depth terms"""
    p, s = _mms_gpu(name, 2, nsteps=40)
    uv_g, eta_g = _nodal(s, p["mesh"])
    fields = {"momentum_source": p["momentum_source"], "volume_source": p["volume_source"]}
    if p["coriolis"] is not None:
        fields["coriolis"] = p["coriolis"]
    if p["viscosity_vertex"] is not None:
        fields["viscosity_h"] = p["viscosity_vertex"][p["mesh"].cells]
    opts = dict(use_nonlinear_equations=True, use_lax_friedrichs_velocity=True)
    opts.update(p["setup"]["options"])
    orc = O.SWEOracle(p["mesh"], p["bath"], options=opts, fields=fields, bnd_conditions=p["bnd"], g_grav=K.MMS["g"])
    uv, eta = p["uv"].copy(), p["elev"].copy()
    st = O.ShuOsherStepper(orc, [uv, eta], p["dt"])
    for i in range(40):
        st.advance(i * p["dt"])
    assert _rel(eta_g, eta) < 1e-10 and _rel(uv_g, uv) < 1e-10


# ---------------------------------------------------------------- tracer h-advection
def _hadv_gpu(refinement):
    from thetis_b200.shim import Constant
    mesh = K.hadv_mesh(refinement)
    s = _solver(mesh, K.HADV["depth"], use_nonlinear_equations=False, use_lax_friedrichs_velocity=True,
                use_lax_friedrichs_tracer=False, horizontal_velocity_scale=abs(K.HADV["u"]),
                simulation_end_time=K.HADV["t_end"], simulation_export_time=K.HADV["t_end"] / 8.0,
                tracer_timestepper_type="SSPRK33")
    s.options.lax_friedrichs_velocity_scaling_factor = Constant(1.0)
    s.create_function_spaces()
    s.options.tracer_advective_velocity_factor = Constant(1.0)
    s.options.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d")
    s.options.use_limiter_for_tracers = True
    uv_bc = Constant(np.array([K.HADV["u"], 0.0]))
    s.bnd_functions["tracer"] = {1: {"value": Constant(0.0), "uv": uv_bc}, 2: {"value": Constant(0.0), "uv": uv_bc}}
    s.bnd_functions["momentum"] = {1: {"uv": uv_bc}, 2: {"uv": uv_bc}}
    s.create_equations()
    uv0 = np.zeros((mesh.n_cells, 3, 2))
    uv0[..., 0] = K.HADV["u"]
    c0 = K.project_dg1(mesh, K.hadv_exact(0.0))
    s.assign_initial_conditions(uv=_dg(s.function_spaces.U_2d, uv0), tracer=_dg(s.function_spaces.Q_2d, c0))
    assert abs(s.dt - K.hadv_timestep(mesh)) < 1e-12 * s.dt           # the automatic time step rule (solver2d.py:214-241)
    # custom time loop that solves the tracer equation only (:96-103): no SWE step, no limiter
    ti = s.timestepper.timesteppers.tracer_2d
    t = 0.0
    while t < K.HADV["t_end"] - 1e-8:
        ti.advance(t)
        t += s.dt
    ti.sync_to_host()
    c = s.fields.tracer_2d.dat.data_ro.reshape(mesh.n_cells, 3).copy()
    area = K.HADV["lx"] * 6.0e3 / refinement
    return mesh, c, t, O.l2_error(mesh, c, K.hadv_exact(t)) / np.sqrt(area)


def test_tracer_h_advection_convergence_on_gpu():
    refs = [1, 2, 3]
    errs = []
    for r in refs:
        mesh, c, t, err = _hadv_gpu(r)
        errs.append(err)
        if r == 1:
            # same run on the CPU oracle
            swe = O.SWEOracle(mesh, K.HADV["depth"], options=dict(use_nonlinear_equations=False))
            trc = O.TracerOracle(swe, bnd_conditions={m: {"value": 0.0, "uv": (K.HADV["u"], 0.0)} for m in (1, 2)},
                                 fields={"tracer_advective_velocity_factor": 1.0})
            uv = np.zeros((mesh.n_cells, 3, 2))
            uv[..., 0] = K.HADV["u"]
            trc.set_velocity(uv, np.zeros((mesh.n_cells, 3)))
            co = K.project_dg1(mesh, K.hadv_exact(0.0))
            dt = K.hadv_timestep(mesh)
            st = O.ShuOsherStepper(trc, [co], dt)
            tt = 0.0
            while tt < K.HADV["t_end"] - 1e-8:
                st.advance(tt)
                tt += dt
            assert _rel(c, co) < 1e-10
    assert K.convergence_slope(refs, errs) > 2 * (1 - 0.2), errs


# ---------------------------------------------------------------- tracer diffusion with a diff_flux boundary
def _diff_flux_series(x, t, lx, nu, D, n_terms=100):
    """The truncated Fourier-series solution of test/tracerEq/test_bcs_2d.py:6-83 (c_t = nu c_xx, c_x(0) = -D,
    c_x(lx) = 0, c(x, 0) = 0), with the coefficients of I(x) = D (lx - x)^2 / (2 lx) in closed form:
    a_0 = D lx / 3, a_k = 2 D lx / (k pi)^2; the source -nu D / lx only feeds the mean."""
    ic = D * 0.5 * (lx - x) ** 2 / lx
    expr = 0.5 * (2.0 * (-nu * D / lx)) * t + 0.5 * (D * lx / 3.0) - ic
    for k in range(1, n_terms):
        expr = expr + 2.0 * D * lx / (k * np.pi) ** 2 * np.exp(-nu * (k * np.pi / lx) ** 2 * t) * np.cos(k * np.pi * x / lx)
    return -expr


def _diff_flux_gpu(refinement, nsteps=None):
    from thetis_b200.shim import Constant
    from thetis_b200.mesh import rectangle_mesh
    lx, ly, nu, D, depth = 10.0, 1.0, 0.1, 0.2, 40.0
    mesh = rectangle_mesh(40 * refinement, 4, lx, ly)
    dt = 0.1 / refinement
    s = _solver(mesh, depth, timestep=dt, simulation_export_time=0.1, simulation_end_time=1.0 - 0.5 * dt,
                tracer_only=True, tracer_timestepper_type="SSPRK33", horizontal_velocity_scale=0.0)
    s.options.horizontal_diffusivity_scale = Constant(nu)
    s.options.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d", diffusivity=Constant(nu))
    s.options.use_limiter_for_tracers = True                           # tracer_element_family == 'dg' (:118)
    s.bnd_functions["tracer_2d"] = {1: {"diff_flux": D * nu}}           # :123
    if nsteps is not None:
        s.options.tracer_timestepper_options.use_automatic_timestep = False
        s.options.swe_timestepper_options.use_automatic_timestep = False
        s.options.timestep = 2.0e-4
        s.options.simulation_end_time = nsteps * 2.0e-4 - 1e-9
    s.assign_initial_conditions()
    s.iterate()
    c = s.fields.tracer_2d.dat.data_ro.reshape(mesh.n_cells, 3).copy()
    return mesh, s, c, (lx, ly, nu, D, depth)


def test_tracer_diff_flux_boundary_convergence_on_gpu():
    """test/tracerEq/test_bcs_2d.py:86-148 with SSPRK33 / dg: refinements 1, 2, 4, successive L2 error ratios > 2
    against the Fourier-series solution (the explicit stepper runs with the reference's automatic time step)"""
    errs = []
    for r in (1, 2, 4):
        mesh, s, c, (lx, ly, nu, D, depth) = _diff_flux_gpu(r)
        assert s.simulation_time >= 1.0 - 0.5 * 0.1 / r - 1e-9
        errs.append(O.l2_error(mesh, c, lambda X, Y: _diff_flux_series(X, 1.0, lx, nu, D)))
    assert errs[0] / errs[1] > 2 and errs[1] / errs[2] > 2, errs


def test_tracer_diff_flux_boundary_steps_match_oracle():
    """the same set-up, 60 steps against the oracle (SIPG diffusion + 'diff_flux' boundary term + limiter)"""
    mesh, s, c_g, (lx, ly, nu, D, depth) = _diff_flux_gpu(1, nsteps=60)
    swe = O.SWEOracle(mesh, depth)
    trc = O.TracerOracle(swe, bnd_conditions={1: {"diff_flux": D * nu}}, fields={"diffusivity_h": nu})
    trc.set_velocity(np.zeros((mesh.n_cells, 3, 2)), np.zeros((mesh.n_cells, 3)))
    c = np.zeros((mesh.n_cells, 3))
    st = O.ShuOsherStepper(trc, [c], 2.0e-4)
    for i in range(60):
        st.advance(i * 2.0e-4)
        c[...] = O.vertex_based_limiter(mesh, c)
    assert np.abs(c).max() > 1e-6
    assert _rel(c_g, c) < 1e-10
