"""
GPU parity of whole time steps, driven through the reference-shaped surface
(FlowSolver2d / SSPRK33 / VertexBasedP1DGLimiter mirrors -> C-ABI -> CUDA) against
the numpy oracle stepping the same mesh, dt and forcings.

Tolerances (fp64, relative to the field's max-norm): 1e-12 per step for polynomial
integrands, 1e-10 after O(10^2..10^3) steps, 1e-9 with sqrt/cbrt-heavy terms.
"""
import numpy as np
import pytest

from thetis_b200.mesh import (rectangle_mesh, periodic_rectangle_mesh, unit_square_mesh, delaunay_mesh, sfc_renumber,
                              FACET_NODES)
from oracle import swe_oracle as O

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _solver(mesh, bath, **opts):
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, as_shim_mesh
    sm = as_shim_mesh(mesh)
    P1 = FunctionSpace(sm, "CG", 1)
    b = Function(P1, name="Bathymetry")
    if callable(bath):
        b.interpolate(bath)
    else:
        b.assign(bath)
    s = solver2d.FlowSolver2d(sm, b)
    s.options.swe_timestepper_type = "SSPRK33"
    s.options.swe_timestepper_options.use_automatic_timestep = False
    s.options.tracer_timestepper_options.use_automatic_timestep = False
    s.options.no_exports = True
    s.options.update(opts)
    return s, P1


def _fields_nodal(solver, mesh):
    """host solution in oracle layout (cells of `mesh`, local CCW nodes)"""
    uv = solver.fields.uv_2d.dat.data_ro.reshape(mesh.n_cells, 3, 2).copy()
    eta = solver.fields.elev_2d.dat.data_ro.reshape(mesh.n_cells, 3).copy()
    return uv, eta


def test_channel_demo_config1():
    """BASELINE config 1: demo_2d_channel 40x25, nonlinear, closed, Gaussian hump, SSPRK33, 60 steps"""
    mesh = rectangle_mesh(40, 25, 40e3, 2e3)
    dt, nsteps = 0.5, 60          # cells are 1000 m x 80 m: CFL-stable step
    s, P1 = _solver(mesh, 20.0, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=dt * 20)
    ic = lambda x, y: 2.0 * np.exp(-((x - 20e3) / 4000.0) ** 2)
    s.assign_initial_conditions(elev=ic)
    s.iterate()
    assert s.iteration == nsteps
    uv_g, eta_g = _fields_nodal(s, mesh)
    orc = O.SWEOracle(mesh, 20.0)
    eta = O.interpolate(mesh, ic)
    uv = np.zeros(eta.shape + (2,))
    st = O.ShuOsherStepper(orc, [uv, eta], dt)
    for i in range(nsteps):
        st.advance(i * dt)
    assert _rel(eta_g, eta) < 1e-10 and _rel(uv_g, uv) < 1e-10
    # print_state norms come from the device reduction kernel
    assert abs(s.last_norms[0] - O.l2_norm(mesh, eta)) / O.l2_norm(mesh, eta) < 1e-12
    assert abs(s.last_norms[1] - O.l2_norm(mesh, uv)) / O.l2_norm(mesh, uv) < 1e-12


def test_eta_norm_printed_at_t0():
    """demos/demo_2d_channel.py:95: 'eta norm: 6251.2574' on the 25x2 mesh, now from the device reduction"""
    mesh = rectangle_mesh(25, 2, 40e3, 2e3)
    s, _ = _solver(mesh, 20.0, timestep=1.0, simulation_end_time=1.0)
    s.assign_initial_conditions(elev=lambda x, y: 2.0 * np.exp(-((x - 20e3) / 4000.0) ** 2))
    s.print_state(0.0)
    assert abs(s.last_norms[0] - 6251.2574) < 5e-5


def test_wave_equation_config2_convergence_and_parity():
    """BASELINE config 2 (waveEq2d standing wave, linear): fp64 convergence on 32^2..128^2 + parity vs oracle at 32^2"""
    lx, depth, g = 44294.46, 50.0, 9.81
    c = np.sqrt(g * depth)
    T = lx / c
    errs = []
    for n in (32, 64, 128):
        mesh = rectangle_mesh(n, n, lx, lx)
        nsteps = 40 * n
        dt = T / nsteps
        s, _ = _solver(mesh, depth, timestep=dt, simulation_end_time=T - 0.1 * dt, simulation_export_time=T,
                       use_nonlinear_equations=False)
        ic = lambda x, y: -np.cos(2 * np.pi * x / lx)
        s.assign_initial_conditions(elev=ic)
        s.iterate()
        assert s.iteration == nsteps
        uv_g, eta_g = _fields_nodal(s, mesh)
        errs.append(O.l2_error(mesh, eta_g, ic) / lx)
        if n == 32:
            orc = O.SWEOracle(mesh, depth, options=dict(use_nonlinear_equations=False))
            eta = O.interpolate(mesh, ic)
            uv = np.zeros(eta.shape + (2,))
            st = O.ShuOsherStepper(orc, [uv, eta], dt)
            for i in range(nsteps):
                st.advance(i * dt)
            assert _rel(eta_g, eta) < 1e-10
    assert np.log2(errs[0] / errs[1]) > 1.8 and np.log2(errs[1] / errs[2]) > 1.8


def test_stommel_config3_unstructured():
    """BASELINE config 3 style: unstructured mesh, linear, beta-plane Coriolis, wind stress, linear drag"""
    L = 1.0e6
    mesh = delaunay_mesh(1500, L, L, seed=0)
    dt, nsteps = 100.0, 40
    s, P1 = _solver(mesh, 1000.0, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=dt * nsteps,
                    use_nonlinear_equations=False)
    from thetis_b200.shim import Function, FunctionSpace, Constant
    f = Function(P1).interpolate(lambda x, y: 1e-4 + 2e-11 * y)
    P1v = FunctionSpace(s.mesh2d, "CG", 1, value_size=2)
    tau = Function(P1v).interpolate(lambda x, y: (0.1 * np.sin(np.pi * (y / L - 0.5)), 0 * y))
    s.options.coriolis_frequency = f
    s.options.wind_stress = tau
    s.options.linear_drag_coefficient = Constant(1e-6)
    s.assign_initial_conditions()
    s.iterate()
    uv_g, eta_g = _fields_nodal(s, mesh)
    x = mesh.coords[mesh.cells]
    orc = O.SWEOracle(mesh, 1000.0, options=dict(use_nonlinear_equations=False),
                      fields={"coriolis": 1e-4 + 2e-11 * x[..., 1],
                              "wind_stress": np.stack([0.1 * np.sin(np.pi * (x[..., 1] / L - 0.5)), 0 * x[..., 1]], -1),
                              "linear_drag_coefficient": 1e-6})
    eta = np.zeros(x.shape[:2])
    uv = np.zeros(eta.shape + (2,))
    st = O.ShuOsherStepper(orc, [uv, eta], dt)
    for i in range(nsteps):
        st.advance(i * dt)
    assert _rel(eta_g, eta) < 1e-10 and _rel(uv_g, uv) < 1e-10


def test_time_dependent_boundary_forcing():
    """update_forcings mutates a Constant and a Function between stages (rungekutta.py:933-934)"""
    from thetis_b200.shim import Function, Constant
    mesh = rectangle_mesh(20, 6, 20e3, 6e3)
    dt, nsteps = 4.0, 30
    s, P1 = _solver(mesh, lambda x, y: 15.0 + 0.0002 * x, timestep=dt, simulation_end_time=dt * nsteps,
                    simulation_export_time=dt * nsteps)
    tide = Function(P1, name="tide")
    un = Constant(0.0)
    s.bnd_functions["shallow_water"] = {1: {"elev": tide, "uv": Constant((0.0, 0.0))}, 2: {"un": un}}
    s.options.manning_drag_coefficient = Constant(0.02)
    times = []

    def update_forcings(t):
        times.append(t)
        tide.interpolate(lambda x, y: 0.5 * np.sin(2 * np.pi * t / 600.0) * (1 + y / 6e3))
        un.assign(0.05 * np.sin(2 * np.pi * t / 300.0))

    s.assign_initial_conditions()
    s.iterate(update_forcings=update_forcings)
    # forcings are requested at t + c_i dt, c = [0, 1, 1/2] (rungekutta.py:346,933)
    assert np.allclose(times[:3], [0.0, dt, 0.5 * dt])
    uv_g, eta_g = _fields_nodal(s, mesh)
    x = mesh.coords[mesh.cells]
    bnd = {1: {"elev": None, "uv": (0.0, 0.0)}, 2: {"un": 0.0}}
    orc = O.SWEOracle(mesh, 15.0 + 0.0002 * x[..., 0], fields={"manning_drag_coefficient": 0.02}, bnd_conditions=bnd)

    def uf(t):
        bnd[1]["elev"] = 0.5 * np.sin(2 * np.pi * t / 600.0) * (1 + x[..., 1] / 6e3)
        bnd[2]["un"] = 0.05 * np.sin(2 * np.pi * t / 300.0)
        orc.bnd = bnd

    eta = np.zeros(x.shape[:2])
    uv = np.zeros(eta.shape + (2,))
    st = O.ShuOsherStepper(orc, [uv, eta], dt)
    for i in range(nsteps):
        st.advance(i * dt, uf)
    assert _rel(eta_g, eta) < 1e-9 and _rel(uv_g, uv) < 1e-9


def test_rossby_soliton_short_parity_and_g_mutation():
    """test/swe2d/test_rossby_wave.py set-up (g_grav mutated to 1, f = y, periodic in x, 'uv' BCs), 40 steps"""
    from thetis_b200.equations import physical_constants
    from thetis_b200.shim import Function, Constant
    r = 12
    lx, ly = 48.0, 24.0
    mesh = periodic_rectangle_mesh(2 * r, r, lx, ly, origin=(-lx / 2, -ly / 2))
    g_old = float(physical_constants["g_grav"])
    physical_constants["g_grav"].assign(1.0)
    try:
        dt, nsteps = 0.96 / r, 40
        s, P1 = _solver(mesh, 1.0, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=5.0)
        s.create_function_spaces()
        s.options.coriolis_frequency = Function(s.function_spaces.P1_2d).interpolate(lambda x, y: y)
        for tag in s.mesh2d.exterior_facets.unique_markers:
            s.bnd_functions["shallow_water"][int(tag)] = {"uv": Constant((0.0, 0.0))}
        ic_e = lambda x, y: 0.15 * np.exp(-(x / 3.0) ** 2) * y * np.exp(-0.5 * y * y)
        ic_u = lambda x, y: (0.1 * np.exp(-(x / 3.0) ** 2) * (1 - y * y) * np.exp(-0.5 * y * y), 0.02 * x * np.exp(-(x / 3) ** 2 - 0.5 * y * y))
        s.assign_initial_conditions(elev=ic_e, uv=ic_u)
        s.iterate()
    finally:
        physical_constants["g_grav"].assign(g_old)
    uv_g, eta_g = _fields_nodal(s, mesh)
    x = mesh.coords[mesh.cells]
    bnd = {m: {"uv": (0.0, 0.0)} for m in mesh.unique_markers()}
    orc = O.SWEOracle(mesh, 1.0, fields={"coriolis": x[..., 1]}, bnd_conditions=bnd, g_grav=1.0)
    eta = ic_e(x[..., 0], x[..., 1])
    uv = np.stack(ic_u(x[..., 0], x[..., 1]), -1)
    st = O.ShuOsherStepper(orc, [uv, eta], dt)
    for i in range(nsteps):
        st.advance(i * dt)
    assert _rel(eta_g, eta) < 1e-10 and _rel(uv_g, uv) < 1e-10


@pytest.mark.parametrize("kind", ["linear", "jump"])
@pytest.mark.parametrize("direction", ["x", "y", "xy"])
def test_limiter_matches_oracle_and_reference_criteria(kind, direction):
    """test/slopelimiter/test_slopelimiter.py on UnitSquareMesh(5,5) through VertexBasedP1DGLimiter.apply"""
    from thetis_b200.limiter import VertexBasedP1DGLimiter
    from thetis_b200.shim import Function, FunctionSpace, as_shim_mesh
    mesh = unit_square_mesh(5, 5)
    sm = as_shim_mesh(mesh)
    p1dg = FunctionSpace(sm, "DG", 1)
    x = mesh.coords[mesh.cells]
    coord = {"x": x[..., 0], "y": x[..., 1], "xy": x[..., 0] + 0.5 * x[..., 1] - 0.25}[direction]
    q0 = coord if kind == "linear" else 0.5 + 0.5 * np.tanh(20 * (coord - 0.5))
    tracer = Function(p1dg, name="tracer")
    tracer.dat.data[:] = q0.reshape(-1)
    VertexBasedP1DGLimiter(p1dg).apply(tracer)
    q = tracer.dat.data_ro.reshape(-1, 3)
    ref = O.vertex_based_limiter(mesh, q0)
    assert np.abs(q - ref).max() < 1e-14
    area = mesh.cell_area()
    if direction != "xy":
        if kind == "linear":
            assert np.sqrt((area[:, None] * (q - q0) ** 2).sum()) < 1e-12
        else:
            assert abs((area * q.mean(1)).sum() - (area * q0.mean(1)).sum()) < 1e-12
            assert q.min() > -2e-5


def test_coupled_swe_tracer_limiter_config4():
    """BASELINE config 4 style: SWE(SSPRK33) -> tracer(SSPRK33) -> limiter each step, vs the oracle in the same order"""
    from thetis_b200.shim import Constant
    lx, ly = 18e3, 4e3
    mesh = rectangle_mesh(18, 4, lx, ly)
    dt, nsteps = 5.0, 40
    bath_fn = lambda x, y: 10.0 + 2.0 * np.cos(2 * np.pi * x / lx)
    s, P1 = _solver(mesh, bath_fn, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=dt * nsteps)
    s.options.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d")
    s.options.use_limiter_for_tracers = True
    s.bnd_functions["shallow_water"] = {1: {"elev": Constant(0.2)}}
    s.bnd_functions["tracer"] = {1: {"value": Constant(4.0)}}
    ic_e = lambda x, y: 1.0 * np.cos(np.pi * x / lx)
    ic_c = lambda x, y: 4.5 + 2.0 * (np.abs(x - lx / 2) < 3e3)
    s.assign_initial_conditions(elev=ic_e, tracer=ic_c)
    s.iterate()
    uv_g, eta_g = _fields_nodal(s, mesh)
    c_g = s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)
    x = mesh.coords[mesh.cells]
    orc = O.SWEOracle(mesh, bath_fn(x[..., 0], x[..., 1]), bnd_conditions={1: {"elev": 0.2}})
    trc = O.TracerOracle(orc, bnd_conditions={1: {"value": 4.0}})
    eta = ic_e(x[..., 0], x[..., 1])
    uv = np.zeros(eta.shape + (2,))
    c = ic_c(x[..., 0], x[..., 1]) + 0.0
    ss = O.ShuOsherStepper(orc, [uv, eta], dt)
    ts = O.ShuOsherStepper(trc, [c], dt)
    for i in range(nsteps):
        ss.advance(i * dt)
        trc.set_velocity(uv, eta)
        ts.advance(i * dt)
        c[...] = O.vertex_based_limiter(mesh, c)
    assert _rel(eta_g, eta) < 1e-10 and _rel(uv_g, uv) < 1e-10
    assert _rel(c_g, c) < 1e-10
    assert c_g.max() <= 6.5 + 1e-9 and c_g.min() >= 4.0 - 1e-9


def test_tracer_only_solid_body_rotation():
    """demos/demo_2d_tracer.py style: tracer_only, prescribed velocity on the host, no limiter"""
    mesh = unit_square_mesh(16, 16)
    dt, nsteps = np.pi / 300.0, 20
    s, P1 = _solver(mesh, 1.0, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=1.0,
                    tracer_only=True, use_limiter_for_tracers=False)
    s.options.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d")
    ic_c = lambda x, y: 1.0 + np.exp(-((x - 0.25) ** 2 + (y - 0.5) ** 2) / 0.01)
    ic_u = lambda x, y: (0.5 - y, x - 0.5)
    s.assign_initial_conditions(uv=ic_u, tracer_2d=ic_c)
    s.iterate()
    c_g = s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)
    x = mesh.coords[mesh.cells]
    orc = O.SWEOracle(mesh, 1.0)
    trc = O.TracerOracle(orc)
    uv = np.stack(ic_u(x[..., 0], x[..., 1]), -1)
    trc.set_velocity(uv, np.zeros(x.shape[:2]))
    c = ic_c(x[..., 0], x[..., 1])
    ts = O.ShuOsherStepper(trc, [c], dt)
    for i in range(nsteps):
        ts.advance(i * dt)
    assert _rel(c_g, c) < 1e-11


def test_unsupported_configurations_raise():
    from thetis_b200.shim import Constant
    mesh = rectangle_mesh(4, 4, 1.0, 1.0)
    s, _ = _solver(mesh, 1.0, timestep=0.01, simulation_end_time=0.01)
    s.options.nikuradse_bed_roughness = Constant(1.0)
    s.options.manning_drag_coefficient = Constant(0.02)
    with pytest.raises(Exception, match="Cannot set both Nikuradse"):      # shallowwater_eq.py:690-694
        s.assign_initial_conditions()
    s2, _ = _solver(mesh, 1.0, timestep=0.01, simulation_end_time=0.01)
    s2.bnd_functions["shallow_water"] = {1: {"bogus": Constant(1.0)}}
    with pytest.raises(Exception, match="Invalid boundary tag"):
        s2.assign_initial_conditions()
    s3, _ = _solver(mesh, 1.0, timestep=0.01, simulation_end_time=0.01)
    s3.options.swe_timestepper_type = "CrankNicolson"
    with pytest.raises(NotImplementedError):
        s3.assign_initial_conditions()


def test_fused_stage_integrals_match_separate_reduction():
    """tb_stage_integrals: the diagnostics reduced in the last RK stage's epilogue equal tb_swe_integrals of the new
    state (different summation order: 1e-13), through SSPRK33 and through the Butcher-form ERKLSPUM2"""
    import torch
    from thetis_b200 import rungekutta
    mesh = delaunay_mesh(1200, 2.0e4, 1.5e4, seed=5)
    for name in ("SSPRK33", "ERKLSPUM2"):
        s, P1 = _solver(mesh, lambda x, y: 12.0 + 3.0 * np.sin(x / 3e3), timestep=0.5, simulation_end_time=1.0,
                        swe_timestepper_type=name)
        s.assign_initial_conditions(elev=lambda x, y: 0.4 * np.cos(x / 2e3) * np.sin(y / 3e3),
                                    uv=lambda x, y: (0.1 * np.sin(y / 2e3), 0.05 * np.cos(x / 4e3)))
        ts = s.timestepper
        assert isinstance(ts, getattr(rungekutta, name))
        fused = torch.zeros(4, dtype=torch.float64, device=ts.engine.device)
        ts.fused_norms = fused
        for i in range(3):
            ts.advance(i * 0.5)
        sep = torch.zeros(4, dtype=torch.float64, device=ts.engine.device)
        ts.engine.swe_integrals(ts.device_state(), sep)
        f, g = fused.cpu().numpy(), sep.cpu().numpy()
        assert np.all(np.abs(f - g) <= 1e-13 * np.abs(g)), (f, g)
        assert g[0] > 0 and g[1] > 0 and g[3] > 0


# ---------------------------------------------------------------- round 2: boundary gaps
def test_ufl_expression_boundary_data_north_sea_style():
    """examples/north_sea/model_config.py:181-192: 'elev' = elev_ramp * elev_tide_2d with
    elev_ramp = conditional(bnd_time < ramp_t, bnd_time / ramp_t, 1.0), bnd_time and the tidal Function re-assigned in
    update_forcings.  The expression is affine in its Function operand, so nodal evaluation is exact."""
    from thetis_b200.shim import Constant, Function, FunctionSpace, conditional
    mesh = rectangle_mesh(16, 6, 16e3, 6e3)
    dt, nsteps = 4.0, 30
    bath_fn = lambda x, y: 12.0 + 2.0 * np.cos(x / 3e3)
    s, P1 = _solver(mesh, bath_fn, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=dt * nsteps)
    s.options.manning_drag_coefficient = Constant(0.02)
    elev_tide = Function(P1, name="Tidal elevation")
    bnd_time = Constant(0.0)
    ramp_t = 60.0
    elev_ramp = conditional(bnd_time < ramp_t, bnd_time / ramp_t, 1.0)
    s.bnd_functions["shallow_water"] = {1: {"elev": elev_ramp * elev_tide, "uv": Constant(np.array([0.0, 0.0]))}}
    tide = lambda t: (lambda x, y: 0.5 * np.sin(2 * np.pi * t / 300.0 + y / 4e3))

    def update_forcings(t):
        bnd_time.assign(t)
        elev_tide.interpolate(tide(t))
    update_forcings(0.0)
    s.assign_initial_conditions(elev=lambda x, y: 0.0 * x)
    s.iterate(update_forcings=update_forcings)
    uv_g, eta_g = _fields_nodal(s, mesh)
    x = mesh.coords[mesh.cells]
    orc = O.SWEOracle(mesh, bath_fn(x[..., 0], x[..., 1]), fields={"manning_drag_coefficient": 0.02},
                      bnd_conditions={1: {"elev": 0.0, "uv": (0.0, 0.0)}})
    eta = np.zeros(x.shape[:2])
    uv = np.zeros(eta.shape + (2,))

    def uf(t):
        ramp = t / ramp_t if t < ramp_t else 1.0
        orc.bnd = {1: {"elev": ramp * tide(t)(x[..., 0], x[..., 1]), "uv": (0.0, 0.0)}}
    st = O.ShuOsherStepper(orc, [uv, eta], dt)
    for i in range(nsteps):
        st.advance(i * dt, uf)
    assert np.abs(eta).max() > 1e-2
    assert _rel(eta_g, eta) < 1e-10 and _rel(uv_g, uv) < 1e-10


def test_nonaffine_expression_is_refused():
    from thetis_b200.shim import Constant, Function
    mesh = rectangle_mesh(4, 4, 1.0, 1.0)
    s, P1 = _solver(mesh, 1.0, timestep=0.01, simulation_end_time=0.01)
    f = Function(P1).interpolate(lambda x, y: x)
    s.bnd_functions["shallow_water"] = {1: {"elev": f * f}}
    with pytest.raises(NotImplementedError, match="affine"):
        s.assign_initial_conditions()


def test_two_tracers_with_different_boundary_markers_do_not_share_slots():
    """each tracer equation is built from its own bnd_conditions dict (solver2d.py:580-598) although both integrators
    share one device context: salt has an inflow 'value' on marker 1, temp has no entry there (closed form)"""
    from thetis_b200.shim import Constant
    lx, ly = 12e3, 3e3
    mesh = rectangle_mesh(12, 3, lx, ly)
    dt, nsteps = 5.0, 20
    s, P1 = _solver(mesh, 10.0, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=dt * nsteps)
    s.options.add_tracer_2d("salt_2d", "Salinity", "Salinity2d")
    s.options.add_tracer_2d("temp_2d", "Temperature", "Temperature2d")
    s.options.use_limiter_for_tracers = False
    s.bnd_functions["shallow_water"] = {1: {"un": Constant(-0.3)}, 2: {"elev": Constant(0.0)}}
    s.bnd_functions["salt"] = {1: {"value": Constant(35.0)}}
    s.bnd_functions["temp"] = {2: {"value": Constant(3.0)}}
    ic_s = lambda x, y: 30.0 + 2.0 * np.sin(x / 2e3)
    ic_t = lambda x, y: 10.0 + 1.0 * np.cos(x / 1.5e3)
    s.assign_initial_conditions(salt=ic_s, temp=ic_t)
    s.iterate()
    sal_g = s.fields.salt_2d.dat.data_ro.reshape(-1, 3)
    tmp_g = s.fields.temp_2d.dat.data_ro.reshape(-1, 3)
    x = mesh.coords[mesh.cells]
    orc = O.SWEOracle(mesh, 10.0, bnd_conditions={1: {"un": -0.3}, 2: {"elev": 0.0}})
    t_s = O.TracerOracle(orc, bnd_conditions={1: {"value": 35.0}})
    t_t = O.TracerOracle(orc, bnd_conditions={2: {"value": 3.0}})
    eta = np.zeros(x.shape[:2])
    uv = np.zeros(eta.shape + (2,))
    cs, ct = ic_s(x[..., 0], x[..., 1]) + 0.0, ic_t(x[..., 0], x[..., 1]) + 0.0
    ss = O.ShuOsherStepper(orc, [uv, eta], dt)
    st_s, st_t = O.ShuOsherStepper(t_s, [cs], dt), O.ShuOsherStepper(t_t, [ct], dt)
    for i in range(nsteps):
        ss.advance(i * dt)
        for trc, stp in ((t_s, st_s), (t_t, st_t)):
            trc.set_velocity(uv, eta)
            stp.advance(i * dt)
    assert _rel(sal_g, cs) < 1e-10 and _rel(tmp_g, ct) < 1e-10


def test_nikuradse_run_through_flowsolver():
    from thetis_b200.shim import Constant
    mesh = rectangle_mesh(10, 4, 5e3, 2e3)
    dt, nsteps = 2.0, 25
    s, P1 = _solver(mesh, 3.0, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=dt * nsteps)
    s.options.nikuradse_bed_roughness = Constant(0.05)
    ic = lambda x, y: 0.3 * np.cos(np.pi * x / 5e3)
    s.assign_initial_conditions(elev=ic)
    s.iterate()
    uv_g, eta_g = _fields_nodal(s, mesh)
    orc = O.SWEOracle(mesh, 3.0, fields={"nikuradse_bed_roughness": 0.05})
    eta = O.interpolate(mesh, ic)
    uv = np.zeros(eta.shape + (2,))
    st = O.ShuOsherStepper(orc, [uv, eta], dt)
    for i in range(nsteps):
        st.advance(i * dt)
    assert _rel(eta_g, eta) < 1e-10 and _rel(uv_g, uv) < 1e-10


def test_wetting_drying_alpha_function_through_flowsolver():
    """wetting_and_drying_alpha as a P1 Function (solver2d.py:279-287)"""
    from thetis_b200.shim import Function
    mesh = rectangle_mesh(12, 5, 6e3, 2.5e3)
    dt, nsteps = 1.0, 20
    bath_fn = lambda x, y: 1.0 + 1.5 * np.cos(np.pi * x / 6e3)       # dry towards x = lx
    s, P1 = _solver(mesh, bath_fn, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=dt * nsteps,
                    use_wetting_and_drying=True)
    al_fn = lambda x, y: 0.3 + 0.2 * np.sin(x / 1e3) ** 2
    s.options.wetting_and_drying_alpha = Function(P1).interpolate(al_fn)
    ic = lambda x, y: 0.2 * np.cos(np.pi * x / 6e3)
    s.assign_initial_conditions(elev=ic)
    s.iterate()
    uv_g, eta_g = _fields_nodal(s, mesh)
    x = mesh.coords[mesh.cells]
    orc = O.SWEOracle(mesh, bath_fn(x[..., 0], x[..., 1]),
                      options=dict(use_wetting_and_drying=True, wetting_and_drying_alpha=al_fn(x[..., 0], x[..., 1])))
    eta = O.interpolate(mesh, ic)
    uv = np.zeros(eta.shape + (2,))
    st = O.ShuOsherStepper(orc, [uv, eta], dt)
    for i in range(nsteps):
        st.advance(i * dt)
    assert _rel(eta_g, eta) < 1e-9 and _rel(uv_g, uv) < 1e-9


@pytest.mark.parametrize("which", ["unstructured", "periodic", "north_sea"])
def test_patch_staged_limiter_multi_patch_meshes(which):
    """the patch-staged limiter kernel (bounds in shared memory, vertex halo gathered per patch) against the oracle's
    global gather on meshes with many patches: unstructured with boundaries, x-periodic (vertices identified across
    the seam), the tagged North Sea coastline; noisy data so that most cells are limited"""
    import os
    from thetis_b200.limiter import VertexBasedP1DGLimiter
    from thetis_b200.shim import Function, FunctionSpace, as_shim_mesh
    from thetis_b200.mesh import load_npz_mesh
    if which == "unstructured":
        mesh = delaunay_mesh(3000, 5.0, 4.0, seed=2)
    elif which == "periodic":
        mesh = periodic_rectangle_mesh(40, 21, 8.0, 4.0)
    else:
        mesh = load_npz_mesh(os.path.join(os.path.dirname(__file__), "golden", "north_sea_mesh.npz"))
    sm = as_shim_mesh(mesh)
    p1dg = FunctionSpace(sm, "DG", 1)
    rng = np.random.default_rng(4)
    x = mesh.coords[mesh.cells]
    L = np.ptp(mesh.coords, axis=0)
    q0 = np.tanh(4 * (x[..., 0] - mesh.coords[:, 0].mean()) / L[0]) + 0.3 * rng.standard_normal(x.shape[:2])
    tracer = Function(p1dg, name="tracer")
    tracer.dat.data[:] = q0.reshape(-1)
    VertexBasedP1DGLimiter(p1dg).apply(tracer)
    q = tracer.dat.data_ro.reshape(-1, 3)
    ref = O.vertex_based_limiter(mesh, q0)
    assert np.abs(ref - q0).max() > 0.05                       # the limiter did act
    assert np.abs(q - ref).max() < 1e-14


def test_step_graph_with_banked_boundary_data_is_bit_identical():
    """One CUDA graph per step on the reference-facing path: update_forcings still runs for every stage, the tidal
    elevation of stage i goes into bank i of the device boundary arrays, one replay does the step.  Bit-identical to
    the stage-by-stage path; a Constant changing mid-run makes the integrator fall back (and stay correct)."""
    import torch
    from harness.workloads import north_sea_mesh, north_sea_setup
    from harness.runs import SingleSWE
    from thetis_b200.solver2d import physical_constants
    mesh = north_sea_mesh(1)
    setup = north_sea_setup(mesh, wetting_drying=True)
    a = SingleSWE(mesh, setup, wd=True)
    b = SingleSWE(mesh, setup, wd=True)
    b.enable_step_graph()
    assert b.ts.step_graph is not None
    for _ in range(6):
        a.step_e2e()
        b.step_e2e()
    assert b.ts.step_graph is not None                       # still on the graph path
    ua, ea = a.state_nodal()
    ub, eb = b.state_nodal()
    assert np.array_equal(ua, ub) and np.array_equal(ea, eb)
    assert np.abs(ea - setup["eta0"]).max() > 1e-6           # the tide did enter
    g_old = float(physical_constants["g_grav"])
    try:
        physical_constants["g_grav"].assign(9.5)             # not a boundary array: the graph cannot honour it
        for _ in range(3):
            a.step_e2e()
            b.step_e2e()
    finally:
        physical_constants["g_grav"].assign(g_old)
    assert b.ts.step_graph is None                           # fell back to the stage-by-stage path
    ua, ea = a.state_nodal()
    ub, eb = b.state_nodal()
    assert np.array_equal(ua, ub) and np.array_equal(ea, eb)
