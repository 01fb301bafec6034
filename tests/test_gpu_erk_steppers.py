"""
GPU parity of whole runs for the SURVEY 8f rows, driven through the reference-shaped surface
(FlowSolver2d mirror -> integrator classes -> C-ABI -> CUDA) against the numpy oracle:
Butcher-form ERK integrators (rungekutta.py:762-867, 959-980), `timeintegrator.ForwardEuler`
(timeintegrator.py:115-165), horizontal viscosity / tracer diffusion in time stepping, the
conservative tracer form and the device-resident conservation callbacks (callback.py:301-484).
Tolerances (fp64, relative to the field's max-norm): 1e-10 after O(10^2) steps.
"""
import numpy as np
import pytest
from scipy import stats
from scipy.special import erf

from thetis_b200.mesh import rectangle_mesh, delaunay_mesh
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
    s.options.swe_timestepper_options.use_automatic_timestep = False
    s.options.tracer_timestepper_options.use_automatic_timestep = False
    s.options.no_exports = True
    s.options.update(opts)
    return s, P1


def _nodal(solver, mesh):
    uv = solver.fields.uv_2d.dat.data_ro.reshape(mesh.n_cells, 3, 2).copy()
    eta = solver.fields.elev_2d.dat.data_ro.reshape(mesh.n_cells, 3).copy()
    return uv, eta


@pytest.mark.parametrize("name", ["ERKLSPUM2", "ERKLPUM2", "ERKMidpoint", "ERKEuler"])
def test_butcher_form_integrators_with_forcing(name):
    """ERKGeneric: forcings at t + c_i dt, tendencies combined with the a / b rows of the reference's tableaux"""
    from thetis_b200.shim import Function, Constant
    from thetis_b200 import rungekutta
    mesh = rectangle_mesh(20, 6, 20e3, 6e3)
    dt, nsteps = 2.0, 40
    s, P1 = _solver(mesh, lambda x, y: 15.0 + 0.0002 * x, timestep=dt, simulation_end_time=dt * nsteps,
                    simulation_export_time=dt * nsteps, swe_timestepper_type=name)
    tide = Function(P1, name="tide")
    un = Constant(0.0)
    s.bnd_functions["shallow_water"] = {1: {"elev": tide, "uv": Constant((0.0, 0.0))}, 2: {"un": un}}
    s.options.horizontal_viscosity = Constant(20.0)
    times = []

    def update_forcings(t):
        times.append(t)
        tide.interpolate(lambda x, y: 0.5 * np.sin(2 * np.pi * t / 600.0) * (1 + y / 6e3))
        un.assign(0.05 * np.sin(2 * np.pi * t / 300.0))

    s.assign_initial_conditions()
    cls = getattr(rungekutta, name)
    assert isinstance(s.timestepper, cls)
    s.iterate(update_forcings=update_forcings)
    a, b, c, cfl = O.ERK_TABLEAUX[name]
    assert np.allclose(times[:len(c)], [ci * dt for ci in c])
    assert s.timestepper.cfl_coeff == cfl and s.timestepper.n_stages == len(b)
    uv_g, eta_g = _nodal(s, mesh)
    x = mesh.coords[mesh.cells]
    bnd = {1: {"elev": None, "uv": (0.0, 0.0)}, 2: {"un": 0.0}}
    orc = O.SWEOracle(mesh, 15.0 + 0.0002 * x[..., 0], fields={"viscosity_h": 20.0}, bnd_conditions=bnd)

    def uf(t):
        bnd[1]["elev"] = 0.5 * np.sin(2 * np.pi * t / 600.0) * (1 + x[..., 1] / 6e3)
        bnd[2]["un"] = 0.05 * np.sin(2 * np.pi * t / 300.0)
        orc.bnd = bnd

    eta = np.zeros(x.shape[:2])
    uv = np.zeros(eta.shape + (2,))
    st = O.ButcherStepper(orc, [uv, eta], dt, a, b, c)
    for i in range(nsteps):
        st.advance(i * dt, uf)
    assert _rel(eta_g, eta) < 1e-10 and _rel(uv_g, uv) < 1e-10


def test_golden_tableaux_match_product_classes():
    """tableaux of the product classes == numbers obtained by executing the reference's class bodies
    (tests/golden/make_shuosher_golden.py)"""
    import json
    import os
    from thetis_b200 import rungekutta
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "shuosher_ssprk33.json")))
    for name in ("ERKLSPUM2", "ERKLPUM2", "ERKMidpoint"):
        g = gold[name + "Abstract"]
        cls = getattr(rungekutta, name)
        assert np.array_equal(np.array(cls.a, float), np.array(g["a"]))
        assert list(map(float, cls.b)) == g["b"] and list(map(float, cls.c)) == g["c"]
        assert cls.cfl_coeff == g["cfl_coeff"]


def test_forward_euler_lagged_fields_semantics():
    """timeintegrator.ForwardEuler: update_forcings(t + dt); Function coefficients lag one step (fields_old),
    Constants and boundary data are live"""
    from thetis_b200.shim import Function, Constant
    mesh = rectangle_mesh(12, 6, 12e3, 6e3)
    dt, nsteps = 2.0, 25
    s, P1 = _solver(mesh, 12.0, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=dt * nsteps,
                    swe_timestepper_type="ForwardEuler")
    drag = Function(P1, name="linear drag")          # Function-valued coefficient: lags
    elev = Constant(0.0)                              # boundary Constant: live
    s.options.linear_drag_coefficient = drag
    s.bnd_functions["shallow_water"] = {1: {"elev": elev}}
    times = []

    def update_forcings(t):
        times.append(t)
        drag.assign(1e-3 * (1 + np.sin(t / 20.0)))
        elev.assign(0.3 * np.sin(2 * np.pi * t / 200.0))

    s.assign_initial_conditions()
    s.iterate(update_forcings=update_forcings)
    assert np.allclose(times[:3], [dt, 2 * dt, 3 * dt])
    uv_g, eta_g = _nodal(s, mesh)
    x = mesh.coords[mesh.cells]
    bnd = {1: {"elev": 0.0}}
    orc = O.SWEOracle(mesh, 12.0, fields={"linear_drag_coefficient": 0.0}, bnd_conditions=bnd)
    eta = np.zeros(x.shape[:2])
    uv = np.zeros(eta.shape + (2,))
    drag_old = 0.0                                   # fields_old after initialize(): the initial Function (zero)
    for i in range(nsteps):
        t = i * dt
        bnd[1]["elev"] = 0.3 * np.sin(2 * np.pi * (t + dt) / 200.0)
        orc.bnd = bnd
        orc.fields["linear_drag_coefficient"] = drag_old
        ku, ke = orc.tendency(uv, eta, dt=dt)
        uv += ku
        eta += ke
        drag_old = 1e-3 * (1 + np.sin((t + dt) / 20.0))    # update_fields_old at the end of advance()
    assert _rel(eta_g, eta) < 1e-11 and _rel(uv_g, uv) < 1e-11


@pytest.mark.parametrize("stepper", ["SSPRK33", "ForwardEuler"])
def test_horizontal_diffusion_reference_kat(stepper):
    """test/tracerEq/test_h-diffusion_mes_2d.py: erf front, refinements [1, 2, 3], convergence rate > 1.8,
    through the tracer integrator on the GPU (and parity with the oracle on the coarsest mesh)"""
    lx, depth, mu = 20e3, 30.0, 1.0e3
    t0, t1 = 1000.0, 3000.0
    ana = lambda X, t: -erf((X - lx / 2) / np.sqrt(4 * mu * t))
    errs = []
    from thetis_b200.shim import Constant
    for ref in (1, 2, 3):
        ly = 5e3 / ref
        mesh = rectangle_mesh(8 * ref + 1, 1, lx, ly)
        s, P1 = _solver(mesh, depth, use_nonlinear_equations=False, simulation_end_time=t1,
                        simulation_export_time=(t1 - t0) / 8, swe_timestepper_type=stepper,
                        tracer_timestepper_type=stepper, horizontal_velocity_scale=Constant(1.0))
        s.options.swe_timestepper_options.use_automatic_timestep = True
        s.options.tracer_timestepper_options.use_automatic_timestep = True
        s.options.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d", diffusivity=Constant(mu))
        s.options.use_limiter_for_tracers = True
        s.create_equations()
        # L2 projection of the initial condition onto P1DG, like assign_initial_conditions' project()
        x = mesh.coords[mesh.cells]
        lam, w = O.cell_quadrature("dunavant6")
        xq = np.einsum("qa,ca->cq", lam, x[..., 0])
        mref = np.einsum("q,qa,qb->ab", w, lam, lam)
        c0 = np.linalg.solve(mref, np.einsum("q,qa,cq->ca", w, lam, ana(xq, t0)).T).T.copy()
        s.assign_initial_conditions()
        s.fields.tracer_2d.dat.data[:] = c0.reshape(-1)
        ti = s.timestepper.timesteppers["tracer_2d"]
        ti.initialize(s.fields.tracer_2d)
        dt = s.dt
        t, n = t0, 0
        while t < t1 - 1e-8:                         # custom loop advancing the tracer only, as the reference test does
            ti.advance(t)
            t += dt
            n += 1
        ti.sync_to_host()
        c_g = s.fields.tracer_2d.dat.data_ro.reshape(-1, 3).copy()
        errs.append(O.l2_error(mesh, c_g, lambda X, Y: ana(X, t)) / np.sqrt(lx * ly))
        if ref == 1:
            swe = O.SWEOracle(mesh, depth, options=dict(use_nonlinear_equations=False))
            trc = O.TracerOracle(swe, fields={"diffusivity_h": mu})
            trc.set_velocity(np.zeros(x.shape), np.zeros(x.shape[:2]))
            c = c0.copy()
            if stepper == "SSPRK33":
                st = O.ShuOsherStepper(trc, [c], dt)
            else:
                st = O.ButcherStepper(trc, [c], dt, *O.ERK_TABLEAUX["ERKEuler"][:3])
            for i in range(n):
                st.advance(t0 + i * dt)
            assert _rel(c_g, c) < 1e-10
    slope = stats.linregress(np.log10(1.0 / np.array([1.0, 2.0, 3.0])), np.log10(errs)).slope
    assert slope > 1.8, (slope, errs)


@pytest.mark.parametrize("conservative", [False, True])
def test_coupled_run_with_viscosity_diffusion_and_callbacks(conservative):
    """SWE (viscosity, open boundary) -> tracer (diffusion, optional conservative form) [-> limiter], with the
    device-resident volume / tracer-mass / overshoot callbacks evaluated at every export"""
    from thetis_b200.shim import Constant
    from thetis_b200 import callback
    lx, ly = 18e3, 4e3
    mesh = rectangle_mesh(18, 4, lx, ly)
    dt, nsteps = 5.0, 40
    bath_fn = lambda x, y: 10.0 + 2.0 * np.cos(2 * np.pi * x / lx)
    s, P1 = _solver(mesh, bath_fn, timestep=dt, simulation_end_time=dt * nsteps, simulation_export_time=dt * 10,
                    swe_timestepper_type="SSPRK33", tracer_timestepper_type="SSPRK33",
                    check_volume_conservation_2d=True, check_tracer_conservation=True, check_tracer_overshoot=True)
    s.options.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d", diffusivity=Constant(15.0),
                            use_conservative_form=conservative)
    s.options.use_limiter_for_tracers = not conservative
    s.options.horizontal_viscosity = Constant(25.0)
    s.options.use_grad_div_viscosity_term = True
    ic_e = lambda x, y: 1.0 * np.cos(np.pi * x / lx)
    ic_c = lambda x, y: 4.5 + 2.0 * np.exp(-((x - lx / 2) / 3e3) ** 2)
    s.assign_initial_conditions(elev=ic_e, tracer=ic_c)
    s.iterate()
    uv_g, eta_g = _nodal(s, mesh)
    c_g = s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)
    x = mesh.coords[mesh.cells]
    bn = bath_fn(x[..., 0], x[..., 1])
    orc = O.SWEOracle(mesh, bn, options=dict(use_grad_div_viscosity_term=True), fields={"viscosity_h": 25.0})
    trc = O.TracerOracle(orc, fields={"diffusivity_h": 15.0}, options=dict(use_conservative_form=conservative))
    eta = ic_e(x[..., 0], x[..., 1])
    uv = np.zeros(eta.shape + (2,))
    c = ic_c(x[..., 0], x[..., 1]) + 0.0
    ss = O.ShuOsherStepper(orc, [uv, eta], dt)
    ts = O.ShuOsherStepper(trc, [c], dt)
    area = mesh.cell_area()
    mref = (np.ones((3, 3)) + np.eye(3)) / 12.0
    vols, masses = [], []

    def diag():
        vols.append((area * (eta + bn).mean(1)).sum())
        masses.append((area * c.mean(1)).sum() if conservative
                      else (area * np.einsum("ca,ab,cb->c", bn + eta, mref, c)).sum())

    diag()
    for i in range(nsteps):
        ss.advance(i * dt)
        trc.set_velocity(uv, eta)
        ts.advance(i * dt)
        if not conservative:
            c[...] = O.vertex_based_limiter(mesh, c)
        if (i + 1) % 10 == 0:
            diag()
    assert _rel(eta_g, eta) < 1e-10 and _rel(uv_g, uv) < 1e-10 and _rel(c_g, c) < 1e-10
    cbs = {cb.name: cb for cb in s.callbacks["export"]}
    vol_cb, mass_cb, over_cb = cbs["volume2d"], cbs["tracer_2d mass"], cbs["tracer_2d overshoot"]
    assert isinstance(mass_cb, callback.ConservativeTracerMassConservation2DCallback if conservative
                      else callback.TracerMassConservation2DCallback)
    assert len(vol_cb.history) == len(vols) == 5
    for (t, (v, rel)), vo in zip(vol_cb.history, vols):
        assert abs(v - vo) / vo < 1e-12
    for (t, (m, rel)), mo in zip(mass_cb.history, masses):
        assert abs(m - mo) / mo < 1e-11
    # closed basin: volume is conserved to rounding (the tracer integral only up to the weakly imposed u.n = 0)
    assert abs(vol_cb.history[-1][1][1]) < 1e-13
    assert abs(mass_cb.history[-1][1][1]) < 1e-4
    mn, mx, under, over = over_cb.history[-1][1]
    assert abs(mn - c.min()) < 1e-9 and abs(mx - c.max()) < 1e-9
    assert over == max(mx - over_cb.initial_value[1], 0.0)


def test_viscous_spin_down_unstructured():
    """unstructured mesh, variable viscosity Function + Manning: 30 SSPRK33 steps against the oracle"""
    from thetis_b200.shim import Function, Constant
    L = 2.0e4
    mesh = delaunay_mesh(900, L, L, seed=3)
    dt, nsteps = 0.5, 30
    s, P1 = _solver(mesh, lambda x, y: 20.0 + 5.0 * np.sin(x / 4e3), timestep=dt, simulation_end_time=dt * nsteps,
                    simulation_export_time=dt * nsteps, swe_timestepper_type="SSPRK33")
    nu = Function(P1).interpolate(lambda x, y: 40.0 * (1 + 0.5 * np.cos(y / 3e3)))
    s.options.horizontal_viscosity = nu
    s.options.manning_drag_coefficient = Constant(0.02)
    s.options.sipg_factor = Constant(2.0)
    ic_e = lambda x, y: 0.5 * np.exp(-((x - L / 2) ** 2 + (y - L / 2) ** 2) / (3e3) ** 2)
    s.assign_initial_conditions(elev=ic_e)
    s.iterate()
    uv_g, eta_g = _nodal(s, mesh)
    x = mesh.coords[mesh.cells]
    orc = O.SWEOracle(mesh, 20.0 + 5.0 * np.sin(x[..., 0] / 4e3), options=dict(sipg_factor=2.0),
                      fields={"viscosity_h": 40.0 * (1 + 0.5 * np.cos(x[..., 1] / 3e3)),
                              "manning_drag_coefficient": 0.02})
    eta = ic_e(x[..., 0], x[..., 1])
    uv = np.zeros(eta.shape + (2,))
    st = O.ShuOsherStepper(orc, [uv, eta], dt)
    for i in range(nsteps):
        st.advance(i * dt)
    assert _rel(eta_g, eta) < 1e-10 and _rel(uv_g, uv) < 1e-10


def test_modesplit_equations_through_the_integrator():
    """An explicit integrator built on a ModeSplit2DEquations descriptor: advection off, fields the equation has no
    term for (Manning, wind, viscosity) are ignored exactly like the reference's term list does"""
    from thetis_b200.shim import Function, FunctionSpace, MixedFunctionSpace, Constant, as_shim_mesh
    from thetis_b200.equations import ModeSplit2DEquations, DepthExpression
    from thetis_b200.options import ModelOptions2d
    from thetis_b200 import rungekutta
    mesh = rectangle_mesh(16, 10, 16e3, 10e3)
    sm = as_shim_mesh(mesh)
    P1 = FunctionSpace(sm, "CG", 1)
    bath = Function(P1).interpolate(lambda x, y: 20.0 + 3.0 * np.sin(x / 3e3))
    U = FunctionSpace(sm, "DG", 1, value_size=2)
    H = FunctionSpace(sm, "DG", 1)
    V = MixedFunctionSpace([U, H])
    sol = Function(V)
    uvf, ef = sol.subfunctions
    ef.interpolate(lambda x, y: 0.3 * np.cos(np.pi * x / 16e3))
    opts = ModelOptions2d()
    eq = ModeSplit2DEquations(V, DepthExpression(bath), opts)
    fields = {"coriolis": Constant(1.0e-4), "manning_drag_coefficient": Constant(0.05),
              "wind_stress": Constant((0.3, 0.1)), "viscosity_h": Constant(100.0),
              "momentum_source": Constant((1e-5, -2e-5))}
    dt, nsteps = 5.0, 20
    ts = rungekutta.ERKLPUM2(eq, sol, fields, dt, opts.swe_timestepper_options, {})
    for i in range(nsteps):
        ts.advance(i * dt)
    uv_g = uvf.dat.data_ro.reshape(mesh.n_cells, 3, 2).copy()
    e_g = ef.dat.data_ro.reshape(mesh.n_cells, 3).copy()
    x = mesh.coords[mesh.cells]
    orc = O.SWEOracle(mesh, 20.0 + 3.0 * np.sin(x[..., 0] / 3e3), options=dict(include_momentum_advection=False),
                      fields={"coriolis": 1.0e-4, "momentum_source": (1e-5, -2e-5)})
    eta = 0.3 * np.cos(np.pi * x[..., 0] / 16e3)
    uv = np.zeros(eta.shape + (2,))
    st = O.ButcherStepper(orc, [uv, eta], dt, *O.ERK_TABLEAUX["ERKLPUM2"][:3])
    for i in range(nsteps):
        st.advance(i * dt)
    assert _rel(e_g, eta) < 1e-11 and _rel(uv_g, uv) < 1e-11


def test_coupled_two_stage_rk_2d_mode_loop():
    """the 2-D side of CoupledTwoStageRK (coupled_timeintegrator.py:563-715): ModeSplit2DEquations
    stepped stage by stage through `swe2d.solve_stage`, with the coupling term split_residual_2d (+ momentum_source_2d)
    re-assigned by a (synthetic) 3-D side after every stage from the stage solution it reads on the host"""
    from thetis_b200.shim import Function, FunctionSpace, MixedFunctionSpace, Constant, as_shim_mesh
    from thetis_b200.equations import ModeSplit2DEquations, DepthExpression
    from thetis_b200.options import ModelOptions2d
    from thetis_b200.solver2d import AttrDict
    from thetis_b200.coupled_timeintegrator import CoupledTwoStageRK2D
    from thetis_b200 import rungekutta
    mesh = rectangle_mesh(14, 9, 14e3, 9e3)
    sm = as_shim_mesh(mesh)
    P1 = FunctionSpace(sm, "CG", 1)
    bath_fn = lambda x, y: 25.0 + 4.0 * np.sin(x / 2.5e3) * np.cos(y / 3e3)
    bath = Function(P1).interpolate(bath_fn)
    U, H = FunctionSpace(sm, "DG", 1, value_size=2), FunctionSpace(sm, "DG", 1)
    V = MixedFunctionSpace([U, H])
    sol = Function(V)
    uvf, ef = sol.subfunctions
    ic = lambda x, y: 0.4 * np.cos(np.pi * x / 14e3) * np.cos(np.pi * y / 9e3)
    ef.interpolate(ic)
    opts = ModelOptions2d()
    opts.coriolis_frequency = Constant(1.0e-4)
    opts.momentum_source_2d = Constant((2e-5, -1e-5))
    dt, nsteps = 6.0, 15
    solver = AttrDict()
    solver.options = opts
    solver.fields = AttrDict(solution_2d=sol, split_residual_2d=Function(U, name="split_residual_2d"))
    solver.equations = AttrDict(sw=ModeSplit2DEquations(V, DepthExpression(bath), opts))
    solver.dt = dt
    solver.bnd_functions = {"shallow_water": {1: {"elev": Constant(0.1)}}}
    calls = []

    class Mode3D:
        """stand-in for the reference's 3-D side: reads the 2-D stage solution, returns the coupling term"""
        def prepare_stage(self, i, t, uf3d):
            calls.append(("prepare", i))

        def solve_stage(self, i):
            calls.append(("solve", i))

        def update_2d_coupling(self, last):
            # split_residual_2d = uv_dav_2d / dt (:65-70); here a damping of the 2-D stage velocity
            solver.fields.split_residual_2d.dat.data[...] = -0.02 / dt * uvf.dat.data_ro + 1e-6 * (1 + int(last))
            calls.append(("couple", last))

    ti = CoupledTwoStageRK2D(solver, mode3d=Mode3D())
    assert isinstance(ti.timesteppers.swe2d, rungekutta.SSPRK22) and ti.n_stages == 2
    for i in range(nsteps):
        ti.advance(i * dt)
    assert calls[:6] == [("prepare", 0), ("solve", 0), ("couple", False), ("prepare", 1), ("solve", 1), ("couple", True)]
    uv_g = uvf.dat.data_ro.reshape(mesh.n_cells, 3, 2).copy()
    e_g = ef.dat.data_ro.reshape(mesh.n_cells, 3).copy()
    # the same loop on the oracle
    x = mesh.coords[mesh.cells]
    orc = O.SWEOracle(mesh, bath_fn(x[..., 0], x[..., 1]), options=dict(include_momentum_advection=False),
                      fields={"coriolis": 1.0e-4, "momentum_source": np.zeros((mesh.n_cells, 3, 2)) + (2e-5, -1e-5)},
                      bnd_conditions={1: {"elev": 0.1}})
    eta = ic(x[..., 0], x[..., 1]) + 0.0
    uv = np.zeros(eta.shape + (2,))
    st = O.ShuOsherStepper(orc, [uv, eta], dt, a=[[0, 0], [1.0, 0]], b=[0.5, 0.5], c=[0, 1.0])
    for i in range(nsteps):
        for k in range(2):
            st.solve_stage(k, i * dt)
            orc.fields["momentum_source"] = (-0.02 / dt * uv + 1e-6 * (1 + int(k == 1))) + (2e-5, -1e-5)
    assert np.abs(uv).max() > 1e-3
    assert _rel(e_g, eta) < 1e-11 and _rel(uv_g, uv) < 1e-11


def test_nonblocking_export_staging():
    """8f-3: stage_export() copies the solution to pinned host memory on a side stream while the time loop keeps
    running; what the handle returns is the solution AT THE TIME OF THE CALL, bit-identical to a blocking sync"""
    import torch
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, as_shim_mesh
    mesh = delaunay_mesh(4000, 2.0e4, 1.5e4, seed=11)
    sm = as_shim_mesh(mesh)
    b = Function(FunctionSpace(sm, "CG", 1)).interpolate(lambda x, y: 15.0 + 3.0 * np.sin(x / 3e3))
    s = solver2d.FlowSolver2d(sm, b)
    s.options.swe_timestepper_options.use_automatic_timestep = False
    s.options.update(dict(timestep=1.0, simulation_end_time=10.0, no_exports=True))
    s.assign_initial_conditions(elev=lambda x, y: 0.3 * np.cos(x / 2e3) * np.sin(y / 3e3))
    ts = s.timestepper
    handles, blocking = [], []
    for i in range(6):
        ts.advance(i * 1.0)
        handles.append(ts.stage_export())               # returns at once; the loop goes on
        if i in (1, 3):                                 # reference copies, taken the blocking way
            ts.sync_to_host()
            blocking.append((i, s.fields.uv_2d.dat.data_ro.copy(), s.fields.elev_2d.dat.data_ro.copy()))
        if i >= 1:
            # double buffering: a handle stays valid until the second-next stage_export()
            uv_h, eta_h = handles[i - 1].wait()
            for j, uv_b, eta_b in blocking:
                if j == i - 1:
                    assert np.array_equal(uv_h, uv_b.reshape(uv_h.shape)) and np.array_equal(eta_h, eta_b)
    assert handles[-1].wait()[1].shape == s.fields.elev_2d.dat.data_ro.shape
    assert all(h.ready() for h in handles)


def test_flowsolver_export_consumers_run_behind_the_time_loop():
    """FlowSolver2d.add_export_consumer: every export is staged without blocking and delivered in order with the
    fields of ITS export time (compared with a second, blocking, run)"""
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, as_shim_mesh

    def make():
        mesh = rectangle_mesh(20, 8, 2.0e4, 8e3)
        sm = as_shim_mesh(mesh)
        b = Function(FunctionSpace(sm, "CG", 1)).interpolate(lambda x, y: 12.0 + 2.0 * np.cos(x / 3e3))
        s = solver2d.FlowSolver2d(sm, b)
        s.options.swe_timestepper_options.use_automatic_timestep = False
        s.options.update(dict(timestep=2.0, simulation_end_time=40.0, simulation_export_time=8.0, no_exports=True))
        s.assign_initial_conditions(elev=lambda x, y: 0.3 * np.cos(np.pi * x / 2.0e4))
        return s
    got = []
    s1 = make()
    s1.add_export_consumer(lambda t, i, arr: got.append((t, i, arr["uv_2d"].copy(), arr["elev_2d"].copy())))
    s1.iterate()
    ref = []
    s2 = make()
    s2.iterate(export_func=lambda: ref.append((s2.simulation_time, s2.i_export, s2.fields.uv_2d.dat.data_ro.copy(),
                                               s2.fields.elev_2d.dat.data_ro.copy())))
    assert [g[:2] for g in got] == [r[:2] for r in ref] and len(got) == 6        # t = 0, 8, ..., 40
    for g, r in zip(got, ref):
        assert np.array_equal(g[2], r[2].reshape(g[2].shape)) and np.array_equal(g[3], r[3])
