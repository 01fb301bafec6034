"""
The drop-in boundary exercised with the REFERENCE'S OWN OBJECTS, on the CPU.

`thetis_b200.rungekutta.SSPRK33 / ERKLSPUM2 / ... / ForwardEuler` are constructed exactly as
`FlowSolver2d.get_swe_timestepper` constructs the reference classes (solver2d.py:541-572):

    integrator(equation, solution, fields, dt, options, bnd_conditions)

with `equation` an instance of the reference's `thetis.shallowwater_eq.ShallowWaterEquations` (imported from the
reference tree, tests/golden/refenv.py), `solution` / `fields` / `bnd_conditions` Firedrake-shaped Functions and Constants
(tests/golden/ufl_lite.py) and the reference's own `physical_constants`.  Everything the host classes do -- mesh
extraction from the Firedrake-shaped mesh, classification of Constants / P1 / P1DG data, node maps through the SFC
renumbering, boundary slots and arrays, per-stage `update_forcings`, version tracking, stage coefficients, buffer
rotation, write-back into `solution.dat.data` -- runs for real; only the C-ABI engine is replaced by
tests/oracle_engine.py, which evaluates each stage with the CPU oracle from what was SENT to it.  The result after N
steps must equal what the reference's own integrator produced from the same objects
(tests/golden/reference_residuals.npz, `step/*`).  The kernels behind the real engine are tied to the same oracle by
the `-m gpu` tests.

Needs the reference tree (build container); skipped elsewhere.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/thetis"),
                                reason="the reference tree only exists in the build container")


@pytest.fixture
def ref(monkeypatch):
    """reference modules on the numpy UFL stand-in + the generator's set-up helpers; the stand-in modules are removed
    from sys.modules afterwards (other tests rely on `import firedrake` failing)"""
    for p in (HERE, os.path.join(HERE, "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in
             ("firedrake", "ufl", "mpi4py", "pyop2", "pyadjoint", "thetis")}
    sys.modules.pop("make_reference_residual_golden", None)
    import refenv
    import make_reference_residual_golden as G          # installs the stand-ins and imports the reference modules
    from thetis_b200 import adaptor
    from oracle_engine import OracleEngine
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: type("S", (), {"synchronize": lambda s: None})())
    engines = []

    def get_engine(self):
        if self.engine is None:
            self.engine = OracleEngine(self.mesh)
            engines.append(self.engine)
        return self.engine
    monkeypatch.setattr(adaptor.MeshAdaptor, "get_engine", get_engine)
    yield types.SimpleNamespace(G=G, engines=engines)
    refenv.uninstall()
    sys.modules.pop("make_reference_residual_golden", None)
    sys.modules.update(saved)


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


STEP_NAMES = ["ssprk33_tidal_constant", "ssprk33_tidal_function_manning", "ssprk33_closed_linear",
              "ssprk33_tidal_ufl_expression",
              "erklspum2_tidal_function", "erklpum2_tidal_constant", "erkmidpoint_viscous", "erkeuler_closed",
              "forward_euler_lagged_drag"]


@pytest.mark.parametrize("name", STEP_NAMES)
def test_b200_integrator_on_reference_objects_reproduces_the_reference_integrator(ref, name):
    import reference_cases as RC
    from thetis_b200 import rungekutta as B
    G = ref.G
    gold = np.load(os.path.join(HERE, "golden", "reference_residuals.npz"))
    spec = RC.STEP_CASES[name]
    case = RC.SWE_CASES[spec["case"]]
    st = G.Setup(case)
    eq, fields, bnd, o = G.swe_equation(st)                     # the reference's ShallowWaterEquations instance
    assert type(eq).__module__ == "thetis.shallowwater_eq"
    seed = 80 + list(RC.STEP_CASES).index(name)
    sol, uv0, eta0 = st.swe_solution(seed)
    assert np.array_equal(uv0, gold[f"step/{name}/uv0"])
    expr_parts = G.install_elev_expression(bnd, spec) if spec["forcing"] == "elev_expression" else None
    topt = types.SimpleNamespace(ad_block_tag=None, solver_parameters={})
    cls = getattr(B, spec.get("integrator", "SSPRK33"))
    ti = cls(eq, sol, fields, spec["dt"], topt, bnd)            # the reference's constructor signature
    ti.initialize(sol)
    U = G.U
    base, drag, drag_base = {}, None, None
    if spec["forcing"] in ("elev_const", "elev_function"):
        for mk, funcs in bnd.items():
            if "elev" in funcs:
                el = funcs["elev"]
                base[mk] = float(el) if isinstance(el, U.Constant) else el.dat.data.copy()
    elif spec["forcing"] == "lagged_drag":
        drag = fields["linear_drag_coefficient"]
        drag_base = drag.dat.data.copy()

    def update_forcings(t):                                     # what a user script does: assign to its own objects
        f = RC.forcing_factor(t)
        if expr_parts is not None:
            G.update_elev_expression(expr_parts, t)
        for mk, b in base.items():
            el = bnd[mk]["elev"]
            if isinstance(el, U.Constant):
                el.assign(b * f)
            else:
                el.dat.data[...] = b * f
                el.dat.dat_version += 1
        if drag is not None:
            drag.dat.data[...] = drag_base * f
            drag.dat.dat_version += 1

    t = 0.0
    for _ in range(spec["n_steps"]):
        ti.advance(t, update_forcings if spec["forcing"] else None)
        t += spec["dt"]
    nt = st.m2.n_cells
    uv = sol.subfunctions[0].dat.data.reshape(nt, 3, 2)         # written back in place by the drop-in
    eta = sol.subfunctions[1].dat.data.reshape(nt, 3)
    eu, ee = _rel(uv, gold[f"step/{name}/uv"]), _rel(eta, gold[f"step/{name}/eta"])
    assert eu < 1e-12 and ee < 1e-12, (eu, ee)
    eng = ref.engines[-1]
    assert eng.n_stage_launches == spec["n_steps"] * ti.n_stages     # one fused launch per stage, nothing else


@pytest.mark.parametrize("name", ["coupled_ssprk33_advection", "coupled_ssprk33_diffusion_unstructured"])
def test_b200_classes_inside_the_reference_coupled_integrator(ref, name):
    """The reference's own `coupled_timeintegrator_2d.GeneralCoupledTimeIntegrator2D` (executed from the reference
    tree) drives the B200 classes handed to it as `integrators` -- the route a Thetis user takes after
    `thetis_b200.install()`: SWE step, then the tracer step on the new velocity, both through the reference's
    `advance()`; the result must equal the run in which the reference drove its own SSPRK33 classes."""
    import reference_cases as RC
    from thetis_b200 import rungekutta as B
    G = ref.G
    gold = np.load(os.path.join(HERE, "golden", "reference_residuals.npz"))
    spec = RC.COUPLED_CASES[name]
    swe_case, tr_case = RC.SWE_CASES[spec["swe"]], RC.TRACER_CASES[spec["tracer"]]
    st = G.Setup(swe_case)
    seed = 120 + list(RC.COUPLED_CASES).index(name)
    solver = G._FakeSolver(st, swe_case, tr_case, spec["dt"], seed)
    cti = G.MODS["coupled_timeintegrator_2d"].GeneralCoupledTimeIntegrator2D(
        solver, {"shallow_water": B.SSPRK33, "tracer": B.SSPRK33})
    assert type(cti.timesteppers.swe2d).__module__ == "thetis_b200.rungekutta"
    cti.initialize(solver.fields.solution_2d)
    bnd = solver.bnd_functions["shallow_water"]
    base = {mk: float(f["elev"]) for mk, f in bnd.items() if "elev" in f} if spec["forcing"] else {}

    def update_forcings(t):
        for mk, b in base.items():
            bnd[mk]["elev"].assign(b * RC.forcing_factor(t))

    t = 0.0
    for _ in range(spec["n_steps"]):
        cti.advance(t, update_forcings if spec["forcing"] else None)
        t += spec["dt"]
    nt = st.m2.n_cells
    sol = solver.fields.solution_2d
    got = dict(uv=sol.subfunctions[0].dat.data.reshape(nt, 3, 2), eta=sol.subfunctions[1].dat.data.reshape(nt, 3),
               c=solver.fields.tracer_2d.dat.data.reshape(nt, 3))
    for k, v in got.items():
        e = _rel(v, gold[f"coupled/{name}/{k}"])
        assert e < 1e-12, (k, e)


def _swe_case_names():
    sys.path.insert(0, HERE)
    import reference_cases as RC
    return list(RC.SWE_CASES)


@pytest.mark.parametrize("name", _swe_case_names())
def test_b200_tendency_on_reference_objects_equals_the_reference_terms(ref, name):
    """Every SWE set-up of tests/reference_cases.py handed to the drop-in as the reference's own equation object plus
    Firedrake-shaped Constants / P1 / P1DG Functions: one forward-Euler step of size 1 (`ERKEuler`, so that
    u1 - u0 = M^-1 R(u0)) must reproduce the tendency the reference's term classes produced.  Covers the adaptor's
    classification of every coefficient and boundary-datum form on objects it did not create."""
    import reference_cases as RC
    from thetis_b200 import rungekutta as B
    G = ref.G
    gold = np.load(os.path.join(HERE, "golden", "reference_residuals.npz"))
    case = RC.SWE_CASES[name]
    st = G.Setup(case)
    g_old = float(G.PC["g_grav"])
    G.PC["g_grav"].assign(case.get("g", g_old))                   # the reference's own Constant, as a script mutates it
    try:
        eq, fields, bnd, o = G.swe_equation(st)
        sol, uv0, eta0 = st.swe_solution(list(RC.SWE_CASES).index(name))
        ti = B.ERKEuler(eq, sol, fields, 1.0, types.SimpleNamespace(ad_block_tag=None, solver_parameters={}), bnd)
        ti.initialize(sol)
        ti.advance(0.0)
    finally:
        G.PC["g_grav"].assign(g_old)
    nt = st.m2.n_cells
    ku = sol.subfunctions[0].dat.data.reshape(nt, 3, 2) - uv0
    ke = sol.subfunctions[1].dat.data.reshape(nt, 3) - eta0
    eu, ee = _rel(ku, gold[f"swe/{name}/ku"]), _rel(ke, gold[f"swe/{name}/ke"])
    assert eu < 1e-11 and ee < 1e-11, (eu, ee)        # the difference u1 - u0 costs a few digits


def _tracer_case_names():
    sys.path.insert(0, HERE)
    import reference_cases as RC
    return list(RC.TRACER_CASES)


@pytest.mark.parametrize("name", _tracer_case_names())
def test_b200_tracer_tendency_on_reference_objects_equals_the_reference_terms(ref, name):
    """The same for the reference's own `TracerEquation2D` (both forms): fields named as
    `FlowSolver2d.get_tracer_timestepper` names them (solver2d.py:575-598), velocity / elevation as host Functions."""
    import reference_cases as RC
    from thetis_b200 import rungekutta as B
    G = ref.G
    U = G.U
    gold = np.load(os.path.join(HERE, "golden", "reference_residuals.npz"))
    case = RC.TRACER_CASES[name]
    st = G.Setup(case)
    depth, opts, o = st.depth_and_options()
    eq = G.treq.TracerEquation2D("tracer_2d", st.H, depth, opts, None)
    sol, uv, eta = st.swe_solution(50 + list(RC.TRACER_CASES).index(name))
    c0 = gold[f"tracer/{name}/c"]
    q = U.Function(st.H, name="tracer_2d")
    q.dat.data[...] = c0.reshape(-1)
    fields = {"uv_2d": sol.subfunctions[0], "elev_2d": sol.subfunctions[1],
              "tracer_advective_velocity_factor": U.Constant(1.0), "lax_friedrichs_tracer_scaling_factor": U.Constant(1.0)}
    for fname, spec in case.get("fields", {}).items():
        fields[f"{fname}-tracer_2d" if fname in ("source", "diffusivity_h") else fname] = st.obj(spec)
    ti = B.ERKEuler(eq, q, fields, 1.0, types.SimpleNamespace(ad_block_tag=None, solver_parameters={}), st.bnd())
    ti.initialize(q)
    ti.advance(0.0)
    kc = q.dat.data.reshape(-1, 3) - c0
    e = _rel(kc, gold[f"tracer/{name}/kc"])
    assert e < 1e-10, e                              # c ~ 1, dc ~ 1e-3: the difference costs three digits


def test_install_rebinds_the_reference_modules_and_the_coupled_run_still_matches(ref):
    """`thetis_b200.install(thetis)` on the REAL `thetis.rungekutta` / `thetis.timeintegrator` / `thetis.limiter` modules
    (imported from the reference tree): the attributes `FlowSolver2d.create_timestepper` looks up at call time
    (solver2d.py:662-672, :535-539) now resolve to the B200 classes, and the reference's coupled integrator fed from
    those attributes reproduces the reference-only run."""
    import importlib
    import reference_cases as RC
    import thetis_b200
    from thetis_b200 import rungekutta as B, limiter as BL
    G = ref.G
    thetis = sys.modules["thetis"]
    importlib.import_module("thetis.limiter")
    rk_mod, ti_mod, lim_mod = thetis.rungekutta, thetis.timeintegrator, thetis.limiter
    saved = {(m, n): getattr(m, n) for m, n in [(rk_mod, "SSPRK33"), (rk_mod, "ERKLSPUM2"), (rk_mod, "ERKLPUM2"),
                                                (rk_mod, "ERKMidpoint"), (rk_mod, "ERKEuler"),
                                                (ti_mod, "ForwardEuler"), (lim_mod, "VertexBasedP1DGLimiter")]}
    assert saved[(rk_mod, "SSPRK33")].__module__ == "thetis.rungekutta"
    try:
        thetis_b200.install(thetis, sync_policy="every_step")
        assert issubclass(rk_mod.SSPRK33, B.SSPRK33) and issubclass(rk_mod.ERKLSPUM2, B.ERKLSPUM2)
        assert issubclass(ti_mod.ForwardEuler, B.ForwardEuler)
        assert issubclass(lim_mod.VertexBasedP1DGLimiter, BL.VertexBasedP1DGLimiter)
        steppers = {"SSPRK33": rk_mod.SSPRK33}                      # what solver2d.py:662-672 builds at call time
        name = "coupled_ssprk33_advection"
        spec = RC.COUPLED_CASES[name]
        swe_case, tr_case = RC.SWE_CASES[spec["swe"]], RC.TRACER_CASES[spec["tracer"]]
        st = G.Setup(swe_case)
        solver = G._FakeSolver(st, swe_case, tr_case, spec["dt"], 120)
        cti = thetis.coupled_timeintegrator_2d.GeneralCoupledTimeIntegrator2D(
            solver, {"shallow_water": steppers["SSPRK33"], "tracer": steppers["SSPRK33"]})
        cti.initialize(solver.fields.solution_2d)
        bnd = solver.bnd_functions["shallow_water"]
        base = {mk: float(f["elev"]) for mk, f in bnd.items() if "elev" in f}
        t = 0.0
        for _ in range(spec["n_steps"]):
            cti.advance(t, lambda tt: [bnd[mk]["elev"].assign(b * RC.forcing_factor(tt)) for mk, b in base.items()])
            t += spec["dt"]
        gold = np.load(os.path.join(HERE, "golden", "reference_residuals.npz"))
        nt = st.m2.n_cells
        assert _rel(solver.fields.tracer_2d.dat.data.reshape(nt, 3), gold[f"coupled/{name}/c"]) < 1e-12
        assert _rel(solver.fields.solution_2d.subfunctions[1].dat.data.reshape(nt, 3), gold[f"coupled/{name}/eta"]) < 1e-12
    finally:
        for (m, n), v in saved.items():
            setattr(m, n, v)


def test_b200_limiter_on_a_reference_shaped_space(ref):
    """`VertexBasedP1DGLimiter(p1dg_space).apply(field)` with a Firedrake-shaped space / Function it did not create:
    the node maps through the SFC renumbering must bring back, dof by dof, what the oracle limiter gives in the
    caller's own cell order (the limiter arithmetic itself is pinned elsewhere)."""
    import reference_cases as RC
    from thetis_b200.limiter import VertexBasedP1DGLimiter
    from oracle.swe_oracle import vertex_based_limiter
    G = ref.G
    st = G.Setup(RC.SWE_CASES["nonlinear_no_lf"])                   # unstructured mesh
    x = st.m2.coords[st.m2.cells]
    rng = np.random.default_rng(3)
    q0 = np.tanh((x[..., 0] - 2.5e3) / 300.0) + 0.3 * rng.standard_normal(x.shape[:2])     # over- and undershoots
    f = G.U.Function(st.H, name="tracer_2d")
    f.dat.data[...] = q0.reshape(-1)
    VertexBasedP1DGLimiter(st.H).apply(f)
    want = vertex_based_limiter(st.m2, q0)
    assert np.abs(want - q0).max() > 0.05                           # the limiter did something
    assert np.abs(f.dat.data.reshape(-1, 3) - want).max() < 1e-13


@pytest.mark.parametrize("name", ["wetting_drying_manning", "wetting_drying_alpha_p1"])
def test_displaced_mass_wetting_drying_step_on_reference_objects(ref, name):
    """`SSPRK33(..., wd_mass='displaced')` on the reference's own ShallowWaterEquations instance with wetting-drying:
    the option reaches the engine, every stage advances the reference's mass functional (the oracle-backed engine does
    what TB_OPT_WD_DISPLACED_MASS does in the kernel) and the run equals the oracle's DisplacedMassShuOsherStepper; the
    displaced volume int (eta + f) is conserved on the closed set-up; the Butcher-form classes refuse the option."""
    import reference_cases as RC
    from thetis_b200 import rungekutta as B
    from thetis_b200 import _lib as L
    from oracle import swe_oracle as O
    import test_oracle_reference_residuals as T
    G = ref.G
    case = RC.SWE_CASES[name]
    st = G.Setup(case)
    eq, fields, bnd, o = G.swe_equation(st)
    sol, uv0, eta0 = st.swe_solution(7)
    topt = types.SimpleNamespace(ad_block_tag=None, solver_parameters={})
    dt, n_steps = 2.0, 3
    ti = B.SSPRK33(eq, sol, fields, dt, topt, bnd, wd_mass="displaced")
    eng = ref.engines[-1]
    assert eng.opt[L.OPT_WD_DISPLACED_MASS] == 1.0
    for i in range(n_steps):
        ti.advance(i * dt)
    nt = st.m2.n_cells
    uv = sol.subfunctions[0].dat.data.reshape(nt, 3, 2)
    eta = sol.subfunctions[1].dat.data.reshape(nt, 3)
    orc = T._swe_oracle(case, st.m2)
    wuv, weta = uv0.copy(), eta0.copy()
    ost = O.DisplacedMassShuOsherStepper(orc, [wuv, weta], dt)
    for i in range(n_steps):
        ost.advance(i * dt)
    assert _rel(uv, wuv) < 1e-11 and _rel(eta, weta) < 1e-11
    # ... and differs from the plain-mass step by far more than that
    puv, peta = uv0.copy(), eta0.copy()
    pst = O.ShuOsherStepper(orc, [puv, peta], dt)
    for i in range(n_steps):
        pst.advance(i * dt)
    assert _rel(eta, peta) > 1e-4
    if not case.get("bnd"):
        v0, v1 = orc.displaced_mass(eta0).sum(), orc.displaced_mass(eta).sum()
        assert abs(v1 - v0) <= 1e-12 * abs(v0)
    with pytest.raises(NotImplementedError, match="displaced"):
        B.ERKLSPUM2(eq, st.swe_solution(7)[0], fields, dt, topt, bnd, wd_mass="displaced")
    # default: plain mass, the option is sent as 0
    B.SSPRK33(eq, st.swe_solution(7)[0], fields, dt, topt, bnd)
    assert ref.engines[-1].opt[L.OPT_WD_DISPLACED_MASS] == 0.0
    # behind an unmodified FlowSolver2d (which constructs the integrators itself): the module-level default
    B.WD_MASS_DEFAULT = "displaced"
    try:
        B.SSPRK33(eq, st.swe_solution(7)[0], fields, dt, topt, bnd)
        assert ref.engines[-1].opt[L.OPT_WD_DISPLACED_MASS] == 1.0
    finally:
        B.WD_MASS_DEFAULT = "plain"


def test_reference_diagnostics_equal_the_closed_forms_the_device_reductions_are_checked_against(ref):
    """`utility.comp_volume_2d` / `comp_tracer_mass_2d` (utility.py:422-443; behind VolumeConservation2DCallback and
    TracerMassConservation2DCallback, callback.py:350-389) executed from the reference tree on the UFL stand-in, against
    the closed P1 forms tests/test_gpu_tracer_terms.py holds the device reductions (tb_swe_integrals /
    tb_tracer_integrals, fused print_state norms) to: int (eta + b) = sum A mean(eta + b), int H c = sum A H^T M c,
    M = (I + 1 1^T) / 12, and the L2 norms of print_state (solver2d.py:955-956)."""
    import reference_cases as RC
    from oracle import swe_oracle as O
    G = ref.G
    case = RC.SWE_CASES["nonlinear_lf_closed"]
    st = G.Setup(case)
    depth, opts, o = st.depth_and_options()
    sol, uv, eta = st.swe_solution(3)
    m = st.m2
    area = m.cell_area()
    bn = RC.nodal_value(case["bath"], m)
    vol_ref = G.util.comp_volume_2d(sol.subfunctions[1], depth.bathymetry_2d)
    vol = (area * (eta + bn).mean(1)).sum()
    assert abs(vol_ref - vol) <= 1e-13 * abs(vol)
    rng = np.random.default_rng(5)
    c = 1.0 + 0.3 * rng.standard_normal(eta.shape)
    q = G.U.Function(st.H, name="tracer_2d")
    q.dat.data[...] = c.reshape(-1)
    mass_ref = G.util.comp_tracer_mass_2d(q, depth.get_total_depth(sol.subfunctions[1]))
    mref = (np.ones((3, 3)) + np.eye(3)) / 12.0
    mass = (area * np.einsum("ca,ab,cb->c", bn + eta, mref, c)).sum()
    assert abs(mass_ref - mass) <= 1e-13 * abs(mass)
    # print_state norms: firedrake.norm(f) = sqrt(assemble(inner(f, f) * dx))
    U = G.U
    n_h = np.sqrt(U.assemble(U.inner(sol.subfunctions[1], sol.subfunctions[1]) * U.dx))
    n_u = np.sqrt(U.assemble(U.inner(sol.subfunctions[0], sol.subfunctions[0]) * U.dx))
    assert abs(n_h - O.l2_norm(m, eta)) <= 1e-13 * n_h and abs(n_u - O.l2_norm(m, uv)) <= 1e-13 * n_u


def test_install_falls_back_to_the_reference_class_outside_the_accelerated_path(ref):
    """`steppers['SSPRK33']` also serves equations this library does not accelerate (solver2d.py:662-700: sediment,
    Exner; boundary data it cannot take).  After install(fallback=True) such a construction returns an instance of the
    reference's own class built from the same arguments (with a warning); a supported one returns the B200 class; by
    default the NotImplementedError propagates; errors the reference raises too are never swallowed."""
    import importlib
    import warnings
    import reference_cases as RC
    import thetis_b200
    from thetis_b200 import rungekutta as B
    G = ref.G
    thetis = sys.modules["thetis"]
    importlib.import_module("thetis.limiter")
    rk_mod, ti_mod, lim_mod = thetis.rungekutta, thetis.timeintegrator, thetis.limiter
    names = [(rk_mod, "SSPRK33"), (rk_mod, "ERKLSPUM2"), (rk_mod, "ERKLPUM2"), (rk_mod, "ERKMidpoint"), (rk_mod, "ERKEuler"),
             (ti_mod, "ForwardEuler"), (lim_mod, "VertexBasedP1DGLimiter")]
    saved = {(m, n): getattr(m, n) for m, n in names}
    ref_ssprk33 = saved[(rk_mod, "SSPRK33")]
    topt = types.SimpleNamespace(ad_block_tag="", solver_parameters={})
    try:
        thetis_b200.install(thetis, fallback=True)
        thetis_b200.install(thetis, fallback=True)                   # twice: the reference class is not lost
        assert rk_mod.SSPRK33._reference_class is ref_ssprk33
        # 1. supported set-up: the B200 class
        case = RC.SWE_CASES["nonlinear_lf_closed"]
        st = G.Setup(case)
        eq, fields, bnd, o = G.swe_equation(st)
        sol, uv0, eta0 = st.swe_solution(1)
        ti = rk_mod.SSPRK33(eq, sol, fields, 2.0, topt, bnd)
        assert isinstance(ti, B.SSPRK33) and ti.sync_policy == "every_step"
        # 2. a Function-valued boundary 'drag' is outside the accelerated path: the reference integrator takes over
        case2 = dict(case, bnd={1: {"drag": ("p1", "lin_drag")}})
        st2 = G.Setup(case2)
        eq2, fields2, bnd2, _ = G.swe_equation(st2)
        sol2, uv2, eta2 = st2.swe_solution(1)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            ti2 = rk_mod.SSPRK33(eq2, sol2, fields2, 2.0, topt, bnd2)
        assert type(ti2) is ref_ssprk33 and not isinstance(ti2, B.SSPRK33)
        assert any("falls back to the reference class" in str(x.message) for x in w)
        eng = ref.engines[-1]
        assert eng.swe_stepper is None                               # the half-built B200 stepper deregistered itself
        ti2.advance(0.0)                                             # ... and the reference object works
        assert np.abs(sol2.subfunctions[1].dat.data.reshape(eta2.shape) - eta2).max() > 1e-6
        # 3. an equation class the library does not know at all
        class SedimentEquation:                                      # noqa: N801 (stands for thetis.sediment_eq_2d's)
            pass
        made = {}

        class RefStandIn:
            def __init__(self, *a, **k):
                made["args"] = a
        rk_mod.SSPRK33._reference_class = RefStandIn                  # look at what the fallback is called with
        thetis_b200.install(thetis, fallback=True)                   # rebinding keeps the remembered class
        out = rk_mod.SSPRK33(SedimentEquation(), sol, fields, 2.0, topt, {})
        assert isinstance(out, RefStandIn) and made["args"][3] == 2.0
        # 4. the default (fallback off): the error propagates
        for (m, n), v in saved.items():
            setattr(m, n, v)
        thetis_b200.install(thetis)
        with pytest.raises(NotImplementedError, match="drag"):
            rk_mod.SSPRK33(eq2, st2.swe_solution(1)[0], fields2, 2.0, topt, bnd2)
        # 5. what the reference rejects too is never swallowed
        for (m, n), v in saved.items():
            setattr(m, n, v)
        thetis_b200.install(thetis, fallback=True)
        with pytest.raises(Exception, match="Invalid boundary tag"):
            rk_mod.SSPRK33(eq, st.swe_solution(1)[0], fields, 2.0, topt, {1: {"elevation": G.U.Constant(0.0)}})
    finally:
        for (m, n), v in saved.items():
            setattr(m, n, v)
