"""
Field-level pin of the CPU oracle against numbers produced by EXECUTING THE REFERENCE'S OWN SOURCE.

tests/golden/reference_residuals.npz holds, for every case of tests/reference_cases.py, M^-1 R(u) obtained by importing
thetis/shallowwater_eq.py, tracer_eq_2d.py, equation.py, utility.py and rungekutta.py from /root/reference and running
their `residual()` / `mass_term()` / `SSPRK33.advance()` on `tests/golden/ufl_lite.py`, a numpy stand-in for the
Firedrake / UFL operators those files use (generator: tests/golden/make_reference_residual_golden.py).  Here the CPU
oracle (oracle/swe_oracle.py; its C port for the cases its interface covers) is fed the same inputs and must reproduce those numbers: every
shallow-water term, all open-boundary combinations with Constant and Function data, wetting-drying depth, the three
drag laws, SIPG viscosity in its four variants, P1DG coefficient fields, the tracer terms in both forms, and whole
SSPRK33 steps with time-dependent forcing.

fp64, tolerance 1e-12 relative to the max-norm of the compared field (observed: <= 1e-14): the only differences are the
order of floating-point operations.  The CUDA path is tied to the same oracle at <= 1e-12 ... 1e-10 by the `-m gpu` tests.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import reference_cases as RC                                      # noqa: E402
from oracle.swe_oracle import SWEOracle, TracerOracle, ShuOsherStepper   # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_residuals.npz"))
TOL = 1e-12


def _options(case, mesh):
    o = dict(case.get("options", {}))
    if case.get("equation") == "modesplit":
        o["include_momentum_advection"] = False
    al = o.get("wetting_and_drying_alpha")
    if isinstance(al, tuple):
        o["wetting_and_drying_alpha"] = RC.nodal_value(al, mesh)
    return o


def _swe_oracle(case, mesh):
    fields = {k: RC.nodal_value(v, mesh) for k, v in case.get("fields", {}).items()}
    bnd = {mk: {tag: RC.nodal_value(v, mesh) for tag, v in funcs.items()} for mk, funcs in case.get("bnd", {}).items()}
    return SWEOracle(mesh, RC.nodal_value(case["bath"], mesh), options=_options(case, mesh), fields=fields,
                     bnd_conditions=bnd, g_grav=case.get("g", 9.81))


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_every_case_has_reference_output():
    have = {k.split("/")[1] for k in GOLD.files if k.startswith("swe/")}
    assert have == set(RC.SWE_CASES)
    assert {k.split("/")[1] for k in GOLD.files if k.startswith("tracer/")} == set(RC.TRACER_CASES)
    assert {k.split("/")[1] for k in GOLD.files if k.startswith("step/")} == set(RC.STEP_CASES)
    assert {k.split("/")[1] for k in GOLD.files if k.startswith("coupled/")} == set(RC.COUPLED_CASES)


@pytest.mark.parametrize("name", list(RC.SWE_CASES))
def test_swe_tendency_equals_the_reference_terms(name):
    case = RC.SWE_CASES[name]
    mesh = RC.build_mesh(case["mesh"])
    seed = list(RC.SWE_CASES).index(name)
    uv, eta = RC.state(mesh, seed, *case.get("amp", (0.5, 0.3)))
    # the inputs the reference side saw (guards against drift of the case definitions)
    assert np.array_equal(uv, GOLD[f"swe/{name}/uv"]) and np.array_equal(eta, GOLD[f"swe/{name}/eta"])
    ku, ke = _swe_oracle(case, mesh).tendency(uv, eta)
    eu, ee = _rel(ku, GOLD[f"swe/{name}/ku"]), _rel(ke, GOLD[f"swe/{name}/ke"])
    assert eu < TOL and ee < TOL, (eu, ee)


@pytest.mark.parametrize("name", [n for n, c in RC.SWE_CASES.items() if c.get("options", {}).get("use_wetting_and_drying")])
def test_wetting_drying_mass_functional_equals_the_reference_mass_term(name):
    """`ShallowWaterEquations.mass_term(solution)` with wetting-drying (shallowwater_eq.py:917-920: the plain mass
    plus BathymetryDisplacementMassTerm, :834-850) executed from the reference tree, against the oracle's
    `displaced_mass`; and the cell-local Newton solve inverts it (used by DisplacedMassShuOsherStepper)."""
    case = RC.SWE_CASES[name]
    mesh = RC.build_mesh(case["mesh"])
    eta = GOLD[f"swe/{name}/eta"]
    orc = _swe_oracle(case, mesh)
    F = orc.displaced_mass(eta)
    me = np.linalg.solve(orc.mass, F[..., None])[..., 0]
    assert _rel(me, GOLD[f"swe/{name}/me"]) < TOL
    assert _rel(me, eta) > 1e-3                        # the displacement term is really there
    back = orc.solve_displaced_mass(F, np.zeros_like(eta))
    assert _rel(back, eta) < 1e-11


@pytest.mark.parametrize("name", [n for n in RC.SWE_CASES if n.startswith("boundary_drag")])
def test_boundary_drag_cases_are_sensitive_to_the_term(name):
    """BoundaryDragTerm (shallowwater_eq.py:704-726): dropping the 'drag' tags must move the tendency far beyond the
    parity tolerance, otherwise the cases above would pin nothing"""
    case = RC.SWE_CASES[name]
    mesh = RC.build_mesh(case["mesh"])
    uv, eta = (GOLD[f"swe/{name}/{k}"] for k in ("uv", "eta"))
    nodrag = dict(case, bnd={mk: {t: v for t, v in funcs.items() if t != "drag"} for mk, funcs in case["bnd"].items()})
    ku, _ = _swe_oracle(nodrag, mesh).tendency(uv, eta)
    assert _rel(ku, GOLD[f"swe/{name}/ku"]) > 1e-4


@pytest.mark.parametrize("name", list(RC.TRACER_CASES))
def test_tracer_tendency_equals_the_reference_terms(name):
    case = RC.TRACER_CASES[name]
    mesh = RC.build_mesh(case["mesh"])
    uv, eta, c = (GOLD[f"tracer/{name}/{k}"] for k in ("uv", "eta", "c"))
    seed = 50 + list(RC.TRACER_CASES).index(name)
    uv2, eta2 = RC.state(mesh, seed)
    assert np.array_equal(uv, uv2) and np.array_equal(eta, eta2)
    o = _options(case, mesh)
    swe = SWEOracle(mesh, RC.nodal_value(case["bath"], mesh),
                    options={k: v for k, v in o.items() if k in ("use_wetting_and_drying", "wetting_and_drying_alpha")})
    fields = {k: RC.nodal_value(v, mesh) for k, v in case.get("fields", {}).items()}
    bnd = {mk: {tag: RC.nodal_value(v, mesh) for tag, v in funcs.items()} for mk, funcs in case.get("bnd", {}).items()}
    tr = TracerOracle(swe, bnd_conditions=bnd, fields=fields,
                      options={k: v for k, v in o.items()
                               if k in ("use_lax_friedrichs_tracer", "use_conservative_form", "sipg_factor_tracer")})
    tr.set_velocity(uv, eta)
    (kc,) = tr.tendency(c)
    e = _rel(kc, GOLD[f"tracer/{name}/kc"])
    assert e < TOL, e


@pytest.mark.parametrize("name", list(RC.STEP_CASES))
def test_whole_steps_equal_the_reference_integrators(name):
    """whole steps: the reference's rungekutta.SSPRK33 / ERKLSPUM2 / ERKLPUM2 / ERKMidpoint / ERKEuler and
    timeintegrator.ForwardEuler `advance()` (update_forcings at the stage times, stage combinations, one mass solve per
    stage) against the oracle's steppers on the same forcing"""
    spec = RC.STEP_CASES[name]
    case = RC.SWE_CASES[spec["case"]]
    mesh = RC.build_mesh(case["mesh"])
    seed = 80 + list(RC.STEP_CASES).index(name)
    uv, eta = RC.state(mesh, seed, *case.get("amp", (0.5, 0.3)))
    assert np.array_equal(uv, GOLD[f"step/{name}/uv0"]) and np.array_equal(eta, GOLD[f"step/{name}/eta0"])
    orc = _swe_oracle(case, mesh)
    base = {mk: funcs["elev"] for mk, funcs in orc.bnd.items() if "elev" in funcs} if spec["forcing"] else {}

    def update_forcings(t):
        f = RC.forcing_factor(t)
        if spec["forcing"] == "elev_expression":      # elev_ramp * elev_tide_2d, evaluated by hand
            f = f * (t / spec["ramp_t"] if t < spec["ramp_t"] else 1.0)
        for mk, b in base.items():
            orc.bnd[mk] = dict(orc.bnd[mk], elev=(b * f if isinstance(b, np.ndarray) else float(b) * f))

    kind = spec.get("integrator", "SSPRK33")
    dt, t = spec["dt"], 0.0
    if kind == "ForwardEuler":
        # timeintegrator.py:115-165: update_forcings(t + dt), then one Euler step assembled with `fields_old`: the
        # Function-valued drag coefficient is the one of the END OF THE PREVIOUS step, boundary data are live
        drag_base = orc.fields["linear_drag_coefficient"].copy()
        stp = ShuOsherStepper(orc, [uv, eta], dt, a=[[0]], b=[1.0], c=[0])
        for _ in range(spec["n_steps"]):
            drag_new = drag_base * RC.forcing_factor(t + dt)      # what update_forcings assigns to the Function
            stp.advance(t)                                       # ... but the step still reads fields_old
            orc.fields["linear_drag_coefficient"] = drag_new      # update_fields_old
            t += dt
    else:
        if kind == "SSPRK33":
            stp = ShuOsherStepper(orc, [uv, eta], dt)
        else:
            from oracle.swe_oracle import ButcherStepper, ERK_TABLEAUX
            stp = ButcherStepper(orc, [uv, eta], dt, *ERK_TABLEAUX[kind])
        for _ in range(spec["n_steps"]):
            stp.advance(t, update_forcings if spec["forcing"] else None)
            t += dt
    eu, ee = _rel(uv, GOLD[f"step/{name}/uv"]), _rel(eta, GOLD[f"step/{name}/eta"])
    assert eu < TOL and ee < TOL, (eu, ee)


@pytest.mark.parametrize("name", list(RC.COUPLED_CASES))
def test_coupled_swe_then_tracer_steps_equal_the_reference(name):
    """coupled_timeintegrator_2d.CoupledTimeIntegrator2D.advance executed from the reference tree: the SWE step, then
    the tracer step reading the NEW velocity / elevation, `update_forcings` handed to both integrators"""
    spec = RC.COUPLED_CASES[name]
    swe_case, tr_case = RC.SWE_CASES[spec["swe"]], RC.TRACER_CASES[spec["tracer"]]
    mesh = RC.build_mesh(swe_case["mesh"])
    uv, eta, c = (GOLD[f"coupled/{name}/{k}"].copy() for k in ("uv0", "eta0", "c0"))
    orc = _swe_oracle(swe_case, mesh)
    o = _options(tr_case, mesh)
    fields = {k: RC.nodal_value(v, mesh) for k, v in tr_case.get("fields", {}).items()}
    bnd = {mk: {tag: RC.nodal_value(v, mesh) for tag, v in funcs.items()} for mk, funcs in tr_case.get("bnd", {}).items()}
    tr = TracerOracle(orc, bnd_conditions=bnd, fields=fields,
                      options={k: v for k, v in o.items()
                               if k in ("use_lax_friedrichs_tracer", "use_conservative_form", "sipg_factor_tracer")})
    tr.set_velocity(uv, eta)                  # the same arrays the SWE stepper updates in place (solver2d.py:580-583)
    base = {mk: funcs["elev"] for mk, funcs in orc.bnd.items() if "elev" in funcs} if spec["forcing"] else {}

    def update_forcings(t):
        for mk, b in base.items():
            orc.bnd[mk] = dict(orc.bnd[mk], elev=float(b) * RC.forcing_factor(t))

    uf = update_forcings if spec["forcing"] else None
    s_swe = ShuOsherStepper(orc, [uv, eta], spec["dt"])
    s_tr = ShuOsherStepper(tr, [c], spec["dt"])
    t = 0.0
    for _ in range(spec["n_steps"]):
        s_swe.advance(t, uf)
        s_tr.advance(t, uf)
        t += spec["dt"]
    for got, key in ((uv, "uv"), (eta, "eta"), (c, "c")):
        e = _rel(got, GOLD[f"coupled/{name}/{key}"])
        assert e < TOL, (key, e)


# cases the C port's interface covers (per-vertex bathymetry / Coriolis / Manning / wind, constant linear drag, constant
# boundary data, wetting-drying): oracle/swe_oracle.c is what `bench.py --impl reference` and `cpu_baseline` time
_C_CASES = ["linear_constant_depth_closed", "linear_variable_depth_ragged", "nonlinear_lf_closed", "nonlinear_no_lf",
            "nonlinear_lf_scaling", "stommel_terms_unstructured", "wind_const_vector_nonlinear",
            "open_bc_const_1_nonlinear", "open_bc_const_2_nonlinear", "open_bc_const_1_linear", "open_bc_const_2_linear",
            "wetting_drying_manning", "wetting_drying_open_flux"]


def _vertex(spec, mesh):
    if spec is None:
        return None
    if spec[0] == "const":
        return spec[1]
    assert spec[0] == "p1"
    return np.asarray(RC.FUNCS[spec[1]](mesh.coords[:, 0], mesh.coords[:, 1]), dtype=float)


@pytest.mark.parametrize("name", _C_CASES)
def test_c_oracle_tendency_equals_the_reference_terms(name):
    from oracle.c_oracle import COracle, records_from_nodal, nodal_from_records
    case = RC.SWE_CASES[name]
    mesh = RC.build_mesh(case["mesh"])
    o, f = case.get("options", {}), dict(case.get("fields", {}))
    lf_sigma = f.pop("lax_friedrichs_velocity_scaling_factor", ("const", 1.0))[1]
    lin = f.pop("linear_drag_coefficient", ("const", 0.0))[1]
    cor, man, wind = (_vertex(f.pop(k, None), mesh) for k in ("coriolis", "manning_drag_coefficient", "wind_stress"))
    assert not f, f"not in the C port's interface: {sorted(f)}"
    bnd = {mk: {tag: v[1] for tag, v in funcs.items()} for mk, funcs in case.get("bnd", {}).items()}
    co = COracle(mesh, _vertex(case["bath"], mesh), nonlinear=o.get("use_nonlinear_equations", True),
                 lf_on=o.get("use_lax_friedrichs_velocity", True), g=case.get("g", 9.81), coriolis=cor, manning=man,
                 linear_drag=lin, bnd=bnd, norm_smoother=o.get("norm_smoother", 0.0), lf_sigma=lf_sigma,
                 wd_on=o.get("use_wetting_and_drying", False), wd_alpha=o.get("wetting_and_drying_alpha", 0.5),
                 wind_stress=wind)
    uv, eta = GOLD[f"swe/{name}/uv"], GOLD[f"swe/{name}/eta"]
    ku, ke = nodal_from_records(co.tendency(records_from_nodal(uv, eta)))
    eu, ee = _rel(ku, GOLD[f"swe/{name}/ku"]), _rel(ke, GOLD[f"swe/{name}/ke"])
    assert eu < TOL and ee < TOL, (eu, ee)


@pytest.mark.skipif(not os.path.isdir("/root/reference/thetis"), reason="the reference tree only exists in the build container")
def test_committed_fixture_is_what_the_reference_source_produces(tmp_path):
    """Re-runs the generator (in a subprocess: it registers stand-in `firedrake` / `thetis` modules) and compares its
    output with the committed fixture, array by array."""
    import subprocess
    out = tmp_path / "regen.npz"
    gen = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_residual_golden.py")
    r = subprocess.run([sys.executable, gen, "--out", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    new = np.load(out)
    assert set(new.files) == set(GOLD.files)
    for k in GOLD.files:
        assert np.array_equal(new[k], GOLD[k]), k
