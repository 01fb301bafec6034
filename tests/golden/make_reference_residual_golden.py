"""
Generates tests/golden/reference_residuals.npz by EXECUTING THE REFERENCE'S OWN SOURCE on the cases of
tests/reference_cases.py:

* thetis/shallowwater_eq.py   ShallowWaterEquations / ModeSplit2DEquations .residual('all', ...)  -- every term class,
                              get_bnd_functions, impose_dynamic_bnd -- and Equation.mass_term     (equation.py:99-105)
* thetis/utility.py           DepthExpression, compute_boundary_length, tensor_jump, element_continuity
* thetis/tracer_eq_2d.py      TracerEquation2D (non-conservative and conservative terms)
* thetis/rungekutta.py        SSPRK33 = ERKGenericShuOsher + SSPRK33Abstract, the Butcher-form ERKGeneric schemes
                              (ERKLSPUM2, ERKLPUM2, ERKMidpoint, ERKEuler) and thetis/timeintegrator.py ForwardEuler:
                              whole steps through `advance()`
* thetis/coupled_timeintegrator_2d.py   GeneralCoupledTimeIntegrator2D.advance: SWE step, then the tracer step with the
                              new velocity (driven through a stand-in for the FlowSolver2d attributes it reads)

are imported from /root/reference (tests/golden/refenv.py) and run on `ufl_lite`, the numpy stand-in for the
Firedrake / UFL operators those files use (Firedrake itself is not installable here).  The stored numbers are
M^-1 R(u) per case -- the mass system solved with the reference's own mass form -- and the states after N steps.
Nothing of the reference is copied into the repository but these numbers.

    python tests/golden/make_reference_residual_golden.py [--out file.npz]     (build container only: needs /root/reference)

Wetting-drying: the reference's mass term gains a term that is nonlinear in the trial function
(shallowwater_eq.py:917-920), so `SSPRK33` + wetting-drying is not a reference code path; for those cases
the RESIDUAL is the reference's, the mass form is the plain P1DG mass of `Equation.mass_term` (the extension DESIGN.md
section 6 defines), and no step case uses wetting-drying.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import refenv                      # noqa: E402
import ufl_lite as U               # noqa: E402
import reference_cases as RC       # noqa: E402

MODS = refenv.install()
util, sweq, treq, rk, eqn = (MODS[k] for k in ("utility", "shallowwater_eq", "tracer_eq_2d", "rungekutta", "equation"))
PC = MODS["physical_constants"]


class Setup:
    """Reference-side objects of one case"""

    def __init__(self, case):
        self.case = case
        self.m2 = RC.build_mesh(case["mesh"])
        self.mesh = U.Mesh(self.m2)
        self.mesh.boundary_len = util.compute_boundary_length(self.mesh)          # utility.py:821-832
        self.P1 = util.get_functionspace(self.mesh, "CG", 1)
        self.P1v = util.get_functionspace(self.mesh, "CG", 1, vector=True)
        self.H = util.get_functionspace(self.mesh, "DG", 1, name="H_2d")
        self.Uv = util.get_functionspace(self.mesh, "DG", 1, vector=True, name="U_2d")
        self.V = U.MixedFunctionSpace([self.Uv, self.H])

    def obj(self, spec):
        """field spec -> Constant / Function as a user script would pass it"""
        kind = spec[0]
        if kind == "const":
            return U.Constant(spec[1])
        m = self.m2
        if kind == "p1":
            vals = np.asarray(RC.FUNCS[spec[1]](m.coords[:, 0], m.coords[:, 1]), dtype=float)
            sp = self.P1v if vals.ndim == 2 else self.P1
            f = U.Function(sp, name=spec[1])
            if m.periodic:
                f.dat.data[m.topo] = vals                     # the P1 space identifies the periodic vertices
            else:
                f.dat.data[...] = vals
            return f
        if kind == "dg":
            vals = RC.nodal_value(spec, m)
            f = U.Function(self.Uv if vals.ndim == 3 else self.H)
            f.dat.data[...] = vals.reshape(f.dat.data.shape)
            return f
        raise ValueError(kind)

    def depth_and_options(self):
        o = dict(use_nonlinear_equations=True, use_lax_friedrichs_velocity=True, use_wetting_and_drying=False,
                 wetting_and_drying_alpha=0.5, norm_smoother=0.0, use_grad_div_viscosity_term=False,
                 use_grad_depth_viscosity_term=True, sipg_factor=1.0, use_lax_friedrichs_tracer=False,
                 sipg_factor_tracer=1.0, use_conservative_form=False)
        o.update(self.case.get("options", {}))
        al = o["wetting_and_drying_alpha"]
        al = self.obj(al) if isinstance(al, tuple) else U.Constant(al)
        bath = self.obj(self.case["bath"])
        if isinstance(bath, U.Constant):                      # solver2d is always given a Function
            bath = U.Function(self.P1).assign(float(bath))
        depth = util.DepthExpression(bath, use_nonlinear_equations=o["use_nonlinear_equations"],
                                     use_wetting_and_drying=o["use_wetting_and_drying"],
                                     wetting_and_drying_alpha=al)
        opts = util.AttrDict(use_nonlinear_equations=o["use_nonlinear_equations"],
                             use_lax_friedrichs_velocity=o["use_lax_friedrichs_velocity"],
                             use_grad_div_viscosity_term=o["use_grad_div_viscosity_term"],
                             use_grad_depth_viscosity_term=o["use_grad_depth_viscosity_term"],
                             sipg_factor=U.Constant(o["sipg_factor"]), norm_smoother=U.Constant(o["norm_smoother"]),
                             use_lax_friedrichs_tracer=o["use_lax_friedrichs_tracer"],
                             sipg_factor_tracer=U.Constant(o["sipg_factor_tracer"]), use_supg_tracer=False,
                             tracer={"tracer_2d": types.SimpleNamespace(use_conservative_form=o["use_conservative_form"])})
        return depth, opts, o

    def bnd(self):
        return {mk: {tag: self.obj(v) for tag, v in funcs.items()} for mk, funcs in self.case.get("bnd", {}).items()}

    def swe_solution(self, seed):
        uv, eta = RC.state(self.m2, seed, *self.case.get("amp", (0.5, 0.3)))
        sol = U.Function(self.V, name="solution_2d")
        sol.subfunctions[0].dat.data[...] = uv.reshape(-1, 2)
        sol.subfunctions[1].dat.data[...] = eta.reshape(-1)
        return sol, uv, eta


def _solve_mass(mass_form, rhs_form, space):
    out = U.Function(space)
    U.LinearVariationalSolver(U.LinearVariationalProblem(mass_form, rhs_form, out)).solve()
    return out


def swe_equation(st):
    depth, opts, o = st.depth_and_options()
    if st.case.get("equation") == "modesplit":
        eq = sweq.ModeSplit2DEquations(st.V, depth, opts)
    else:
        eq = sweq.ShallowWaterEquations(st.V, depth, opts)
    fields = {"lax_friedrichs_velocity_scaling_factor": U.Constant(1.0)}          # solver2d.py:546-558 default
    for name, spec in st.case.get("fields", {}).items():
        fields[name] = st.obj(spec)
    return eq, fields, st.bnd(), o


def run_swe(name, case, seed):
    st = Setup(case)
    g_old = float(PC["g_grav"])
    PC["g_grav"].assign(case.get("g", g_old))                 # test_rossby_wave.py:154-155 does the same
    try:
        eq, fields, bnd, o = swe_equation(st)
        sol, uv, eta = st.swe_solution(seed)
        F = eq.residual("all", sol, sol, fields, fields, bnd)                     # rungekutta.py:901-904
        if o["use_wetting_and_drying"]:
            mass = eqn.Equation.mass_term(eq, eq.trial)       # plain P1DG mass (see the module docstring)
        else:
            mass = eq.mass_term(eq.trial)
        k = _solve_mass(mass, F, st.V)
        nt = st.m2.n_cells
        out = dict(uv=uv, eta=eta, ku=k.subfunctions[0].dat.data.reshape(nt, 3, 2).copy(),
                   ke=k.subfunctions[1].dat.data.reshape(nt, 3).copy())
        if o["use_wetting_and_drying"]:
            # the reference's OWN mass functional of the state with wetting-drying (shallowwater_eq.py:917-920:
            # Equation.mass_term + BathymetryDisplacementMassTerm), stored as M^-1 F(u) with the plain P1DG mass M
            m = _solve_mass(mass, eq.mass_term(sol), st.V)
            out["me"] = m.subfunctions[1].dat.data.reshape(nt, 3).copy()
            assert np.allclose(m.subfunctions[0].dat.data.reshape(nt, 3, 2), uv, rtol=0, atol=1e-13)
    finally:
        PC["g_grav"].assign(g_old)
    return out


def run_tracer(name, case, seed):
    st = Setup(case)
    depth, opts, o = st.depth_and_options()
    eq = treq.TracerEquation2D("tracer_2d", st.H, depth, opts, None)
    sol, uv, eta = st.swe_solution(seed)
    rng = np.random.default_rng(seed + 100)
    x = st.m2.coords[st.m2.cells]
    c = 1.0 + 0.5 * np.sin(x[..., 0] / 900.0) * np.cos(x[..., 1] / 700.0) + 0.05 * rng.standard_normal(x.shape[:2])
    q = U.Function(st.H, name="tracer_2d")
    q.dat.data[...] = c.reshape(-1)
    fields = {"uv_2d": sol.subfunctions[0], "elev_2d": sol.subfunctions[1],
              "tracer_advective_velocity_factor": U.Constant(1.0),
              "lax_friedrichs_tracer_scaling_factor": U.Constant(1.0)}
    for fname, spec in case.get("fields", {}).items():
        key = f"{fname}-tracer_2d" if fname in ("source", "diffusivity_h") else fname
        fields[key] = st.obj(spec)
    F = eq.residual("all", q, q, fields, fields, st.bnd())
    k = _solve_mass(eq.mass_term(eq.trial), F, st.H)
    nt = st.m2.n_cells
    return dict(uv=uv, eta=eta, c=c, kc=k.dat.data.reshape(nt, 3).copy())


def install_elev_expression(bnd, spec):
    """Replace every 'elev' Function of `bnd` by the UFL expression elev_ramp * elev_tide_2d of
    examples/north_sea/model_config.py:181-192; returns what update_forcings has to assign."""
    bnd_time, ramp_t = U.Constant(0.0), U.Constant(spec["ramp_t"])
    elev_ramp = U.conditional(bnd_time < ramp_t, bnd_time / ramp_t, 1.0)
    tides = []
    for mk, funcs in bnd.items():
        if "elev" in funcs:
            tide = funcs["elev"]
            tides.append((tide, tide.dat.data.copy()))
            funcs["elev"] = elev_ramp * tide
    return bnd_time, tides


def update_elev_expression(parts, t):
    bnd_time, tides = parts
    bnd_time.assign(t)
    for tide, base in tides:
        tide.dat.data[...] = base * RC.forcing_factor(t)
        tide.dat.dat_version += 1


def run_steps(name, spec, seed):
    case = RC.SWE_CASES[spec["case"]]
    st = Setup(case)
    eq, fields, bnd, o = swe_equation(st)
    assert not o["use_wetting_and_drying"]
    sol, uv, eta = st.swe_solution(seed)
    expr_parts = install_elev_expression(bnd, spec) if spec["forcing"] == "elev_expression" else None
    topt = types.SimpleNamespace(ad_block_tag=None, solver_parameters={})
    kind = spec.get("integrator", "SSPRK33")
    cls = MODS["timeintegrator"].ForwardEuler if kind == "ForwardEuler" else getattr(rk, kind)
    ti = cls(eq, sol, fields, spec["dt"], topt, bnd)
    ti.initialize(sol)
    base = {}
    drag = fields.get("linear_drag_coefficient") if spec["forcing"] == "lagged_drag" else None
    drag_base = drag.dat.data.copy() if drag is not None else None
    if spec["forcing"] == "elev_const":
        for mk, funcs in bnd.items():
            if "elev" in funcs:
                base[mk] = float(funcs["elev"])
    elif spec["forcing"] == "elev_function":
        for mk, funcs in bnd.items():
            if "elev" in funcs:
                base[mk] = funcs["elev"].dat.data.copy()

    def update_forcings(t):
        f = RC.forcing_factor(t)
        if expr_parts is not None:
            update_elev_expression(expr_parts, t)
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
    return dict(uv0=uv, eta0=eta, uv=sol.subfunctions[0].dat.data.reshape(nt, 3, 2).copy(),
                eta=sol.subfunctions[1].dat.data.reshape(nt, 3).copy())


class _FakeSolver:
    """What coupled_timeintegrator_2d.CoupledTimeIntegrator2D reads from a FlowSolver2d; the two factory methods
    restate solver2d.py:540-598 (which fields go to which integrator)."""

    def __init__(self, st, swe_case, tr_case, dt, seed):
        self.dt = dt
        self.solve_tracer, self.sediment_model = True, None
        depth, opts, o = st.depth_and_options()
        trc = Setup(tr_case)
        trc.m2, trc.mesh, trc.P1, trc.P1v, trc.H, trc.Uv, trc.V = st.m2, st.mesh, st.P1, st.P1v, st.H, st.Uv, st.V
        _, topts, to = trc.depth_and_options()
        for k in ("use_lax_friedrichs_tracer", "sipg_factor_tracer", "tracer"):
            opts[k] = topts[k]
        self.eq_sw = sweq.ShallowWaterEquations(st.V, depth, opts)
        self.eq_tr = treq.TracerEquation2D("tracer_2d", st.H, depth, opts, None)
        self.swe_fields = {"lax_friedrichs_velocity_scaling_factor": U.Constant(1.0)}
        for name, spec in swe_case.get("fields", {}).items():
            self.swe_fields[name] = st.obj(spec)
        self.bnd_functions = {"shallow_water": st.bnd(), "tracer_2d": trc.bnd()}
        sol, self.uv0, self.eta0 = st.swe_solution(seed)
        rng = np.random.default_rng(seed + 100)
        x = st.m2.coords[st.m2.cells]
        self.c0 = 1.0 + 0.5 * np.sin(x[..., 0] / 900.0) * np.cos(x[..., 1] / 700.0) + 0.05 * rng.standard_normal(x.shape[:2])
        q = U.Function(st.H, name="tracer_2d")
        q.dat.data[...] = self.c0.reshape(-1)
        self.fields = util.AttrDict(solution_2d=sol, tracer_2d=q)
        tf = {k: trc.obj(v) for k, v in tr_case.get("fields", {}).items()}
        sed = types.SimpleNamespace(solve_suspended_sediment=False, solve_exner=False)
        self.options = util.AttrDict(
            tracer_only=False, tracer_fields=["tracer_2d"], sediment_model_options=sed, tracer_picard_iterations=1,
            use_limiter_for_tracers=False,
            lax_friedrichs_tracer_scaling_factor=tf.get("lax_friedrichs_tracer_scaling_factor", U.Constant(1.0)),
            tracer_advective_velocity_factor=tf.get("tracer_advective_velocity_factor", U.Constant(1.0)),
            tracer={"tracer_2d": types.SimpleNamespace(diffusivity=tf.get("diffusivity_h"), source=tf.get("source"))},
            swe_timestepper_options=types.SimpleNamespace(ad_block_tag=None, solver_parameters={}),
            tracer_timestepper_options=types.SimpleNamespace(ad_block_tag=None, solver_parameters={}))

    def get_swe_timestepper(self, integrator):                 # solver2d.py:541-572
        return integrator(self.eq_sw, self.fields.solution_2d, self.swe_fields, self.dt,
                          self.options.swe_timestepper_options, self.bnd_functions["shallow_water"])

    def get_tracer_timestepper(self, integrator, system):      # solver2d.py:575-598
        uv, elev = self.fields.solution_2d.subfunctions
        fields = {"elev_2d": elev, "uv_2d": uv,
                  "lax_friedrichs_tracer_scaling_factor": self.options.lax_friedrichs_tracer_scaling_factor,
                  "tracer_advective_velocity_factor": self.options.tracer_advective_velocity_factor}
        for label in system.split(","):
            fields[f"diffusivity_h-{label}"] = self.options.tracer[label].diffusivity
            fields[f"source-{label}"] = self.options.tracer[label].source
        return integrator(self.eq_tr, self.fields[system], fields, self.dt, self.options.tracer_timestepper_options,
                          self.bnd_functions.get(system, {}))


def run_coupled(name, spec, seed):
    swe_case, tr_case = RC.SWE_CASES[spec["swe"]], RC.TRACER_CASES[spec["tracer"]]
    assert swe_case["mesh"] == tr_case["mesh"] and swe_case["bath"] == tr_case["bath"]
    st = Setup(swe_case)
    solver = _FakeSolver(st, swe_case, tr_case, spec["dt"], seed)
    cti = MODS["coupled_timeintegrator_2d"].GeneralCoupledTimeIntegrator2D(
        solver, {"shallow_water": rk.SSPRK33, "tracer": rk.SSPRK33})
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
    return dict(uv0=solver.uv0, eta0=solver.eta0, c0=solver.c0,
                uv=sol.subfunctions[0].dat.data.reshape(nt, 3, 2).copy(),
                eta=sol.subfunctions[1].dat.data.reshape(nt, 3).copy(),
                c=solver.fields.tracer_2d.dat.data.reshape(nt, 3).copy())


def main():
    out = {}
    for i, (name, case) in enumerate(RC.SWE_CASES.items()):
        r = run_swe(name, case, seed=i)
        for k, v in r.items():
            out[f"swe/{name}/{k}"] = v
        print(f"swe    {name:44s} |ku| {np.abs(r['ku']).max():.3e}  |ke| {np.abs(r['ke']).max():.3e}")
    for i, (name, case) in enumerate(RC.TRACER_CASES.items()):
        r = run_tracer(name, case, seed=50 + i)
        for k, v in r.items():
            out[f"tracer/{name}/{k}"] = v
        print(f"tracer {name:44s} |kc| {np.abs(r['kc']).max():.3e}")
    for i, (name, spec) in enumerate(RC.STEP_CASES.items()):
        r = run_steps(name, spec, seed=80 + i)
        for k, v in r.items():
            out[f"step/{name}/{k}"] = v
        print(f"step   {name:44s} |uv| {np.abs(r['uv']).max():.3e}  |eta| {np.abs(r['eta']).max():.3e}")
    for i, (name, spec) in enumerate(RC.COUPLED_CASES.items()):
        r = run_coupled(name, spec, seed=120 + i)
        for k, v in r.items():
            out[f"coupled/{name}/{k}"] = v
        print(f"coupled {name:43s} |uv| {np.abs(r['uv']).max():.3e}  |c| {np.abs(r['c']).max():.3e}")
    path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else os.path.join(HERE, "reference_residuals.npz")
    np.savez_compressed(path, **out)
    print(path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
