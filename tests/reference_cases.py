"""
Case definitions shared by tests/golden/make_reference_residual_golden.py (which EXECUTES the reference's own term
classes on them, see tests/golden/ufl_lite.py) and tests/test_oracle_reference_residuals.py (which feeds the same
inputs to the CPU oracle and compares with the stored reference output).

A case is a plain dict:
    mesh      (builder name, args)                           -> thetis_b200.mesh.Mesh2D
    options   use_nonlinear_equations, use_lax_friedrichs_velocity, use_wetting_and_drying, wetting_and_drying_alpha,
              norm_smoother, use_grad_div_viscosity_term, use_grad_depth_viscosity_term, sipg_factor
    bath      field spec
    fields    {name of solver2d.py:546-558: field spec}
    bnd       {marker: {tag: field spec}}
    g         optional gravity (test_rossby_wave.py:154-155 mutates physical_constants['g_grav'])
Field specs:  ("const", value) | ("p1", name of a function of (x, y) below) | ("dg", seed, base, amplitude[, ncomp])
              -- a Constant, a continuous P1 Function, a genuinely discontinuous P1DG Function.
"""
import numpy as np

from thetis_b200.mesh import rectangle_mesh, periodic_rectangle_mesh, delaunay_mesh

LX, LY = 5.0e3, 4.0e3

FUNCS = {
    "bath_wavy": lambda x, y: 20.0 + 3.0 * np.sin(x / 900.0) * np.cos(y / 700.0),
    "bath_slope": lambda x, y: 3.0 - 5.0 * x / LX,                       # dries out on the right
    "bath_deep": lambda x, y: 1000.0 + 50.0 * np.cos(x / 2.0e3),
    "coriolis": lambda x, y: 1.0e-4 + 2.0e-8 * y,
    "coriolis_y": lambda x, y: y / 1.0e3,
    "manning": lambda x, y: 0.02 + 0.01 * np.cos(x / 1.5e3) * np.sin(y / 1.1e3),
    "lin_drag": lambda x, y: 1.0e-3 * (1.0 + 0.5 * np.sin(x / 1.0e3)),
    "wind": lambda x, y: np.stack([0.1 * np.sin(np.pi * (y / LY - 0.5)), 0.02 * np.cos(x / 1.2e3)], -1),
    "pressure": lambda x, y: 1.0e5 + 300.0 * np.sin(x / 1.3e3) * np.cos(y / 0.9e3),
    "msrc": lambda x, y: np.stack([1.0e-4 * np.cos(y / 800.0), -2.0e-4 * np.sin(x / 700.0)], -1),
    "vsrc": lambda x, y: 1.0e-4 * np.sin(x / 600.0 + y / 900.0),
    "visc": lambda x, y: 50.0 + 20.0 * np.sin(x / 1.1e3) * np.sin(y / 1.4e3),
    "nikuradse": lambda x, y: 0.05 + 0.02 * np.cos(x / 1.0e3),
    "elev_bc": lambda x, y: 0.4 * np.sin(y / 1.0e3) + 0.1 * np.cos(x / 1.7e3),
    "uv_bc": lambda x, y: np.stack([0.2 + 0.1 * np.sin(y / 900.0), -0.1 * np.cos(x / 1.1e3)], -1),
    "un_bc": lambda x, y: 0.15 * np.cos(y / 1.2e3),
    "flux_bc": lambda x, y: 900.0 + 200.0 * np.sin(y / 1.0e3),
    "wd_alpha": lambda x, y: 0.3 + 0.2 * np.cos(x / 1.3e3) ** 2,
    "diff": lambda x, y: 5.0 + 2.0 * np.sin(x / 1.2e3) * np.cos(y / 1.0e3),
    "tsrc": lambda x, y: 1.0e-3 * np.cos(x / 800.0) * np.sin(y / 600.0),
    "value_bc": lambda x, y: 1.0 + 0.3 * np.sin(y / 700.0),
}


def build_mesh(spec):
    kind, args = spec[0], spec[1:]
    if kind == "rect":
        return rectangle_mesh(*args)
    if kind == "periodic":
        return periodic_rectangle_mesh(*args)
    if kind == "delaunay":
        return delaunay_mesh(*args[:3], seed=args[3])
    raise ValueError(kind)


def nodal_value(spec, mesh):
    """Field spec -> float / tuple (Constant) or nodal P1DG array (nt, 3[, k]) in the mesh's own cell order."""
    kind = spec[0]
    if kind == "const":
        return spec[1]
    if kind == "p1":
        x = mesh.coords[mesh.cells]
        return np.asarray(FUNCS[spec[1]](x[..., 0], x[..., 1]), dtype=float)
    if kind == "dg":
        seed, base, amp = spec[1], spec[2], spec[3]
        ncomp = spec[4] if len(spec) > 4 else None
        rng = np.random.default_rng(seed)
        shape = (mesh.n_cells, 3) + ((ncomp,) if ncomp else ())
        return base + amp * rng.standard_normal(shape)
    raise ValueError(kind)


def state(mesh, seed, amp_u=0.5, amp_e=0.3):
    """smooth + random P1DG state (uv (nt, 3, 2), eta (nt, 3)); the random part makes every facet jump non-zero"""
    rng = np.random.default_rng(seed)
    x = mesh.coords[mesh.cells]
    L = np.ptp(mesh.coords, axis=0).max()
    k = 2 * np.pi / L
    u = amp_u * np.sin(k * x[..., 0] + 0.3) * np.cos(k * x[..., 1]) + 0.05 * rng.standard_normal(x.shape[:2])
    v = -amp_u * np.cos(2 * k * x[..., 0]) * np.sin(k * x[..., 1] + 0.1) + 0.05 * rng.standard_normal(x.shape[:2])
    e = amp_e * np.cos(k * x[..., 0]) * np.sin(k * x[..., 1]) + 0.02 * rng.standard_normal(x.shape[:2])
    return np.stack([u, v], -1), e


RECT = ("rect", 5, 4, LX, LY)
RAGGED = ("rect", 3, 7, 3.0e3, 7.0e3)
DELAUNAY = ("delaunay", 40, LX, LY, 2)
PERIODIC = ("periodic", 6, 4, 6.0e3, 4.0e3)
C = lambda v: ("const", v)          # noqa: E731
P = lambda name: ("p1", name)       # noqa: E731

_OPEN_1 = {1: {"elev": C(0.3), "uv": C((0.2, -0.1))}, 2: {"elev": C(-0.2), "un": C(0.15)},
           3: {"elev": C(0.1), "flux": C(800.0)}, 4: {"elev": C(0.25)}}
_OPEN_2 = {1: {"uv": C((0.1, 0.05))}, 2: {"un": C(-0.2)}, 3: {"flux": C(-600.0)}}
_OPEN_F1 = {1: {"elev": P("elev_bc"), "uv": P("uv_bc")}, 2: {"elev": P("elev_bc"), "un": P("un_bc")},
            3: {"elev": P("elev_bc"), "flux": P("flux_bc")}, 4: {"elev": P("elev_bc")}}
_OPEN_F2 = {1: {"uv": P("uv_bc")}, 2: {"un": P("un_bc")}, 3: {"flux": P("flux_bc")}}

SWE_CASES = {
    "linear_constant_depth_closed": dict(mesh=RECT, options=dict(use_nonlinear_equations=False), bath=C(50.0)),
    "linear_variable_depth_ragged": dict(mesh=RAGGED, options=dict(use_nonlinear_equations=False), bath=P("bath_wavy")),
    "nonlinear_lf_closed": dict(mesh=RECT, bath=P("bath_wavy")),
    "nonlinear_no_lf": dict(mesh=DELAUNAY, options=dict(use_lax_friedrichs_velocity=False), bath=P("bath_wavy")),
    "nonlinear_lf_scaling": dict(mesh=RECT, bath=P("bath_wavy"),
                                 fields={"lax_friedrichs_velocity_scaling_factor": C(0.7)}),
    "periodic_coriolis_uv_walls_g1": dict(mesh=PERIODIC, bath=C(1.0), g=1.0, fields={"coriolis": P("coriolis_y")},
                                          bnd={1: {"uv": C((0.0, 0.0))}, 2: {"uv": C((0.0, 0.0))}}, amp=(0.05, 0.05)),
    "stommel_terms_unstructured": dict(mesh=DELAUNAY, options=dict(use_nonlinear_equations=False), bath=P("bath_deep"),
                                       fields={"coriolis": P("coriolis"), "wind_stress": P("wind"),
                                               "linear_drag_coefficient": C(1.0e-6)}),
    "manning_pressure_sources": dict(mesh=RECT, options=dict(norm_smoother=0.05), bath=P("bath_wavy"),
                                     fields={"manning_drag_coefficient": P("manning"),
                                             "atmospheric_pressure": P("pressure"), "momentum_source": P("msrc"),
                                             "volume_source": P("vsrc")}),
    "quadratic_drag_const_linear_drag_field": dict(mesh=DELAUNAY, bath=P("bath_wavy"),
                                                   fields={"quadratic_drag_coefficient": C(2.5e-3),
                                                           "linear_drag_coefficient": P("lin_drag")}),
    "wind_const_vector_nonlinear": dict(mesh=RECT, bath=P("bath_wavy"), fields={"wind_stress": C((0.1, -0.05)),
                                                                               "coriolis": C(1.2e-4)}),
    "open_bc_const_1_nonlinear": dict(mesh=RECT, bath=P("bath_wavy"), bnd=_OPEN_1),
    "open_bc_const_2_nonlinear": dict(mesh=RECT, bath=P("bath_wavy"), bnd=_OPEN_2),
    "open_bc_const_1_linear": dict(mesh=RECT, options=dict(use_nonlinear_equations=False), bath=P("bath_wavy"),
                                   bnd=_OPEN_1),
    "open_bc_const_2_linear": dict(mesh=RECT, options=dict(use_nonlinear_equations=False), bath=P("bath_wavy"),
                                   bnd=_OPEN_2),
    "open_bc_functions_1": dict(mesh=RECT, bath=P("bath_wavy"), bnd=_OPEN_F1,
                                fields={"manning_drag_coefficient": C(0.03), "coriolis": P("coriolis")}),
    "open_bc_functions_2_unstructured": dict(mesh=DELAUNAY, bath=P("bath_wavy"), bnd=_OPEN_F2),
    "wetting_drying_manning": dict(mesh=RECT, options=dict(use_wetting_and_drying=True, wetting_and_drying_alpha=0.4),
                                   bath=P("bath_slope"), fields={"manning_drag_coefficient": C(0.02)},
                                   bnd={1: {"elev": C(0.5)}}),
    "wetting_drying_open_flux": dict(mesh=RECT, options=dict(use_wetting_and_drying=True, wetting_and_drying_alpha=0.5),
                                     bath=P("bath_slope"), bnd={1: {"elev": C(0.4), "flux": C(300.0)}, 3: {"flux": C(-100.0)}}),
    "wetting_drying_alpha_p1": dict(mesh=RECT, options=dict(use_wetting_and_drying=True,
                                                           wetting_and_drying_alpha=P("wd_alpha")),
                                    bath=P("bath_slope"), fields={"manning_drag_coefficient": P("manning")}),
    "nikuradse": dict(mesh=RECT, bath=P("bath_wavy"), fields={"nikuradse_bed_roughness": P("nikuradse")}),
    "p1dg_coriolis_manning_sources": dict(mesh=RECT, bath=P("bath_wavy"),
                                          fields={"coriolis": ("dg", 5, 1.0e-4, 2.0e-5),
                                                  "manning_drag_coefficient": ("dg", 6, 0.03, 0.003),
                                                  "momentum_source": ("dg", 7, 0.0, 1.0e-4, 2),
                                                  "volume_source": ("dg", 8, 0.0, 1.0e-4),
                                                  "atmospheric_pressure": ("dg", 9, 1.0e5, 100.0)}),
}
for _gd in (False, True):
    for _gh in (False, True):
        SWE_CASES[f"viscosity_graddiv{int(_gd)}_graddepth{int(_gh)}"] = dict(
            mesh=DELAUNAY, options=dict(use_grad_div_viscosity_term=_gd, use_grad_depth_viscosity_term=_gh,
                                        sipg_factor=1.5),
            bath=P("bath_wavy"), fields={"viscosity_h": P("visc"), "coriolis": C(1.0e-4)},
            bnd={1: {"uv": C((0.1, 0.05))}, 2: {"un": C(-0.2)}, 3: {"elev": C(0.1), "flux": C(500.0)}, 4: {"elev": C(0.2)}})
SWE_CASES["viscosity_const_linear"] = dict(mesh=RECT, options=dict(use_nonlinear_equations=False), bath=P("bath_wavy"),
                                           fields={"viscosity_h": C(30.0)})
SWE_CASES["viscosity_wetting_drying_grad_depth"] = dict(
    mesh=RECT, options=dict(use_wetting_and_drying=True, wetting_and_drying_alpha=0.4, use_grad_div_viscosity_term=True),
    bath=P("bath_slope"), fields={"viscosity_h": P("visc"), "manning_drag_coefficient": C(0.025)},
    bnd={1: {"elev": P("elev_bc"), "uv": P("uv_bc")}})

# ModeSplit2DEquations (shallowwater_eq.py:931-966): pressure gradient, Coriolis, momentum source, atmospheric pressure
# and the continuity terms only -- `equation="modesplit"` selects that class on the reference side and
# include_momentum_advection=False on the oracle side (fields the class has no term for must be absent)
SWE_CASES["modesplit_open_bcs"] = dict(mesh=RECT, bath=P("bath_wavy"), equation="modesplit",
                                       fields={"coriolis": P("coriolis"), "momentum_source": P("msrc"),
                                               "atmospheric_pressure": P("pressure"), "volume_source": P("vsrc")},
                                       bnd={1: {"elev": C(0.3), "uv": C((0.2, -0.1))}, 2: {"un": C(-0.2)}})
SWE_CASES["modesplit_closed_unstructured"] = dict(mesh=DELAUNAY, bath=P("bath_wavy"), equation="modesplit",
                                                  fields={"coriolis": C(1.0e-4), "momentum_source": ("dg", 11, 0.0, 1.0e-4, 2)})

# BoundaryDragTerm (shallowwater_eq.py:704-726): 'drag' on a closed marker, combined with open tags, on every marker of
# an unstructured mesh, and with the linear equations (the term does not depend on use_nonlinear_equations)
SWE_CASES["boundary_drag_closed_and_open"] = dict(
    mesh=RECT, bath=P("bath_wavy"),
    bnd={1: {"drag": C(0.05)}, 2: {"elev": C(-0.2), "un": C(0.15), "drag": C(0.02)}, 3: {"uv": C((0.1, 0.05)), "drag": C(0.1)}})
SWE_CASES["boundary_drag_unstructured_manning"] = dict(
    mesh=DELAUNAY, bath=P("bath_wavy"), fields={"manning_drag_coefficient": C(0.03), "coriolis": P("coriolis")},
    bnd={m: {"drag": C(0.01 * m)} for m in (1, 2, 3, 4)})
SWE_CASES["boundary_drag_linear"] = dict(mesh=RAGGED, options=dict(use_nonlinear_equations=False), bath=P("bath_wavy"),
                                         bnd={1: {"drag": C(0.05)}, 4: {"elev": C(0.1), "drag": C(0.03)}})

# ModeSplit2DEquations accepts the tag (impose_dynamic_bnd, :286-296) but adds no BoundaryDragTerm (:953-957)
SWE_CASES["modesplit_ignores_boundary_drag"] = dict(mesh=RECT, bath=P("bath_wavy"), equation="modesplit",
                                                    fields={"coriolis": C(1.0e-4)},
                                                    bnd={1: {"drag": C(0.05)}, 2: {"elev": C(0.2), "drag": C(0.1)}})

TRACER_CASES = {
    "advection_closed_source": dict(mesh=RECT, bath=P("bath_wavy"), fields={"source": P("tsrc")}),
    "advection_bcs_lf": dict(mesh=RECT, bath=P("bath_wavy"), options=dict(use_lax_friedrichs_tracer=True),
                             fields={"lax_friedrichs_tracer_scaling_factor": C(0.8),
                                     "tracer_advective_velocity_factor": C(0.9)},
                             bnd={1: {"value": C(1.5), "uv": C((0.3, 0.1))}, 2: {"un": C(-0.2)},
                                  3: {"value": P("value_bc"), "flux": C(700.0), "elev": C(0.2)}, 4: {"value": C(0.5)}}),
    "advection_unstructured_function_bcs": dict(mesh=DELAUNAY, bath=P("bath_wavy"),
                                                bnd={1: {"value": P("value_bc"), "uv": P("uv_bc")}, 2: {"un": P("un_bc")}}),
    "diffusion_sipg": dict(mesh=DELAUNAY, bath=P("bath_wavy"), options=dict(sipg_factor_tracer=1.5),
                           fields={"diffusivity_h": P("diff")}, bnd={1: {"diff_flux": C(0.02)}, 2: {"value": C(1.2)}}),
    "diffusion_const_source": dict(mesh=RECT, bath=P("bath_wavy"), fields={"diffusivity_h": C(3.0), "source": C(1.0e-3)}),
    "conservative_form": dict(mesh=RECT, bath=P("bath_wavy"), options=dict(use_conservative_form=True),
                              fields={"source": P("tsrc")},
                              bnd={1: {"value": C(1.5), "uv": C((0.3, 0.1))}, 2: {"un": C(-0.2)}}),
    "conservative_wetting_drying": dict(mesh=RECT, bath=P("bath_slope"),
                                        options=dict(use_conservative_form=True, use_wetting_and_drying=True,
                                                     wetting_and_drying_alpha=0.4),
                                        fields={"source": C(2.0e-3), "diffusivity_h": C(2.0)}),
}

# whole SSPRK33 steps through the reference's own rungekutta.SSPRK33: (SWE case name, dt, number of steps, what
# update_forcings does)
STEP_CASES = {
    "ssprk33_tidal_constant": dict(case="open_bc_const_1_nonlinear", dt=4.0, n_steps=5, forcing="elev_const"),
    "ssprk33_tidal_function_manning": dict(case="open_bc_functions_1", dt=4.0, n_steps=5, forcing="elev_function"),
    "ssprk33_closed_linear": dict(case="linear_variable_depth_ragged", dt=6.0, n_steps=4, forcing=None),
    # Butcher-form integrators (rungekutta.py:762-867, tableaux :350-392) and timeintegrator.ForwardEuler (:115-165,
    # whose Function-valued coefficients lag one step: fields_old)
    "erklspum2_tidal_function": dict(case="open_bc_functions_1", dt=4.0, n_steps=4, forcing="elev_function",
                                     integrator="ERKLSPUM2"),
    "erklpum2_tidal_constant": dict(case="open_bc_const_1_nonlinear", dt=5.0, n_steps=4, forcing="elev_const",
                                    integrator="ERKLPUM2"),
    "erkmidpoint_viscous": dict(case="viscosity_graddiv1_graddepth1", dt=1.0, n_steps=4, forcing="elev_const",
                                integrator="ERKMidpoint"),
    "erkeuler_closed": dict(case="nonlinear_lf_closed", dt=2.0, n_steps=5, forcing=None, integrator="ERKEuler"),
    # boundary elevation given as a UFL EXPRESSION over a Constant-valued ramp and a Function, the way
    # examples/north_sea/model_config.py:181-192 writes it: elev_ramp * elev_tide_2d with
    # elev_ramp = conditional(bnd_time < ramp_t, bnd_time / ramp_t, 1.0); update_forcings assigns bnd_time and the tide
    "ssprk33_tidal_ufl_expression": dict(case="open_bc_functions_1", dt=4.0, n_steps=5, forcing="elev_expression",
                                         ramp_t=12.0),
    "forward_euler_lagged_drag": dict(case="quadratic_drag_const_linear_drag_field", dt=2.0, n_steps=5,
                                      forcing="lagged_drag", integrator="ForwardEuler"),
}


# SWE step, then tracer step with the NEW velocity / elevation (coupled_timeintegrator_2d.py:94-105, the tracer's
# `uv_2d` / `elev_2d` are the sub-functions of solution_2d, solver2d.py:580-598), update_forcings passed to both
COUPLED_CASES = {
    "coupled_ssprk33_advection": dict(swe="open_bc_const_1_nonlinear", tracer="advection_bcs_lf", dt=4.0, n_steps=4,
                                      forcing="elev_const"),
    "coupled_ssprk33_diffusion_unstructured": dict(swe="nonlinear_no_lf", tracer="diffusion_sipg", dt=2.0, n_steps=4,
                                                   forcing=None),
}


def forcing_factor(t):
    """time dependence applied by `update_forcings(t)` in the step cases"""
    return 1.0 + 0.5 * np.sin(2.0 * np.pi * t / 40.0)
