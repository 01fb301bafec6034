"""
B200 drop-in for `thetis.rungekutta.SSPRK33` (= ERKGenericShuOsher + SSPRK33Abstract,
/root/reference thetis/rungekutta.py:326-347, 870-956) behind the reference's own
time-integrator interface (`thetis.timeintegrator.TimeIntegrator`,
thetis/timeintegrator.py:13-73):

    cls(equation, solution, fields, dt, options, bnd_conditions, terms_to_add='all')
    .initialize(solution)  .advance(t, update_forcings=None)  .set_dt(dt)
    .solve_stage(i, t, update_forcings=None)  .n_stages  .cfl_coeff  .name

The per-stage residual assembly, the P1DG mass solve and the Shu-Osher update
(rungekutta.py:929-946) are ONE CUDA kernel launch (tb_swe_stage /
tb_tracer_stage); Firedrake/PETSc are not touched on the hot path.  State lives
on the device; `solution` (host) is refreshed according to ``sync_policy``.

There is no CPU fallback: construction raises if the configuration is outside
the accelerated path (NotImplementedError) or if the CUDA library is missing.
"""
from __future__ import annotations

import numpy as np
import torch

import weakref

from . import _lib as L
from .adaptor import get_adaptor, is_constant, is_function, is_expression, expression_leaves, constant_value
from .engine import Engine

__all__ = ["SSPRK33", "SSPRK22", "ExportStage", "ERKGenericShuOsher", "ERKGeneric", "ERKLSPUM2", "ERKLPUM2", "ERKMidpoint", "ERKEuler",
           "ForwardEuler", "butcher_to_shuosher_form", "CFL_UNCONDITIONALLY_STABLE"]

CFL_UNCONDITIONALLY_STABLE = np.inf
# which mass functional the explicit wetting-drying step advances when neither the `wd_mass` argument nor
# `options.explicit_wetting_and_drying_mass` says so: 'plain' | 'displaced' (DESIGN.md section 6).  FlowSolver2d builds
# the integrators itself (solver2d.py:572-573), so a script that wants the displaced-mass step behind an unmodified
# Thetis sets  thetis_b200.rungekutta.WD_MASS_DEFAULT = 'displaced'  before create_timestepper().
WD_MASS_DEFAULT = "plain"


def butcher_to_shuosher_form(a, b):
    """
    Shu-Osher form of an explicit Butcher tableau; same construction as
    thetis/rungekutta.py:13-87 (sub-diagonal entries of [a; b] become beta).
    """
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if np.diag(a).any():
        raise NotImplementedError("implicit Runge-Kutta schemes are outside the accelerated path")
    lower = np.vstack((a, b))[1:, :]
    n = lower.shape[0]
    beta_sub = np.diag(np.diag(lower))
    alpha = np.zeros((n + 1, n + 1))
    alpha[1:, 1:] = np.eye(n) - beta_sub @ np.linalg.inv(lower)
    alpha[:, 0] = 1.0 - alpha.sum(axis=1)
    beta = np.zeros((n + 1, n + 1))
    beta[1:, :-1] = beta_sub
    alpha[np.abs(alpha) < 1e-13] = 0.0
    beta[np.abs(beta) < 1e-13] = 0.0
    assert np.allclose(alpha.sum(axis=1), 1.0)
    return alpha, beta


_SWE_FIELDS = {
    "coriolis": L.F_CORIOLIS, "manning_drag_coefficient": L.F_MANNING,
    "quadratic_drag_coefficient": L.F_QUAD_DRAG, "linear_drag_coefficient": L.F_LINEAR_DRAG,
    "wind_stress": L.F_WIND_STRESS, "atmospheric_pressure": L.F_ATM_PRESSURE,
    "momentum_source": L.F_MOMENTUM_SOURCE, "volume_source": L.F_VOLUME_SOURCE,
    "viscosity_h": L.F_VISCOSITY, "nikuradse_bed_roughness": L.F_NIKURADSE,
}
# coefficients of facet terms: must be continuous (P1); the others may be genuinely discontinuous P1DG fields
_CONTINUOUS_ONLY = {"bathymetry", "viscosity_h", "diffusivity_h", "wetting_and_drying_alpha"}
_MODESPLIT_FIELDS = ("coriolis", "momentum_source", "atmospheric_pressure", "volume_source")
_SWE_TAGS = {"elev": L.BC_ELEV, "uv": L.BC_UV, "un": L.BC_UN, "flux": L.BC_FLUX, "drag": L.BC_DRAG}
_TRACER_TAGS = {"elev": L.BC_ELEV, "uv": L.BC_UV, "un": L.BC_UN, "flux": L.BC_FLUX, "value": L.BC_VALUE,
                "diff_flux": L.BC_DIFF_FLUX}
_CONST_SLOT = {"elev": 0, "uv": 1, "un": 3, "flux": 4, "value": 5, "diff_flux": 6, "drag": 7}


def _opt(options, name, default=None):
    if options is None:
        return default
    if isinstance(options, dict):
        return options.get(name, default)
    return getattr(options, name, default)


def _version(obj):
    """Cheap change stamp of a Function/Constant (PyOP2 dat_version when available); for an expression the tuple of
    its operands' stamps."""
    d = getattr(obj, "dat", None)
    if d is not None:        # Functions and Constants: the common, cheap case first
        v = getattr(d, "dat_version", None)
        if v is not None:
            return ("dat", v)
        return None          # unknown: always refresh
    if is_expression(obj):
        vs = tuple((id(o), _version(o)) for o in expression_leaves(obj))
        return None if any(v[1] is None for v in vs) else ("expr", vs)
    v = getattr(obj, "version", None)
    return None if v is None else ("const", v)


class ERKGenericShuOsher:
    """
    Explicit Runge-Kutta integrator in Shu-Osher form on the device.  Only
    schemes whose stages combine the step's initial state u^(0) and the latest
    stage are supported (SSPRK33, forward Euler, SSPRK22): the fused kernel
    computes  u_out = a0*u^(0) + a1*u^(i) + beta*dt*M^-1 R(u^(i)).
    """
    a = None
    b = None
    c = None
    cfl_coeff = 1.0
    n_buffers = 3
    butcher_form = False
    fused_norms = None      # device tensor (4 doubles): when set, the last SWE stage also reduces int eta^2, int |u|^2,
                            # int eta, int (eta + bathymetry) of the new solution into it (tb_stage_integrals)

    def __init__(self, equation, solution, fields, dt, options=None, bnd_conditions=None, terms_to_add="all",
                 sync_policy="every_step", wd_mass=None):
        # wd_mass: 'plain' (default) | 'displaced' -- which mass functional the explicit wetting-drying step advances
        # (DESIGN.md section 6); also read from options.explicit_wetting_and_drying_mass when not given
        self.wd_mass = wd_mass if wd_mass is not None else _opt(options, "explicit_wetting_and_drying_mass", None)
        if self.wd_mass is None:
            self.wd_mass = WD_MASS_DEFAULT
        if self.wd_mass not in ("plain", "displaced"):
            raise ValueError(f"wd_mass must be 'plain' or 'displaced', not {self.wd_mass!r}")
        self.equation = equation
        self.solution = solution
        self.fields = fields if fields is not None else {}
        self.dt = float(dt)
        self.options = options
        self.bnd_conditions = bnd_conditions if bnd_conditions is not None else {}
        self.name = "-".join([self.__class__.__name__, self.equation.__class__.__name__])
        self.solver_parameters = _opt(options, "solver_parameters", {})   # accepted, unused: the mass solve is exact
        if terms_to_add not in ("all", ["implicit", "explicit", "source"]) and set(terms_to_add) != {"implicit", "explicit", "source"}:
            raise NotImplementedError("term subsets are outside the accelerated path")
        self.a = np.array(self.a, dtype=float)
        self.b = np.array(self.b, dtype=float)
        self.c = np.array(self.c, dtype=float)
        self.n_stages = len(self.b)
        if self.butcher_form:
            self.n_buffers = self.n_stages + 2           # solution, stage solution, one tendency per stage
        else:
            self.alpha, self.beta = butcher_to_shuosher_form(self.a, self.b)
            for i in range(self.n_stages):
                if np.any(self.alpha[i + 1][1:i] != 0.0):
                    raise NotImplementedError("Shu-Osher form uses intermediate stages other than u^(0) and the latest")
        self.sync_policy = sync_policy          # 'every_step' | 'manual'
        self._kind = self._equation_kind()
        mesh_obj = self._function_space().mesh()
        self.adaptor = get_adaptor(mesh_obj)
        self.engine = self.adaptor.get_engine()
        self.halo = self.adaptor.halo          # None on a single GPU
        self._host_stale = False
        self._last_host_version = None
        self._field_versions = {}
        self._bc_versions = {}
        self._stamps = {}
        try:
            self._setup_buffers()
            self._check_supported()
            self._push_static()
            self.initialize(solution)
        except Exception:
            self._unregister()      # a caller may fall back to the reference class: leave no half-built stepper behind
            raise

    # ------------------------------------------------------------------ set-up
    def _function_space(self):
        fs = getattr(self.equation, "function_space", None)
        return fs if fs is not None else self.solution.function_space()

    def _equation_kind(self):
        n = self.equation.__class__.__name__
        self._modesplit = n == "ModeSplit2DEquations"
        if n in ("ShallowWaterEquations", "ModeSplit2DEquations"):
            return "swe"
        if n == "TracerEquation2D":
            return "tracer"
        raise NotImplementedError(f"{n} is outside the accelerated path (ShallowWaterEquations, ModeSplit2DEquations, "
                                  "TracerEquation2D)")

    def _setup_buffers(self):
        eng, ad = self.engine, self.adaptor
        dev = eng.device
        if self._kind == "swe":
            uv_f, eta_f = self.solution.subfunctions
            nm_u = ad.dg_node_map(uv_f.function_space())
            nm_e = ad.dg_node_map(eta_f.function_space())
            if not np.array_equal(nm_u, nm_e):
                raise NotImplementedError("velocity and elevation P1DG spaces must share their node numbering")
            self.node_map = torch.as_tensor(nm_u.reshape(-1)).to(dev)
            n_nodes = int(ad.dat_ro(eta_f).shape[0])
            nb = self.n_buffers
            self.buf = self.halo.alloc(9, nbuf=nb) if self.halo is not None else [eng.new_state() for _ in range(nb)]
            self._d_uv = torch.empty((n_nodes, 2), dtype=torch.float64, device=dev)
            self._d_eta = torch.empty(n_nodes, dtype=torch.float64, device=dev)
            self._h_uv = torch.empty((n_nodes, 2), dtype=torch.float64).pin_memory()
            self._h_eta = torch.empty(n_nodes, dtype=torch.float64).pin_memory()
            eng.swe_stepper = self
        else:
            nm = ad.dg_node_map(self.solution.function_space())
            self.node_map = torch.as_tensor(nm.reshape(-1)).to(dev)
            n_nodes = int(ad.dat_ro(self.solution).shape[0])
            nb = self.n_buffers
            self.buf = self.halo.alloc(3, nbuf=nb) if self.halo is not None else [eng.new_tracer() for _ in range(nb)]
            self._d_q = torch.empty(n_nodes, dtype=torch.float64, device=dev)
            self._h_q = torch.empty(n_nodes, dtype=torch.float64).pin_memory()
            self._own_swe_state = None
            if not hasattr(eng, "tracer_steppers"):
                eng.tracer_steppers = {}
            # keyed by id() for the lookup, validated through a weak reference (an id can be reused after GC)
            eng.tracer_steppers[id(self.solution)] = (weakref.ref(self.solution), self)

    def _unregister(self):
        """Remove this integrator from the registries of the shared device engine (failed construction)."""
        eng = self.engine
        if getattr(eng, "swe_stepper", None) is self:
            eng.swe_stepper = None
        reg = getattr(eng, "tracer_steppers", None)
        if reg:
            for k in [k for k, (_, st) in reg.items() if st is self]:
                del reg[k]
        if getattr(eng, "_tracer_cfg_owner", None) is self:
            eng._tracer_cfg_owner = None

    def _check_supported(self):
        if self._kind == "swe":
            if self.fields.get("nikuradse_bed_roughness") is not None and (
                    self.fields.get("manning_drag_coefficient") is not None
                    or self.fields.get("quadratic_drag_coefficient") is not None):
                raise Exception("Cannot set both Nikuradse drag and Manning / dimensionless drag parameter")
            for m, funcs in self.bnd_conditions.items():
                for k, v in (funcs or {}).items():
                    if k not in _SWE_TAGS:
                        raise Exception(f'Invalid boundary tag "{k}" specified on boundary {m}')
                    if k == "drag" and not is_constant(v):
                        # BoundaryDragTerm (shallowwater_eq.py:704-726): the coefficient is a kernel parameter
                        raise NotImplementedError("spatially varying boundary 'drag' is outside the accelerated path")
            if getattr(self.equation, "tidal_farms", None):
                raise NotImplementedError("tidal turbines are outside the accelerated path")
        else:
            opts = self.equation.options
            if _opt(opts, "use_supg_tracer", False):
                raise NotImplementedError("SUPG is outside the accelerated path")
            diff = self._tracer_field("diffusivity_h")
            for m, funcs in self.bnd_conditions.items():
                for k, v in (funcs or {}).items():
                    if k == "diff_flux" and not is_constant(v):
                        raise NotImplementedError("spatially varying 'diff_flux' is outside the accelerated path")
                    if k == "value" and diff is not None and not is_constant(v):
                        raise NotImplementedError(
                            "diffusive boundary flux with a spatially varying 'value' is outside the accelerated path")

    # ------------------------------------------------------------------ configuration upload
    def _depth(self):
        return self.equation.depth

    def _tracer_label(self):
        """Label of the tracer this equation advances ('tracer_2d', 'salinity_2d', ...).  The reference's
        `TracerEquation2D` does not keep its `system` argument; every term it adds carries the label
        (`TracerTerm.label`, tracer_eq_2d.py:62, keys 'HorizontalAdvectionTerm_<label>' in `equation.terms`), whereas
        `equation.labels` is the term-key -> 'explicit'/'source' dict of `Equation` (equation.py:69), not a list."""
        lab = getattr(self.equation, "system", None)
        if lab is None:
            terms = getattr(self.equation, "terms", None)
            found = []
            for term in (terms.values() if hasattr(terms, "values") else ()):
                tl = getattr(term, "label", None)
                if tl is not None and tl not in found:
                    found.append(tl)
            if len(found) > 1:
                raise NotImplementedError("mixed systems of several tracers are outside the accelerated path")
            if found:
                return found[0]
            labels = getattr(self.equation, "labels", None)
            if isinstance(labels, (list, tuple)) and labels:
                lab = labels[0]
        if isinstance(lab, str) and "," in lab:
            raise NotImplementedError("mixed systems of several tracers are outside the accelerated path")
        return lab

    def _tracer_field(self, prefix):
        """fields['<prefix>-<label>'] (solver2d.py:590-592); falls back to any key with that prefix."""
        lab = self._tracer_label()
        if lab is not None and f"{prefix}-{lab}" in self.fields:
            return self.fields[f"{prefix}-{lab}"]
        for k, v in self.fields.items():
            if k.startswith(prefix) and v is not None:
                return v
        return None

    def _push_static(self):
        """Options, bathymetry and everything else read once at construction."""
        eng = self.engine
        depth = self._depth()
        eqo = self.equation.options
        eng.set_option(L.OPT_NONLINEAR, bool(depth.use_nonlinear_equations))
        eng.set_option(L.OPT_WETTING_DRYING, bool(depth.use_wetting_and_drying))
        if depth.use_wetting_and_drying:
            al = depth.wetting_and_drying_alpha
            if is_constant(al):
                eng.set_option(L.OPT_WD_ALPHA, float(constant_value(al)[0]))
                if self._kind == "swe":
                    eng.set_field(L.F_WD_ALPHA, None)
            elif self._kind == "swe":
                # P1 Function (use_automatic_wetting_and_drying_alpha, solver2d.py:279-287): one vertex column
                self._set_field(L.F_WD_ALPHA, al, "wetting_and_drying_alpha")
            else:
                raise NotImplementedError("tracer equation with a spatially varying wetting_and_drying_alpha is "
                                          "outside the accelerated path")
        self._set_field(L.F_BATHYMETRY, depth.bathymetry_2d, "bathymetry")
        for m, ln in self.adaptor.boundary_len.items():
            eng.set_boundary_length(m, ln)
        if self._kind == "swe":
            displaced = self.wd_mass == "displaced" and bool(depth.use_wetting_and_drying)
            if displaced and (self.butcher_form or not depth.use_nonlinear_equations):
                raise NotImplementedError("wd_mass='displaced' advances the reference's wetting-drying mass functional "
                                          "stage by stage: Shu-Osher integrators with nonlinear equations only")
            eng.set_option(L.OPT_WD_DISPLACED_MASS, displaced)
            eng.set_option(L.OPT_LAX_FRIEDRICHS, bool(_opt(eqo, "use_lax_friedrichs_velocity", True)))
            eng.set_option(L.OPT_GRAD_DIV_VISCOSITY, bool(_opt(eqo, "use_grad_div_viscosity_term", False)))
            eng.set_option(L.OPT_GRAD_DEPTH_VISCOSITY, bool(_opt(eqo, "use_grad_depth_viscosity_term", True)))
            eng.set_option(L.OPT_MOMENTUM_ADVECTION, not self._modesplit)
        else:
            eng.set_option(L.OPT_LF_TRACER, bool(_opt(eqo, "use_lax_friedrichs_tracer", False)))
            cons = False
            tr = _opt(eqo, "tracer", None)
            lab = self._tracer_label()
            if tr and lab in tr:
                cons = bool(getattr(tr[lab], "use_conservative_form", False))
            self._conservative = cons
        self._push_dynamic(force=True)

    def _stamp(self, obj):
        """Cheap change stamp of a datum: the value itself for None / plain numbers, (identity, version) for
        Constants and Functions that carry a version counter; None = unknown (always re-read)."""
        if obj is None or isinstance(obj, (int, float)):
            return ("v", obj)
        ver = _version(obj)
        return None if ver is None else (id(obj), ver)

    def _push_option(self, key, opt_id, obj, default):
        """Scalar option from None | number | Constant; re-read only when its stamp changed."""
        st = self._stamp(obj)
        if st is not None and self._stamps.get(key) == st:
            return
        self.engine.set_option(opt_id, default if obj is None else float(constant_value(obj)[0]))
        self._stamps[key] = st

    def _set_field(self, fid, value, key):
        st = self._stamp(value)
        if st is not None and self._stamps.get(("field", key)) == st:
            return
        self._stamps[("field", key)] = st
        eng = self.engine
        if value is None:
            if self._field_versions.get(key, "unset") != "none":
                eng.set_field(fid, None)
                self._field_versions[key] = "none"
            return
        if is_constant(value):
            v = constant_value(value)
            stamp = ("c",) + tuple(v.tolist())
            if self._field_versions.get(key) != stamp:
                eng.set_field(fid, v if v.size > 1 else float(v[0]))
                self._field_versions[key] = stamp
            return
        if is_function(value) or is_expression(value):
            # Functions, and expressions that are affine in their Function operands (evaluated nodally: exact)
            ver = _version(value)
            stamp = ("f", id(value), ver)
            if ver is None or self._field_versions.get(key) != stamp:
                if key in _CONTINUOUS_ONLY:
                    eng.set_field(fid, self.adaptor.vertex_values(value))
                else:
                    eng.set_field(fid, self.adaptor.coefficient_values(value)[1])    # P1 column or P1DG cell nodes
                self._field_versions[key] = stamp
            return
        raise NotImplementedError(f"coefficient {key!r}: unsupported value {type(value).__name__}")

    def _push_dynamic(self, force=False, functions=True):
        """Everything `update_forcings` may have changed: Constants are re-read, Functions re-uploaded if touched.
        ``functions=False`` leaves Function-valued `fields` entries at their last uploaded values (the lagged
        `fields_old` of timeintegrator.ForwardEuler); boundary data are always live."""
        eng = self.engine
        if self._kind == "swe" and functions and not force and self._watch_fast():
            return
        if self._kind == "swe":
            gc = getattr(self.equation, "physical_constants", None)
            if gc is None:
                try:
                    from thetis.physical_constants import physical_constants as gc   # live Thetis install
                except ImportError:
                    gc = None
            if gc is not None:
                self._push_option("g", L.OPT_G_GRAV, gc["g_grav"], 9.81)
                self._push_option("rho0", L.OPT_RHO0, gc["rho0"], 1000.0)
                if "von_karman" in gc:
                    self._push_option("kappa", L.OPT_VON_KARMAN, gc["von_karman"], 0.4)
            eqo = self.equation.options
            self._push_option("eps", L.OPT_NORM_SMOOTHER, _opt(eqo, "norm_smoother", None), 0.0)
            self._push_option("lf", L.OPT_LF_SCALING, self.fields.get("lax_friedrichs_velocity_scaling_factor"), 1.0)
            self._push_option("sipg", L.OPT_SIPG_FACTOR, _opt(eqo, "sipg_factor", None), 1.0)
            for name, fid in _SWE_FIELDS.items():
                val = self.fields.get(name)
                if self._modesplit and name not in _MODESPLIT_FIELDS:
                    val = None      # ModeSplit2DEquations has no term that reads this field (shallowwater_eq.py:953-957)
                if functions or val is None or not is_function(val):
                    self._set_field(fid, val, name)
            self._push_bcs(0, _SWE_TAGS)
            if functions:
                self._watch_build(gc)
        else:
            # several tracer integrators share one device context: equation-specific switches are re-sent every stage
            # (cached by value in the engine, so only changes reach the library) and this integrator's own upload
            # stamps are dropped whenever another tracer integrator configured the context in between
            if getattr(eng, "_tracer_cfg_owner", None) is not self:
                if getattr(eng, "_tracer_cfg_owner", None) is not None:
                    self._field_versions = {}
                    self._bc_versions = {}
                    self._stamps = {}
                eng._tracer_cfg_owner = self
            eng.set_option(L.OPT_TRACER_CONSERVATIVE, self._conservative)
            self._push_option("sipg", L.OPT_SIPG_FACTOR_TRACER, _opt(self.equation.options, "sipg_factor_tracer", None), 1.0)
            diff = self._tracer_field("diffusivity_h")
            if functions or diff is None or not is_function(diff):
                self._set_field(L.F_DIFFUSIVITY, diff, "diffusivity_h")
            self._push_option("lf", L.OPT_LF_TRACER_SCALING, self.fields.get("lax_friedrichs_tracer_scaling_factor"), 1.0)
            cf = self.fields.get("tracer_advective_velocity_factor")
            if cf is not None and not is_constant(cf):
                raise NotImplementedError("spatially varying tracer_advective_velocity_factor is outside the accelerated path")
            self._push_option("corr", L.OPT_TRACER_VEL_FACTOR, cf, 1.0)
            src = self._tracer_field("source")
            if functions or src is None or not is_function(src):
                self._set_field(L.F_TRACER_SOURCE, src, "tracer_source")
            self._push_bcs(1, _TRACER_TAGS)
        eng.sync_fields()      # pending coefficient uploads go out now (the stage may be a CUDA-graph replay)

    # -- watch list: after the first full pass, a stage only looks at the version counters of the objects that CAN
    # change (Constants, Functions, leaves of expressions) and re-pushes just the consumers of those that did.  The
    # reference builds its forms once from the `fields` / `bnd_conditions` entries present at construction, so
    # entries replaced later are not part of the contract; a cheap identity signature still catches that and falls
    # back to the full pass.
    def _watch_signature(self):
        sig = [id(v) for v in self.fields.values()]
        for funcs in self.bnd_conditions.values():
            if funcs:
                sig.extend(id(v) for v in funcs.values())
        return tuple(sig)

    def _watch_build(self, gc):
        watch = []          # [object, last version, [actions], dat, feeds something other than a boundary array]
        index = {}
        self._bc_array_items = []

        def add(obj, action, array_only=False):
            if obj is None or isinstance(obj, (int, float, tuple, list, np.ndarray)):
                return
            leaves = expression_leaves(obj) if is_expression(obj) else [obj]
            for leaf in leaves:
                ent = index.get(id(leaf))
                if ent is None:
                    d = getattr(leaf, "dat", None)
                    ent = index[id(leaf)] = [leaf, _version(leaf), [], d if hasattr(d, "dat_version") else None, False]
                    watch.append(ent)
                ent[2].append(action)
                if not array_only:
                    ent[4] = True
        eqo = self.equation.options
        if gc is not None:
            add(gc["g_grav"], lambda: self._push_option("g", L.OPT_G_GRAV, gc["g_grav"], 9.81))
            add(gc["rho0"], lambda: self._push_option("rho0", L.OPT_RHO0, gc["rho0"], 1000.0))
            if "von_karman" in gc:
                add(gc["von_karman"], lambda: self._push_option("kappa", L.OPT_VON_KARMAN, gc["von_karman"], 0.4))
        for key, opt_id, obj, default in (("eps", L.OPT_NORM_SMOOTHER, _opt(eqo, "norm_smoother", None), 0.0),
                                          ("lf", L.OPT_LF_SCALING, self.fields.get("lax_friedrichs_velocity_scaling_factor"), 1.0),
                                          ("sipg", L.OPT_SIPG_FACTOR, _opt(eqo, "sipg_factor", None), 1.0)):
            add(obj, lambda key=key, opt_id=opt_id, obj=obj, default=default: self._push_option(key, opt_id, obj, default))
        for name, fid in _SWE_FIELDS.items():
            val = self.fields.get(name)
            if self._modesplit and name not in _MODESPLIT_FIELDS:
                continue
            add(val, lambda fid=fid, val=val, name=name: self._set_field(fid, val, name))
        for marker, funcs in self.bnd_conditions.items():
            for tag, val in (funcs or {}).items():
                if is_constant(val):
                    # a Constant is a kernel parameter of the marker's slot: re-push the slot
                    add(val, lambda marker=marker, funcs=funcs: self._push_bc_marker(0, _SWE_TAGS, marker, funcs))
                else:
                    # Function / expression data: only that array travels (the tidal elevation of every stage)
                    add(val, lambda marker=marker, tag=tag, val=val: self._push_bc_array(0, _SWE_TAGS, marker, tag, val),
                        array_only=True)
                    if tag in _SWE_TAGS:
                        self._bc_array_items.append((marker, tag, val))
        self._watch = watch
        self._watch_sig = self._watch_signature()
        # a datum without a version counter must be re-read every stage: no fast path then
        self._watch_ok = all(ent[1] is not None for ent in watch)

    def _watch_fast(self):
        """True when the stage's dynamic inputs were refreshed through the watch list (else: do the full pass)."""
        if not getattr(self, "_watch_ok", False) or self._watch_signature() != self._watch_sig:
            return False
        dirty = False
        for ent in self._watch:
            d = ent[3]
            v = ("dat", d.dat_version) if d is not None else _version(ent[0])
            if v != ent[1]:
                ent[1] = v
                for act in ent[2]:
                    act()
                dirty = True
        if dirty:
            self.engine.sync_fields()
        return True

    def _push_bcs(self, eq, tags):
        eng = self.engine
        if eq == 1 and not self._stamps.get("bc_cleared"):
            # the device context is shared by every tracer on the mesh while each tracer equation is built from its
            # own bnd_conditions dict (solver2d.py:580-598): markers without an entry here must not inherit another
            # tracer's opcode / data.  (_stamps is dropped whenever another tracer configured the context.)
            for marker in self.adaptor.mesh.unique_markers():
                if self.bnd_conditions.get(marker) is None:
                    eng.clear_bc(1, marker)
            self._stamps["bc_cleared"] = True
        for marker, funcs in self.bnd_conditions.items():
            if funcs is None:
                continue
            self._push_bc_marker(eq, tags, marker, funcs)

    def _push_bc_marker(self, eq, tags, marker, funcs):
        eng = self.engine
        # fast path: nothing in this marker's dict changed since the last stage
        sig = []
        for tag, val in funcs.items():
            st = self._stamp(val)
            if st is None:
                sig = None
                break
            sig.append((tag, st))
        if sig is not None:
            sig = tuple(sig)
            if self._stamps.get(("bc", eq, marker)) == sig:
                return
        self._stamps[("bc", eq, marker)] = sig
        op = 0
        consts = np.zeros(8)
        arrays = []
        for tag, val in funcs.items():
            if tag not in tags:
                if eq == 0:
                    raise Exception(f'Invalid boundary tag "{tag}" specified on boundary {marker}')
                continue
            if tag == "drag" and self._modesplit:
                continue        # ModeSplit2DEquations accepts the tag but has no BoundaryDragTerm (shallowwater_eq.py:953-957)
            op |= tags[tag]
            if is_constant(val):
                v = constant_value(val)
                s = _CONST_SLOT[tag]
                consts[s:s + v.size] = v
            elif is_function(val) or is_expression(val):
                arrays.append((tag, val))     # expressions: affine in their Functions, evaluated nodally (exact)
            else:
                raise NotImplementedError(f"boundary datum {tag!r} on marker {marker}: unsupported value "
                                          f"{type(val).__name__}")
        stamp = (op,) + tuple(consts.tolist())
        key = (eq, marker)
        if self._bc_versions.get(key) != stamp:
            eng.set_bc(eq, marker, op, consts)
            self._bc_versions[key] = stamp
            for tag, _ in arrays:
                self._bc_versions.pop((eq, marker, tag), None)
        for tag, val in arrays:
            self._push_bc_array(eq, tags, marker, tag, val)

    def _push_bc_array(self, eq, tags, marker, tag, val, bank=0):
        """Upload one Function- / expression-valued boundary datum if it changed (pinned ring + async H2D).  ``bank``:
        the copy of the boundary arrays RK stage ``bank`` of a graph-replayed step reads (tb_set_bc_bank)."""
        ver = _version(val)
        akey = (eq, marker, tag) if bank == 0 else (eq, marker, tag, bank)
        astamp = (id(val), ver)
        if ver is None or self._bc_versions.get(akey) != astamp:
            self.engine.set_bc_bank(bank)
            self.engine.set_bc_array(eq, marker, tags[tag], self.adaptor.bfacet_values(val, marker))
            self.engine.set_bc_bank(0)
            self._bc_versions[akey] = astamp
            # keep the marker-level fast-path signature of _push_bc_marker consistent with what is on the device
            self._stamps.pop(("bc", eq, marker), None)

    # ------------------------------------------------------------------ host <-> device
    def _solution_version(self):
        if self._kind == "swe":
            return tuple(_version(f) for f in self.solution.subfunctions)
        return (_version(self.solution),)

    def upload(self):
        """H2D: host `solution` -> device state (buffer 0)."""
        eng = self.engine
        if self._kind == "swe":
            uv_f, eta_f = self.solution.subfunctions
            self._h_uv.numpy()[...] = self.adaptor.dat_ro(uv_f).reshape(-1, 2)
            self._h_eta.numpy()[...] = self.adaptor.dat_ro(eta_f)
            self._d_uv.copy_(self._h_uv, non_blocking=True)
            self._d_eta.copy_(self._h_eta, non_blocking=True)
            eng.state_from_fields(self._d_uv, self._d_eta, self.node_map, self.buf[0])
        else:
            self._h_q.numpy()[...] = self.adaptor.dat_ro(self.solution)
            self._d_q.copy_(self._h_q, non_blocking=True)
            eng.tracer_from_field(self._d_q, self.node_map, self.buf[0])
        if self.halo is not None:
            self.halo.exchange(self.buf[0])        # ghost records come from their owners, not from the host copy
        self._host_stale = False
        self._last_host_version = self._solution_version()

    def _visible_state(self):
        """The device buffer a host observer should see: the solution, or -- between the stages of a step, where the
        reference leaves the stage solution in `self.solution` (rungekutta.py:936-946) -- the latest stage."""
        return self._cur if getattr(self, "_mid_step", False) else self.buf[0]

    def sync_to_host(self):
        """D2H: device state -> `solution.dat.data` (in place; sub-function views stay valid).  Blocking; exports that
        can run behind the time loop use `stage_export()` / `ExportStage.wait()` instead."""
        if not self._host_stale:
            return
        eng = self.engine
        src = self._visible_state()
        if self._kind == "swe":
            eng.state_to_fields(src, self.node_map, self._d_uv, self._d_eta)
            self._h_uv.copy_(self._d_uv, non_blocking=True)
            self._h_eta.copy_(self._d_eta, non_blocking=True)
            torch.cuda.current_stream(eng.device).synchronize()
            uv_f, eta_f = self.solution.subfunctions
            ad = self.adaptor
            ad.dat_rw(uv_f)[...] = self._h_uv.numpy().reshape(ad.dat_ro(uv_f).shape)
            ad.dat_rw(eta_f)[...] = self._h_eta.numpy()
        else:
            eng.tracer_to_field(src, self.node_map, self._d_q)
            self._h_q.copy_(self._d_q, non_blocking=True)
            torch.cuda.current_stream(eng.device).synchronize()
            self.adaptor.dat_rw(self.solution)[...] = self._h_q.numpy()
        self._host_stale = False
        self._last_host_version = self._solution_version()

    def stage_export(self):
        """
        Non-blocking export staging (the D2H the reference does implicitly whenever `export()` / `print_state` read
        `solution.dat.data`, solver2d.py:1132-1142): the current device solution is converted to the Thetis layout and
        copied into one of two pinned host buffers on a SIDE stream; the time loop keeps launching stages on the main
        stream.  Returns an `ExportStage`; `wait()` blocks on its event only and hands out numpy views of the pinned
        buffers (valid until the second-next `stage_export()`).  The host `solution` is not touched.
        """
        eng = self.engine
        dev = eng.device
        ring = self.__dict__.get("_export_ring")
        if ring is None:
            side = torch.cuda.Stream(device=dev)
            slots = []
            for _ in range(2):
                if self._kind == "swe":
                    d = (torch.empty_like(self._d_uv), torch.empty_like(self._d_eta))
                    h = (torch.empty(self._d_uv.shape, dtype=torch.float64).pin_memory(),
                         torch.empty(self._d_eta.shape, dtype=torch.float64).pin_memory())
                else:
                    d = (torch.empty_like(self._d_q),)
                    h = (torch.empty(self._d_q.shape, dtype=torch.float64).pin_memory(),)
                slots.append(dict(d=d, h=h, done=torch.cuda.Event(), snap=torch.empty_like(self.buf[0])))
            ring = self._export_ring = dict(side=side, slots=slots, next=0)
        slot = ring["slots"][ring["next"]]
        ring["next"] ^= 1
        main = torch.cuda.current_stream(dev)
        slot["done"].synchronize()                      # the copy that used this slot two exports ago has landed
        # snapshot on the main stream (device-to-device, HBM speed): later stages may overwrite the solution buffer
        slot["snap"].copy_(self._visible_state(), non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(main)
        side = ring["side"]
        side.wait_event(ready)
        with torch.cuda.stream(side):
            if self._kind == "swe":
                eng.state_to_fields(slot["snap"], self.node_map, slot["d"][0], slot["d"][1])
            else:
                eng.tracer_to_field(slot["snap"], self.node_map, slot["d"][0])
            for dd, hh in zip(slot["d"], slot["h"]):
                hh.copy_(dd, non_blocking=True)
            slot["done"].record(side)
        return ExportStage(slot["done"], tuple(h.numpy() for h in slot["h"]))

    def device_state(self):
        """Device tensor holding the current solution (cell records)."""
        return self.buf[0]

    def mark_device_modified(self):
        """Another device operator (the limiter) changed buffer 0 in place."""
        self._host_stale = True

    def _host_changed(self):
        v = self._solution_version()
        if any(x is None for x in v):
            return not self._host_stale and self.sync_policy == "every_step"
        return v != self._last_host_version

    # ------------------------------------------------------------------ TimeIntegrator API
    def initialize(self, solution):
        """Assign initial conditions (timeintegrator.py:33-39): upload `solution` to the device."""
        if solution is not self.solution and solution is not None:
            self.solution.assign(solution)
        self.upload()

    def set_dt(self, dt):
        """Update time step (timeintegrator.py:70-73)."""
        self.dt = float(dt)

    def update_solver(self):
        """Kept for API compatibility (rungekutta.py:913-919): the mass solve is closed-form, nothing to rebuild."""
        return None

    def _swe_state_for_tracer(self):
        eng = self.engine
        sw = eng.swe_stepper
        uv_f = self.fields.get("uv_2d")
        if sw is not None and uv_f is not None and any(uv_f is s for s in sw.solution.subfunctions):
            if not sw._host_stale and sw._host_changed():
                sw.upload()
            if sw.halo is not None and not (self.halo is not None and self.halo.fused):
                sw.halo.wait_ghosts()         # the tracer kernel reads the SWE records of its halo cells (a fused
                                              # tracer launch waits for them itself: one epoch sequence)
            return sw.device_state()
        # tracer-only run: velocity / elevation come from host Functions
        if uv_f is None:
            raise NotImplementedError("tracer equation without uv_2d is trivial; not on the accelerated path")
        if self._own_swe_state is None:
            self._own_swe_state = self.halo.alloc(9, nbuf=1)[0] if self.halo is not None else eng.new_state()
        el_f = self.fields.get("elev_2d")
        uvd = torch.as_tensor(np.ascontiguousarray(self.adaptor.dat_ro(uv_f).reshape(-1, 2))).to(eng.device)
        if el_f is not None:
            ed = torch.as_tensor(np.ascontiguousarray(self.adaptor.dat_ro(el_f))).to(eng.device)
        else:
            ed = torch.zeros(uvd.shape[0], dtype=torch.float64, device=eng.device)
        eng.state_from_fields(uvd, ed, self.node_map, self._own_swe_state)
        if self.halo is not None:
            self.halo.exchange(self._own_swe_state)
        return self._own_swe_state

    def solve_stage(self, i_stage, t, update_forcings=None):
        """Solve i-th stage and leave it in the device solution (rungekutta.py:929-946)."""
        if update_forcings is not None:
            update_forcings(t + self.c[i_stage] * self.dt)
        if i_stage == 0 and not self._host_stale and self._host_changed():
            self.upload()                      # the host copy was modified since the last sync
        self._push_dynamic()
        graphs = getattr(self, "stage_graphs", None)
        if graphs:
            graphs[i_stage].replay()           # same launches, captured once (the forcing arrays are read at replay time)
            if i_stage == self.n_stages - 1:
                self._host_stale = True
        else:
            self._launch_stage(i_stage)
        if self.sync_policy == "every_stage":
            self.sync_to_host()                # the reference leaves every stage solution in `solution` (:936-946)

    def _launch_stage(self, i_stage):
        """One fused kernel launch: residual + mass inverse + Shu-Osher update of stage i."""
        eng = self.engine
        a0 = float(self.alpha[i_stage + 1][0]) if i_stage > 0 else 0.0
        a1 = float(self.alpha[i_stage + 1][i_stage]) if i_stage > 0 else float(self.alpha[1][0])
        bdt = float(self.beta[i_stage + 1][i_stage]) * self.dt
        last = i_stage == self.n_stages - 1
        A, B, Cb = self.buf
        # stage buffers rotate A -> B -> C -> ... and the last stage writes back into A (u0 may alias u_out)
        if i_stage == 0:
            src, dst = A, B
        else:
            src = self._cur
            dst = A if last else (Cb if src is B else B)
        self._cur = dst
        u0 = A if i_stage > 0 else None
        if self._kind == "swe":
            # optional fused diagnostics of the new solution (print_state norms / volume) in the last stage's epilogue
            fused = last and self.fused_norms is not None
            if fused:
                eng.stage_integrals(True)
            if self.halo is not None:
                self.halo.swe_stage(a0, a1, bdt, src, u0, dst)        # + one halo exchange per stage
            else:
                eng.swe_stage(a0, a1, bdt, src, u0, dst)
            if fused:
                eng.stage_integrals(False)
                eng.stage_integrals_finish(self.fused_norms)          # rank-local; callers all-reduce on a distributed mesh
        elif self.halo is not None:
            self.halo.tracer_stage(a0, a1, bdt, src, u0, dst, self._swe_state_for_tracer())
        else:
            eng.tracer_stage(a0, a1, bdt, src, u0, dst, self._swe_state_for_tracer())
        self._mid_step = not last
        self._host_stale = True
        if last:
            if dst is not A:
                self.buf[0], self.buf[1] = self.buf[1], self.buf[0]

    # -- one CUDA graph per STEP on the reference-facing path.  `update_forcings(t + c_i dt)` still runs before every
    # stage (rungekutta.py:933-934), but its effect is gathered first: the Function-valued boundary data it assigned
    # for stage i go into bank i of the device arrays, then ONE graph replays all stages, stage i reading bank i.
    # Anything else that changes inside a step (a Constant, a coefficient field, the entries of the dicts) cannot be
    # replayed from a graph with baked parameters: the step is then done stage by stage as usual and the graph dropped.
    def enable_step_graph(self):
        if self._kind != "swe" or self.butcher_form or self.n_stages < 2 or self.n_stages > 4 or \
                self.sync_policy == "every_stage":
            raise NotImplementedError("step graph: Shu-Osher SWE integrators with 2-4 stages, sync_policy != 'every_stage'")
        self.step_graph = None
        self._push_dynamic()                                  # full pass: builds the watch list
        for i in range(self.n_stages):                        # allocate / fill every bank before the capture
            if self._push_banked(i):
                raise NotImplementedError("step graph: a datum without a version counter is re-read every stage")
        saved = self.buf[0].clone()
        self.advance_device()                                 # warm-up outside the capture (lazy allocations) ...
        self.buf[0].copy_(saved)                              # ... on a copy: the solution is not advanced
        torch.cuda.synchronize(self.engine.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(self.n_stages):
                self.engine.set_bc_bank(i)
                self._launch_stage(i)
            self.engine.set_bc_bank(0)
        self.step_graph = g
        return g

    def _push_banked(self, i_stage):
        """Stage i of a graph-replayed step: upload the boundary arrays that differ from what bank i holds.  Returns
        True when something the graph cannot honour changed (the caller falls back to the stage-by-stage path)."""
        if not getattr(self, "_watch_ok", False) or self._watch_signature() != self._watch_sig:
            return True
        for ent in self._watch:
            d = ent[3]
            v = ("dat", d.dat_version) if d is not None else _version(ent[0])
            if v != ent[1]:
                if ent[4]:
                    return True            # a Constant / coefficient field moved: not an array-only change
                ent[1] = v
        for marker, tag, val in self._bc_array_items:
            self._push_bc_array(0, _SWE_TAGS, marker, tag, val, bank=i_stage)
        return False

    def _advance_step_graph(self, t, update_forcings):
        for i in range(self.n_stages):
            if update_forcings is not None:
                update_forcings(t + self.c[i] * self.dt)
            if i == 0 and not self._host_stale and self._host_changed():
                self.upload()
            if self._push_banked(i):
                return False
        self.step_graph.replay()
        self._mid_step = False
        self._host_stale = True
        return True

    def advance_device(self):
        """One step with every input already resident on the device (no forcing refresh, no host sync)."""
        for i in range(self.n_stages):
            self._launch_stage(i)

    def advance(self, t, update_forcings=None):
        """Advances equations for one time step (rungekutta.py:948-952)."""
        if getattr(self, "step_graph", None) is not None:
            if self._advance_step_graph(t, update_forcings):
                if self.sync_policy == "every_step":
                    self.sync_to_host()
                return
            self.step_graph = None          # something other than boundary arrays changed: stage by stage from now on
        for i in range(self.n_stages):
            self.solve_stage(i, t, update_forcings)
        if self.sync_policy == "every_step":
            self.sync_to_host()


class ExportStage:
    """Handle of one non-blocking export (`stage_export`): `wait()` -> numpy views of the pinned host copy."""

    def __init__(self, event, arrays):
        self.event = event
        self._arrays = arrays

    def ready(self):
        return self.event.query()

    def wait(self):
        self.event.synchronize()
        return self._arrays


class SSPRK22(ERKGenericShuOsher):
    """
    SSP(2,2):  u1 = u0 + dt F(u0),  u = (u0 + u1 + dt F(u1)) / 2,  forcings at t and t + dt -- the explicit scheme
    `timeintegrator.SSPRK22ALE` applies to the 3-D fields (timeintegrator.py:609-660, c = [0, 1]), here as an
    explicit `integrator_2d` for the external mode of the two-stage coupled loop
    (thetis_b200.coupled_timeintegrator.CoupledTwoStageRK2D).
    """
    a = [[0, 0], [1.0, 0]]
    b = [0.5, 0.5]
    c = [0, 1.0]
    cfl_coeff = 1.0


class SSPRK33(ERKGenericShuOsher):
    """3rd order SSP(3,3): tableau of thetis/rungekutta.py:342-347."""
    a = [[0, 0, 0], [1.0, 0, 0], [0.25, 0.25, 0]]
    b = [1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0]
    c = [0, 1.0, 0.5]
    cfl_coeff = 1.0


class ForwardEuler(ERKGenericShuOsher):
    """
    `thetis.timeintegrator.ForwardEuler` (timeintegrator.py:115-165) through the same stage kernel.  The reference
    calls ``update_forcings(t + dt)`` and assembles with `fields_old`: Function-valued coefficients are the ones of
    the end of the previous step (`update_fields_old`, :163-165), Constants and boundary data are live.
    """
    a = [[0]]
    b = [1.0]
    c = [0]
    cfl_coeff = 1.0

    def initialize(self, solution):
        super().initialize(solution)
        self._push_dynamic()                         # update_fields_old (timeintegrator.py:152-153)

    def solve_stage(self, i_stage, t, update_forcings=None):
        if update_forcings is not None:
            update_forcings(t + self.dt)             # timeintegrator.py:158-159
        if not self._host_stale and self._host_changed():
            self.upload()
        self._push_dynamic(functions=False)          # fields_old: Functions lag one step
        self._launch_stage(0)
        self._push_dynamic()                         # update_fields_old (:165)


class ERKGeneric(ERKGenericShuOsher):
    """
    Generic explicit Runge-Kutta integrator in Butcher form, `thetis.rungekutta.ERKGeneric`
    (rungekutta.py:762-867): stage i evaluates  k_i = dt M^-1 R(u_old + sum_j a_ij k_j)  with the forcings at
    t + c_i dt; the step ends with  u = u_old + sum_j b_j k_j  (get_final_solution).  Each tendency is one launch of
    the fused stage kernel (a0 = a1 = 0); the stage combinations are one streaming `tb_lincomb` launch each and the
    last tendency is fused with the final combination.
    Buffers: buf[0] = solution (u_old between steps), buf[1] = stage solution / scratch, buf[2+i] = k_i; the last
    tendency buffer receives the new solution and swaps roles with buf[0] at the end of every step.
    """
    butcher_form = True

    def initialize(self, solution):
        """rungekutta.py:811-814"""
        super().initialize(solution)
        self._initialized = True

    def _stage_input(self, i_stage):
        """update_solution (rungekutta.py:816-828): u_old + sum_{j<i} a_ij k_j"""
        U, S = self.buf[0], self.buf[1]
        terms = [(1.0, U)] + [(float(self.a[i_stage][j]), self.buf[2 + j]) for j in range(i_stage)
                               if self.a[i_stage][j] != 0.0]
        if len(terms) == 1:
            return U
        self.engine.lincomb(terms, S)
        return S

    def _launch_tendency(self, src, a0, u0, bdt, dst):
        eng = self.engine
        if self._kind == "swe":
            if self.halo is not None:
                # not the fused push: the ghost blocks of the tendency buffers are also read by the stage
                # combinations (tb_lincomb), which the per-peer flags of the fused launch do not order
                self.halo.swe_stage(a0, 0.0, bdt, src, u0, dst, fused=False)
            else:
                eng.swe_stage(a0, 0.0, bdt, src, u0, dst)
        else:
            eng.tracer_stage(a0, 0.0, bdt, src, u0, dst, self._swe_state_for_tracer())
            if self.halo is not None:
                self.halo.exchange(dst)

    def _launch_stage(self, i_stage):
        last = i_stage == self.n_stages - 1
        src = self._stage_input(i_stage)
        if not last:
            self._launch_tendency(src, 0.0, None, self.dt, self.buf[2 + i_stage])
            return
        # get_final_solution (rungekutta.py:841-852) fused with the last tendency:
        #   u = [u_old + sum_{j<s-1} b_j k_j] + b_{s-1} dt M^-1 R(u_{s-1})
        # The result goes into the last tendency buffer K (never into U): on a distributed mesh the peers push their
        # new records into the ghost block of the output while this rank may still be reading the ghost block of U
        # for its own stage input, so U must stay intact until the next exchange.  K and U then swap roles.
        U = self.buf[0]
        K = self.buf[2 + i_stage]
        rec = 9 if self._kind == "swe" else 3
        n_own = self.engine.n_owned_pad * rec         # the ghost block of K is written by the exchange only
        terms = [(1.0, U)] + [(float(self.b[j]), self.buf[2 + j]) for j in range(i_stage) if self.b[j] != 0.0]
        if float(self.b[i_stage]) == 0.0:
            self.engine.lincomb(terms, K, length=n_own)
            if self.halo is not None:
                self.halo.exchange(K)
        else:
            if len(terms) > 1:
                self.engine.lincomb(terms, K, length=n_own)
                base = K                              # u0 may alias u_out: each CTA reads its patch of u0 before writing
            else:
                base = U
            fused = self._kind == "swe" and self.fused_norms is not None
            if fused:
                self.engine.stage_integrals(True)
            self._launch_tendency(src, 1.0, base, float(self.b[i_stage]) * self.dt, K)
            if fused:
                self.engine.stage_integrals(False)
                self.engine.stage_integrals_finish(self.fused_norms)
        self.buf[0], self.buf[2 + i_stage] = K, U
        self._host_stale = True

    def solve_stage(self, i_stage, t, update_forcings=None):
        """rungekutta.py:855-867: stage solution first, then the forcings at t + c_i dt, then the tendency."""
        if i_stage == 0 and not self._host_stale and self._host_changed():
            self.upload()
        if update_forcings is not None:
            update_forcings(t + self.c[i_stage] * self.dt)
        self._push_dynamic()
        self._launch_stage(i_stage)


class ERKLSPUM2(ERKGeneric):
    """3-stage 2nd-order ERK of Higueras et al. (2014), tableau of rungekutta.py:360-365."""
    a = [[0, 0, 0], [5.0 / 6.0, 0, 0], [11.0 / 24.0, 11.0 / 24.0, 0]]
    b = [24.0 / 55.0, 1.0 / 5.0, 4.0 / 11.0]
    c = [0, 5.0 / 6.0, 11.0 / 12.0]
    cfl_coeff = 1.2


class ERKLPUM2(ERKGeneric):
    """3-stage 2nd-order ERK of Higueras et al. (2014), tableau of rungekutta.py:379-384."""
    a = [[0, 0, 0], [0.5, 0, 0], [0.5, 0.5, 0]]
    b = [1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0]
    c = [0, 0.5, 1.0]
    cfl_coeff = 2.0


class ERKMidpoint(ERKGeneric):
    """Explicit midpoint rule, tableau of rungekutta.py:388-392."""
    a = [[0.0, 0.0], [0.5, 0.0]]
    b = [0.0, 1.0]
    c = [0.0, 0.5]
    cfl_coeff = 1.0


class ERKEuler(ERKGeneric):
    """Forward Euler in Butcher form (rungekutta.py:979, ForwardEulerAbstract :142-149)."""
    a = [[0]]
    b = [1.0]
    c = [0]
    cfl_coeff = 1.0
