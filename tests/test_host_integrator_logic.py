"""
Host logic of the reference-facing integrators WITHOUT a GPU: the device engine is replaced by a recording stand-in
(same method names as `thetis_b200.engine.Engine`), so that what the integrators SEND to the C-ABI -- and when -- is
checked on the CPU: classification of Constants / Functions / P1DG fields / expressions, the per-stage watch list,
boundary-data banks of the step graph, per-tracer boundary slots, Nikuradse / wetting-drying-alpha routing.
The numerical effect of every call is covered by the -m gpu parity tests; nothing here touches the oracle.
"""
import numpy as np
import pytest
import torch

from thetis_b200 import _lib as L
from thetis_b200 import mesh as M


class RecordingEngine:
    """Stand-in for Engine: records configuration calls, launches nothing."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.device = torch.device("cpu")
        self.n_cells = self.n_owned = mesh.n_cells
        self.n_owned_pad = ((mesh.n_cells + 127) // 128) * 128
        self.state_len = self.n_owned_pad * 9
        self.tracer_len = self.n_owned_pad * 3
        self.calls = []
        self.swe_stepper = None
        self.bank = 0
        self.options = {}

    def _rec(self, name, *a):
        self.calls.append((name,) + a)

    def new_state(self):
        return torch.zeros(self.state_len, dtype=torch.float64)

    def new_tracer(self):
        return torch.zeros(self.tracer_len, dtype=torch.float64)

    def set_option(self, opt, value):
        if self.options.get(opt) != float(value):
            self.options[opt] = float(value)
            self._rec("set_option", opt, float(value))

    def set_field(self, field, value):
        if value is None:
            kind = "none"
        else:
            a = np.asarray(value, dtype=float)
            kind = ("const" if a.ndim == 0 or a.shape == (2,) and self.mesh.n_vertices != 2 else
                    "cell" if a.ndim >= 2 and a.shape[:2] == (self.mesh.n_cells, 3) else "vertex")
        self._rec("set_field", field, kind, None if value is None else np.array(value, dtype=float, copy=True))

    def set_bc(self, eq, marker, opcode, consts=None):
        self._rec("set_bc", eq, marker, opcode, np.array(consts, dtype=float, copy=True))

    def set_bc_array(self, eq, marker, tag, values):
        self._rec("set_bc_array", eq, marker, tag, self.bank, np.array(values, dtype=float, copy=True))

    def set_bc_bank(self, bank):
        self.bank = bank

    def clear_bc(self, eq, marker):
        self._rec("clear_bc", eq, marker)

    def set_boundary_length(self, marker, length):
        pass

    def sync_fields(self):
        self._rec("sync_fields")

    def swe_stage(self, a0, a1, bdt, src, u0, dst):
        self._rec("swe_stage", round(a0, 12), round(a1, 12), bdt, self.bank)

    def tracer_stage(self, a0, a1, bdt, src, u0, dst, swe):
        self._rec("tracer_stage", round(a0, 12), round(a1, 12), bdt)

    def stage_integrals(self, on):
        pass

    def state_from_fields(self, *a):
        self._rec("state_from_fields")

    def state_to_fields(self, *a):
        self._rec("state_to_fields")

    def tracer_from_field(self, *a):
        self._rec("tracer_from_field")

    def tracer_to_field(self, *a):
        self._rec("tracer_to_field")

    def named(self, name):
        return [c for c in self.calls if c[0] == name]


@pytest.fixture
def setup(monkeypatch):
    """FlowSolver2d mirror whose mesh adaptor hands out a RecordingEngine."""
    from thetis_b200 import solver2d, adaptor
    from thetis_b200.shim import Function, FunctionSpace, as_shim_mesh
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: type("S", (), {"synchronize": lambda s: None})())
    engines = []

    def get_engine(self):
        if self.engine is None:
            self.engine = RecordingEngine(self.mesh)
            engines.append(self.engine)
        return self.engine
    monkeypatch.setattr(adaptor.MeshAdaptor, "get_engine", get_engine)

    def make(nx=6, ny=4, **opts):
        mesh = M.rectangle_mesh(nx, ny, 600.0, 400.0)
        sm = as_shim_mesh(mesh)
        P1 = FunctionSpace(sm, "CG", 1)
        b = Function(P1).assign(10.0)
        s = solver2d.FlowSolver2d(sm, b)
        s.options.swe_timestepper_options.use_automatic_timestep = False
        s.options.tracer_timestepper_options.use_automatic_timestep = False
        s.options.update(dict(timestep=1.0, simulation_end_time=3.0, no_exports=True))
        s.options.update(opts)
        return s, P1, mesh
    return make, engines


def test_static_configuration_and_shu_osher_coefficients(setup):
    make, engines = setup
    from thetis_b200.shim import Constant
    s, P1, mesh = make()
    s.options.manning_drag_coefficient = Constant(0.03)
    s.bnd_functions["shallow_water"] = {1: {"elev": Constant(0.5), "un": Constant(-0.1)}}
    s.assign_initial_conditions()
    eng = engines[0]
    fields = {c[1]: c[2] for c in eng.named("set_field")}
    assert fields[L.F_BATHYMETRY] == "vertex" and fields[L.F_MANNING] == "const" and fields[L.F_CORIOLIS] == "none"
    (bc,) = eng.named("set_bc")
    assert bc[1:4] == (0, 1, L.BC_ELEV | L.BC_UN) and bc[4][0] == 0.5 and bc[4][3] == -0.1
    eng.calls.clear()
    s.timestepper.advance(0.0)
    st = eng.named("swe_stage")
    # SSPRK33 in Shu-Osher form (rungekutta.py:342-347 through butcher_to_shuosher_form): (a0, a1, beta dt)
    assert [c[1:4] for c in st] == [(0.0, 1.0, 1.0), (0.75, 0.25, 0.25), (round(1 / 3, 12), round(2 / 3, 12), 2 / 3)]
    assert not eng.named("set_field") and not eng.named("set_bc")            # nothing changed: nothing re-sent


def test_watch_list_resends_only_what_changed(setup):
    make, engines = setup
    from thetis_b200.shim import Constant, Function
    s, P1, mesh = make()
    tide = Function(P1)
    ramp = Constant(0.0)
    wind = Function(__import__("thetis_b200.shim", fromlist=["FunctionSpace"]).FunctionSpace(s.mesh2d, "CG", 1, value_size=2))
    s.options.wind_stress = wind
    s.bnd_functions["shallow_water"] = {1: {"elev": ramp * tide, "uv": Constant(np.array([0.0, 0.0]))}}
    s.assign_initial_conditions()
    eng, ts = engines[0], s.timestepper
    seen = []

    def update_forcings(t):
        seen.append(t)
        ramp.assign(min(t / 2.0, 1.0))
        tide.interpolate(lambda x, y: np.sin(t) + 0 * x)
    ts.advance(0.0, update_forcings)
    assert seen == [0.0, 1.0, 0.5]                                            # t + c_i dt (rungekutta.py:933-934)
    eng.calls.clear()
    ts.advance(1.0, update_forcings)
    arr = eng.named("set_bc_array")
    assert len(arr) == 3 and all(a[1:4] == (0, 1, L.BC_ELEV) for a in arr)   # one upload per stage, of that tag only
    rows = mesh.bf_marker == 1
    for a, t in zip(arr, (1.0, 2.0, 1.5)):
        assert np.allclose(a[5][rows], min(t / 2.0, 1.0) * np.sin(t))        # the expression, evaluated at the facet nodes
    assert not eng.named("set_field") and not eng.named("set_bc")            # wind untouched: not re-sent
    eng.calls.clear()
    wind.interpolate(lambda x, y: (0.1 + 0 * x, 0 * x))
    ts.advance(2.0, update_forcings)
    sf = eng.named("set_field")
    assert len(sf) == 1 and sf[0][1] == L.F_WIND_STRESS and sf[0][2] == "vertex"


def test_discontinuous_fields_nikuradse_and_wd_alpha_routing(setup):
    make, engines = setup
    from thetis_b200.shim import Constant, Function, FunctionSpace
    s, P1, mesh = make(use_wetting_and_drying=True)
    H = FunctionSpace(s.mesh2d, "DG", 1)
    cor = Function(H)
    cor.dat.data[:] = np.random.default_rng(0).standard_normal(cor.dat.data_ro.shape)      # genuinely discontinuous
    s.options.coriolis_frequency = cor
    s.options.nikuradse_bed_roughness = Constant(0.05)
    s.options.wetting_and_drying_alpha = Function(P1).interpolate(lambda x, y: 0.3 + x / 6e3)
    s.assign_initial_conditions()
    eng = engines[0]
    fields = {c[1]: c for c in eng.named("set_field")}
    assert fields[L.F_CORIOLIS][2] == "cell" and fields[L.F_CORIOLIS][3].shape == (mesh.n_cells, 3)
    assert fields[L.F_NIKURADSE][2] == "const" and fields[L.F_WD_ALPHA][2] == "vertex"
    # a discontinuous coefficient of a facet term is refused before anything reaches the device
    s2, P1b, _ = make()
    s2.options.horizontal_viscosity = cor
    with pytest.raises(NotImplementedError, match="continuous"):
        s2.assign_initial_conditions()
    s3, _, _ = make()
    s3.options.nikuradse_bed_roughness = Constant(0.05)
    s3.options.quadratic_drag_coefficient = Constant(0.0025)
    with pytest.raises(Exception, match="Cannot set both Nikuradse"):
        s3.assign_initial_conditions()


def test_boundary_drag_tag_routing(setup):
    """'drag' (BoundaryDragTerm, shallowwater_eq.py:704-726) is a kernel parameter of the marker's slot: bit
    TB_BC_DRAG + consts[7], alone (the boundary stays closed) or next to open tags; a changed Constant re-sends just
    that slot; a Function-valued coefficient is refused at construction; an unknown tag raises like the reference
    (shallowwater_eq.py:291-293)."""
    make, engines = setup
    from thetis_b200.shim import Constant, Function
    s, P1, mesh = make()
    cd = Constant(0.05)
    s.bnd_functions["shallow_water"] = {1: {"drag": cd}, 2: {"elev": Constant(0.5), "drag": Constant(0.02)}}
    s.assign_initial_conditions()
    eng = engines[0]
    bcs = {c[2]: c for c in eng.named("set_bc")}
    assert bcs[1][3] == L.BC_DRAG and bcs[1][4][7] == 0.05 and not bcs[1][4][:7].any()
    assert bcs[2][3] == L.BC_ELEV | L.BC_DRAG and bcs[2][4][0] == 0.5 and bcs[2][4][7] == 0.02
    eng.calls.clear()
    s.timestepper.advance(0.0)
    assert not eng.named("set_bc")
    cd.assign(0.08)
    s.timestepper.advance(1.0)
    (bc,) = eng.named("set_bc")
    assert bc[2] == 1 and bc[3] == L.BC_DRAG and bc[4][7] == 0.08
    s2, P1b, _ = make()
    s2.bnd_functions["shallow_water"] = {1: {"drag": Function(P1b).assign(0.05)}}
    with pytest.raises(NotImplementedError, match="drag"):
        s2.assign_initial_conditions()
    s3, _, _ = make()
    s3.bnd_functions["shallow_water"] = {1: {"dragg": Constant(0.05)}}
    with pytest.raises(Exception, match="Invalid boundary tag"):
        s3.assign_initial_conditions()


def test_each_tracer_sees_only_its_own_boundary_conditions(setup):
    make, engines = setup
    from thetis_b200.shim import Constant
    s, P1, mesh = make(use_limiter_for_tracers=False)
    s.options.add_tracer_2d("salt_2d", "Salinity", "Salinity2d")
    s.options.add_tracer_2d("temp_2d", "Temperature", "Temperature2d")
    s.bnd_functions["salt"] = {1: {"value": Constant(35.0)}}
    s.bnd_functions["temp"] = {2: {"value": Constant(3.0)}}
    s.assign_initial_conditions()
    eng = engines[0]
    eng.calls.clear()
    s.timestepper.advance(0.0)
    # walk the call stream: at every tracer stage the device slots must equal that tracer's own dict
    slots, seen = {}, []
    for c in eng.calls:
        if c[0] == "clear_bc" and c[1] == 1:
            slots.pop(c[2], None)
        elif c[0] == "set_bc" and c[1] == 1:
            slots[c[2]] = c[4][5]
        elif c[0] == "tracer_stage":
            seen.append(dict(slots))
    assert len(seen) == 6
    assert seen[:3] == [{1: 35.0}] * 3 and seen[3:] == [{2: 3.0}] * 3


def test_step_graph_banks_and_fallback(setup):
    """_push_banked: the boundary array of stage i goes into bank i; a Constant moving inside a step is not something a
    replayed graph can honour, so the integrator reports it and falls back to the stage-by-stage path"""
    make, engines = setup
    from thetis_b200.shim import Constant, Function
    from thetis_b200.solver2d import physical_constants
    s, P1, mesh = make()
    tide = Function(P1)
    s.bnd_functions["shallow_water"] = {1: {"elev": tide, "uv": Constant(np.array([0.0, 0.0]))}}
    s.assign_initial_conditions()
    eng, ts = engines[0], s.timestepper
    ts._push_dynamic()                                   # full pass builds the watch list

    class FakeGraph:
        replays = 0

        def replay(self):
            FakeGraph.replays += 1
    ts.step_graph = FakeGraph()

    def update_forcings(t):
        tide.interpolate(lambda x, y: t + 0 * x)
    eng.calls.clear()
    ts.advance(0.0, update_forcings)
    assert FakeGraph.replays == 1 and not eng.named("swe_stage")
    arr = eng.named("set_bc_array")
    assert [a[4] for a in arr] == [0, 1, 2]              # stage i -> bank i
    rows = mesh.bf_marker == 1
    assert [float(a[5][rows][0, 0]) for a in arr] == [0.0, 1.0, 0.5]
    assert eng.bank == 0
    g_old = float(physical_constants["g_grav"])
    try:
        physical_constants["g_grav"].assign(9.0)
        eng.calls.clear()
        ts.advance(1.0, update_forcings)
    finally:
        physical_constants["g_grav"].assign(g_old)
    assert ts.step_graph is None and FakeGraph.replays == 1
    assert len(eng.named("swe_stage")) == 3 and ("set_option", L.OPT_G_GRAV, 9.0) in eng.calls
