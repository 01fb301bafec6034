"""
TEST DOUBLE for `thetis_b200.engine.Engine` (the ctypes wrapper of the C-ABI): same method names, state buffers as CPU
torch tensors in the device layout (cell records in the adaptor's cell order), and every stage evaluated by the CPU
oracle from exactly the configuration the host classes SENT -- options, coefficient fields (Constant / vertex column /
cell-node array), boundary slots (opcode + constants) and boundary arrays.

Purpose: run the reference-facing host classes (`thetis_b200.rungekutta.*`: adaptor classification, node maps,
version tracking, boundary pushes, Shu-Osher / Butcher coefficients, buffer rotation, host synchronisation) END TO END
on the CPU, so that their output can be compared with what the reference's own integrators produce
(tests/test_dropin_with_reference_objects.py).  The kernels themselves are tied to the same oracle by the `-m gpu`
tests.  Lives under tests/: the product never imports it and has no CPU path.
"""
import numpy as np
import torch

from thetis_b200 import _lib as L
from thetis_b200.mesh import FACET_NODES
from oracle.swe_oracle import SWEOracle, TracerOracle

_FIELD_NAMES = {
    L.F_CORIOLIS: "coriolis", L.F_MANNING: "manning_drag_coefficient", L.F_QUAD_DRAG: "quadratic_drag_coefficient",
    L.F_LINEAR_DRAG: "linear_drag_coefficient", L.F_WIND_STRESS: "wind_stress",
    L.F_ATM_PRESSURE: "atmospheric_pressure", L.F_MOMENTUM_SOURCE: "momentum_source",
    L.F_VOLUME_SOURCE: "volume_source", L.F_VISCOSITY: "viscosity_h", L.F_NIKURADSE: "nikuradse_bed_roughness",
}
_SWE_TAGS = {L.BC_ELEV: ("elev", 0, 1), L.BC_UV: ("uv", 1, 2), L.BC_UN: ("un", 3, 1), L.BC_FLUX: ("flux", 4, 1)}
_TRACER_TAGS = dict(_SWE_TAGS)
_SWE_TAGS[L.BC_DRAG] = ("drag", 7, 1)          # BoundaryDragTerm: shallow water only
_TRACER_TAGS.update({L.BC_VALUE: ("value", 5, 1), L.BC_DIFF_FLUX: ("diff_flux", 6, 1)})


class OracleEngine:
    def __init__(self, mesh):
        self.mesh = mesh
        self.device = torch.device("cpu")
        self.n_cells = self.n_owned = mesh.n_cells
        self.n_owned_pad = ((mesh.n_cells + 127) // 128) * 128
        self.state_len = self.n_owned_pad * 9
        self.tracer_len = self.n_owned_pad * 3
        self.swe_stepper = None
        self.opt = {L.OPT_G_GRAV: 9.81, L.OPT_RHO0: 1000.0, L.OPT_NONLINEAR: 1.0, L.OPT_LAX_FRIEDRICHS: 1.0,
                    L.OPT_LF_SCALING: 1.0, L.OPT_NORM_SMOOTHER: 0.0, L.OPT_WETTING_DRYING: 0.0, L.OPT_WD_ALPHA: 0.5,
                    L.OPT_SIPG_FACTOR: 1.0, L.OPT_GRAD_DIV_VISCOSITY: 0.0, L.OPT_GRAD_DEPTH_VISCOSITY: 1.0,
                    L.OPT_MOMENTUM_ADVECTION: 1.0, L.OPT_VON_KARMAN: 0.4, L.OPT_LF_TRACER: 0.0,
                    L.OPT_LF_TRACER_SCALING: 1.0, L.OPT_TRACER_VEL_FACTOR: 1.0, L.OPT_SIPG_FACTOR_TRACER: 1.0,
                    L.OPT_TRACER_CONSERVATIVE: 0.0}
        self.fields = {}
        self.bc = {0: {}, 1: {}}             # eq -> marker -> (opcode, consts)
        self.bc_arrays = {}                  # (eq, marker, tag bit) -> (nb, 2[, 2]) over all exterior facets
        self.n_stage_launches = 0

    # ---------------------------------------------------------------- configuration (what the C-ABI would receive)
    def set_option(self, opt, value):
        self.opt[opt] = float(value)

    def set_field(self, field, value):
        self.fields[field] = None if value is None else np.array(value, dtype=float, copy=True)

    def set_bc(self, eq, marker, opcode, consts=None):
        # like tb_set_bc: the slot is REPLACED and its array flags are reset (arrays must be sent again afterwards)
        self.bc[eq][marker] = (int(opcode), np.array(consts, dtype=float, copy=True))
        for k in [k for k in self.bc_arrays if k[0] == eq and k[1] == marker]:
            del self.bc_arrays[k]

    def clear_bc(self, eq, marker):
        self.bc[eq].pop(marker, None)
        for k in [k for k in self.bc_arrays if k[0] == eq and k[1] == marker]:
            del self.bc_arrays[k]

    def set_bc_array(self, eq, marker, tag, values):
        self.bc_arrays[(eq, marker, tag)] = np.array(values, dtype=float, copy=True)

    def set_bc_bank(self, bank):
        assert bank == 0, "the step-graph banks are a CUDA-graph device; not used on this path"

    def set_boundary_length(self, marker, length):
        pass                                 # the oracle computes the same lengths from the mesh (checked below)

    def sync_fields(self):
        pass

    def stage_integrals(self, on):
        pass

    # ---------------------------------------------------------------- buffers
    def new_state(self):
        return torch.zeros(self.state_len, dtype=torch.float64)

    def new_tracer(self):
        return torch.zeros(self.tracer_len, dtype=torch.float64)

    def _rec(self, buf, width):
        return buf.numpy().reshape(-1, width)[: self.n_cells]

    def state_from_fields(self, d_uv, d_eta, node_map, buf):
        nm = node_map.numpy().reshape(-1, 3)
        r = self._rec(buf, 9)
        r[:, :6] = d_uv.numpy()[nm].reshape(-1, 6)
        r[:, 6:] = d_eta.numpy()[nm]

    def state_to_fields(self, buf, node_map, d_uv, d_eta):
        nm = node_map.numpy().reshape(-1, 3)
        r = self._rec(buf, 9)
        d_uv.numpy()[nm] = r[:, :6].reshape(-1, 3, 2)
        d_eta.numpy()[nm] = r[:, 6:]

    def tracer_from_field(self, d_q, node_map, buf):
        self._rec(buf, 3)[...] = d_q.numpy()[node_map.numpy().reshape(-1, 3)]

    def tracer_to_field(self, buf, node_map, d_q):
        d_q.numpy()[node_map.numpy().reshape(-1, 3)] = self._rec(buf, 3)

    def lincomb(self, terms, out, length=None):
        n = out.numel() if length is None else int(length)
        acc = torch.zeros(n, dtype=torch.float64)
        for c, t in terms:
            acc += float(c) * t[:n]
        out[:n] = acc

    # ---------------------------------------------------------------- the oracle behind the stages
    def _nodal(self, a):
        """what the library does with a field: Constant, vertex column (per geometric vertex) or cell-node array"""
        m = self.mesh
        if a is None:
            return None
        if a.ndim == 0:
            return float(a)
        if a.shape == (2,) and m.n_vertices != 2:
            return tuple(a.tolist())
        if a.ndim >= 2 and a.shape[:2] == (m.n_cells, 3):
            return a
        assert a.shape[0] == m.n_vertices, a.shape
        return a[m.cells]

    def _bnd(self, eq, tags):
        m = self.mesh
        out = {}
        for marker, (op, consts) in self.bc[eq].items():
            funcs = {}
            for bit, (name, slot, n) in tags.items():
                if not op & bit:
                    continue
                arr = self.bc_arrays.get((eq, marker, bit))
                if arr is None:
                    funcs[name] = float(consts[slot]) if n == 1 else tuple(consts[slot:slot + n].tolist())
                else:
                    # only the rows of THIS marker are meaningful (the library reads no others for this slot)
                    sel = m.bf_marker == marker
                    full = np.zeros((m.n_cells, 3) + arr.shape[2:])
                    for side in range(2):
                        full[m.bf_cell[sel], FACET_NODES[m.bf_lf[sel], side]] = arr[sel, side]
                    funcs[name] = full
            out[marker] = funcs
        return out

    def swe_oracle(self):
        o, f = self.opt, self.fields
        al = self._nodal(f.get(L.F_WD_ALPHA))
        options = dict(use_nonlinear_equations=bool(o[L.OPT_NONLINEAR]),
                       use_lax_friedrichs_velocity=bool(o[L.OPT_LAX_FRIEDRICHS]),
                       use_wetting_and_drying=bool(o[L.OPT_WETTING_DRYING]),
                       wetting_and_drying_alpha=o[L.OPT_WD_ALPHA] if al is None else al,
                       norm_smoother=o[L.OPT_NORM_SMOOTHER],
                       use_grad_div_viscosity_term=bool(o[L.OPT_GRAD_DIV_VISCOSITY]),
                       use_grad_depth_viscosity_term=bool(o[L.OPT_GRAD_DEPTH_VISCOSITY]),
                       sipg_factor=o[L.OPT_SIPG_FACTOR], include_momentum_advection=bool(o[L.OPT_MOMENTUM_ADVECTION]))
        fields = {name: self._nodal(f.get(fid)) for fid, name in _FIELD_NAMES.items() if f.get(fid) is not None}
        fields["lax_friedrichs_velocity_scaling_factor"] = o[L.OPT_LF_SCALING]
        fields["von_karman"] = o[L.OPT_VON_KARMAN]
        return SWEOracle(self.mesh, self._nodal(f[L.F_BATHYMETRY]), options=options, fields=fields,
                         bnd_conditions=self._bnd(0, _SWE_TAGS), g_grav=o[L.OPT_G_GRAV], rho0=o[L.OPT_RHO0])

    def swe_stage(self, a0, a1, bdt, src, u0, dst):
        self.n_stage_launches += 1
        s = self._rec(src, 9)
        uv, eta = s[:, :6].reshape(-1, 3, 2).copy(), s[:, 6:].copy()
        orc = self.swe_oracle()
        ku, ke = orc.tendency(uv, eta)
        new = a1 * s + bdt * np.concatenate([ku.reshape(-1, 6), ke], axis=1)
        if u0 is not None:
            new = new + a0 * self._rec(u0, 9)
        if self.opt.get(L.OPT_WD_DISPLACED_MASS) and orc.options["use_wetting_and_drying"]:
            # TB_OPT_WD_DISPLACED_MASS: the stage advances the reference's mass functional of the elevation
            assert abs(a0 + a1 - 1.0) < 1e-9
            _, Re = orc.residual(uv, eta)
            target = a1 * orc.displaced_mass(eta) + bdt * Re
            if u0 is not None:
                target = target + a0 * orc.displaced_mass(self._rec(u0, 9)[:, 6:].copy())
            new[:, 6:] = orc.solve_displaced_mass(target, eta)
        self._rec(dst, 9)[...] = new

    def limiter_apply_to(self, c_in, c_out):
        from oracle.swe_oracle import vertex_based_limiter
        self._rec(c_out, 3)[...] = vertex_based_limiter(self.mesh, self._rec(c_in, 3).copy())

    def limiter_apply(self, c):
        self.limiter_apply_to(c, c)

    def tracer_stage(self, a0, a1, bdt, src, u0, dst, swe_state):
        self.n_stage_launches += 1
        o, f = self.opt, self.fields
        tf = {"tracer_advective_velocity_factor": o[L.OPT_TRACER_VEL_FACTOR],
              "lax_friedrichs_tracer_scaling_factor": o[L.OPT_LF_TRACER_SCALING]}
        if f.get(L.F_TRACER_SOURCE) is not None:
            tf["source"] = self._nodal(f[L.F_TRACER_SOURCE])
        if f.get(L.F_DIFFUSIVITY) is not None:
            tf["diffusivity_h"] = self._nodal(f[L.F_DIFFUSIVITY])
        tr = TracerOracle(self.swe_oracle(), bnd_conditions=self._bnd(1, _TRACER_TAGS), fields=tf,
                          options=dict(use_lax_friedrichs_tracer=bool(o[L.OPT_LF_TRACER]),
                                       use_conservative_form=bool(o[L.OPT_TRACER_CONSERVATIVE]),
                                       sipg_factor_tracer=o[L.OPT_SIPG_FACTOR_TRACER]))
        w = self._rec(swe_state, 9)
        tr.set_velocity(w[:, :6].reshape(-1, 3, 2).copy(), w[:, 6:].copy())
        c = self._rec(src, 3).copy()
        (kc,) = tr.tendency(c)
        new = a1 * c + bdt * kc
        if u0 is not None:
            new = new + a0 * self._rec(u0, 3)
        self._rec(dst, 3)[...] = new


class DistributedOracleEngine(OracleEngine):
    """
    The same double for ONE RANK of a distributed run: `mesh` = owned cells followed by ghost cells whose outer facets
    have no neighbour on this rank (INT32_MIN in `nbr`).  Patch size 1, so the ghost block follows the owned cells
    directly (`n_owned_pad == n_owned`) and a state buffer is one contiguous array of cell records.  Every stage is
    evaluated on ALL local cells with the unknown facets closed by a fake marker: the ghost results are meaningless
    and must be overwritten by the halo exchange that follows each stage -- which is exactly what is under test
    (`thetis_b200.parallel.HaloPlan` and the integrators' distributed code paths, on the CPU over gloo).
    """
    FAKE = 999

    def __init__(self, mesh, n_owned, boundary_len):
        from thetis_b200.mesh import Mesh2D
        nbr = mesh.nbr.astype(np.int64)
        unknown = nbr == np.iinfo(np.int32).min
        assert not unknown[:n_owned].any()
        nfake = int(unknown.sum())
        nbr[unknown] = -(1 + mesh.n_bfacets + np.arange(nfake))
        cu, fu = np.nonzero(unknown)
        closed = Mesh2D(coords=mesh.coords, cells=mesh.cells, topo=mesh.topo, periodic=mesh.periodic)
        closed.nbr, closed.nbr_lf = nbr.astype(np.int32), mesh.nbr_lf
        closed.bf_cell = np.concatenate([mesh.bf_cell, cu]).astype(np.int32)
        closed.bf_lf = np.concatenate([mesh.bf_lf, fu]).astype(np.int8)
        closed.bf_marker = np.concatenate([mesh.bf_marker, np.full(nfake, self.FAKE)]).astype(np.int32)
        super().__init__(closed)
        self.n_real_bfacets = mesh.n_bfacets
        self.n_owned = self.n_owned_pad = int(n_owned)
        self.patch_size, self.n_patches = 1, int(n_owned)
        self.state_len, self.tracer_len = mesh.n_cells * 9, mesh.n_cells * 3
        self.boundary_len = {int(k): float(v) for k, v in boundary_len.items()}
        self.n_gathers = 0

    def set_boundary_length(self, marker, length):
        self.boundary_len[int(marker)] = float(length)         # global lengths (flux boundary data divide by them)

    def set_bc_array(self, eq, marker, tag, values):
        v = np.array(values, dtype=float, copy=True)
        pad = np.zeros((self.mesh.n_bfacets - v.shape[0],) + v.shape[1:])
        super().set_bc_array(eq, marker, tag, np.concatenate([v, pad]))

    def swe_oracle(self):
        orc = super().swe_oracle()
        orc.boundary_len = dict(self.boundary_len)
        orc.boundary_len[self.FAKE] = 1.0
        return orc

    # the real kernels advance owned cells only and leave the ghost block of their output STALE: the double poisons
    # it, so that one missing or misdirected exchange reaches an owned cell as NaN in the very next stage (without
    # this, the redundancy of a vertex overlap hides a single dropped exchange)
    def _poison_ghosts(self, buf, width):
        buf.view(-1, width)[self.n_owned:] = float("nan")

    def swe_stage(self, a0, a1, bdt, src, u0, dst):
        super().swe_stage(a0, a1, bdt, src, u0, dst)
        self._poison_ghosts(dst, 9)

    def tracer_stage(self, a0, a1, bdt, src, u0, dst, swe_state):
        super().tracer_stage(a0, a1, bdt, src, u0, dst, swe_state)
        self._poison_ghosts(dst, 3)

    def limiter_apply_to(self, c_in, c_out):
        super().limiter_apply_to(c_in, c_out)
        self._poison_ghosts(c_out, 3)

    def gather_cells(self, state, idx, rec, out):
        self.n_gathers += 1
        out[: idx.numel()] = state.view(-1, rec)[idx.long()]
