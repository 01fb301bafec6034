"""
Device engine: one `tb_ctx` (include/thetis_b200.h) per mesh, torch tensors as
device buffers, CUDA stream taken from torch.  Everything numerical happens in
the CUDA library; this file only moves pointers around.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np
import torch

from . import _lib as L
from .mesh import Mesh2D, sfc_renumber

__all__ = ["Engine", "get_engine"]

_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
if _raw_stream is None:
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream


def _ptr(t):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


class Engine:
    """
    Owns the device context of one (sub)mesh.

    :arg mesh: `Mesh2D`, cells in the order the device should use (call
        `sfc_renumber` first for locality); cells [0, n_owned) are advanced,
        the rest are ghosts of a partition.
    """

    def __init__(self, mesh: Mesh2D, n_owned=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("thetis_b200 needs a CUDA device (there is no CPU fallback on this path)")
        self.lib = L.load()
        self.mesh = mesh
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self._dev_index = self.device.index
        self._fields_pending = True
        self.n_cells = mesh.n_cells
        self.n_owned = mesh.n_cells if n_owned is None else int(n_owned)
        tm = L.TbMesh()
        self._keep = dict(
            coords=np.ascontiguousarray(mesh.coords, dtype=np.float64),
            cells=np.ascontiguousarray(mesh.cells, dtype=np.int32),
            nbr=np.ascontiguousarray(mesh.nbr, dtype=np.int32),
            nbr_lf=np.ascontiguousarray(mesh.nbr_lf, dtype=np.int8),
            bf_marker=np.ascontiguousarray(mesh.bf_marker, dtype=np.int32),
            topo=np.ascontiguousarray(mesh.topo, dtype=np.int32),
        )
        tm.n_cells = self.n_cells
        tm.n_owned = self.n_owned
        tm.n_vertices = mesh.n_vertices
        tm.n_bfacets = mesh.bf_marker.shape[0]
        for k, a in self._keep.items():
            setattr(tm, k, a.ctypes.data)
        ctx = C.c_void_p()
        rc = self.lib.tb_create(C.byref(ctx), C.byref(tm), self.device.index)
        if rc != 0:
            raise L.TbError(f"tb_create failed ({rc}): {self.lib.tb_last_error(None).decode()}")
        self.ctx = ctx
        self._finalizer = weakref.finalize(self, self.lib.tb_destroy, ctx)
        self.state_len = int(self.lib.tb_state_len(ctx))
        self.tracer_len = int(self.lib.tb_tracer_len(ctx))
        self.patch_size = int(self.lib.tb_patch_size(ctx))
        self.n_patches = int(self.lib.tb_n_patches(ctx))
        self.n_owned_pad = self.n_patches * self.patch_size
        for m, ln in mesh.boundary_length().items():
            self.set_boundary_length(m, ln)
        self._opt_cache = {}
        self.swe_stepper = None      # set by the SWE integrator so tracer integrators can find the live state
        self._identity_map = None

    # ------------------------------------------------------------ helpers
    @property
    def stream(self):
        # raw handle of torch's current stream on this device (the C-level getter: ~10x cheaper than building a
        # torch.cuda.Stream object, and this is read on every call into the library)
        return C.c_void_p(_raw_stream(self._dev_index))

    def _ck(self, rc):
        L.check(self.ctx, rc)

    def new_state(self):
        return torch.zeros(self.state_len, dtype=torch.float64, device=self.device)

    def new_tracer(self):
        return torch.zeros(self.tracer_len, dtype=torch.float64, device=self.device)

    def launch_count(self):
        return int(self.lib.tb_launch_count(self.ctx))

    # ------------------------------------------------------------ configuration
    def set_option(self, opt, value):
        value = float(value)
        if self._opt_cache.get(opt) == value:      # options are re-read every stage; only changes reach the library
            return
        self._ck(self.lib.tb_set_option(self.ctx, opt, value))
        self._opt_cache[opt] = value

    def set_field(self, field, value):
        """value: None | scalar/sequence (Constant) | ndarray over geometric vertices (nv,) / (nv, 2) (P1) |
        ndarray over cell nodes (n_cells, 3) / (n_cells, 3, 2) (discontinuous P1DG, cell terms only)."""
        self._fields_pending = True
        if value is None:
            self._ck(self.lib.tb_clear_field(self.ctx, field))
            return
        a = np.ascontiguousarray(np.asarray(value, dtype=np.float64))
        ncomp = 2 if field in (L.F_WIND_STRESS, L.F_MOMENTUM_SOURCE) else 1
        if a.ndim >= 2 and a.shape[0] == self.mesh.n_cells and a.shape[1] == 3 and a.ndim == (3 if ncomp == 2 else 2):
            a = np.ascontiguousarray(a[: self.n_owned])
            self._ck(self.lib.tb_set_field_cell(self.ctx, field, _np_ptr(a), ncomp, self.stream))
            return
        if a.ndim == 0 or (a.ndim == 1 and a.shape[0] == ncomp and a.shape[0] != self.mesh.n_vertices):
            a = np.atleast_1d(a)
            self._ck(self.lib.tb_set_field_const(self.ctx, field, _np_ptr(a), int(a.shape[0])))
        else:
            if a.shape[0] != self.mesh.n_vertices:
                raise ValueError("vertex field has wrong length")
            self._ck(self.lib.tb_set_field_vertex(self.ctx, field, _np_ptr(a), ncomp))

    def sync_fields(self):
        """Apply pending coefficient changes on the current stream (needed in front of a CUDA-graph replay; plain
        stage launches do it themselves)."""
        if not self._fields_pending:
            return
        self._ck(self.lib.tb_sync_fields(self.ctx, self.stream))
        self._fields_pending = False

    def set_bc_bank(self, bank):
        """bank of the Function-valued SWE boundary data the next uploads fill / the next stage launches read"""
        if getattr(self, "_bc_bank", 0) != bank:
            self._ck(self.lib.tb_set_bc_bank(self.ctx, int(bank)))
            self._bc_bank = bank

    def clear_bc(self, eq, marker):
        self._ck(self.lib.tb_clear_bc(self.ctx, eq, int(marker)))

    def set_boundary_length(self, marker, length):
        self._ck(self.lib.tb_set_boundary_length(self.ctx, int(marker), float(length)))

    def set_bc(self, eq, marker, opcode, consts=None):
        """consts: {elev, uv_x, uv_y, un, flux, value[, diff_flux]}"""
        c = np.zeros(8, dtype=np.float64)
        if consts is not None:
            consts = np.asarray(consts, dtype=np.float64).reshape(-1)
            c[:consts.shape[0]] = consts
        self._ck(self.lib.tb_set_bc(self.ctx, eq, int(marker), int(opcode), _np_ptr(c)))

    def set_bc_array(self, eq, marker, tag, values):
        """values: host array over ALL exterior facets, (nb, 2) or (nb, 2, 2) for 'uv'."""
        a = np.ascontiguousarray(values, dtype=np.float64)
        ncomp = 2 if tag == L.BC_UV else 1
        self._ck(self.lib.tb_set_bc_array(self.ctx, eq, int(marker), int(tag), _np_ptr(a), ncomp, self.stream))

    def set_cell_quadrature(self, lam, w):
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        self._ck(self.lib.tb_set_cell_quadrature(self.ctx, int(w.shape[0]), _np_ptr(lam), _np_ptr(w)))

    def set_patch_range(self, first=0, count=-1):
        self._ck(self.lib.tb_set_patch_range(self.ctx, int(first), int(count)))

    def set_patch_list(self, lst=None):
        """lst: device int32 tensor of patch ids for the next SWE stage launches, or None for all patches."""
        if lst is None:
            self._ck(self.lib.tb_set_patch_list(self.ctx, None, 0))
        else:
            self._ck(self.lib.tb_set_patch_list(self.ctx, _ptr(lst), int(lst.numel())))

    # ------------------------------------------------------------ hot path
    def swe_stage(self, a0, a1, b_dt, u_in, u0, u_out):
        self._ck(self.lib.tb_swe_stage(self.ctx, a0, a1, b_dt, _ptr(u_in), _ptr(u0), _ptr(u_out), self.stream))

    def swe_stage_fused(self, a0, a1, b_dt, u_in, u0, u_out, push_dst):
        """stage kernel over all patches, boundary patches first, halo push from the epilogue (tb_swe_stage_fused)"""
        self._ck(self.lib.tb_swe_stage_fused(self.ctx, a0, a1, b_dt, _ptr(u_in), _ptr(u0), _ptr(u_out), _ptr(push_dst),
                                             self.stream))

    def halo_fused_setup(self, order, push_ptr, push_cell, recv_peers, remote_flags, flags_ptr):
        h = L.TbHaloFused()
        keep = [np.ascontiguousarray(order, np.int32), np.ascontiguousarray(push_ptr, np.int32),
                np.ascontiguousarray(push_cell if len(push_cell) else [0], np.int32)]
        h.n_bpatch = len(push_ptr) - 1
        h.patch_order, h.push_ptr, h.push_cell = (a.ctypes.data for a in keep)
        h.n_recv, h.n_send = len(recv_peers), len(remote_flags)
        for i, q in enumerate(recv_peers):
            h.recv_peer[i] = int(q)
        for i, a in enumerate(remote_flags):
            h.remote_flag[i] = int(a)
        h.flags = int(flags_ptr)
        self._ck(self.lib.tb_halo_fused_setup(self.ctx, C.byref(h)))

    def halo_fused_wait(self):
        self._ck(self.lib.tb_halo_fused_wait(self.ctx, self.stream))

    def halo_fused_status(self):
        ep, err = C.c_int64(), C.c_int32()
        self._ck(self.lib.tb_halo_fused_status(self.ctx, C.byref(ep), C.byref(err)))
        return int(ep.value), int(err.value)

    def swe_tendency(self, u, k_out):
        self._ck(self.lib.tb_swe_tendency(self.ctx, _ptr(u), _ptr(k_out), self.stream))

    def tracer_stage(self, a0, a1, b_dt, c_in, c0, c_out, swe_state):
        self._ck(self.lib.tb_tracer_stage(self.ctx, a0, a1, b_dt, _ptr(c_in), _ptr(c0), _ptr(c_out),
                                          _ptr(swe_state), self.stream))

    def tracer_stage_fused(self, a0, a1, b_dt, c_in, c0, c_out, swe_state, push_dst):
        self._ck(self.lib.tb_tracer_stage_fused(self.ctx, a0, a1, b_dt, _ptr(c_in), _ptr(c0), _ptr(c_out),
                                                _ptr(swe_state), _ptr(push_dst), self.stream))

    def limiter_apply_to_fused(self, c_in, c_out, push_dst):
        self._ck(self.lib.tb_limiter_apply_to_fused(self.ctx, _ptr(c_in), _ptr(c_out), _ptr(push_dst), self.stream))

    def limiter_apply(self, c):
        self._ck(self.lib.tb_limiter_apply(self.ctx, _ptr(c), self.stream))

    def limiter_apply_to(self, c_in, c_out):
        """out of place: owned cells of c_out = limited c_in (one patch-staged kernel, no copy back)"""
        self._ck(self.lib.tb_limiter_apply_to(self.ctx, _ptr(c_in), _ptr(c_out), self.stream))

    def swe_integrals(self, state, out):
        """out (device, 4 doubles): int eta^2, int |u|^2, int eta, int (eta + bathymetry)"""
        self._ck(self.lib.tb_swe_integrals(self.ctx, _ptr(state), _ptr(out), self.stream))

    def stage_integrals(self, enable):
        """while enabled, swe_stage launches also reduce the integrals of the state they write (per patch)"""
        self._ck(self.lib.tb_stage_integrals(self.ctx, int(bool(enable))))

    def stage_integrals_finish(self, out):
        """out (device, 4 doubles) = sum of the per-patch values of the last integrals-enabled stage launches"""
        self._ck(self.lib.tb_stage_integrals_finish(self.ctx, _ptr(out), self.stream))

    def tracer_integrals(self, c, swe_state, out):
        """out (device, 4 doubles): int c, int H c, min c, max c"""
        self._ck(self.lib.tb_tracer_integrals(self.ctx, _ptr(c), _ptr(swe_state), _ptr(out), self.stream))

    def lincomb(self, terms, out, length=None):
        """out = sum w*x over ``terms`` = [(w, tensor), ...] (at most 6) for the first ``length`` doubles (default:
        all of ``out``); out may alias an operand."""
        n = len(terms)
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for _, t in terms])
        ws = (C.c_double * n)(*[float(w) for w, _ in terms])
        ln = int(out.numel()) if length is None else int(length)
        self._ck(self.lib.tb_lincomb(self.ctx, n, ptrs, ws, _ptr(out), ln, self.stream))

    def gather_cells(self, state, idx, rec_len, buf):
        self._ck(self.lib.tb_gather_cells(self.ctx, _ptr(state), _ptr(idx), int(idx.numel()), rec_len, _ptr(buf),
                                          self.stream))

    def scatter_cells(self, buf, idx, rec_len, state):
        self._ck(self.lib.tb_scatter_cells(self.ctx, _ptr(buf), _ptr(idx), int(idx.numel()), rec_len, _ptr(state),
                                           self.stream))

    def push_cells(self, state, idx, dst_ptrs, rec_len):
        self._ck(self.lib.tb_push_cells(self.ctx, _ptr(state), _ptr(idx), _ptr(dst_ptrs), int(idx.numel()), rec_len,
                                        self.stream))

    # ------------------------------------------------------------ layout conversion
    def identity_node_map(self):
        if self._identity_map is None:
            self._identity_map = torch.arange(3 * self.n_owned, dtype=torch.int32, device=self.device)
        return self._identity_map

    def state_from_fields(self, uv_dev, eta_dev, node_map, state):
        self._ck(self.lib.tb_state_from_fields(self.ctx, _ptr(uv_dev), _ptr(eta_dev), _ptr(node_map), _ptr(state),
                                               self.stream))

    def state_to_fields(self, state, node_map, uv_dev, eta_dev):
        self._ck(self.lib.tb_state_to_fields(self.ctx, _ptr(state), _ptr(node_map), _ptr(uv_dev), _ptr(eta_dev),
                                             self.stream))

    def tracer_from_field(self, q_dev, node_map, c):
        self._ck(self.lib.tb_tracer_from_field(self.ctx, _ptr(q_dev), _ptr(node_map), _ptr(c), self.stream))

    def tracer_to_field(self, c, node_map, q_dev):
        self._ck(self.lib.tb_tracer_to_field(self.ctx, _ptr(c), _ptr(node_map), _ptr(q_dev), self.stream))

    # nodal-array convenience (tests): uv (n_owned,3,2), eta (n_owned,3) in this mesh's cell order
    def upload_nodal(self, uv, eta, state=None):
        state = self.new_state() if state is None else state
        uvd = torch.as_tensor(np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)).to(self.device)
        ed = torch.as_tensor(np.ascontiguousarray(eta, dtype=np.float64).reshape(-1)).to(self.device)
        self.state_from_fields(uvd, ed, self.identity_node_map(), state)
        return state

    def download_nodal(self, state):
        uvd = torch.empty((3 * self.n_owned, 2), dtype=torch.float64, device=self.device)
        ed = torch.empty(3 * self.n_owned, dtype=torch.float64, device=self.device)
        self.state_to_fields(state, self.identity_node_map(), uvd, ed)
        return uvd.cpu().numpy().reshape(self.n_owned, 3, 2), ed.cpu().numpy().reshape(self.n_owned, 3)

    def upload_tracer(self, q, c=None):
        c = self.new_tracer() if c is None else c
        qd = torch.as_tensor(np.ascontiguousarray(q, dtype=np.float64).reshape(-1)).to(self.device)
        self.tracer_from_field(qd, self.identity_node_map(), c)
        return c

    def download_tracer(self, c):
        qd = torch.empty(3 * self.n_owned, dtype=torch.float64, device=self.device)
        self.tracer_to_field(c, self.identity_node_map(), qd)
        return qd.cpu().numpy().reshape(self.n_owned, 3)


_ENGINES = weakref.WeakKeyDictionary()


def get_engine(mesh_key, build):
    """Engine cache keyed by the (Firedrake or shim) mesh object; ``build()`` makes a new `Engine`."""
    eng = _ENGINES.get(mesh_key)
    if eng is None:
        eng = build()
        _ENGINES[mesh_key] = eng
    return eng
