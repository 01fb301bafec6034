"""
ctypes binding of the C-ABI declared in include/thetis_b200.h.

The product path has no CPU fallback: if the shared library is missing or a
call fails this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("THETIS_B200_LIB") or os.path.join(_HERE, "libthetis_b200.so")   # override: developer A/B builds

TB_OK = 0
# tb_option
OPT_G_GRAV, OPT_RHO0, OPT_NONLINEAR, OPT_LAX_FRIEDRICHS, OPT_LF_SCALING, OPT_NORM_SMOOTHER, \
    OPT_WETTING_DRYING, OPT_WD_ALPHA, OPT_LF_TRACER, OPT_LF_TRACER_SCALING, OPT_TRACER_VEL_FACTOR, \
    OPT_FORCE_GENERIC_KERNEL, OPT_SIPG_FACTOR, OPT_SIPG_FACTOR_TRACER, OPT_GRAD_DIV_VISCOSITY, \
    OPT_GRAD_DEPTH_VISCOSITY, OPT_TRACER_CONSERVATIVE, OPT_MOMENTUM_ADVECTION, OPT_VON_KARMAN, \
    OPT_WD_DISPLACED_MASS = range(20)
# tb_field
F_BATHYMETRY, F_CORIOLIS, F_MANNING, F_QUAD_DRAG, F_LINEAR_DRAG, F_WIND_STRESS, F_ATM_PRESSURE, \
    F_MOMENTUM_SOURCE, F_VOLUME_SOURCE, F_TRACER_SOURCE, F_VISCOSITY, F_DIFFUSIVITY, F_NIKURADSE, F_WD_ALPHA = range(14)
BC_ELEV, BC_UV, BC_UN, BC_FLUX, BC_VALUE, BC_DIFF_FLUX, BC_DRAG = 1, 2, 4, 8, 16, 64, 128


class TbMesh(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int64), ("n_owned", C.c_int64), ("n_vertices", C.c_int64), ("n_bfacets", C.c_int64),
        ("coords", C.c_void_p), ("cells", C.c_void_p), ("nbr", C.c_void_p), ("nbr_lf", C.c_void_p),
        ("bf_marker", C.c_void_p), ("topo", C.c_void_p),
    ]


class TbHaloFused(C.Structure):
    _fields_ = [
        ("n_bpatch", C.c_int64), ("patch_order", C.c_void_p), ("push_ptr", C.c_void_p), ("push_cell", C.c_void_p),
        ("n_recv", C.c_int32), ("n_send", C.c_int32), ("recv_peer", C.c_int32 * 16), ("remote_flag", C.c_uint64 * 16),
        ("flags", C.c_uint64),
    ]


# every symbol include/thetis_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.c_double
_I = C.c_int
_L = C.c_int64
SIGNATURES = {
    "tb_create": (_I, [C.POINTER(_P), C.POINTER(TbMesh), _I]),
    "tb_destroy": (_I, [_P]),
    "tb_last_error": (C.c_char_p, [_P]),
    "tb_version": (_I, []),
    "tb_state_len": (_L, [_P]),
    "tb_tracer_len": (_L, [_P]),
    "tb_patch_size": (_L, [_P]),
    "tb_n_patches": (_L, [_P]),
    "tb_set_option": (_I, [_P, _I, _D]),
    "tb_set_field_const": (_I, [_P, _I, _P, _I]),
    "tb_set_field_vertex": (_I, [_P, _I, _P, _I]),
    "tb_set_field_cell": (_I, [_P, _I, _P, _I, _P]),
    "tb_clear_field": (_I, [_P, _I]),
    "tb_sync_fields": (_I, [_P, _P]),
    "tb_clear_bc": (_I, [_P, _I, _I]),
    "tb_set_bc_bank": (_I, [_P, _I]),
    "tb_set_bc": (_I, [_P, _I, _I, _I, _P]),
    "tb_set_bc_array": (_I, [_P, _I, _I, _I, _P, _I, _P]),
    "tb_set_boundary_length": (_I, [_P, _I, _D]),
    "tb_set_cell_quadrature": (_I, [_P, _I, _P, _P]),
    "tb_swe_stage": (_I, [_P, _D, _D, _D, _P, _P, _P, _P]),
    "tb_swe_tendency": (_I, [_P, _P, _P, _P]),
    "tb_swe_stage_fused": (_I, [_P, _D, _D, _D, _P, _P, _P, _P, _P]),
    "tb_halo_fused_setup": (_I, [_P, C.POINTER(TbHaloFused)]),
    "tb_halo_fused_wait": (_I, [_P, _P]),
    "tb_halo_fused_status": (_I, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "tb_tracer_stage": (_I, [_P, _D, _D, _D, _P, _P, _P, _P, _P]),
    "tb_limiter_apply": (_I, [_P, _P, _P]),
    "tb_limiter_apply_to": (_I, [_P, _P, _P, _P]),
    "tb_limiter_apply_to_fused": (_I, [_P, _P, _P, _P, _P]),
    "tb_tracer_stage_fused": (_I, [_P, _D, _D, _D, _P, _P, _P, _P, _P, _P]),
    "tb_state_from_fields": (_I, [_P, _P, _P, _P, _P, _P]),
    "tb_state_to_fields": (_I, [_P, _P, _P, _P, _P, _P]),
    "tb_tracer_from_field": (_I, [_P, _P, _P, _P, _P]),
    "tb_tracer_to_field": (_I, [_P, _P, _P, _P, _P]),
    "tb_swe_integrals": (_I, [_P, _P, _P, _P]),
    "tb_tracer_integrals": (_I, [_P, _P, _P, _P, _P]),
    "tb_lincomb": (_I, [_P, _I, _P, _P, _P, _L, _P]),
    "tb_stage_integrals": (_I, [_P, _I]),
    "tb_stage_integrals_finish": (_I, [_P, _P, _P]),
    "tb_gather_cells": (_I, [_P, _P, _P, _L, _I, _P, _P]),
    "tb_scatter_cells": (_I, [_P, _P, _P, _L, _I, _P, _P]),
    "tb_push_cells": (_I, [_P, _P, _P, _P, _L, _I, _P]),
    "tb_set_patch_range": (_I, [_P, _L, _L]),
    "tb_set_patch_list": (_I, [_P, _P, _L]),
    "tb_launch_count": (_L, [_P]),
    "tb_selftest_math": (_I, [_P, _P, _P, _L, _P]),
}

_lib = None


def load():
    """Load libthetis_b200.so (built in-tree by thetis_b200.build); raise if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library must be built (python -m thetis_b200.build); "
            "there is no CPU fallback on this path")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class TbError(RuntimeError):
    pass


def check(ctx, rc):
    if rc != TB_OK:
        lib = load()
        msg = lib.tb_last_error(ctx)
        raise TbError(f"thetis_b200 error {rc}: {msg.decode() if msg else ''}")
