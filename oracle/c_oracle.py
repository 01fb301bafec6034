"""
CPU ORACLE (C port) loader -- TEST / BASELINE INFRASTRUCTURE ONLY.
Builds oracle/swe_oracle.c with gcc (-O3 -fopenmp) into oracle/_build/ and exposes it via ctypes.
Only tests/, __graft_entry__ and bench.py's cpu_baseline / --impl reference legs may use it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "swe_oracle.c")
LIB = os.path.join(HERE, "_build", "libswe_oracle.so")


def _host_signature():
    """Instruction-set flags of this host: the library is built with -march=native and travels with the repo to other
    machines (gpurun snapshot), where a different CPU must trigger a rebuild instead of an illegal instruction."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha1(" ".join(sorted(flags.split(":")[-1].split())).encode()).hexdigest()


def build(force=False):
    sig_file, sig = LIB + ".host", _host_signature()
    same_host = os.path.exists(sig_file) and open(sig_file).read().strip() == sig
    if not force and same_host and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    tmp = f"{LIB}.tmp{os.getpid()}"              # built aside and renamed: concurrent processes never load a partial file
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", tmp, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        # -march=native may be unavailable on exotic hosts
        cmd.remove("-march=native")
        r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
            import warnings
            warnings.warn("gcc could not rebuild the C oracle for this host; using the library built elsewhere:\n"
                          + r.stderr[-500:], RuntimeWarning)
            return LIB
        raise RuntimeError("gcc failed building the C oracle:\n" + r.stderr)
    os.replace(tmp, LIB)
    with open(sig_file, "w") as f:
        f.write(sig + "\n")
    return LIB


class _Problem(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("n_vertices", C.c_int64), ("n_bfacets", C.c_int64),
                ("coords", C.c_void_p), ("cells", C.c_void_p), ("nbr", C.c_void_p), ("nbr_lf", C.c_void_p),
                ("bf_opcode", C.c_void_p), ("bf_elev", C.c_void_p), ("bf_uv", C.c_void_p), ("bf_un", C.c_void_p),
                ("bf_flux", C.c_void_p), ("bf_len", C.c_void_p), ("bath", C.c_void_p), ("coriolis", C.c_void_p),
                ("manning", C.c_void_p), ("linear_drag", C.c_double), ("g", C.c_double), ("lf_sigma", C.c_double),
                ("eps2", C.c_double), ("nonlinear", C.c_int), ("lf_on", C.c_int), ("wd_on", C.c_int),
                ("wd_alpha", C.c_double), ("wind", C.c_void_p), ("rho0", C.c_double)]


class COracle:
    """
    :arg mesh: Mesh2D;  bath/coriolis/manning: per-vertex arrays (or scalars);  bnd: {marker: {tag: const}};
    bf_elev: optional (nb, 2) external elevation per exterior facet node (overrides the constant).
    State = cell records (nt, 9): [u0x u0y u1x u1y u2x u2y e0 e1 e2].
    """

    def __init__(self, mesh, bath, nonlinear=True, lf_on=True, g=9.81, coriolis=None, manning=None, linear_drag=0.0,
                 bnd=None, bf_elev=None, norm_smoother=0.0, lf_sigma=1.0, threads=None, wd_on=False, wd_alpha=0.5,
                 wind_stress=None, rho0=1000.0):
        self.lib = C.CDLL(build())
        self.lib.swe_oracle_threads.restype = C.c_int
        self.lib.swe_oracle_set_threads.argtypes = [C.c_int]
        if threads:
            # explicit: an inherited OMP_NUM_THREADS (torchrun sets 1) must not decide the baseline's core count
            self.lib.swe_oracle_set_threads(int(threads))
        nv, nt, nb = mesh.n_vertices, mesh.n_cells, mesh.n_bfacets
        full = lambda v: None if v is None else np.ascontiguousarray(np.broadcast_to(np.asarray(v, float), (nv,)))
        k = self._keep = dict(
            coords=np.ascontiguousarray(mesh.coords, np.float64), cells=np.ascontiguousarray(mesh.cells, np.int32),
            nbr=np.ascontiguousarray(mesh.nbr, np.int32), nbr_lf=np.ascontiguousarray(mesh.nbr_lf, np.int8),
            bath=full(bath), coriolis=full(coriolis), manning=full(manning),
            wind=None if wind_stress is None else np.ascontiguousarray(np.broadcast_to(np.asarray(wind_stress, float), (nv, 2))),
            bf_opcode=np.zeros(max(nb, 1), np.int32), bf_elev=np.zeros((max(nb, 1), 2)), bf_uv=np.zeros((max(nb, 1), 4)),
            bf_un=np.zeros((max(nb, 1), 2)), bf_flux=np.zeros((max(nb, 1), 2)), bf_len=np.ones(max(nb, 1)))
        tags = {"elev": 1, "uv": 2, "un": 4, "flux": 8}
        blen = mesh.boundary_length()
        for m, funcs in (bnd or {}).items():
            sel = mesh.bf_marker == m
            op = 0
            for t, v in funcs.items():
                op |= tags[t]
                if t == "elev": k["bf_elev"][sel] = v
                if t == "uv": k["bf_uv"][sel] = np.tile(np.asarray(v, float), 2)
                if t == "un": k["bf_un"][sel] = v
                if t == "flux": k["bf_flux"][sel] = v
            k["bf_opcode"][sel] = op
            k["bf_len"][sel] = blen[m]
        if bf_elev is not None:
            k["bf_elev"][:nb] = bf_elev
        p = self.p = _Problem()
        p.n_cells, p.n_vertices, p.n_bfacets = nt, nv, nb
        for name in ("coords", "cells", "nbr", "nbr_lf", "bf_opcode", "bf_elev", "bf_uv", "bf_un", "bf_flux", "bf_len",
                     "bath", "coriolis", "manning", "wind"):
            a = k[name]
            setattr(p, name, None if a is None else a.ctypes.data)
        p.linear_drag, p.g, p.lf_sigma, p.eps2 = float(linear_drag), float(g), float(lf_sigma), float(norm_smoother) ** 2
        p.nonlinear, p.lf_on = int(nonlinear), int(lf_on)
        p.wd_on, p.wd_alpha = int(wd_on), float(wd_alpha)
        p.rho0 = float(rho0)
        self.nt = nt

    def threads(self):
        return int(self.lib.swe_oracle_threads())

    def set_bf_elev(self, arr):
        self._keep["bf_elev"][:arr.shape[0]] = arr

    def stage(self, a0, a1, bdt, u, u0=None):
        out = np.empty_like(u)
        self.lib.swe_oracle_stage(C.byref(self.p), C.c_double(a0), C.c_double(a1), C.c_double(bdt),
                                  C.c_void_p(u.ctypes.data), C.c_void_p(u0.ctypes.data) if u0 is not None else None,
                                  C.c_void_p(out.ctypes.data))
        return out

    def tendency(self, u):
        return self.stage(0.0, 0.0, 1.0, u)

    def ssprk33(self, state, dt, nsteps):
        work = np.empty(2 * state.size)
        self.lib.swe_oracle_ssprk33(C.byref(self.p), C.c_double(dt), C.c_int(nsteps), C.c_void_p(state.ctypes.data),
                                    C.c_void_p(work.ctypes.data))
        return state


def host_threads():
    """Threads the timed CPU baseline may use: the cores this process is allowed to run on."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def records_from_nodal(uv, eta):
    return np.ascontiguousarray(np.concatenate([uv.reshape(uv.shape[0], 6), eta], axis=1))


def nodal_from_records(rec):
    return rec[:, :6].reshape(-1, 3, 2).copy(), rec[:, 6:].copy()
