"""
The drop-in on a mesh that Firedrake distributed itself, end to end on the CPU (2 gloo ranks): `SSPRK33` (SWE, then
tracer) and the vertex-based limiter are constructed on Firedrake-SHAPED objects of one MPI rank -- a mesh with owned
cells + a vertex overlap, clockwise cells, private vertex numbering, `cell_set.size / total_size`, `cell_node_map().values / values_with_halo`, a global
DG0 `lgmap`, `Function.dat.data(_ro) / data(_ro)_with_halos`, `mesh.comm` -- exactly as
`thetis.solver2d.FlowSolver2d.get_swe_timestepper` would construct it under `mpiexec -n 2`
(solver2d.py:542-573).  The adaptor must build the halo plan from the communicator, the integrator must read the
overlap rows of the host Functions, exchange ghosts every stage and write owned AND overlap rows back; the result
must equal the single-rank run on the global mesh.  Engine = the oracle-backed double (tests/oracle_engine.py).

What this cannot pin: that real Firedrake objects behave like the look-alike (tests/test_firedrake_lookalike_mesh.py).
"""
import os
import socket
import sys
import types
from datetime import timedelta

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

LX, LY, NSTEPS, DT = 8e3, 6e3, 3, 15.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _src_mesh():
    from thetis_b200.mesh import rectangle_mesh
    return rectangle_mesh(8, 6, LX, LY)


def _bath(x, y):
    return 12.0 + 2.0 * np.cos(2 * np.pi * x / LX) + 1e-4 * y


def _eta0(x, y):
    return 0.4 * np.cos(np.pi * x / LX) * np.cos(np.pi * y / LY)


class _Mixed:
    """`solution_2d`: a mixed Function with (uv_2d, elev_2d) sub-functions"""

    def __init__(self, space, uv, eta):
        self._fs, self.subfunctions = space, (uv, eta)

    def function_space(self):
        return self._fs


def _tide(x, y, t):
    return 0.2 + 0.1 * np.sin(2 * np.pi * t / 600.0) * (1.0 + y / LY)


def _stepper(mesh_obj, mixed_space, uv_f, eta_f, bath_f, tide_f):
    from thetis_b200 import rungekutta
    from thetis_b200.equations import ShallowWaterEquations, DepthExpression
    from thetis_b200.options import ModelOptions2d
    from thetis_b200.shim import Constant
    o = ModelOptions2d()
    eq = ShallowWaterEquations(mixed_space, DepthExpression(bath_f, True, False), o)
    # marker 1: a Function-valued tidal elevation re-assigned by update_forcings before every stage (the North-Sea
    # pattern, examples/north_sea/model_config.py:188-192) + a flux datum (divides by the GLOBAL boundary length)
    bnd = {1: {"elev": tide_f, "flux": Constant(-300.0)}, 2: {"elev": Constant(0.0), "uv": Constant((0.05, 0.0))}}
    fields = {"manning_drag_coefficient": Constant(0.02), "lax_friedrichs_velocity_scaling_factor": Constant(1.0)}
    return rungekutta.SSPRK33(eq, _Mixed(mixed_space, uv_f, eta_f), fields, DT, o.swe_timestepper_options, bnd,
                              sync_policy="every_step")


def _q0(x, y):
    return 4.5 + 2.0 * ((np.abs(x - LX / 2) < 2.5e3) & (np.abs(y - LY / 2) < 2e3))


def _tracer_stepper(space, uv_f, eta_f, q_f, bath_f):
    """what FlowSolver2d.get_tracer_timestepper builds (solver2d.py:576-598) + the limiter of solver2d.py:535-539"""
    from thetis_b200 import rungekutta
    from thetis_b200.equations import TracerEquation2D, DepthExpression
    from thetis_b200.limiter import VertexBasedP1DGLimiter
    from thetis_b200.options import ModelOptions2d
    from thetis_b200.shim import Constant
    o = ModelOptions2d()
    o.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d")
    eq = TracerEquation2D("tracer_2d", space, DepthExpression(bath_f, True, False), o, uv_f)
    fields = {"elev_2d": eta_f, "uv_2d": uv_f, "tracer_advective_velocity_factor": Constant(1.0)}
    ti = rungekutta.SSPRK33(eq, q_f, fields, DT, o.tracer_timestepper_options, {1: {"value": Constant(4.0)}},
                            sync_policy="every_step")
    return ti, VertexBasedP1DGLimiter(space)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=timedelta(seconds=300))
    try:
        import test_firedrake_lookalike_mesh as LK
        from test_distributed_host_path_gloo import _install_double
        from thetis_b200 import parallel as PA
        _install_double()
        fd = types.ModuleType("firedrake")
        fd.FunctionSpace = lambda mesh, family, degree: mesh.dg0_space() if (family, degree) == ("DG", 0) else mesh.p1_space()
        sys.modules["firedrake"] = fd
        # the GPU transport is not available here: plain all-to-all, no overlap stream (what 'auto' falls back to)
        orig = PA.plan_from_local_mesh
        PA.plan_from_local_mesh = lambda *a, **k: orig(*a, **{**k, "transport": "nccl", "overlap": False, "fused": False})
        src = _src_mesh()
        c = src.cell_centroids()
        owner = ((c[:, 0] > 0.45 * LX).astype(np.int32) + (c[:, 1] > 0.6 * LY)) % world
        fm = LK._DistributedLookalike(src, owner, rank, world, seed=4)
        fm.comm.allgather = PA.torch_allgather            # mpi4py's comm.allgather, here over the gloo group
        n_local, n_owned = fm.cell_set.total_size, fm.n_owned
        nodes = np.arange(3 * n_local, dtype=np.int64).reshape(n_local, 3)
        xy = src.coords[fm.local_cells_global_vertices].reshape(-1, 2)          # P1DG node coordinates, local order
        space = LK._HaloSpace(fm, "Discontinuous Lagrange", nodes, n_owned)
        mixed = types.SimpleNamespace(mesh=lambda: fm)

        def func(data):
            f = LK._Function(space, None)
            f.dat = LK._HaloDat(np.ascontiguousarray(data, dtype=float), 3 * n_owned)
            return f
        uv_f, eta_f = func(np.zeros((3 * n_local, 2))), func(_eta0(xy[:, 0], xy[:, 1]))
        bath_f = func(_bath(xy[:, 0], xy[:, 1]))
        # the overlap rows of the host solution start out WRONG (a stale PyOP2 halo): the owners' values must win
        eta_f.dat.data_with_halos[3 * n_owned:] = 77.0
        tide_f = func(_tide(xy[:, 0], xy[:, 1], 0.0))

        def update_forcings(t):
            tide_f.dat.data_with_halos[...] = _tide(xy[:, 0], xy[:, 1], t)
            tide_f.dat.dat_version += 1                       # PyOP2 bumps the version on write access
        ti = _stepper(fm, mixed, uv_f, eta_f, bath_f, tide_f)
        assert ti.adaptor.with_halos and ti.halo is not None and ti.halo.world == world and ti.adaptor.n_owned == n_owned
        q_f = func(_q0(xy[:, 0], xy[:, 1]))
        tr, lim = _tracer_stepper(space, uv_f, eta_f, q_f, bath_f)
        assert tr.adaptor is ti.adaptor and lim.halo is ti.halo            # one adaptor / plan / engine per mesh
        t = 0.0
        for _ in range(NSTEPS):
            ti.advance(t, update_forcings)                                  # coupled_timeintegrator_2d.py:94-105
            tr.advance(t, update_forcings)
            lim.apply(q_f)
            t += DT
        from thetis_b200 import _lib as L
        assert (0, 1, L.BC_ELEV) in ti.engine.bc_arrays      # the tidal Function went out as a boundary array
        out[rank] = (fm.gids.copy(), n_owned, fm.local_cells_global_vertices.copy(),
                     uv_f.dat.data_ro_with_halos.reshape(n_local, 3, 2).copy(),
                     eta_f.dat.data_ro_with_halos.reshape(n_local, 3).copy(),
                     int(ti.engine.n_gathers), q_f.dat.data_ro_with_halos.reshape(n_local, 3).copy())
    finally:
        dist.destroy_process_group()


def _single_rank(path):
    from test_distributed_host_path_gloo import _install_double
    from thetis_b200.shim import Function, FunctionSpace, as_shim_mesh
    _install_double()
    sm = as_shim_mesh(_src_mesh())
    dg = FunctionSpace(sm, "DG", 1)
    uv_f = Function(FunctionSpace(sm, "DG", 1, value_size=2))
    eta_f = Function(dg).interpolate(_eta0)
    bath_f = Function(dg).interpolate(_bath)
    tide_f = Function(dg).interpolate(lambda x, y: _tide(x, y, 0.0))
    ti = _stepper(sm, types.SimpleNamespace(mesh=lambda: sm), uv_f, eta_f, bath_f, tide_f)
    q_f = Function(dg).interpolate(_q0)
    tr, lim = _tracer_stepper(dg, uv_f, eta_f, q_f, bath_f)
    uf = lambda tt: tide_f.interpolate(lambda x, y: _tide(x, y, tt))
    t = 0.0
    for _ in range(NSTEPS):
        ti.advance(t, uf)
        tr.advance(t, uf)
        lim.apply(q_f)
        t += DT
    np.savez(path, uv=np.array(uv_f.dat.data_ro).reshape(-1, 3, 2), eta=np.array(eta_f.dat.data_ro).reshape(-1, 3),
             q=np.array(q_f.dat.data_ro).reshape(-1, 3))


def test_ssprk33_on_a_distributed_firedrake_shaped_mesh():
    import subprocess
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    ref = os.path.join(HERE, "_dist_ref_fd_%d.npz" % os.getpid())
    try:
        subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); "
                        "import test_distributed_firedrake_lookalike_gloo as T; T._single_rank(%r)" % (HERE, ref)],
                       check=True, cwd=os.path.dirname(HERE), timeout=600)
        g = np.load(ref)
        uv1, eta1, q1 = g["uv"], g["eta"], g["q"]
    finally:
        if os.path.exists(ref):
            os.remove(ref)
    src = _src_mesh()
    assert np.abs(uv1).max() > 1e-3 and np.abs(q1 - 4.5).max() > 0.5
    owned_all = []
    for r in range(world):
        gids, n_owned, lcgv, uv, eta, n_gathers, q = out[r]
        assert n_gathers >= 7 * NSTEPS + 2                     # uploads, every SWE / tracer stage and the limiter exchanged
        owned_all.append(gids[:n_owned])
        # local (cell, node a) sits at global vertex lcgv[c, a] = node j of the global cell gids[c]
        j = np.argmax(src.cells[gids][:, None, :] == lcgv[:, :, None], axis=2)
        idx = gids[:, None]
        for name, a, b in (("uv", uv, uv1[idx, j]), ("eta", eta, eta1[idx, j]), ("tracer", q, q1[idx, j])):
            for what, sl in (("owned", slice(0, n_owned)), ("overlap", slice(n_owned, None))):
                err = np.abs(a[sl] - b[sl]).max() / np.abs(b).max()
                assert err < 1e-12, (r, name, what, err)         # overlap rows of the host Functions are current too
    assert np.array_equal(np.sort(np.concatenate(owned_all)), np.arange(src.n_cells))
