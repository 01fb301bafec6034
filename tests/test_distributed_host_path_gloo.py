"""
The DISTRIBUTED code paths of the host classes end to end on the CPU: `SSPRK33` (SWE and tracer), the limiter, the
coupled integrator and `HaloPlan` (buffer allocation with the ghost block behind the owned cells, send lists,
per-stage exchange over torch.distributed) run on 2 and 3 gloo ranks against the oracle-backed engine double
(tests/oracle_engine.py: DistributedOracleEngine poisons the ghost block of every stage output like the real kernels
leave it stale, so one missing or misdirected exchange reaches an owned cell as NaN in the next stage), and must
reproduce the single-rank run of the same classes.

Two ways to the halo plan: `distribute_mesh` (the library's own partitioner, the route of every multi-GPU
measurement) and `plan_from_local_mesh` (a mesh that arrives distributed: scattered ownership, scrambled local
numbering, one all-gather).  Reference analogue: the same asserts under 2 MPI ranks,
test/swe2d/test_steady_state_channel.py:6.  The kernels are tied to the same oracle by the `-m gpu` tests.
"""
import os
import socket
import sys
from datetime import timedelta

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

LX, LY, NSTEPS, DT = 18e3, 8e3, 4, 20.0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mesh():
    from thetis_b200.mesh import rectangle_mesh, sfc_renumber
    return sfc_renumber(rectangle_mesh(12, 6, LX, LY))


def _install_double():
    """Route `MeshAdaptor.get_engine` to the oracle-backed double (no CUDA in this process)."""
    from thetis_b200 import adaptor
    from oracle_engine import OracleEngine, DistributedOracleEngine
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.cuda.current_stream = lambda *a, **k: type("S", (), {"synchronize": lambda s: None})()

    def get_engine(self):
        if self.engine is None:
            if self.halo is not None:
                self.engine = DistributedOracleEngine(self.mesh, self.n_owned, self.boundary_len)
                self.halo.attach(self.engine)
            else:
                self.engine = OracleEngine(self.mesh)
        return self.engine
    adaptor.MeshAdaptor.get_engine = get_engine


def _solver(mesh_obj, scheme="ssprk33"):
    """SWE (nonlinear, Lax-Friedrichs, Manning drag, open boundary with elev + flux data: the flux datum divides by
    the GLOBAL boundary length) -> tracer (inflow value) -> vertex-based limiter."""
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, Constant, as_shim_mesh, ShimMesh
    sm = mesh_obj if isinstance(mesh_obj, ShimMesh) else as_shim_mesh(mesh_obj)
    b = Function(FunctionSpace(sm, "CG", 1)).interpolate(lambda x, y: 10.0 + 2.0 * np.cos(2 * np.pi * x / LX))
    s = solver2d.FlowSolver2d(sm, b)
    o = s.options
    o.swe_timestepper_options.use_automatic_timestep = False
    o.tracer_timestepper_options.use_automatic_timestep = False
    o.timestep = DT
    o.simulation_end_time = DT * NSTEPS
    o.simulation_export_time = DT * NSTEPS
    o.manning_drag_coefficient = Constant(0.02)
    kw = {}
    if scheme == "erk_viscous":
        # Butcher-form ERK (tendency buffers + tb_lincomb: ghost blocks are read outside the Shu-Osher stage order),
        # SIPG viscosity as a P1 field and tracer diffusion (neighbour gradients across the partition cut)
        o.swe_timestepper_type = o.tracer_timestepper_type = "ERKLSPUM2"
        o.horizontal_viscosity = Function(FunctionSpace(sm, "CG", 1)).interpolate(
            lambda x, y: 20.0 * (1.0 + 0.3 * np.sin(y / 2e3)))
        o.use_grad_div_viscosity_term = True
        kw["diffusivity"] = Constant(12.0)
    o.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d", **kw)
    o.use_limiter_for_tracers = True
    s.bnd_functions["shallow_water"] = {1: {"elev": Constant(0.2), "flux": Constant(-400.0)},
                                        2: {"elev": Constant(0.0), "uv": Constant((0.05, 0.0))}}
    s.bnd_functions["tracer"] = {1: {"value": Constant(4.0)}}
    s.assign_initial_conditions(elev=lambda x, y: 0.5 * np.cos(np.pi * x / LX),
                                tracer=lambda x, y: 4.5 + 2.0 * ((np.abs(x - LX / 2) < 3e3) & (np.abs(y - LY / 2) < 2e3)))
    return s


def _run(s):
    t = 0.0
    for _ in range(NSTEPS):
        s.timestepper.advance(t)
        t += DT
    s.timestepper.sync_to_host()
    f = s.fields
    return (np.array(f.uv_2d.dat.data_ro).reshape(-1, 3, 2), np.array(f.elev_2d.dat.data_ro).reshape(-1, 3),
            np.array(f.tracer_2d.dat.data_ro).reshape(-1, 3))


def _local_view(mesh, owner, rank, seed):
    """Rank-local view with a vertex overlap, scrambled cell / vertex numbering, no neighbour table."""
    from thetis_b200.mesh import Mesh2D, FACET_NODES
    from thetis_b200.parallel import build_overlap_connectivity
    rng = np.random.default_rng(100 * seed + rank)
    owned = np.nonzero(owner == rank)[0]
    ptr, idx = mesh.vertex_to_cell_csr()
    tv = np.unique(mesh.topo[mesh.cells[owned]])
    cand = np.unique(np.concatenate([idx[ptr[v]:ptr[v + 1]] for v in tv]))
    ghost = cand[owner[cand] != rank]
    gids = np.concatenate([rng.permutation(owned), rng.permutation(ghost)]).astype(np.int64)
    cells_g = mesh.cells[gids]
    vused = rng.permutation(np.unique(cells_g))
    vloc = np.full(mesh.n_vertices, -1, dtype=np.int64)
    vloc[vused] = np.arange(vused.shape[0])
    lm = Mesh2D(coords=mesh.coords[vused], cells=vloc[cells_g].astype(np.int32),
                topo=np.unique(mesh.topo[vused], return_inverse=True)[1].astype(np.int32))
    ext = {}
    cl, fl = np.nonzero(mesh.nbr[gids] < 0)
    for c_loc, f in zip(cl, fl):
        a, b = lm.topo[lm.cells[c_loc, FACET_NODES[f]]]
        ext[(int(a), int(b))] = int(mesh.bf_marker[-(mesh.nbr[gids[c_loc], f] + 1)])
    build_overlap_connectivity(lm, ext)
    return lm, owned.shape[0], gids


SCHEMES = ("ssprk33", "erk_viscous")


def _worker(rank, world, port, route, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=timedelta(seconds=300))
    try:
        _install_double()
        from thetis_b200 import parallel as PA
        from thetis_b200.shim import ShimMesh
        mesh = _mesh()
        res = {}
        for scheme in SCHEMES:                  # a fresh mesh object (adaptor, engine, halo plan) per run
            if route == "partition":
                sm = PA.distribute_mesh(mesh, rank, world, halo="vertex", transport="nccl", overlap=False, fused=False)
            else:
                c = mesh.cell_centroids()
                rng = np.random.default_rng(5)
                pts = c[rng.choice(mesh.n_cells, world, replace=False)]
                owner = np.argmin(((c[:, None, :] - pts[None]) ** 2).sum(-1), axis=1).astype(np.int32)
                lm, n_owned, gids = _local_view(mesh, owner, rank, seed=3)
                plan, part = PA.plan_from_local_mesh(lm, n_owned, gids, halo="vertex", transport="nccl",
                                                     overlap=False, fused=False)
                sm = ShimMesh(part.mesh)
                sm.boundary_len = dict(part.mesh.meta["global_boundary_len"])
                sm.halo_plan = plan
            s = _solver(sm, scheme)
            uv, eta, q = _run(s)
            plan = sm.halo_plan
            eng = s.timestepper.timesteppers["swe2d"].engine
            n = plan.part.n_owned
            assert plan.transport == "nccl" and not plan.fused and not plan.overlap
            res[scheme] = (sm.topology_mesh.meta["global_cells"][:n].copy(), uv[:n], eta[:n], q[:n],
                           eng.n_stage_launches, eng.n_gathers, int(plan.part.n_ghost))
        out[rank] = res
    finally:
        dist.destroy_process_group()


_SINGLE = {}


def _single_rank(scheme):
    """The same classes on the same double, one rank -- in a process of its own (the double is patched in globally)."""
    import subprocess
    if scheme in _SINGLE:
        return _SINGLE[scheme]
    ref = os.path.join(HERE, "_dist_ref_%d.npz" % os.getpid())
    code = ("import sys; sys.path.insert(0, %r); import numpy as np; import test_distributed_host_path_gloo as T; "
            "T._install_double(); uv, eta, q = T._run(T._solver(T._mesh(), %r)); np.savez(%r, uv=uv, eta=eta, q=q)"
            % (HERE, scheme, ref))
    try:
        subprocess.run([sys.executable, "-c", code], check=True, cwd=os.path.dirname(HERE), timeout=600)
        g = np.load(ref)
        _SINGLE[scheme] = (g["uv"], g["eta"], g["q"])
        return _SINGLE[scheme]
    finally:
        if os.path.exists(ref):
            os.remove(ref)


@pytest.mark.parametrize("route,world", [("partition", 3), ("local", 2)])
def test_distributed_host_classes_reproduce_the_single_rank_run(route, world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), route, out), nprocs=world, join=True)
    assert len(out) == world
    for scheme in SCHEMES:
        uv1, eta1, q1 = _single_rank(scheme)
        assert np.abs(q1 - 4.5).max() > 0.5 and np.abs(uv1).max() > 1e-3          # something happened
        seen = []
        for r in range(world):
            cells, uv, eta, q, n_stage, n_gather, n_ghost = out[r][scheme]
            seen.append(cells)
            assert n_ghost > 0 and n_gather >= 5 * NSTEPS      # every SWE / tracer stage and the limiter is exchanged
            for name, a, b in (("uv", uv, uv1[cells]), ("eta", eta, eta1[cells]), ("tracer", q, q1[cells])):
                err = np.abs(a - b).max() / np.abs(b).max()
                assert err < 1e-12, (route, scheme, r, name, err)                  # NaN (a stale ghost was read) fails too
        assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(uv1.shape[0]))
