"""
Multi-GPU parity: the domain-decomposed run (one process per GPU; fused compute + halo-push launches, or the unfused /
NCCL transports) must reproduce the single-GPU run BIT FOR BIT -- every cell is evaluated from its own side by the
same kernel, so ownership cannot change a result.  The NVLink variants need >= 2 GPUs (skipped otherwise); the
"host" / one_gpu variants put both ranks on ONE GPU with the halo staged through the host over gloo, so a single-GPU
test run still exercises partitioning, ghost cells, the vertex halo of the limiter and the all-reduced diagnostics.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port, one_gpu):
    """one process per GPU over NCCL -- or, with ``one_gpu``, every rank on cuda:0 over gloo (host-staged halo)"""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if one_gpu:
        torch.cuda.set_device(0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))


def _worker(rank, world, port, k, nsteps, transport, out):
    import torch
    import torch.distributed as dist
    _init(rank, world, port, transport == "host")
    if transport == "host":
        transport = "nccl"              # pack + exchange_halo + scatter; gloo stages it through the host
    try:
        from harness.workloads import north_sea_mesh, north_sea_setup
        from harness.runs import PartitionedSWE
        mesh = north_sea_mesh(k)
        setup = north_sea_setup(mesh, wetting_drying=True)
        fused = transport == "symm"
        tr = "symm" if transport.startswith("symm") else transport
        run = PartitionedSWE(mesh, setup, rank, world, wd=True, transport=tr, fused=fused)
        assert run.transport == tr
        assert run.overlap
        assert run.plan.fused == fused
        for _ in range(nsteps - 2):
            run.step_e2e()
        # the last two steps: one plain resident step, one replay of its CUDA graph (forcing frozen on both sides)
        if tr == "symm":
            run.enable_graph()            # runs one warm-up step, then captures
        else:
            run.step_resident()
        run.step_resident()
        torch.cuda.synchronize()
        uv, eta = run.owned_nodal()
        if fused:
            # 3 fused launches per step (the capture itself launches nothing), no flag wait timed out
            epoch, err = run.eng.halo_fused_status()
            assert err == 0 and epoch == 3 * nsteps, (epoch, err)
        out[rank] = (run.part.owned_global.copy(), uv, eta)
    finally:
        dist.destroy_process_group()


# "symm": fused compute + halo push (tb_swe_stage_fused, per-peer epoch flags); "symm-unfused": boundary launch +
# push kernel + cross-rank barrier; "nccl": pack + all-to-all
# "host": both ranks on ONE GPU, halo staged through the host over gloo -- runs on a single-GPU box, so the
# partition / ghost-cell / boundary-patch logic and the ownership-independence of the kernels are always exercised
@pytest.mark.parametrize("transport", ["nccl", "symm", "symm-unfused", "host"])
@pytest.mark.parametrize("world", [2])
def test_partitioned_run_is_bit_identical(world, transport):
    import torch
    import torch.multiprocessing as mp
    if transport != "host" and torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    k, nsteps = 2, 4
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), k, nsteps, transport, out), nprocs=world, join=True)
    from harness.workloads import north_sea_mesh, north_sea_setup
    from harness.runs import SingleSWE
    mesh = north_sea_mesh(k)
    setup = north_sea_setup(mesh, wetting_drying=True)
    single = SingleSWE(mesh, setup, wd=True)
    for _ in range(nsteps - 2):
        single.step_e2e()
    single.step_resident()
    single.step_resident()
    uv1, eta1 = single.state_nodal()
    for r in range(world):
        owned, uv, eta = out[r]
        du = np.abs(uv - uv1[owned]).max()
        de = np.abs(eta - eta1[owned]).max()
        assert du == 0.0 and de == 0.0, (r, du, de, np.abs(uv1).max(), int((np.abs(uv - uv1[owned]).max(axis=(1, 2)) > 0).sum()))
    assert np.isfinite(uv1).all() and np.abs(eta1).max() > 0


def _worker_coupled(rank, world, port, out, one_gpu=False):
    import torch
    import torch.distributed as dist
    _init(rank, world, port, one_gpu)
    try:
        from thetis_b200.parallel import distribute_mesh
        mesh = _coupled_mesh()
        sm = distribute_mesh(mesh, rank, world, halo="vertex")
        lm = sm.topology_mesh
        s = _coupled_solver(sm)
        s.iterate()
        n = sm.halo_plan.part.n_owned
        out[rank] = (lm.meta["global_cells"][:n].copy(),
                     s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2)[:n].copy(),
                     s.fields.elev_2d.dat.data_ro.reshape(-1, 3)[:n].copy(),
                     s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)[:n].copy(), s.last_norms)
    finally:
        dist.destroy_process_group()


def _coupled_mesh():
    from thetis_b200.mesh import rectangle_mesh, sfc_renumber
    return sfc_renumber(rectangle_mesh(36, 16, 18e3, 8e3))


def _coupled_solver(mesh_obj):
    """BASELINE config 4 style: SWE + tracer + limiter (same set-up on 1 GPU and on the distributed mesh)"""
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, Constant, as_shim_mesh, ShimMesh
    sm = mesh_obj if isinstance(mesh_obj, ShimMesh) else as_shim_mesh(mesh_obj)
    lx = 18e3
    b = Function(FunctionSpace(sm, "CG", 1)).interpolate(lambda x, y: 10.0 + 2.0 * np.cos(2 * np.pi * x / lx))
    s = solver2d.FlowSolver2d(sm, b)
    o = s.options
    o.swe_timestepper_options.use_automatic_timestep = False
    o.tracer_timestepper_options.use_automatic_timestep = False
    o.timestep = 2.5
    o.simulation_end_time = 2.5 * 30
    o.simulation_export_time = 2.5 * 10
    o.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d")
    o.use_limiter_for_tracers = True
    s.bnd_functions["shallow_water"] = {1: {"elev": Constant(0.2), "uv": Constant((0.05, 0.0))}}
    s.bnd_functions["tracer"] = {1: {"value": Constant(4.0)}}
    s.assign_initial_conditions(elev=lambda x, y: 0.5 * np.cos(np.pi * x / lx),
                                tracer=lambda x, y: 4.5 + 2.0 * ((np.abs(x - lx / 2) < 3e3) & (np.abs(y - 4e3) < 2e3)))
    return s


@pytest.mark.parametrize("one_gpu", [False, True])
def test_coupled_tracer_limiter_distributed_is_bit_identical(one_gpu):
    """SWE -> tracer -> limiter on 2 ranks (vertex halo for the limiter bounds) == the 1-GPU run, bit for bit; on 2 GPUs
    with the fused compute + halo-push launches, or both ranks on one GPU with the halo staged through the host"""
    import torch
    import torch.multiprocessing as mp
    world = 2
    if not one_gpu and torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_coupled, args=(world, _free_port(), out, one_gpu), nprocs=world, join=True)
    s = _coupled_solver(_coupled_mesh())
    s.iterate()
    uv1 = s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2)
    e1 = s.fields.elev_2d.dat.data_ro.reshape(-1, 3)
    c1 = s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)
    assert np.abs(c1 - 4.5).max() > 0.5 and c1.max() <= 6.5 + 1e-9 and c1.min() >= 4.0 - 1e-9
    for r in range(world):
        cells, uv, e, c, norms = out[r]
        assert np.array_equal(uv, uv1[cells]) and np.array_equal(e, e1[cells]) and np.array_equal(c, c1[cells])
        assert abs(norms[0] - s.last_norms[0]) <= 1e-12 * s.last_norms[0]      # all-reduced print_state norms


# ---------------------------------------------------------------- SURVEY 8f rows on a distributed mesh
def _sipg_mesh():
    from thetis_b200.mesh import delaunay_mesh, sfc_renumber
    return sfc_renumber(delaunay_mesh(1500, 18e3, 8e3, seed=11))


def _sipg_solver(mesh_obj):
    """viscosity + tracer diffusion + Butcher-form ERK + device callbacks: same set-up on 1 GPU and distributed"""
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, Constant, as_shim_mesh, ShimMesh
    sm = mesh_obj if isinstance(mesh_obj, ShimMesh) else as_shim_mesh(mesh_obj)
    lx = 18e3
    P1 = FunctionSpace(sm, "CG", 1)
    b = Function(P1).interpolate(lambda x, y: 10.0 + 2.0 * np.cos(2 * np.pi * x / lx))
    s = solver2d.FlowSolver2d(sm, b)
    o = s.options
    o.swe_timestepper_type = "ERKLSPUM2"
    o.tracer_timestepper_type = "ERKLSPUM2"
    o.swe_timestepper_options.use_automatic_timestep = False
    o.tracer_timestepper_options.use_automatic_timestep = False
    o.timestep = 1.0
    o.simulation_end_time = 1.0 * 24
    o.simulation_export_time = 1.0 * 8
    o.horizontal_viscosity = Function(P1).interpolate(lambda x, y: 20.0 * (1.0 + 0.3 * np.sin(y / 2e3)))
    o.use_grad_div_viscosity_term = True
    o.check_volume_conservation_2d = True
    o.check_tracer_conservation = True
    o.check_tracer_overshoot = True
    o.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d", diffusivity=Constant(12.0))
    o.use_limiter_for_tracers = True
    s.bnd_functions["shallow_water"] = {1: {"elev": Constant(0.2), "uv": Constant((0.05, 0.0))}}
    s.bnd_functions["tracer"] = {1: {"value": Constant(4.0)}}
    s.assign_initial_conditions(elev=lambda x, y: 0.5 * np.cos(np.pi * x / lx),
                                tracer=lambda x, y: 4.5 + 2.0 * np.exp(-((x - lx / 2) ** 2 + (y - 4e3) ** 2) / 2e3 ** 2))
    return s


def _worker_sipg(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from thetis_b200.parallel import distribute_mesh
        sm = distribute_mesh(_sipg_mesh(), rank, world, halo="vertex")
        lm = sm.topology_mesh
        s = _sipg_solver(sm)
        s.iterate()
        n = sm.halo_plan.part.n_owned
        hist = {cb.name: [v for _, v in cb.history] for cb in s.callbacks["export"]}
        out[rank] = (lm.meta["global_cells"][:n].copy(),
                     s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2)[:n].copy(),
                     s.fields.elev_2d.dat.data_ro.reshape(-1, 3)[:n].copy(),
                     s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)[:n].copy(), hist)
    finally:
        dist.destroy_process_group()


def test_sipg_erk_callbacks_distributed_is_bit_identical():
    """viscosity (neighbour gradients from ghost cells), tracer diffusion, ERKLSPUM2 (tb_lincomb over owned + ghost
    records) and the all-reduced device callbacks on 2 GPUs == the 1-GPU run"""
    import torch
    import torch.multiprocessing as mp
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_sipg, args=(world, _free_port(), out), nprocs=world, join=True)
    s = _sipg_solver(_sipg_mesh())
    s.iterate()
    uv1 = s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2)
    e1 = s.fields.elev_2d.dat.data_ro.reshape(-1, 3)
    c1 = s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)
    assert np.isfinite(uv1).all() and np.abs(uv1).max() > 1e-3
    hist1 = {cb.name: [v for _, v in cb.history] for cb in s.callbacks["export"]}
    for r in range(world):
        cells, uv, e, c, hist = out[r]
        assert np.array_equal(uv, uv1[cells]) and np.array_equal(e, e1[cells]) and np.array_equal(c, c1[cells])
        for name in ("volume2d", "tracer_2d mass"):
            for (v, _), (v1, _) in zip(hist[name], hist1[name]):
                assert abs(v - v1) <= 1e-12 * abs(v1)
        assert hist["tracer_2d overshoot"][-1][:2] == hist1["tracer_2d overshoot"][-1][:2]     # min / max are exact
