"""Developer diagnostic: which SURVEY-8f ingredient breaks bit-identity between 1 GPU and 2 GPUs."""
import os, sys, socket
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))


def mesh_():
    from thetis_b200.mesh import delaunay_mesh, sfc_renumber
    return sfc_renumber(delaunay_mesh(1500, 18e3, 8e3, seed=11))


def solver_(mesh_obj, variant):
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, Constant, as_shim_mesh, ShimMesh
    sm = mesh_obj if isinstance(mesh_obj, ShimMesh) else as_shim_mesh(mesh_obj)
    lx = 18e3
    P1 = FunctionSpace(sm, "CG", 1)
    b = Function(P1).interpolate(lambda x, y: 10.0 + 2.0 * np.cos(2 * np.pi * x / lx))
    s = solver2d.FlowSolver2d(sm, b)
    o = s.options
    o.swe_timestepper_type = "ERKLSPUM2" if "erk" in variant else "SSPRK33"
    o.tracer_timestepper_type = o.swe_timestepper_type
    o.swe_timestepper_options.use_automatic_timestep = False
    o.tracer_timestepper_options.use_automatic_timestep = False
    o.timestep = 1.0
    o.simulation_end_time = 1.0 * 6
    o.simulation_export_time = 1.0 * 6
    if "visc" in variant:
        o.horizontal_viscosity = Function(P1).interpolate(lambda x, y: 20.0 * (1.0 + 0.3 * np.sin(y / 2e3)))
        o.use_grad_div_viscosity_term = "graddiv" in variant
        o.use_grad_depth_viscosity_term = "nodepth" not in variant
    o.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d",
                    diffusivity=Constant(12.0) if "diff" in variant else None)
    o.use_limiter_for_tracers = "lim" in variant
    s.bnd_functions["shallow_water"] = {1: {"elev": Constant(0.2), "uv": Constant((0.05, 0.0))}}
    s.bnd_functions["tracer"] = {1: {"value": Constant(4.0)}}
    s.assign_initial_conditions(elev=lambda x, y: 0.5 * np.cos(np.pi * x / lx),
                                tracer=lambda x, y: 4.5 + 2.0 * np.exp(-((x - lx / 2) ** 2 + (y - 4e3) ** 2) / 2e3 ** 2))
    return s


def worker(rank, world, port, variant, out):
    import torch, torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from thetis_b200.parallel import distribute_mesh
        sm = distribute_mesh(mesh_(), rank, world, halo="vertex")
        lm = sm.topology_mesh
        s = solver_(sm, variant)
        s.iterate()
        n = sm.halo_plan.part.n_owned
        out[rank] = (lm.meta["global_cells"][:n].copy(), s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2)[:n].copy(),
                     s.fields.elev_2d.dat.data_ro.reshape(-1, 3)[:n].copy(),
                     s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)[:n].copy())
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    import contextlib, io
    for variant in sys.argv[1:]:
        sk = socket.socket(); sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]; sk.close()
        mgr = mp.Manager(); out = mgr.dict()
        with contextlib.redirect_stdout(io.StringIO()):
            mp.spawn(worker, args=(2, port, variant, out), nprocs=2, join=True)
            s = solver_(mesh_(), variant); s.iterate()
        uv1 = s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2); e1 = s.fields.elev_2d.dat.data_ro.reshape(-1, 3)
        c1 = s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)
        msg = [variant]
        for r in range(2):
            cells, uv, e, c = out[r]
            du = np.abs(uv - uv1[cells]).max(axis=(1, 2)); de = np.abs(e - e1[cells]).max(axis=1); dc = np.abs(c - c1[cells]).max(axis=1)
            msg.append(f"r{r}: du {du.max():.2e} ({(du>0).sum()} cells) de {de.max():.2e} ({(de>0).sum()}) dc {dc.max():.2e} ({(dc>0).sum()}) of {cells.size}"
                       f" first-bad-local {np.nonzero(du>0)[0][:6].tolist()}")
        print(" | ".join(msg), flush=True)
