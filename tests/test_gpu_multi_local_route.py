"""
GPU side of the route for a mesh that ARRIVES distributed (`parallel.plan_from_local_mesh`; Firedrake under mpiexec):
the coupled SWE -> tracer -> limiter run of tests/test_gpu_multi.py on 2 ranks whose cells are owned in scattered
(Voronoi) blocks and numbered at random locally, both ranks on cuda:0 with the halo staged through the host over
gloo, must equal the 1-GPU run bit for bit.

STATUS: written after this round's GPU budget was spent -- NEVER RUN ON HARDWARE when committed.  The same host
classes and the same plan pass on the CPU against the oracle-backed engine double
(tests/test_distributed_host_path_gloo.py), and the plan equals `partition_mesh`'s layout when fed its parts
(tests/test_local_mesh_plan.py); the kernels see nothing new (a local mesh of the same shape as `distribute_mesh`'s).
Marked xfail(strict=False) so that its first hardware run reports XPASS / XFAIL instead of deciding the suite.
"""
import os
import sys
from datetime import timedelta

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="plan_from_local_mesh on the GPU: first hardware run (no GPU budget "
                                                     "was left when it was written); XPASS = verified")]

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=timedelta(seconds=240))
    try:
        import test_gpu_multi as TM
        from test_distributed_host_path_gloo import _local_view
        from thetis_b200 import parallel as PA
        from thetis_b200.shim import ShimMesh
        mesh = TM._coupled_mesh()
        c = mesh.cell_centroids()
        pts = c[np.random.default_rng(5).choice(mesh.n_cells, world, replace=False)]
        owner = np.argmin(((c[:, None, :] - pts[None]) ** 2).sum(-1), axis=1).astype(np.int32)
        lm, n_owned, gids = _local_view(mesh, owner, rank, seed=3)
        plan, part = PA.plan_from_local_mesh(lm, n_owned, gids, halo="vertex")
        sm = ShimMesh(part.mesh)
        sm.boundary_len = dict(part.mesh.meta["global_boundary_len"])
        sm.halo_plan = plan
        s = TM._coupled_solver(sm)
        s.iterate()
        n = part.n_owned
        out[rank] = (part.mesh.meta["global_cells"][:n].copy(),
                     s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2)[:n].copy(),
                     s.fields.elev_2d.dat.data_ro.reshape(-1, 3)[:n].copy(),
                     s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)[:n].copy(), s.last_norms)
    finally:
        dist.destroy_process_group()


def test_coupled_run_on_a_mesh_that_arrives_distributed_is_bit_identical():
    import torch.multiprocessing as mp
    import test_gpu_multi as TM
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, TM._free_port(), out), nprocs=world, join=True)
    s = TM._coupled_solver(TM._coupled_mesh())
    s.iterate()
    uv1 = s.fields.uv_2d.dat.data_ro.reshape(-1, 3, 2)
    e1 = s.fields.elev_2d.dat.data_ro.reshape(-1, 3)
    c1 = s.fields.tracer_2d.dat.data_ro.reshape(-1, 3)
    seen = []
    for r in range(world):
        cells, uv, e, c, norms = out[r]
        seen.append(cells)
        assert np.array_equal(uv, uv1[cells]) and np.array_equal(e, e1[cells]) and np.array_equal(c, c1[cells])
        assert abs(norms[0] - s.last_norms[0]) <= 1e-12 * s.last_norms[0]
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(uv1.shape[0]))
