"""Host-side cost of one e2e step (tiny mesh: the GPU is never the bottleneck).  python scripts/host_overhead.py"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from harness.workloads import north_sea_mesh, north_sea_setup
from harness.runs import SingleSWE
mesh = north_sea_mesh(1)
setup = north_sea_setup(mesh)
run = SingleSWE(mesh, setup)
run.use_fused_norms(True)
run.enable_stage_graphs()
for _ in range(20):
    run.step_e2e()
torch.cuda.synchronize()


def timeit(fn, n=2000):
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t = time.perf_counter() - t0
    torch.cuda.synchronize()
    return t / n * 1e6


print(f"step_e2e (3 stages, tidal forcing every stage): {timeit(run.step_e2e):.1f} us/step")
ts = run.ts
print(f"  update_forcings alone: {timeit(lambda: run.update_forcings(1.0)):.1f} us")
print(f"  _push_dynamic, nothing changed: {timeit(ts._push_dynamic):.1f} us")


def forced_push():
    run.update_forcings(1.0)
    ts._push_dynamic()


print(f"  update_forcings + _push_dynamic (tide upload): {timeit(forced_push):.1f} us")
g = ts.stage_graphs[0]
print(f"  graph replay (1 kernel): {timeit(g.replay):.1f} us")
print(f"  engine.stream property: {timeit(lambda: ts.engine.stream):.2f} us")
print(f"  advance without forcings: {timeit(lambda: ts.advance(0.0, None)):.1f} us/step")
print(f"  norms D2H copy: {timeit(lambda: run._norms_host.copy_(run._norms, non_blocking=True)):.1f} us")
