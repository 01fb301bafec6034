"""cProfile of the e2e step loop (host overhead of the reference-shaped API)."""
import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from harness.workloads import north_sea_mesh, north_sea_setup
from harness.runs import SingleSWE
k = int(sys.argv[1]) if len(sys.argv) > 1 else 19
mesh = north_sea_mesh(k)
setup = north_sea_setup(mesh)
run = SingleSWE(mesh, setup)
if len(sys.argv) > 2 and sys.argv[2] == "graphs":
    run.enable_stage_graphs()
for _ in range(5):
    run.step_e2e()
torch.cuda.synchronize()
if len(sys.argv) > 3 and sys.argv[3] == "fused":
    run.use_fused_norms(True)
    if len(sys.argv) > 2 and sys.argv[2] == "graphs":
        run.enable_stage_graphs()
    for _ in range(5):
        run.step_e2e()
    torch.cuda.synchronize()
# plain wall-clock first (no profiler overhead)
t0 = time.perf_counter()
for _ in range(300):
    run.step_e2e()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"no profiler: host issue time {t_host/300*1e3:.3f} ms/step; incl. GPU drain {(time.perf_counter()-t0)/300*1e3:.3f} ms/step")
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    run.step_e2e()
pr.disable()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host issue time {t_host*10:.3f} ms/step; incl. GPU drain {t_all*10:.3f} ms/step")
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
