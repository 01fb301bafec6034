"""Developer timing of BASELINE config 4 (coupled SWE + tracer + limiter, 2M triangles) on one GPU."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from thetis_b200.mesh import rectangle_mesh, sfc_renumber
from thetis_b200.engine import Engine
import thetis_b200._lib as L
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t0 = time.time()
m = sfc_renumber(rectangle_mesh(n, n, 1.0, 1.0))
print("mesh", m.n_cells, time.time() - t0, flush=True)
eng = Engine(m)
eng.set_field(L.F_BATHYMETRY, 1.0)
x = m.coords[m.cells]
uv = np.stack([0.5 - x[..., 1], x[..., 0] - 0.5], -1)
eta = 0.01 * np.sin(6 * x[..., 0])
c0 = 1.0 + np.exp(-((x[..., 0] - 0.25) ** 2 + (x[..., 1] - 0.5) ** 2) / 0.01)
A = eng.upload_nodal(uv, eta); B = eng.new_state(); C = eng.new_state()
ca = eng.upload_tracer(c0); cb = eng.new_tracer(); cc = eng.new_tracer()
nt = m.n_cells
dt = 1e-4
def step():
    eng.swe_stage(0.0, 1.0, dt, A, None, B); eng.swe_stage(0.75, 0.25, 0.25 * dt, B, A, C); eng.swe_stage(1 / 3, 2 / 3, 2 / 3 * dt, C, A, A)
    eng.tracer_stage(0.0, 1.0, dt, ca, None, cb, A); eng.tracer_stage(0.75, 0.25, 0.25 * dt, cb, ca, cc, A); eng.tracer_stage(1 / 3, 2 / 3, 2 / 3 * dt, cc, ca, ca, A)
    eng.limiter_apply(ca)
def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(step)
print(f"coupled step: {ms:.3f} ms  -> {(9+3)*nt/ms/1e3:.0f} M dof-updates/s")
ms_s = timeit(lambda: eng.swe_stage(0.75, 0.25, dt, B, A, C))
ms_t = timeit(lambda: eng.tracer_stage(0.75, 0.25, dt, cb, ca, cc, A))
ms_l = timeit(lambda: eng.limiter_apply_to(ca, cb))
print(f"swe stage {ms_s:.4f} ms ({228*nt/ms_s/1e6:.0f} GB/s alg)  tracer stage {ms_t:.4f} ms ({(64+72)*nt/ms_t/1e6:.0f} GB/s alg incl. SWE record read)  limiter {ms_l:.4f} ms ({64*nt/ms_l/1e6:.0f} GB/s alg)")
