"""Developer timing of BASELINE config 3 (stommel2d: linear, Coriolis, wind stress, linear drag; ~1 M unstructured triangles)
and of the generic-vs-specialised stage kernel penalty."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from thetis_b200.mesh import delaunay_mesh, sfc_renumber
from thetis_b200.engine import Engine
import thetis_b200._lib as L
Lx = 1.0e6
t0 = time.time()
m = sfc_renumber(delaunay_mesh(500_500, Lx, Lx, seed=0))
print("mesh", m.n_cells, round(time.time() - t0, 1), "s", flush=True)
eng = Engine(m)
Y = m.coords[:, 1]
x = m.coords[m.cells]
uv = np.stack([0.1 * np.sin(x[..., 0] / 1e5), 0.1 * np.cos(x[..., 1] / 1e5)], -1)
eta = 0.1 * np.sin(x[..., 0] / 2e5)
nt = m.n_cells


def timeit(label, alg):
    A = eng.upload_nodal(uv, eta); B = eng.new_state(); C = eng.new_state()
    for _ in range(3):
        eng.swe_stage(0.0, 1.0, 1.0, A, None, B)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        eng.swe_stage(0.0, 1.0, 1.0, A, None, B); eng.swe_stage(0.75, 0.25, 0.25, B, A, C); eng.swe_stage(1 / 3, 2 / 3, 2 / 3, C, A, B)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 120
    print(f"{label}: {ms:.4f} ms/stage  alg {alg*nt/ms/1e6:.0f} GB/s ({alg*nt/ms/1e6/6551.0*100:.1f}%)", flush=True)


eng.set_field(L.F_BATHYMETRY, 1000.0)
eng.set_option(L.OPT_NONLINEAR, 0)
timeit("linear closed (SPEC 1)", 228)
eng.set_option(L.OPT_FORCE_GENERIC_KERNEL, 1)
timeit("linear closed (generic)", 228)
eng.set_option(L.OPT_FORCE_GENERIC_KERNEL, 0)
eng.set_field(L.F_CORIOLIS, 1e-4 + 2e-11 * Y)
eng.set_field(L.F_WIND_STRESS, np.stack([0.1 * np.sin(np.pi * (Y / Lx - 0.5)), 0 * Y], -1))
eng.set_field(L.F_LINEAR_DRAG, 1e-6)
timeit("stommel: linear + Coriolis + wind + linear drag", 228 + 12)
eng.set_option(L.OPT_NONLINEAR, 1)
eng.set_field(L.F_VISCOSITY, 100.0)
timeit("nonlinear + stommel terms + viscosity (generic)", 228 + 12 + 14)
