"""Developer timing of the stage kernel on the ~4M-triangle North Sea mesh (not the bench contract)."""
import os, sys, time, argparse
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from thetis_b200.mesh import load_npz_mesh, refine_uniform, sfc_renumber
from thetis_b200.engine import Engine
import thetis_b200._lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=19)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--mode", default="all", help="all, or a comma-separated list of modes")
ap.add_argument("--nosfc", action="store_true")
a = ap.parse_args()
MODES = a.mode.split(",")


def sel(name):
    return "all" in MODES or name in MODES


t0 = time.time()
m = refine_uniform(load_npz_mesh(os.path.join(os.path.dirname(__file__), "..", "tests/golden/north_sea_mesh.npz")), a.k)
if not a.nosfc:
    m = sfc_renumber(m)
print("mesh", m.n_cells, "cells", time.time() - t0, "s", flush=True)
eng = Engine(m)
print("patches", eng.n_patches, "state MB", eng.state_len * 8 / 1e6, "smem", eng.lib.tb_patch_size(eng.ctx), flush=True)
X, Y = m.coords[:, 0], m.coords[:, 1]
bath = 40.0 + 30.0 * np.sin(X / 2.0e5) * np.cos(Y / 1.5e5)
eng.set_field(L.F_BATHYMETRY, bath)
x = m.coords[m.cells]
eta = 0.5 * np.sin(x[..., 0] / 1e5) * np.cos(x[..., 1] / 1e5)
uv = np.stack([0.3 * np.cos(x[..., 0] / 1e5), 0.2 * np.sin(x[..., 1] / 1e5)], -1)
A = eng.upload_nodal(uv, eta)
B = eng.new_state(); Cc = eng.new_state()
nt = m.n_cells

def run(label, alg_bytes):
    for _ in range(3):
        eng.swe_stage(0.0, 1.0, 0.01, A, None, B)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        eng.swe_stage(0.0, 1.0, 0.01, A, None, B)
        eng.swe_stage(0.75, 0.25, 0.0025, B, A, Cc)
        eng.swe_stage(1 / 3, 2 / 3, 0.00667, Cc, A, B)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (3 * a.steps)
    print(f"{label}: {ms:.4f} ms/stage  {nt/ms/1e6:.2f} Gtri-stage/s  alg {alg_bytes*nt/ms/1e6:.0f} GB/s "
          f"({alg_bytes*nt/ms/1e6/6551.0*100:.1f}% of 6551.0)  {9*nt/(3*ms)/1e3:.0f} Mdof-upd/s", flush=True)

if sel("linear"):
    eng.set_option(L.OPT_NONLINEAR, 0)
    run("linear closed", 228)
if sel("nonlinear"):
    eng.set_option(L.OPT_NONLINEAR, 1)
    run("nonlinear+LF closed", 228)
if sel("northsea"):
    eng.set_option(L.OPT_NONLINEAR, 1)
    eng.set_field(L.F_MANNING, 0.03 + 0 * X)
    eng.set_field(L.F_CORIOLIS, 1.2e-4 + 1e-11 * Y)
    eng.set_bc(0, 100, L.BC_ELEV | L.BC_UV, [0.0, 0, 0, 0, 0, 0])
    eng.set_bc_array(0, 100, L.BC_ELEV, np.zeros((m.n_bfacets, 2)))
    run("north sea (nonlinear+LF+Manning+Coriolis+tide)", 236)
if sel("northsea_wd"):
    eng.set_option(L.OPT_NONLINEAR, 1)
    eng.set_field(L.F_MANNING, 0.03 + 0 * X)
    eng.set_field(L.F_CORIOLIS, 1.2e-4 + 1e-11 * Y)
    eng.set_bc(0, 100, L.BC_ELEV | L.BC_UV, [0.0, 0, 0, 0, 0, 0])
    eng.set_bc_array(0, 100, L.BC_ELEV, np.zeros((m.n_bfacets, 2)))
    eng.set_option(L.OPT_WETTING_DRYING, 1)
    run("north sea + wetting-drying", 236)
if sel("northsea_wd_visc"):
    eng.set_option(L.OPT_NONLINEAR, 1)
    eng.set_field(L.F_MANNING, 0.03 + 0 * X)
    eng.set_field(L.F_CORIOLIS, 1.2e-4 + 1e-11 * Y)
    eng.set_bc(0, 100, L.BC_ELEV | L.BC_UV, [0.0, 0, 0, 0, 0, 0])
    eng.set_bc_array(0, 100, L.BC_ELEV, np.zeros((m.n_bfacets, 2)))
    eng.set_option(L.OPT_WETTING_DRYING, 1)
    eng.set_field(L.F_VISCOSITY, 10.0 + 0 * X)
    run("north sea + wetting-drying + viscosity (SPEC 6)", 240 + 14)
    eng.set_option(L.OPT_FORCE_GENERIC_KERNEL, 1)
    run("north sea + wetting-drying + viscosity (generic)", 240 + 14)
    eng.set_option(L.OPT_FORCE_GENERIC_KERNEL, 0)
