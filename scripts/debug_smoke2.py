import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from thetis_b200.mesh import FACET_NODES
from thetis_b200.workloads import north_sea_mesh, north_sea_setup, tide_values
from thetis_b200.engine import Engine
import thetis_b200._lib as L
from oracle import swe_oracle as O
mesh = north_sea_mesh(k=1)
setup = north_sea_setup(mesh, wetting_drying=False)
cells = mesh.cells
uv, eta = setup["uv0"], setup["eta0"]
def run(label, nonlin=True, lf=True, manning=False, cor=False, bc=False):
    eng = Engine(mesh)
    eng.set_option(L.OPT_NONLINEAR, nonlin); eng.set_option(L.OPT_LAX_FRIEDRICHS, lf)
    eng.set_field(L.F_BATHYMETRY, setup["bath"])
    fields = {}
    if manning:
        eng.set_field(L.F_MANNING, setup["manning"]); fields["manning_drag_coefficient"] = setup["manning"][cells]
    if cor:
        eng.set_field(L.F_CORIOLIS, setup["coriolis"]); fields["coriolis"] = setup["coriolis"][cells]
    bnd = {}
    if bc:
        eng.set_bc(0, 100, L.BC_ELEV | L.BC_UV, [0.3, 0, 0, 0, 0, 0]); bnd = {100: {"elev": 0.3, "uv": (0.0, 0.0)}}
    orc = O.SWEOracle(mesh, setup["bath"][cells], options=dict(use_nonlinear_equations=nonlin, use_lax_friedrichs_velocity=lf), fields=fields, bnd_conditions=bnd)
    ku, ke = orc.tendency(uv, eta)
    st = eng.upload_nodal(uv, eta); k = eng.new_state(); eng.swe_tendency(st, k)
    gu, ge = eng.download_nodal(k)
    du = np.abs(gu - ku).max(axis=(1, 2)); de = np.abs(ge - ke).max(axis=1)
    w = int(np.argmax(du))
    print(label, "err u", du.max() / np.abs(ku).max(), "eta", de.max() / np.abs(ke).max(), "worst", w, "n bad", (du > 1e-9 * np.abs(ku).max()).sum())
    return w, gu, ku, ge, ke
run("linear", nonlin=False)
run("nonlin noLF", lf=False)
w, gu, ku, ge, ke = run("nonlin LF")
print(" gpu", gu[w].ravel(), "\n orc", ku[w].ravel(), "\n nbr", mesh.nbr[w], "coords", mesh.coords[cells[w]].tolist(), "uv", uv[w].tolist())
run("nonlin LF manning", manning=True)
run("nonlin LF cor", cor=True)
run("nonlin LF bc", bc=True)
print("---- combos")
run("man+cor", manning=True, cor=True)
run("man+cor+bc", manning=True, cor=True, bc=True)
def run2():
    from thetis_b200.workloads import tide_values
    eng = Engine(mesh)
    eng.set_option(L.OPT_NONLINEAR, True); eng.set_option(L.OPT_WETTING_DRYING, False)
    eng.set_field(L.F_BATHYMETRY, setup["bath"])
    eng.set_option(L.OPT_LAX_FRIEDRICHS, True)
    eng.set_option(0, 9.81); eng.set_option(1, 1000.0); eng.set_option(5, 0.0); eng.set_option(4, 1.0)
    eng.set_field(1, setup["coriolis"]); eng.set_field(2, setup["manning"])
    for f in range(3, 9): eng.set_field(f, None)
    eng.set_bc(0, 100, 3, np.zeros(6))
    tv = tide_values(setup, 0.0)
    eng.set_bc_array(0, 100, 1, np.zeros((mesh.n_bfacets, 2)))
    eng.set_bc_array(0, 100, 1, tv)
    full = np.zeros((mesh.n_cells, 3))
    from thetis_b200.mesh import FACET_NODES
    for side in range(2):
        full[mesh.bf_cell, FACET_NODES[mesh.bf_lf, side]] = tv[:, side]
    orc = O.SWEOracle(mesh, setup["bath"][cells], fields={"manning_drag_coefficient": setup["manning"][cells], "coriolis": setup["coriolis"][cells]}, bnd_conditions={100: {"elev": full, "uv": (0.0, 0.0)}})
    ku, ke = orc.tendency(uv, eta)
    st = eng.upload_nodal(uv, eta); k = eng.new_state(); eng.swe_tendency(st, k)
    gu, ge = eng.download_nodal(k)
    print("replay of harness calls: err", np.abs(gu - ku).max() / np.abs(ku).max(), np.abs(ge - ke).max() / np.abs(ke).max())
run2()
