import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from thetis_b200.mesh import FACET_NODES
from thetis_b200.workloads import north_sea_mesh, north_sea_setup, tide_values
from thetis_b200.parallel import SingleSWE
import thetis_b200._lib as L
from oracle import swe_oracle as O
mesh = north_sea_mesh(k=1)
setup = north_sea_setup(mesh, wetting_drying=False)
run = SingleSWE(mesh, setup, wd=False)
cells = mesh.cells
tv = tide_values(setup, 0.0)
full = np.zeros((mesh.n_cells, 3))
for side in range(2):
    full[mesh.bf_cell, FACET_NODES[mesh.bf_lf, side]] = tv[:, side]
def oracle(bnd):
    orc = O.SWEOracle(mesh, setup["bath"][cells], fields={"manning_drag_coefficient": setup["manning"][cells], "coriolis": setup["coriolis"][cells]}, bnd_conditions=bnd)
    return orc.tendency(setup["uv0"], setup["eta0"])
eng = run.eng
def check(label, bnd):
    ku, ke = oracle(bnd)
    k = eng.new_state(); eng.swe_tendency(run.ts.device_state(), k)
    gu, ge = eng.download_nodal(k)
    du = np.abs(gu - ku).max(axis=(1, 2)); de = np.abs(ge - ke).max(axis=1)
    bad = np.nonzero((du > 1e-9 * np.abs(ku).max()) | (de > 1e-9 * np.abs(ke).max()))[0]
    print(label, "err", du.max() / np.abs(ku).max(), de.max() / np.abs(ke).max(), "n bad", bad.size)
    if bad.size:
        isb = (mesh.nbr[bad] < 0)
        mk = [sorted(set(mesh.bf_marker[-(mesh.nbr[b][mesh.nbr[b] < 0] + 1)].tolist())) for b in bad[:12]]
        print("   bad cells", bad[:12], "markers", mk)
    return gu, ge
gu, ge = check("harness", {100: {"elev": full, "uv": (0.0, 0.0)}})
su, se = eng.download_nodal(run.ts.device_state())
print("state match", np.abs(su - setup["uv0"]).max(), np.abs(se - setup["eta0"]).max())
# compare external elevation the harness uploaded with the oracle's
hv = run.ts.adaptor.bfacet_values(run.tide)
print("tide values diff", np.abs(hv - tv)[mesh.bf_marker == 100].max(), "nonzero on coast", np.abs(hv[mesh.bf_marker == 200]).max())
eng.set_bc(0, 100, L.BC_ELEV | L.BC_UV, [0, 0, 0, 0, 0, 0]); eng.set_bc_array(0, 100, L.BC_ELEV, tv)
check("direct bc array", {100: {"elev": full, "uv": (0.0, 0.0)}})
