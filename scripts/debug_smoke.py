import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from thetis_b200.mesh import FACET_NODES
from thetis_b200.workloads import north_sea_mesh, north_sea_setup, tide_values
from thetis_b200.parallel import SingleSWE
from oracle import swe_oracle as O
for wd in (False, True):
    mesh = north_sea_mesh(k=1)
    setup = north_sea_setup(mesh, wetting_drying=wd)
    run = SingleSWE(mesh, setup, wd=wd)
    cells = mesh.cells
    tv = tide_values(setup, 0.0)
    full = np.zeros((mesh.n_cells, 3))
    for side in range(2):
        full[mesh.bf_cell, FACET_NODES[mesh.bf_lf, side]] = tv[:, side]
    orc = O.SWEOracle(mesh, setup["bath"][cells],
                      options=dict(use_wetting_and_drying=wd, wetting_and_drying_alpha=setup["wd_alpha"]),
                      fields={"manning_drag_coefficient": setup["manning"][cells], "coriolis": setup["coriolis"][cells]},
                      bnd_conditions={100: {"elev": full, "uv": (0.0, 0.0)}})
    ku, ke = orc.tendency(setup["uv0"], setup["eta0"])
    eng = run.eng
    k = eng.new_state()
    eng.swe_tendency(run.ts.device_state(), k)
    gu, ge = eng.download_nodal(k)
    du = np.abs(gu - ku).max(axis=(1, 2)); de = np.abs(ge - ke).max(axis=1)
    print("wd", wd, "dt", setup["dt"], run.ts.dt, "tend err", du.max() / np.abs(ku).max(), de.max() / np.abs(ke).max(),
          "max tend", np.abs(ku).max(), np.abs(ke).max())
    bad = np.argsort(-du)[:5]
    print(" worst cells", bad, du[bad], "nbr", mesh.nbr[bad].tolist(), "bath", setup["bath"][cells[bad]].tolist())
    print(" nan?", np.isnan(gu).sum(), np.isnan(ku).sum())
