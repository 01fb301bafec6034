"""The C/OpenMP oracle port (CPU baseline) must agree with the numpy oracle.  CPU only."""
import numpy as np
import pytest

from thetis_b200.mesh import rectangle_mesh, delaunay_mesh, read_gmsh, sfc_renumber
from oracle import swe_oracle as O
from oracle.c_oracle import COracle, records_from_nodal, nodal_from_records
import os


def _state(mesh, seed=0):
    rng = np.random.default_rng(seed)
    x = mesh.coords[mesh.cells]
    L = np.ptp(mesh.coords, axis=0).max()
    k = 2 * np.pi / L
    u = 0.5 * np.sin(k * x[..., 0] + 0.3) * np.cos(k * x[..., 1]) + 0.05 * rng.standard_normal(x.shape[:2])
    v = -0.5 * np.cos(2 * k * x[..., 0]) * np.sin(k * x[..., 1]) + 0.05 * rng.standard_normal(x.shape[:2])
    e = 0.3 * np.cos(k * x[..., 0]) * np.sin(k * x[..., 1]) + 0.02 * rng.standard_normal(x.shape[:2])
    return np.stack([u, v], -1), e


@pytest.mark.parametrize("nonlinear", [True, False])
def test_c_oracle_tendency_matches_numpy(nonlinear):
    mesh = sfc_renumber(read_gmsh(os.path.join(os.path.dirname(__file__), "golden", "mini_tagged.msh")))
    X, Y = mesh.coords[:, 0], mesh.coords[:, 1]
    bath = 30.0 + 5 * np.sin(X / 7.0) * np.cos(Y / 5.0)
    cor = 1e-4 + 1e-6 * Y
    man = 0.03 + 0 * X
    rng = np.random.default_rng(1)
    bf_elev = 0.5 + 0.1 * rng.standard_normal((mesh.n_bfacets, 2))
    bnd = {100: {"elev": 0.0, "uv": (0.0, 0.0)}}
    uv, eta = _state(mesh)
    co = COracle(mesh, bath, nonlinear=nonlinear, coriolis=cor, manning=man, linear_drag=1e-4, bnd=bnd, bf_elev=bf_elev)
    k = co.tendency(records_from_nodal(uv, eta))
    ku_c, ke_c = nodal_from_records(k)
    from thetis_b200.mesh import FACET_NODES
    full = np.zeros((mesh.n_cells, 3))
    for side in range(2):
        full[mesh.bf_cell, FACET_NODES[mesh.bf_lf, side]] = bf_elev[:, side]
    cells = mesh.cells
    orc = O.SWEOracle(mesh, bath[cells], options=dict(use_nonlinear_equations=nonlinear),
                      fields={"coriolis": cor[cells], "manning_drag_coefficient": man[cells],
                              "linear_drag_coefficient": 1e-4},
                      bnd_conditions={100: {"elev": full, "uv": (0.0, 0.0)}})
    ku, ke = orc.tendency(uv, eta)
    assert np.abs(ku_c - ku).max() / np.abs(ku).max() < 1e-12
    assert np.abs(ke_c - ke).max() / np.abs(ke).max() < 1e-12


@pytest.mark.parametrize("bc", [{"elev": 0.3, "un": 0.15}, {"elev": -0.2, "flux": 500.0}, {"flux": -300.0}, {"un": 0.1},
                                {"uv": (0.1, 0.2)}, {"elev": 0.2}])
def test_c_oracle_bcs_and_steps(bc):
    mesh = rectangle_mesh(10, 8, 500.0, 400.0)
    bath = 12 + 0.004 * mesh.coords[:, 0]
    uv, eta = _state(mesh, 3)
    co = COracle(mesh, bath, bnd={1: bc, 3: bc})
    rec = records_from_nodal(uv, eta)
    co.ssprk33(rec, 0.2, 5)
    orc = O.SWEOracle(mesh, bath[mesh.cells], bnd_conditions={1: bc, 3: bc})
    st = O.ShuOsherStepper(orc, [uv, eta], 0.2)
    for i in range(5):
        st.advance(i * 0.2)
    u_c, e_c = nodal_from_records(rec)
    assert np.abs(u_c - uv).max() / np.abs(uv).max() < 1e-12
    assert np.abs(e_c - eta).max() / np.abs(eta).max() < 1e-12


def test_c_oracle_wetting_drying():
    mesh = rectangle_mesh(14, 6, 14e3, 1.2e3)
    bath = 3.0 - 5.0 * mesh.coords[:, 0] / 14e3
    uv, eta = _state(mesh, 5)
    co = COracle(mesh, bath, manning=0.02, bnd={1: {"elev": 0.5}}, wd_on=True, wd_alpha=0.4)
    ku_c, ke_c = nodal_from_records(co.tendency(records_from_nodal(uv, eta)))
    orc = O.SWEOracle(mesh, bath[mesh.cells], options=dict(use_wetting_and_drying=True, wetting_and_drying_alpha=0.4),
                      fields={"manning_drag_coefficient": 0.02}, bnd_conditions={1: {"elev": 0.5}})
    ku, ke = orc.tendency(uv, eta)
    assert np.abs(ku_c - ku).max() / np.abs(ku).max() < 1e-12
    assert np.abs(ke_c - ke).max() / np.abs(ke).max() < 1e-12
