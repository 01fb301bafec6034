"""
Synthetic inputs of the BASELINE.json configurations (host, numpy).  Shared by
bench.py, the GPU tests and __graft_entry__.smoke() so they all run the same
workload definitions.  No computation of the hot path happens here.
"""
from __future__ import annotations

import os

import numpy as np

from .mesh import FACET_NODES, load_npz_mesh, refine_uniform, sfc_renumber

__all__ = ["north_sea_mesh", "north_sea_setup", "tide_values", "NORTH_SEA_NPZ"]

NORTH_SEA_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                             "north_sea_mesh.npz")

M2_PERIOD = 44714.0      # s


def north_sea_mesh(k=19, renumber=True):
    """
    BASELINE config 5 mesh: the reference's demos/north_sea.msh (10 920 triangles; arrays in
    tests/golden/north_sea_mesh.npz) k-sectioned: k=19 -> 3 942 120 triangles.
    """
    m = refine_uniform(load_npz_mesh(NORTH_SEA_NPZ), k)
    return sfc_renumber(m) if renumber else m


def north_sea_setup(mesh, wetting_drying=True, seed=1234):
    """
    Synthetic fields for the North Sea tidal configuration (demos/demo_2d_north_sea.py: Manning 0.03,
    f = 2 Omega sin(lat), boundary 100 = tidal elevation Function + uv = 0, boundary 200 closed; the real
    bathymetry is an HDF5 file that cannot be read here -> smooth analytic bathymetry, shallow and slightly
    negative along the coast when wetting-drying is on).
    Returns a dict of numpy arrays over the mesh's geometric vertices / cells.
    """
    from scipy.spatial import cKDTree
    X, Y = mesh.coords[:, 0], mesh.coords[:, 1]
    coast = mesh.bf_marker == 200
    cv = np.unique(mesh.cells[mesh.bf_cell[coast][:, None], FACET_NODES[mesh.bf_lf[coast]]])
    dist, _ = cKDTree(mesh.coords[cv]).query(mesh.coords)
    if wetting_drying:
        bath = np.minimum(-1.0 + 1.5e-3 * dist, 200.0)
    else:
        bath = np.minimum(10.0 + 1.5e-3 * dist, 200.0)
    # UTM30-like northing -> latitude (coarse linear map; only used to give f a realistic variation)
    lat = 48.0 + (Y - Y.min()) / max(np.ptp(Y), 1.0) * 14.0
    coriolis = 2 * 7.292e-05 * np.sin(np.deg2rad(lat))
    manning = np.full_like(X, 3.0e-02)
    rng = np.random.default_rng(seed)
    x = mesh.coords[mesh.cells]
    Lx = max(np.ptp(X), 1.0)
    eta0 = 0.2 * np.sin(2 * np.pi * (x[..., 0] - X.min()) / Lx) + 1e-3 * rng.uniform(-1, 1, x.shape[:2])
    uv0 = np.stack([0.05 * np.cos(2 * np.pi * (x[..., 1] - Y.min()) / Lx), 0.03 * np.sin(2 * np.pi * (x[..., 0] - X.min()) / Lx)], -1)
    # phase of the tidal wave along the open boundary
    p = mesh.coords[mesh.cells[mesh.bf_cell[:, None], FACET_NODES[mesh.bf_lf]]]      # (nb, 2, 2)
    phase = 2 * np.pi * (p[..., 0] - X.min() + p[..., 1] - Y.min()) / (2.0 * Lx)
    # CFL time step, thetis rule (solver2d.py:150-177,237): 0.05 * min(h_elem / (sqrt(g max(b, 0.05)) + U))
    area = mesh.cell_area()
    h_el = np.sqrt(area)
    bc = np.maximum(bath[mesh.cells].max(axis=1), 0.05)
    dt = 0.05 * float((h_el / (np.sqrt(9.81 * bc) + 1.5)).min())
    return dict(bath=bath, coriolis=coriolis, manning=manning, eta0=eta0, uv0=uv0, tide_phase=phase, dt=dt,
                wetting_drying=bool(wetting_drying), wd_alpha=0.5)


def tide_values(setup, t, amplitude=1.0):
    """(nb, 2) external elevation at the nodes of every exterior facet at time t (M2 harmonic)."""
    return amplitude * np.sin(2 * np.pi * t / M2_PERIOD + setup["tide_phase"])
