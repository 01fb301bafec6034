"""
Synthetic inputs of the BASELINE.json configurations (host, numpy).  Shared by
bench.py, the GPU tests and __graft_entry__.smoke() so they all run the same
workload definitions.  No computation of the hot path happens here.
"""
from __future__ import annotations

import os

import numpy as np

from thetis_b200.mesh import FACET_NODES, load_npz_mesh, refine_uniform, sfc_renumber

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


# ---------------------------------------------------------------------- BASELINE configs 1-4
CONFIG_NAMES = {
    1: "demo_2d_channel: 40x25 rectangle (2 000 tri), nonlinear SWE + LF, closed, Gaussian hump",
    2: "waveEq2d standing wave: 512x512 structured (524 288 tri), linear SWE, closed",
    3: "stommel2d: ~1 M Delaunay triangles, linear SWE + Coriolis (beta plane) + wind stress + linear drag",
    4: "demo_2d_tracer: 1000x1000 structured (2 M tri), nonlinear SWE + tracer_eq_2d + VertexBasedP1DGLimiter",
    5: "north_sea tidal model: north_sea.msh k-sectioned (3 942 120 tri at k = 19), Manning + Coriolis + tide + wetting-drying",
}


def config_mesh(cfg, scale=1.0):
    """Global mesh of BASELINE config 1-4, SFC-ordered.  ``scale`` < 1 shrinks the cell count (tests)."""
    from thetis_b200.mesh import rectangle_mesh, delaunay_mesh
    if cfg == 1:
        return sfc_renumber(rectangle_mesh(max(4, int(40 * scale)), max(2, int(25 * scale)), 40e3, 2e3))
    if cfg == 2:
        n = max(8, int(512 * scale))
        return sfc_renumber(rectangle_mesh(n, n, 44294.46, 44294.46))
    if cfg == 3:
        return sfc_renumber(delaunay_mesh(max(200, int(500_500 * scale * scale)), 1.0e6, 1.0e6, seed=0))
    if cfg == 4:
        n = max(8, int(1000 * scale))
        return sfc_renumber(rectangle_mesh(n, n, 1.0, 1.0))
    raise ValueError(cfg)


def _cfl_dt(mesh, depth, u_scale):
    """thetis rule (solver2d.py:150-177,237) on cell values: 0.05 * min(h_elem / (sqrt(g max(b, 0.05)) + U))"""
    h_el = np.sqrt(mesh.cell_area())
    return 0.05 * float((h_el / (np.sqrt(9.81 * max(depth, 0.05)) + u_scale)).min())


def config_solver(cfg, mesh_obj, global_mesh=None):
    """FlowSolver2d mirror configured for BASELINE config 1-4 on `mesh_obj` (a Mesh2D or a distributed shim mesh;
    ``global_mesh`` = the undistributed mesh the time step is derived from)."""
    from thetis_b200 import solver2d
    from thetis_b200.shim import Function, FunctionSpace, Constant, ShimMesh, as_shim_mesh
    sm = mesh_obj if isinstance(mesh_obj, ShimMesh) else as_shim_mesh(mesh_obj)
    gm = global_mesh if global_mesh is not None else sm.topology_mesh
    P1 = FunctionSpace(sm, "CG", 1)
    depth = {1: 20.0, 2: 50.0, 3: 1000.0, 4: 1.0}[cfg]
    b = Function(P1, name="Bathymetry")
    b.dat.data[:] = depth
    s = solver2d.FlowSolver2d(sm, b)
    o = s.options
    o.swe_timestepper_type = "SSPRK33"
    o.swe_timestepper_options.use_automatic_timestep = False
    o.tracer_timestepper_options.use_automatic_timestep = False
    o.simulation_end_time = 1e30
    o.simulation_export_time = 1e30
    if cfg == 1:
        o.timestep = _cfl_dt(gm, depth, 0.1)
        s.assign_initial_conditions(elev=lambda x, y: 2.0 * np.exp(-((x - 20e3) / 4e3) ** 2))
    elif cfg == 2:
        L = 44294.46
        o.use_nonlinear_equations = False
        o.timestep = _cfl_dt(gm, depth, 0.0)
        s.assign_initial_conditions(elev=lambda x, y: -np.cos(2 * np.pi * x / L))
    elif cfg == 3:
        L = 1.0e6
        o.use_nonlinear_equations = False
        o.coriolis_frequency = Function(P1).interpolate(lambda x, y: 1.0e-4 + 2.0e-11 * y)
        o.wind_stress = Function(FunctionSpace(sm, "CG", 1, value_size=2)).interpolate(
            lambda x, y: (0.1 * np.sin(np.pi * (y / L - 0.5)), 0.0 * y))
        o.linear_drag_coefficient = Constant(1.0e-6)
        o.timestep = _cfl_dt(gm, depth, 0.0)
        s.assign_initial_conditions(elev=lambda x, y: 1.0e-3 * np.sin(2 * np.pi * x / L) * np.sin(np.pi * y / L))
    elif cfg == 4:
        o.add_tracer_2d("tracer_2d", "Depth averaged tracer", "Tracer2d")
        o.use_limiter_for_tracers = True
        o.timestep = _cfl_dt(gm, depth, 0.75)

        def bell(x, y):
            r = np.sqrt((x - 0.25) ** 2 + (y - 0.5) ** 2) / 0.15
            cone = np.sqrt((x - 0.5) ** 2 + (y - 0.25) ** 2) / 0.15
            slot = (np.sqrt((x - 0.5) ** 2 + (y - 0.75) ** 2) < 0.15) & ~((np.abs(x - 0.5) < 0.025) & (y < 0.85))
            return 1.0 + 0.25 * (1 + np.cos(np.pi * np.minimum(r, 1.0))) + np.maximum(1.0 - cone, 0.0) + 1.0 * slot
        s.assign_initial_conditions(uv=lambda x, y: (0.5 - y, x - 0.5), tracer=bell)
    else:
        raise ValueError(cfg)
    return s
