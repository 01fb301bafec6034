"""
BASELINE.json configurations at their FULL sizes on one GPU, checked through size-independent properties (the numpy
oracle needs minutes at these sizes): linearity of the linear operator, exact volume conservation in a closed
basin, tracer consistency and the limiter's maximum principle, agreement of the specialised and the generic stage
kernels, and the analytic standing-wave error at 512 x 512.  Everything goes through the C-ABI.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(mesh):
    from thetis_b200.engine import Engine
    return Engine(mesh)


def _rand_state(mesh, seed, amp=1.0):
    rng = np.random.default_rng(seed)
    x = mesh.coords[mesh.cells]
    L = np.ptp(mesh.coords, axis=0).max()
    k = 2 * np.pi / L
    u = amp * (0.3 * np.sin(3 * k * x[..., 0]) * np.cos(2 * k * x[..., 1]) + 0.01 * rng.standard_normal(x.shape[:2]))
    v = amp * (0.2 * np.cos(2 * k * x[..., 0]) * np.sin(4 * k * x[..., 1]) + 0.01 * rng.standard_normal(x.shape[:2]))
    e = amp * (0.4 * np.cos(2 * k * x[..., 0]) * np.cos(k * x[..., 1]) + 0.01 * rng.standard_normal(x.shape[:2]))
    return np.stack([u, v], -1), e


def test_config2_standing_wave_512():
    """waveEq2d (examples/waveEq2d/channel2d_waveEq.py:68): 512 x 512, linear, one period; L2 error against
    -cos(2 pi x / L) cos(2 pi t / T) at 256^2 and 512^2 converges at second order"""
    import thetis_b200._lib as L
    from thetis_b200.mesh import rectangle_mesh, sfc_renumber
    from oracle import swe_oracle as O
    lx, depth, g = 44294.46, 50.0, 9.81
    T = lx / np.sqrt(g * depth)
    errs = []
    for n in (256, 512):
        mesh = sfc_renumber(rectangle_mesh(n, n, lx, lx))
        assert mesh.n_cells == 2 * n * n
        eng = _engine(mesh)
        eng.set_option(L.OPT_NONLINEAR, 0)
        eng.set_field(L.F_BATHYMETRY, depth)
        ic = lambda x, y: -np.cos(2 * np.pi * x / lx)
        eta0 = O.interpolate(mesh, ic)
        A = eng.upload_nodal(np.zeros(eta0.shape + (2,)), eta0)
        B, C = eng.new_state(), eng.new_state()
        nsteps = 20 * n
        dt = T / nsteps
        for _ in range(nsteps):
            eng.swe_stage(0.0, 1.0, dt, A, None, B)
            eng.swe_stage(0.75, 0.25, 0.25 * dt, B, A, C)
            eng.swe_stage(1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0 * dt, C, A, A)
        _, eta = eng.download_nodal(A)
        errs.append(O.l2_error(mesh, eta, ic) / lx)
    assert errs[1] < 1.25e-3                      # test/swe2d/test_standing_wave.py:12,95 threshold, far exceeded
    assert np.log2(errs[0] / errs[1]) > 1.8, errs


def test_config3_stommel_1m_linearity():
    """stommel2d: ~1 M unstructured triangles, linear equations, beta-plane Coriolis, wind stress, linear drag:
    the tendency operator is affine, T(u1 + u2) - T(u1) - T(u2) + T(0) = 0 to rounding"""
    import torch
    import thetis_b200._lib as L
    from thetis_b200.mesh import delaunay_mesh, sfc_renumber
    Lx = 1.0e6
    mesh = sfc_renumber(delaunay_mesh(500_500, Lx, Lx, seed=0))
    assert mesh.n_cells > 990_000
    eng = _engine(mesh)
    eng.set_option(L.OPT_NONLINEAR, 0)
    eng.set_field(L.F_BATHYMETRY, 1000.0)
    Y = mesh.coords[:, 1]
    eng.set_field(L.F_CORIOLIS, 1e-4 + 2e-11 * Y)
    eng.set_field(L.F_WIND_STRESS, np.stack([0.1 * np.sin(np.pi * (Y / Lx - 0.5)), 0 * Y], -1))
    eng.set_field(L.F_LINEAR_DRAG, 1e-6)
    uv1, e1 = _rand_state(mesh, 1)
    uv2, e2 = _rand_state(mesh, 2, amp=0.7)
    ks = []
    for uv, e in ((uv1, e1), (uv2, e2), (uv1 + uv2, e1 + e2), (0 * uv1, 0 * e1)):
        st = eng.upload_nodal(uv, e)
        k = eng.new_state()
        eng.swe_tendency(st, k)
        ks.append(k)
    resid = ks[2] - ks[0] - ks[1] + ks[3]
    scale = ks[2].abs().max().item()
    assert resid.abs().max().item() < 1e-12 * scale


def test_config4_tracer_2m_consistency_and_maximum_principle():
    """demo_2d_tracer size (2 M triangles), coupled SWE -> tracer -> limiter: a constant tracer stays constant
    (test/tracerEq/test_consistency_2d.py:98-107) and the limited solution of a discontinuous initial condition
    stays within the initial bounds up to the small cell-mean overshoot of the upwind scheme (2e-3; the numpy oracle
    gives 4.4e-4 on a 40 x 40 mesh at the same CFL number)"""
    import torch
    import thetis_b200._lib as L
    from thetis_b200.mesh import rectangle_mesh, sfc_renumber
    n = 1000
    mesh = sfc_renumber(rectangle_mesh(n, n, 1.0, 1.0))
    assert mesh.n_cells == 2_000_000
    eng = _engine(mesh)
    eng.set_field(L.F_BATHYMETRY, 1.0)
    x = mesh.coords[mesh.cells]
    uv = np.stack([0.5 - x[..., 1], x[..., 0] - 0.5], -1)
    eta = 0.01 * np.sin(6 * x[..., 0])
    A = eng.upload_nodal(uv, eta)
    dt = 2e-4
    for c0, lo, hi in ((np.full(x.shape[:2], 3.0), 3.0, 3.0),
                       (1.0 + 1.0 * ((np.abs(x[..., 0] - 0.3) < 0.1) & (np.abs(x[..., 1] - 0.5) < 0.1)), 1.0, 2.0)):
        ca = eng.upload_tracer(c0)
        cb, cc = eng.new_tracer(), eng.new_tracer()
        eng.limiter_apply(ca)
        for _ in range(20):
            eng.tracer_stage(0.0, 1.0, dt, ca, None, cb, A)
            eng.tracer_stage(0.75, 0.25, 0.25 * dt, cb, ca, cc, A)
            eng.tracer_stage(1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0 * dt, cc, ca, ca, A)
            eng.limiter_apply(ca)
        out = torch.zeros(4, dtype=torch.float64, device=eng.device)
        eng.tracer_integrals(ca, A, out)
        mn, mx = out[2].item(), out[3].item()
        if lo == hi:
            assert abs(mn - lo) < 1e-11 and abs(mx - hi) < 1e-11, (mn, mx)
        else:
            assert mn > lo - 2e-3 and mx < hi + 2e-3, (mn, mx)


def test_config5_north_sea_4m_properties():
    """north_sea.msh k-sectioned to 3 942 120 triangles: (i) the specialised (SPEC 3) and the generic stage kernels
    agree to rounding on the full workload; (ii) with every boundary closed the nonlinear step conserves volume
    to rounding (VolumeConservation2DCallback criterion) and stays finite"""
    import torch
    import thetis_b200._lib as L
    from harness.workloads import north_sea_mesh, north_sea_setup, tide_values
    mesh = north_sea_mesh(19)
    assert mesh.n_cells == 3_942_120
    setup = north_sea_setup(mesh, wetting_drying=True)
    eng = _engine(mesh)
    eng.set_option(L.OPT_WETTING_DRYING, 1)
    eng.set_option(L.OPT_WD_ALPHA, setup["wd_alpha"])
    eng.set_field(L.F_BATHYMETRY, setup["bath"])
    eng.set_field(L.F_MANNING, setup["manning"])
    eng.set_field(L.F_CORIOLIS, setup["coriolis"])
    eng.set_bc(0, 100, L.BC_ELEV | L.BC_UV, [0.0] * 6)
    eng.set_bc_array(0, 100, L.BC_ELEV, tide_values(setup, 1000.0))
    st = eng.upload_nodal(setup["uv0"], setup["eta0"])
    k1, k2 = eng.new_state(), eng.new_state()
    eng.swe_tendency(st, k1)
    eng.set_option(L.OPT_FORCE_GENERIC_KERNEL, 1)
    eng.swe_tendency(st, k2)
    eng.set_option(L.OPT_FORCE_GENERIC_KERNEL, 0)
    assert torch.isfinite(k1).all()
    assert (k1 - k2).abs().max().item() <= 1e-12 * k1.abs().max().item()
    # closed basin, no wetting-drying (plain P1DG mass matrix): volume conserved to rounding over 10 steps
    eng.set_bc(0, 100, 0, [0.0] * 6)
    eng.set_option(L.OPT_WETTING_DRYING, 0)
    eng.set_field(L.F_BATHYMETRY, np.maximum(setup["bath"], 5.0))
    A, B, C = st, eng.new_state(), eng.new_state()
    o4 = torch.zeros(4, dtype=torch.float64, device=eng.device)
    eng.swe_integrals(A, o4)
    vol0 = o4[3].item()
    dt = 0.1 * setup["dt"]        # the mesh has sliver cells at the coast (inradius / sqrt(area) = 0.06): small step
    for _ in range(10):
        eng.swe_stage(0.0, 1.0, dt, A, None, B)
        eng.swe_stage(0.75, 0.25, 0.25 * dt, B, A, C)
        eng.swe_stage(1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0 * dt, C, A, A)
    eng.swe_integrals(A, o4)
    assert torch.isfinite(A).all()
    assert abs(o4[3].item() - vol0) <= 1e-13 * abs(vol0)


@pytest.mark.parametrize("wd", [True, False])
def test_config5_north_sea_4m_one_step_matches_c_oracle(wd):
    """BASELINE config 5 at full size (3 942 120 triangles): one SSPRK33 step with the tidal elevation re-assigned at
    every stage, GPU (specialised SPEC 3 / SPEC 2 kernels) against oracle/swe_oracle.c on the host cores; fp64,
    relative 1e-10 of each field's max-norm."""
    import thetis_b200._lib as L
    from harness.workloads import north_sea_mesh, north_sea_setup, tide_values
    from oracle import c_oracle as CO
    mesh = north_sea_mesh(19)
    setup = north_sea_setup(mesh, wetting_drying=wd)
    dt = setup["dt"]
    eng = _engine(mesh)
    eng.set_option(L.OPT_WETTING_DRYING, int(wd))
    eng.set_option(L.OPT_WD_ALPHA, setup["wd_alpha"])
    eng.set_field(L.F_BATHYMETRY, setup["bath"])
    eng.set_field(L.F_MANNING, setup["manning"])
    eng.set_field(L.F_CORIOLIS, setup["coriolis"])
    eng.set_bc(0, 100, L.BC_ELEV | L.BC_UV, [0.0] * 6)
    orc = CO.COracle(mesh, setup["bath"], nonlinear=True, lf_on=True, coriolis=setup["coriolis"], manning=setup["manning"],
                     bnd={100: {"elev": 0.0, "uv": (0.0, 0.0)}}, wd_on=wd, wd_alpha=setup["wd_alpha"],
                     threads=CO.host_threads())
    A = eng.upload_nodal(setup["uv0"], setup["eta0"])
    B, C = eng.new_state(), eng.new_state()
    u0 = CO.records_from_nodal(setup["uv0"], setup["eta0"])
    stages = [(0.0, 1.0, 1.0, 0.0), (0.75, 0.25, 0.25, 1.0), (1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0, 0.5)]   # a0, a1, beta, c
    bufs = [(A, None, B), (B, A, C), (C, A, A)]
    u = u0
    t0 = 1000.0
    for (a0, a1, beta, c), (uin, uz, uout) in zip(stages, bufs):
        tv = tide_values(setup, t0 + c * dt)
        eng.set_bc_array(0, 100, L.BC_ELEV, tv)
        eng.swe_stage(a0, a1, beta * dt, uin, uz, uout)
        orc.set_bf_elev(tv)
        u = orc.stage(a0, a1, beta * dt, u, u0 if a0 != 0.0 else None)
    uv_g, eta_g = eng.download_nodal(A)
    uv_c, eta_c = CO.nodal_from_records(u)
    assert np.isfinite(uv_g).all() and np.isfinite(eta_g).all()
    assert np.abs(uv_g - uv_c).max() <= 1e-10 * np.abs(uv_c).max()
    assert np.abs(eta_g - eta_c).max() <= 1e-10 * np.abs(eta_c).max()
