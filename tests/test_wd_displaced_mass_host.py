"""
The displaced-mass elevation update of the stage kernel (thetis_b200/csrc/tb_wd_mass.cuh, TB_OPT_WD_DISPLACED_MASS)
is plain C++ shared by device and host: here the SAME source is compiled with g++ and checked on the CPU against the
oracle's cell-local Newton solve of the reference's wetting-drying mass functional (oracle.SWEOracle.displaced_mass /
solve_displaced_mass, pinned to shallowwater_eq.py:917-920 executed from the reference tree).  fp64, 1e-11.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import reference_cases as RC                               # noqa: E402
from oracle import swe_oracle as O                         # noqa: E402

WRAPPER = r'''
#include "tb_wd_mass.cuh"
extern "C" int wd_update_cells(long long n, double *eta_io, double a0, const double *eta0, double a1, const double *etai,
                               const double *b, const double *al, double a2, const double *qlam, const double *qw, int nq,
                               int *its) {
    int worst = 0;
    for (long long c = 0; c < n; ++c) {
        const int it = tb_wd_displaced_update(eta_io + 3 * c, a0, eta0 ? eta0 + 3 * c : 0, a1, etai + 3 * c, b + 3 * c,
                                              al ? al + 3 * c : 0, a2, qlam, qw, nq);
        its[c] = it;
        if (it < 0) worst = -1; else if (worst >= 0 && it > worst) worst = it;
    }
    return worst;
}
'''


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("wdm")
    src = d / "wdm_host.cpp"
    src.write_text(WRAPPER)
    so = d / "libwdm_host.so"
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-I", os.path.join(ROOT, "thetis_b200", "csrc"),
                    "-o", str(so), str(src)], check=True)
    lb = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)
    lb.wd_update_cells.argtypes = [C.c_longlong, dp, C.c_double, dp, C.c_double, dp, dp, dp, C.c_double, dp, dp, C.c_int,
                                   C.POINTER(C.c_int)]
    lb.wd_update_cells.restype = C.c_int
    return lb


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _update(lib, orc, lin, a0, eta0, a1, etai):
    al = orc._alpha_nodal()
    a2 = 0.0 if al is not None else float(orc.options["wetting_and_drying_alpha"]) ** 2
    out = np.ascontiguousarray(lin, dtype=np.float64).copy()
    its = np.zeros(out.shape[0], dtype=np.int32)
    args = [np.ascontiguousarray(x, dtype=np.float64) if x is not None else None for x in (eta0, etai, orc.bath, al)]
    lam, qw = np.ascontiguousarray(orc.lam), np.ascontiguousarray(orc.qw)
    worst = lib.wd_update_cells(out.shape[0], _ptr(out), a0, _ptr(args[0]), a1, _ptr(args[1]), _ptr(args[2]), _ptr(args[3]),
                                a2, _ptr(lam), _ptr(qw), int(qw.shape[0]), its.ctypes.data_as(C.POINTER(C.c_int)))
    return out, worst


WD_CASES = [n for n, c in RC.SWE_CASES.items() if c.get("options", {}).get("use_wetting_and_drying")]


@pytest.mark.parametrize("name", WD_CASES)
@pytest.mark.parametrize("stage", [0, 1, 2])
def test_device_source_on_the_host_equals_the_oracle_newton_solve(lib, name, stage):
    """one SSPRK33 stage of every wetting-drying case (constant and P1 alpha, cells that are dry, wet and in between):
    the header's update of the plain-mass result `lin` must land on the eta the oracle obtains from
    F(eta_new) = a0 F(eta_0) + a1 F(eta_i) + beta dt R_eta"""
    case = RC.SWE_CASES[name]
    mesh = RC.build_mesh(case["mesh"])
    o = dict(case.get("options", {}))
    if isinstance(o.get("wetting_and_drying_alpha"), tuple):
        o["wetting_and_drying_alpha"] = RC.nodal_value(o["wetting_and_drying_alpha"], mesh)
    orc = O.SWEOracle(mesh, RC.nodal_value(case["bath"], mesh), options=o)
    uv0, eta0 = RC.state(mesh, 3)
    uvi, etai = RC.state(mesh, 4 + stage)
    alpha, beta = O.butcher_to_shuosher_form(O.SSPRK33_A, O.SSPRK33_B)
    a0 = float(alpha[stage + 1][0]) if stage > 0 else 0.0
    a1 = float(alpha[stage + 1][stage]) if stage > 0 else 1.0
    bdt = float(beta[stage + 1][stage]) * 3.0
    _, Re = orc.residual(uvi, etai)
    ke = np.linalg.solve(orc.mass, Re[..., None])[..., 0]
    lin = a0 * eta0 + a1 * etai + bdt * ke                                   # what the stage kernel has formed
    target = a0 * orc.displaced_mass(eta0) + a1 * orc.displaced_mass(etai) + bdt * Re
    want = orc.solve_displaced_mass(target, etai)
    got, worst = _update(lib, orc, lin, a0, eta0 if stage > 0 else None, a1, etai)
    assert 0 < worst <= 12, worst
    assert np.abs(got - want).max() <= 1e-11 * max(1.0, np.abs(want).max())
    assert np.abs(got - lin).max() > 1e-4                                      # the update is not a no-op


def test_no_wetting_drying_limit_and_deep_dry_cells(lib):
    """alpha -> 0 over deep water: f = 0, the update returns lin; dry cells (H = -3 m, alpha = 0.5 m): storage
    coefficient (1 + H / sqrt(H^2 + alpha^2)) / 2 = 0.007, the iteration still converges"""
    mesh = RC.build_mesh(RC.RECT)
    orc = O.SWEOracle(mesh, 50.0, options=dict(use_wetting_and_drying=True, wetting_and_drying_alpha=1e-9))
    _, eta = RC.state(mesh, 1)
    lin = eta + 0.01
    got, worst = _update(lib, orc, lin, 0.0, None, 1.0, eta)
    assert worst > 0 and np.abs(got - lin).max() < 1e-13
    dry = O.SWEOracle(mesh, -3.0, options=dict(use_wetting_and_drying=True, wetting_and_drying_alpha=0.5))
    target = dry.displaced_mass(eta) + 1e-2 * dry.geom.area[:, None] / 3.0
    ke = np.linalg.solve(dry.mass, (target - dry.displaced_mass(eta))[..., None])[..., 0]
    want = dry.solve_displaced_mass(target, eta)
    got, worst = _update(lib, dry, eta + ke, 0.0, None, 1.0, eta)
    assert worst > 0, worst
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    assert np.abs(want - eta).max() > 0.5            # a 1 cm layer of water raises a dry cell's eta by most of a metre


def test_thacker_basin_through_the_device_source(lib):
    """the whole Thacker period (test/swe2d/test_thacker.py) with the elevation of every stage updated by the header's
    function instead of the oracle's Newton solve: 432 steps x 3 stages over cells from 50 m deep to 80 m dry with
    alpha up to 44 m -- every call converges, the run equals the oracle's DisplacedMassShuOsherStepper and meets the
    reference's BackwardEuler threshold"""
    import kat_setups as K
    p = K.thacker_problem(10)
    orc = O.SWEOracle(p["mesh"], p["bath"], options=dict(use_wetting_and_drying=True, wetting_and_drying_alpha=p["alpha"]))
    alpha, beta = O.butcher_to_shuosher_form(O.SSPRK33_A, O.SSPRK33_B)
    dt = 100.0
    eta = p["eta0"].copy()
    uv = np.zeros(eta.shape + (2,))
    weta, wuv = eta.copy(), uv.copy()
    want = O.DisplacedMassShuOsherStepper(orc, [wuv, weta], dt)
    worst_all = 0
    with np.errstate(all="ignore"):
        for step in range(int(round(K.THACKER["t_end"] / dt))):
            uvs, etas = [uv.copy()], [eta.copy()]
            for i in range(3):
                Ru, Re = orc.residual(uvs[i], etas[i])
                ku, ke = orc.solve_mass(dt * Ru, dt * Re)
                a0 = float(alpha[i + 1][0]) if i > 0 else 0.0
                a1 = float(alpha[i + 1][i]) if i > 0 else 1.0
                b = float(beta[i + 1][i])
                lin = a0 * etas[0] + a1 * etas[i] + b * ke
                new_eta, worst = _update(lib, orc, lin, a0, etas[0] if i > 0 else None, a1, etas[i])
                assert worst > 0, (step, i, worst)
                worst_all = max(worst_all, worst)
                uvs.append(a0 * uvs[0] + a1 * uvs[i] + b * ku)
                etas.append(new_eta)
            uv, eta = uvs[3], etas[3]
            want.advance(step * dt)
            if step % 48 == 0:
                assert np.abs(eta - weta).max() <= 1e-8 * np.abs(weta).max(), step
    assert worst_all <= 15, worst_all
    assert np.abs(eta - weta).max() <= 1e-7 * np.abs(weta).max()
    err = K.thacker_error(p, eta)
    assert err < K.THACKER["max_err"][(10, "BackwardEuler")], err


def test_update_finds_the_known_root_on_random_cells(lib):
    """20 000 random cells per alpha kind -- depths from 60 m of water to 60 m above it, alpha from 0.05 to 50 m (constant
    or per node), random Shu-Osher weights -- with targets manufactured from a known answer (eta_true up to 5 m away from
    the start, across the wet-dry kink): the header's damped Newton iteration returns eta_true, satisfies the functional
    equation to rounding, never reports failure and needs at most 40 residual evaluations.  A target below what an empty
    cell holds has no solution: the function then reports -1 and still returns finite numbers."""
    rng = np.random.default_rng(2024)
    n = 20000
    lam, qw = O.cell_quadrature()
    lam, qw = np.ascontiguousarray(lam), np.ascontiguousarray(qw)
    M = (np.ones((3, 3)) + np.eye(3)) / 12.0
    Minv = np.linalg.inv(M)

    def s_of(eta, b, al2):
        H = (b + eta) @ lam.T
        return np.einsum("q,cq,qa->ca", qw, 0.5 * (np.sqrt(H * H + al2) - H), lam)

    def call(lin, a0, eta0, a1, etai, b, al, a2):
        out = np.ascontiguousarray(lin).copy()
        its = np.zeros(out.shape[0], dtype=np.int32)
        args = [np.ascontiguousarray(x) if x is not None else None for x in (eta0, etai, b, al)]
        lib.wd_update_cells(out.shape[0], _ptr(out), a0, _ptr(args[0]), a1, _ptr(args[1]), _ptr(args[2]), _ptr(args[3]),
                            a2, _ptr(lam), _ptr(qw), int(qw.shape[0]), its.ctypes.data_as(C.POINTER(C.c_int)))
        return out, its

    for var_alpha in (False, True):
        b = rng.uniform(-60.0, 60.0, (n, 1)) + rng.uniform(-3.0, 3.0, (n, 3))
        eta0 = rng.uniform(-2.0, 2.0, (n, 3))
        etai = eta0 + rng.uniform(-0.3, 0.3, (n, 3))
        a0 = float(rng.uniform(0.0, 1.0))
        a1 = 1.0 - a0
        if var_alpha:
            al = 10.0 ** rng.uniform(-1.3, 1.7, (n, 3))
            al2 = (al @ lam.T) ** 2
            a2 = 0.0
        else:
            al = None
            a2 = float(10.0 ** rng.uniform(-1.3, 1.7)) ** 2
            al2 = a2
        eta_true = etai + rng.uniform(-1.0, 1.0, (n, 1)) * rng.choice([0.01, 0.3, 5.0], (n, 1)) + rng.uniform(-0.2, 0.2, (n, 3))
        rhs = eta_true @ M + s_of(eta_true, b, al2)
        lin = (rhs - a0 * s_of(eta0, b, al2) - a1 * s_of(etai, b, al2)) @ Minv
        out, its = call(lin, a0, eta0, a1, etai, b, al, a2)
        assert its.min() > 0 and its.max() <= 40, (its.min(), its.max())
        lhs = out @ M + s_of(out, b, al2)
        scale = np.abs(rhs).max(axis=1, keepdims=True) + 1.0
        assert (np.abs(lhs - rhs) / scale).max() < 1e-12
        # the root is unique; where the cell is not almost dry it is also well conditioned
        H_true = (b + eta_true).mean(axis=1)
        well = H_true > -0.5 * np.sqrt(np.max(al2) if var_alpha else a2)
        assert np.abs(out - eta_true)[well].max() < 1e-8
        assert np.median(its) <= 6
    # no solution: more water asked out of dry cells than they hold
    b = np.full((50, 3), -20.0)
    eta = np.zeros((50, 3))
    out, its = call(eta - 0.5, 0.0, None, 1.0, eta, b, None, 0.25)
    assert (its == -1).all() and np.isfinite(out).all()


