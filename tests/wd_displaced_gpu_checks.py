"""
GPU checks of TB_OPT_WD_DISPLACED_MASS, run in a process of their own by tests/test_gpu_wd_displaced_mass.py:

    python tests/wd_displaced_gpu_checks.py stage <case> | refuse | thacker

Exit code 0 = the check passed.  Through the C-ABI (thetis_b200.engine.Engine); the oracle is the checker only.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def _engine(mesh, bath_v, alpha, displaced=True):
    import thetis_b200._lib as L
    from thetis_b200.engine import Engine
    eng = Engine(mesh)
    eng.set_option(L.OPT_NONLINEAR, 1)
    eng.set_option(L.OPT_WETTING_DRYING, 1)
    if isinstance(alpha, np.ndarray):
        eng.set_field(L.F_WD_ALPHA, alpha)                 # P1 field, per vertex
    else:
        eng.set_option(L.OPT_WD_ALPHA, alpha)
    eng.set_field(L.F_BATHYMETRY, bath_v)
    eng.set_option(L.OPT_WD_DISPLACED_MASS, 1 if displaced else 0)
    return eng, L


def _ssprk33(eng, bufs, dt, alpha, beta):
    """one SSPRK33 step with the buffer rotation of thetis_b200.rungekutta._launch_stage; the solution stays in bufs[0]"""
    A, B, Cb = bufs
    eng.swe_stage(0.0, float(alpha[1][0]), float(beta[1][0]) * dt, A, None, B)
    eng.swe_stage(float(alpha[2][0]), float(alpha[2][1]), float(beta[2][1]) * dt, B, A, Cb)
    eng.swe_stage(float(alpha[3][0]), float(alpha[3][2]), float(beta[3][2]) * dt, Cb, A, A)


def check_stage(name):
    """three SSPRK33 steps of a wetting-drying case of tests/reference_cases.py (closed boundaries, no other terms):
    kernel with the displaced-mass epilogue vs oracle.DisplacedMassShuOsherStepper, 1e-10"""
    import reference_cases as RC
    from oracle import swe_oracle as O
    case = RC.SWE_CASES[name]
    mesh = RC.build_mesh(case["mesh"])
    o = case["options"]
    al = o["wetting_and_drying_alpha"]
    bath_v = np.asarray(RC.FUNCS[case["bath"][1]](mesh.coords[:, 0], mesh.coords[:, 1]), dtype=float)
    if isinstance(al, tuple):
        al_v = np.asarray(RC.FUNCS[al[1]](mesh.coords[:, 0], mesh.coords[:, 1]), dtype=float)
        al_o = al_v[mesh.cells]
    else:
        al_v = al_o = float(al)
    orc = O.SWEOracle(mesh, bath_v[mesh.cells], options=dict(use_wetting_and_drying=True, wetting_and_drying_alpha=al_o))
    uv, eta = RC.state(mesh, 11)
    eng, L = _engine(mesh, bath_v, al_v)
    bufs = [eng.upload_nodal(uv, eta), eng.new_state(), eng.new_state()]
    alpha, beta = O.butcher_to_shuosher_form(O.SSPRK33_A, O.SSPRK33_B)
    wuv, weta = uv.copy(), eta.copy()
    st = O.DisplacedMassShuOsherStepper(orc, [wuv, weta], 2.0)
    puv, peta = uv.copy(), eta.copy()
    pst = O.ShuOsherStepper(orc, [puv, peta], 2.0)
    for i in range(3):
        _ssprk33(eng, bufs, 2.0, alpha, beta)
        st.advance(2.0 * i)
        pst.advance(2.0 * i)
    gu, ge = eng.download_nodal(bufs[0])
    eu = np.abs(gu - wuv).max() / np.abs(wuv).max()
    ee = np.abs(ge - weta).max() / np.abs(weta).max()
    print(f"{name}: rel err uv {eu:.2e} eta {ee:.2e}; displaced vs plain {np.abs(weta - peta).max():.2e}")
    assert eu < 1e-10 and ee < 1e-10, (eu, ee)
    assert np.abs(ge - peta).max() > 1e-5            # it is not the plain-mass step
    # option off: the same launches give the plain-mass step again
    eng.set_option(L.OPT_WD_DISPLACED_MASS, 0)
    bufs[0] = eng.upload_nodal(uv, eta)
    for i in range(3):
        _ssprk33(eng, bufs, 2.0, alpha, beta)
    gu, ge = eng.download_nodal(bufs[0])
    assert np.abs(ge - peta).max() < 1e-10 * np.abs(peta).max()


def check_refuse():
    """a tendency evaluation (a0 = a1 = 0) cannot advance a mass functional: the library must refuse it"""
    import reference_cases as RC
    import thetis_b200._lib as L
    mesh = RC.build_mesh(RC.RECT)
    eng, _ = _engine(mesh, np.full(mesh.n_vertices, 2.0), 0.4)
    uv, eta = RC.state(mesh, 1)
    st = eng.upload_nodal(uv, eta)
    try:
        eng.swe_tendency(st, eng.new_state())
    except L.TbError as e:
        assert "Shu-Osher" in str(e), str(e)
        return
    raise AssertionError("tb_swe_tendency accepted TB_OPT_WD_DISPLACED_MASS")


def check_thacker(quick=False):
    """test/swe2d/test_thacker.py on the GPU.  (1) automatic alpha with the 2 m cap lifted (see
    tests/test_oracle_reference_kat.py): the displaced-mass step meets the BackwardEuler threshold of the 10 x 10 mesh at
    dt = 100 s, the plain-mass step misses it.  (2) unless ``quick``: the reference test's OWN alpha (automatic, capped
    at the default 2 m) with the displaced-mass step at dt = 2 s -- 21 600 steps, far too slow for the numpy oracle
    inside a test suite (335 s, measured once: 0.2397), a second on the GPU -- must meet the threshold the reference
    sets for its second-order implicit integrators on that mesh (0.26)."""
    import torch
    import kat_setups as K
    from oracle import swe_oracle as O
    alpha, beta = O.butcher_to_shuosher_form(O.SSPRK33_A, O.SSPRK33_B)

    def run(alpha_max, displaced, dt):
        p = K.thacker_problem(10, alpha_max)
        mesh = p["mesh"]
        bath_v = np.zeros(mesh.n_vertices)
        bath_v[mesh.cells.reshape(-1)] = p["bath"].reshape(-1)
        al_v = np.zeros(mesh.n_vertices)
        al_v[mesh.cells.reshape(-1)] = p["alpha"].reshape(-1)
        eng, _ = _engine(mesh, bath_v, al_v, displaced)
        bufs = [eng.upload_nodal(np.zeros(p["eta0"].shape + (2,)), p["eta0"]), eng.new_state(), eng.new_state()]
        for _ in range(int(round(K.THACKER["t_end"] / dt))):
            _ssprk33(eng, bufs, dt, alpha, beta)
        torch.cuda.synchronize()
        _, ge = eng.download_nodal(bufs[0])
        assert np.isfinite(ge).all()
        return K.thacker_error(p, ge)

    e_disp, e_plain = run(None, True, 100.0), run(None, False, 300.0)
    print("thacker, cap lifted: masked L2 error / l_mesh: displaced mass %.3f, plain mass %.3f" % (e_disp, e_plain))
    assert e_disp < K.THACKER["max_err"][(10, "BackwardEuler")], e_disp
    assert e_plain > 1.0, e_plain
    if not quick:
        e_ref = run(2.0, True, 2.0)
        print("thacker, the reference's own alpha (cap 2 m), displaced mass, dt = 2 s: %.4f" % e_ref)
        assert e_ref < K.THACKER["max_err"][(10, "other")], e_ref


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "stage":
        check_stage(sys.argv[2])
    elif what == "refuse":
        check_refuse()
    elif what == "thacker":
        check_thacker()
    else:
        raise SystemExit(f"unknown check {what!r}")
    print("ok")
