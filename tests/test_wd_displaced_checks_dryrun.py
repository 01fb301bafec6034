"""
CPU dry run of the GPU check script tests/wd_displaced_gpu_checks.py: the engine is replaced by the oracle-backed test
double (tests/oracle_engine.py), so that the LOGIC of the checks (set-up, buffer rotation of the three SSPRK33 stages,
thresholds) is known to be right before they ever see hardware; what stays unverified without a GPU is the kernel path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import wd_displaced_gpu_checks as W                      # noqa: E402
from oracle_engine import OracleEngine                   # noqa: E402
from thetis_b200 import _lib as L                        # noqa: E402


class _Engine(OracleEngine):
    """OracleEngine + the test conveniences of thetis_b200.engine.Engine used by the checks"""

    def upload_nodal(self, uv, eta, state=None):
        state = self.new_state() if state is None else state
        r = self._rec(state, 9)
        r[:, :6] = np.asarray(uv).reshape(-1, 6)
        r[:, 6:] = np.asarray(eta)
        return state

    def download_nodal(self, state):
        r = self._rec(state, 9)
        return r[:, :6].reshape(-1, 3, 2).copy(), r[:, 6:].copy()

    def swe_stage(self, a0, a1, bdt, src, u0, dst):
        # u0 may alias dst in the library (each CTA reads its u0 patch before writing): keep that contract here
        super().swe_stage(a0, a1, bdt, src, None if u0 is None else u0.clone(), dst)

    def swe_tendency(self, u, k):
        if self.opt.get(L.OPT_WD_DISPLACED_MASS):
            raise L.TbError("TB_OPT_WD_DISPLACED_MASS ... needs a Shu-Osher stage (a0 + a1 = 1)")
        raise AssertionError("not used")


def _patched(monkeypatch):
    def _engine(mesh, bath_v, alpha, displaced=True):
        eng = _Engine(mesh)
        eng.set_option(L.OPT_NONLINEAR, 1)
        eng.set_option(L.OPT_WETTING_DRYING, 1)
        if isinstance(alpha, np.ndarray):
            eng.set_field(L.F_WD_ALPHA, alpha)
        else:
            eng.set_option(L.OPT_WD_ALPHA, alpha)
        eng.set_field(L.F_BATHYMETRY, bath_v)
        eng.set_option(L.OPT_WD_DISPLACED_MASS, 1 if displaced else 0)
        return eng, L
    monkeypatch.setattr(W, "_engine", _engine)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)


def test_stage_check_logic(monkeypatch):
    _patched(monkeypatch)
    W.check_stage("wetting_drying_alpha_p1")
    W.check_stage("wetting_drying_manning")


def test_refuse_check_logic(monkeypatch):
    _patched(monkeypatch)
    W.check_refuse()


def test_thacker_check_logic(monkeypatch):
    _patched(monkeypatch)
    W.check_thacker(quick=True)          # the capped-alpha leg takes the numpy oracle 5 minutes: GPU only
