"""Accuracy of the stage kernels' branch-free fp64 helpers (MUFU seed + one third-order correction): <= 2 ulp."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fp64_helpers_are_accurate_to_two_ulp():
    import torch
    from thetis_b200.engine import Engine
    from thetis_b200.mesh import rectangle_mesh
    eng = Engine(rectangle_mesh(2, 2, 1.0, 1.0))
    rng = np.random.default_rng(0)
    # depths / lengths^2 / areas seen on the path: 1e-6 .. 1e12, plus values hugging powers of two
    x = np.concatenate([10.0 ** rng.uniform(-6, 12, 200000), 2.0 ** rng.integers(-20, 40, 2000) * (1 + rng.uniform(-1e-6, 1e-6, 2000)),
                        np.array([0.05, 0.25, 0.5, 1.0, 2.0, 9.81, 20.0, 200.0, 1000.0])])
    xd = torch.as_tensor(x).cuda()
    out = torch.empty(4 * x.size, dtype=torch.float64, device="cuda")
    eng._ck(eng.lib.tb_selftest_math(eng.ctx, C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()), x.size, eng.stream))
    o = out.cpu().numpy().reshape(4, -1)
    xl = x.astype(np.longdouble)
    ref = [1 / np.sqrt(xl), np.sqrt(xl), 1 / xl, xl ** (-np.longdouble(1) / 3)]
    for name, got, r in zip(("rsqrt", "sqrt", "rcp", "rcbrt"), o, ref):
        rel = np.abs((got.astype(np.longdouble) - r) / r).max()
        assert rel < 2 * 2.0 ** -52, (name, float(rel))
