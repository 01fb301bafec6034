"""
Generates tests/golden/reference_kat_fields.npz by EXECUTING the reference's own test functions

* test/swe2d/test_rossby_wave.py:23-130      asymptotic_expansion_uv / asymptotic_expansion_elev
* test/swe2d/test_steady_state_basin_mms.py:16-111   setup7 / setup8 / setup9 (analytic fields, MMS sources)

at fixed sample points.  `from thetis import *` fails here (needs Firedrake), so the two test modules are imported
with a stub `thetis` module whose UFL names are numpy functions: the expressions the reference writes in UFL then
evaluate pointwise.  Nothing is copied into the repo but the resulting numbers.
    python tests/golden/make_reference_kat_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy

REF = "/root/reference/test/swe2d"


class _Mesh:
    def __init__(self, x, y):
        self.xy = (x, y)


class _Space:
    def __init__(self, x, y):
        self._m = _Mesh(x, y)

    def mesh(self):
        return self._m


class _Function:
    def __init__(self, space):
        pass

    def interpolate(self, expr):
        return expr


def _stub():
    t = types.ModuleType("thetis")
    t.numpy = numpy
    for n in ("sqrt", "cos", "sin", "exp", "cosh", "tanh", "sign", "pi"):
        setattr(t, n, getattr(numpy, n))
    t.as_vector = lambda v: numpy.stack([numpy.asarray(c, float) + 0.0 * numpy.asarray(v[0], float) for c in v], -1)
    t.Constant = lambda v: float(v)
    t.SpatialCoordinate = lambda mesh: mesh.xy
    t.Function = _Function
    t.__all__ = [k for k in vars(t) if not k.startswith("_")]
    return t


def _load(name):
    sys.modules["thetis"] = _stub()
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    out = {}
    rng = numpy.random.default_rng(20261017)
    # ---- Rossby soliton initial condition (lx, ly = 48, 24 centred on the origin)
    ros = _load("test_rossby_wave")
    x = rng.uniform(-24.0, 24.0, 400)
    y = rng.uniform(-12.0, 12.0, 400)
    for order in (0, 1):
        for t in (0.0, 7.5):
            uv = ros.asymptotic_expansion_uv(_Space(x, y), order=order, time=t)
            el = ros.asymptotic_expansion_elev(_Space(x, y), order=order, time=t)
            out[f"rossby_o{order}_t{t}_uv"] = numpy.asarray(uv)
            out[f"rossby_o{order}_t{t}_elev"] = numpy.asarray(el)
    out["rossby_x"], out["rossby_y"] = x, y
    # ---- MMS set-ups (lx, ly, depth, f0, nu0, g as in run(), test_steady_state_basin_mms.py:117-127)
    mms = _load("test_steady_state_basin_mms")
    lx, ly, h0, f0, nu0, g = 15e3, 10e3, 10.0, 5e-3, 100.0, 9.81
    px = rng.uniform(0.0, lx, 300)
    py = rng.uniform(0.0, ly, 300)
    out["mms_x"], out["mms_y"] = px, py
    for name in ("setup7", "setup8", "setup9"):
        d = getattr(mms, name)((px, py), lx, ly, h0, f0, nu0, g)
        for k in ("bath_expr", "cori_expr", "visc_expr", "elev_expr", "uv_expr", "res_elev_expr", "res_uv_expr"):
            if k in d:
                out[f"{name}_{k}"] = numpy.asarray(d[k], float)
        # boundary tags (test_steady_state_basin_mms.py:35-39,64-68,99-103) and option overrides, as text
        out[f"{name}_bnd"] = numpy.array(repr({m: sorted(v) for m, v in d["bnd_funcs"].items()}))
        out[f"{name}_options"] = numpy.array(repr(d.get("options", {})))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kat_fields.npz")
    numpy.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
