import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from thetis_b200.engine import Engine
for name in ("set_option", "set_field", "set_bc", "set_bc_array", "set_boundary_length", "set_cell_quadrature"):
    orig = getattr(Engine, name)
    def mk(orig, name):
        def f(self, *a, **k):
            desc = [(x if not isinstance(x, np.ndarray) else ("arr", x.shape, float(x.min()), float(x.max()))) for x in a]
            print("CALL", name, desc, flush=True)
            return orig(self, *a, **k)
        return f
    setattr(Engine, name, mk(orig, name))
from thetis_b200.workloads import north_sea_mesh, north_sea_setup
from thetis_b200.parallel import SingleSWE
mesh = north_sea_mesh(k=1)
setup = north_sea_setup(mesh, wetting_drying=False)
run = SingleSWE(mesh, setup, wd=False)
