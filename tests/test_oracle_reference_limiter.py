"""
Thetis' own part of `VertexBasedP1DGLimiter.compute_bounds` (limiter.py:109-145) pinned to the reference's kernel text:
tests/golden/reference_limiter_bounds.npz holds the vertex bounds after the reference's `my_kernel` -- read out of the
reference file, compiled with gcc and run over the exterior facets like `op2.par_loop` does
(tests/golden/make_reference_limiter_golden.py) -- and the oracle's `limiter_boundary_bounds` must reproduce them bit for
bit (max / min of identical numbers).  The CUDA limiter is tied to the same oracle by the -m gpu tests.
Firedrake's `VertexBasedLimiter` (centroid bounds, limiting) is not in the reference tree and stays recalled.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import reference_cases as RC                                       # noqa: E402
from oracle.swe_oracle import limiter_boundary_bounds, vertex_based_limiter   # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "reference_limiter_bounds.npz"))
CASES = {"rect_5x5": ("rect", 5, 5, 1.0, 1.0), "delaunay_40": ("delaunay", 40, 5.0e3, 4.0e3, 2),
         "periodic_6x4": ("periodic", 6, 4, 6.0e3, 4.0e3)}


@pytest.mark.parametrize("name", list(CASES))
def test_boundary_bounds_equal_the_reference_kernel(name):
    mesh = RC.build_mesh(CASES[name])
    q = GOLD[f"{name}/q"]
    qmax, qmin = GOLD[f"{name}/qmax0"].copy(), GOLD[f"{name}/qmin0"].copy()
    limiter_boundary_bounds(mesh, q, qmax, qmin)
    assert np.array_equal(qmax, GOLD[f"{name}/qmax"]) and np.array_equal(qmin, GOLD[f"{name}/qmin"])
    assert (qmax != GOLD[f"{name}/qmax0"]).sum() >= 10            # the kernel really raised bounds
    # only vertices on the boundary are touched
    tv = mesh.topo[mesh.cells]
    from thetis_b200.mesh import FACET_NODES
    on_bnd = np.unique(tv[mesh.bf_cell[:, None], FACET_NODES[mesh.bf_lf]])
    touched = np.nonzero((qmax != GOLD[f"{name}/qmax0"]) | (qmin != GOLD[f"{name}/qmin0"]))[0]
    assert set(touched) <= set(on_bnd)


def test_the_limiter_uses_that_step():
    """vertex_based_limiter = centroid bounds + limiter_boundary_bounds + limiting: without the boundary step a field
    that is linear up to the wall would be clipped there (test_slopelimiter.py:50-52 is the reference's criterion)"""
    mesh = RC.build_mesh(CASES["rect_5x5"])
    x = mesh.coords[mesh.cells]
    q = 2.0 * x[..., 0] + 0.5
    assert np.abs(vertex_based_limiter(mesh, q) - q).max() < 1e-12


@pytest.mark.skipif(not os.path.isfile("/root/reference/thetis/limiter.py"), reason="the reference tree only exists in the build container")
def test_committed_fixture_is_what_the_reference_kernel_produces(tmp_path):
    out = tmp_path / "regen.npz"
    gen = os.path.join(HERE, "golden", "make_reference_limiter_golden.py")
    r = subprocess.run([sys.executable, gen, "--out", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    new = np.load(out)
    assert set(new.files) == set(GOLD.files)
    for k in GOLD.files:
        assert np.array_equal(new[k], GOLD[k]), k


@pytest.mark.skipif(not os.path.isfile("/root/reference/thetis/utility.py"), reason="the reference tree only exists in the build container")
def test_cell_widths_of_the_automatic_wetting_drying_alpha_equal_the_reference_kernel(tmp_path):
    """`get_cell_widths_2d` (utility.py:716-739; feeds `use_automatic_wetting_and_drying_alpha`, solver2d.py:279-287, which
    the Thacker set-up of tests/kat_setups.py restates as the coordinate ranges of a cell): the reference's kernel text,
    compiled and run cell by cell with access MAX on a DG0 vector initialised to the smallest double, as the reference does"""
    import ast
    import ctypes as C
    tree = ast.parse(open("/root/reference/thetis/utility.py").read())
    src = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "get_cell_widths_2d":
            for sub in ast.walk(node):
                if isinstance(sub, ast.Constant) and isinstance(sub.value, str) and "cell_width_kernel" in sub.value:
                    src = sub.value % {"nodes": 3}                # arity of the P1 coordinate cell_node_map
    assert src is not None
    (tmp_path / "k.c").write_text("#include <math.h>\n" + src + "\n")
    subprocess.run(["gcc", "-O1", "-fPIC", "-shared", "-o", str(tmp_path / "k.so"), str(tmp_path / "k.c"), "-lm"], check=True)
    lib = C.CDLL(str(tmp_path / "k.so"))
    dp = C.POINTER(C.c_double)
    lib.cell_width_kernel.argtypes = [dp, dp]
    import kat_setups as K
    p = K.thacker_problem(10)
    mesh = p["mesh"]
    xc = np.ascontiguousarray(mesh.coords[mesh.cells])               # (nt, 3, 2): the gathered coordinates of a cell
    widths = np.full((mesh.n_cells, 2), np.finfo(0.0).min)
    for c in range(mesh.n_cells):
        w = widths[c].copy()
        lib.cell_width_kernel(xc[c].ctypes.data_as(dp), w.ctypes.data_as(dp))
        widths[c] = np.maximum(widths[c], w)
    assert np.array_equal(widths, np.ptp(xc, axis=1))
