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
