"""
Writes tests/golden/mini_tagged.msh: a small Gmsh-2.2 ASCII mesh with tagged
boundary lines (100 = open, 200 = coast), the format of the reference's
demos/north_sea.msh, for reader / boundary-array tests.
    python tests/golden/make_mini_msh.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from thetis_b200.mesh import delaunay_mesh, FACET_NODES  # noqa: E402

m = delaunay_mesh(180, 50.0, 40.0, seed=7)
out = os.path.join(os.path.dirname(__file__), "mini_tagged.msh")
with open(out, "w") as fh:
    fh.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % m.n_vertices)
    for i, (x, y) in enumerate(m.coords):
        fh.write("%d %.17g %.17g 0\n" % (i + 1, x, y))
    fh.write("$EndNodes\n$Elements\n%d\n" % (m.n_bfacets + m.n_cells))
    k = 1
    for c, lf, mk in zip(m.bf_cell, m.bf_lf, m.bf_marker):
        a, b = m.cells[c, FACET_NODES[lf]]
        tag = 100 if mk == 1 else 200
        fh.write("%d 1 2 %d %d %d %d\n" % (k, tag, tag, a + 1, b + 1))
        k += 1
    for t in m.cells:
        fh.write("%d 2 2 1 1 %d %d %d\n" % (k, t[0] + 1, t[1] + 1, t[2] + 1))
        k += 1
    fh.write("$EndElements\n")
print(out, m.n_cells, m.n_bfacets)
