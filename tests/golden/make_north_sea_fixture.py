"""
Derives tests/golden/north_sea_mesh.npz from the reference's demos/north_sea.msh
(6565 nodes, 10 920 triangles, boundary tags 100 = open ocean / 200 = coast) so
that the GPU box -- where /root/reference does not exist -- can build the
~4 M-triangle BASELINE config 5 mesh by k-section refinement.  Arrays only
(float64 coords, int32 cells CCW, boundary edges + tags); run here:
    python tests/golden/make_north_sea_fixture.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from thetis_b200.mesh import read_gmsh, FACET_NODES  # noqa: E402

src = "/root/reference/demos/north_sea.msh"
m = read_gmsh(src)
edges = np.stack([m.cells[m.bf_cell, FACET_NODES[m.bf_lf, 0]], m.cells[m.bf_cell, FACET_NODES[m.bf_lf, 1]]], 1)
out = os.path.join(os.path.dirname(__file__), "north_sea_mesh.npz")
np.savez_compressed(out, coords=m.coords, cells=m.cells.astype(np.int32), bnd_edges=edges.astype(np.int32),
                    bnd_tags=m.bf_marker.astype(np.int32))
print(out, m.n_cells, m.n_vertices, m.n_bfacets, os.path.getsize(out))
