"""
Generates tests/golden/reference_limiter_bounds.npz by COMPILING AND EXECUTING THE REFERENCE'S OWN KERNEL TEXT:
the C source of `my_kernel` that `VertexBasedP1DGLimiter.compute_bounds` hands to PyOP2 (thetis/limiter.py:123-145) is
read out of the reference file with `ast` (nothing is copied into the repository), instantiated for 2 facet nodes like
the reference does for P1DG triangles, compiled with gcc and run over the exterior facets of the test meshes the way
`op2.par_loop` does: per exterior facet the kernel gets the three vertex bounds of the facet's cell (P1 max / min
fields through `exterior_facet_node_map`, access MAX / MIN), the three P1DG values of that cell (READ), the local facet
number (`exterior_facets.local_facet_dat`) and the table of facet-support dofs (`entity_support_dofs`; FInAT's table for
the P1DG triangle is recalled: the two nodes other than the one opposite the facet, ascending).

What this pins is Thetis' own part of the limiter (row 8a-10 (iii) of SURVEY.md); the centroid bounds and the limiting
itself are Firedrake's `VertexBasedLimiter` and stay recalled.

    python tests/golden/make_reference_limiter_golden.py [--out file.npz]      (build container only: needs /root/reference, gcc)
"""
import ast
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

REF_FILE = "/root/reference/thetis/limiter.py"
CASES = {"rect_5x5": ("rect", 5, 5, 1.0, 1.0), "delaunay_40": ("delaunay", 40, 5.0e3, 4.0e3, 2),
         "periodic_6x4": ("periodic", 6, 4, 6.0e3, 4.0e3)}


def reference_kernel_source():
    """the string assigned to `code` inside VertexBasedP1DGLimiter.compute_bounds, straight from the reference file"""
    tree = ast.parse(open(REF_FILE).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "compute_bounds":
            for sub in ast.walk(node):
                if isinstance(sub, ast.Constant) and isinstance(sub.value, str) and "void my_kernel" in sub.value:
                    return sub.value
    raise RuntimeError("my_kernel not found in " + REF_FILE)


def build_kernel(tmp):
    src = reference_kernel_source() % {"nnodes": 2}                 # limiter.py:137: code % {'nnodes': n_bnd_nodes}
    path = os.path.join(tmp, "my_kernel.c")
    with open(path, "w") as f:
        f.write("#include <math.h>\n" + src + "\n")
    so = os.path.join(tmp, "my_kernel.so")
    subprocess.run(["gcc", "-O1", "-fPIC", "-shared", "-o", so, path, "-lm"], check=True)
    lib = C.CDLL(so)
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint)
    lib.my_kernel.argtypes = [dp, dp, dp, up, up]
    lib.my_kernel.restype = None
    return lib


def run_case(lib, mesh, seed):
    """op2.par_loop(bnd_kernel, exterior_facets.set, max(MAX), min(MIN), field(READ), local_facet_dat, local_facet_idx)"""
    rng = np.random.default_rng(seed)
    nt = mesh.n_cells
    q = rng.standard_normal((nt, 3))
    tv = mesh.topo[mesh.cells]                                      # P1 dof of each cell node (cell_node_map of P1CG)
    nv = int(mesh.topo.max()) + 1
    qmax0 = rng.standard_normal(nv)
    qmin0 = qmax0 - np.abs(rng.standard_normal(nv))
    qmax, qmin = qmax0.copy(), qmin0.copy()
    # entity_support_dofs(P1DG.finat_element, 1): facet e (opposite vertex e) is supported by the other two nodes
    lfi = np.array([[1, 2], [0, 2], [0, 1]], dtype=np.uint32)
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint)
    for c, lf in zip(mesh.bf_cell.astype(np.int64), mesh.bf_lf.astype(np.int64)):
        mx = np.ascontiguousarray(qmax[tv[c]])                      # gathered through exterior_facet_node_map
        mn = np.ascontiguousarray(qmin[tv[c]])
        fl = np.ascontiguousarray(q[c])
        facet = np.array([lf], dtype=np.uint32)
        lib.my_kernel(mx.ctypes.data_as(dp), mn.ctypes.data_as(dp), fl.ctypes.data_as(dp), facet.ctypes.data_as(up),
                      lfi.ctypes.data_as(up))
        np.maximum.at(qmax, tv[c], mx)                              # access MAX / MIN: combined into the global dat
        np.minimum.at(qmin, tv[c], mn)
    return dict(q=q, qmax0=qmax0, qmin0=qmin0, qmax=qmax, qmin=qmin)


def main():
    import reference_cases as RC
    out = os.path.join(HERE, "reference_limiter_bounds.npz")
    if "--out" in sys.argv:
        out = sys.argv[sys.argv.index("--out") + 1]
    data = {}
    with tempfile.TemporaryDirectory() as tmp:
        lib = build_kernel(tmp)
        for i, (name, spec) in enumerate(CASES.items()):
            mesh = RC.build_mesh(spec)
            for k, v in run_case(lib, mesh, 100 + i).items():
                data[f"{name}/{k}"] = v
            print(f"{name}: {mesh.n_cells} cells, {mesh.bf_cell.shape[0]} exterior facets, "
                  f"{int((data[name + '/qmax'] != data[name + '/qmax0']).sum())} max bounds raised")
    np.savez_compressed(out, **data)
    print(out, len(data), "arrays")


if __name__ == "__main__":
    main()
