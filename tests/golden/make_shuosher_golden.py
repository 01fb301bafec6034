"""
Generates tests/golden/shuosher_ssprk33.json by EXECUTING the reference's own
`butcher_to_shuosher_form` (thetis/rungekutta.py:13-87) on the reference's own
SSPRK33 tableau (thetis/rungekutta.py:342-347).  `import thetis` fails here
(needs Firedrake), so the two pure-numpy pieces are lifted with `ast` and run
in isolation -- nothing is copied into the repo but the resulting numbers.
    python tests/golden/make_shuosher_golden.py
"""
import ast
import json
import os

import numpy

SRC = "/root/reference/thetis/rungekutta.py"
tree = ast.parse(open(SRC).read())
ns = {"numpy": numpy}
tableaux = {}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name == "butcher_to_shuosher_form":
        exec(compile(ast.Module(body=[node], type_ignores=[]), SRC, "exec"), ns)
    if isinstance(node, ast.ClassDef) and node.name.endswith("Abstract"):
        vals = {}
        for st in node.body:
            if isinstance(st, ast.Assign) and st.targets[0].id in ("a", "b", "c", "cfl_coeff"):
                try:
                    vals[st.targets[0].id] = eval(compile(ast.Expression(st.value), SRC, "eval"), {"numpy": numpy})
                except Exception:
                    pass
        if {"a", "b", "c"} <= set(vals):
            tableaux[node.name] = vals
out = {}
for name in ("SSPRK33Abstract", "ForwardEulerAbstract"):
    t = tableaux[name]
    a = numpy.array(t["a"], dtype=float)
    b = numpy.array(t["b"], dtype=float)
    alpha, beta = ns["butcher_to_shuosher_form"](a, b)
    out[name] = {"a": a.tolist(), "b": b.tolist(), "c": list(map(float, t["c"])), "cfl_coeff": float(t["cfl_coeff"]),
                 "alpha": alpha.tolist(), "beta": beta.tolist()}
# explicit Butcher-form tableaux stepped by ERKGeneric (thetis/rungekutta.py:350-392, 959-980): the numbers only
for name in ("ERKLSPUM2Abstract", "ERKLPUM2Abstract", "ERKMidpointAbstract"):
    t = tableaux[name]
    out[name] = {"a": numpy.array(t["a"], dtype=float).tolist(), "b": list(map(float, t["b"])),
                 "c": list(map(float, t["c"])), "cfl_coeff": float(t["cfl_coeff"])}
path = os.path.join(os.path.dirname(__file__), "shuosher_ssprk33.json")
json.dump(out, open(path, "w"), indent=1)
print(path)
print(json.dumps(out["SSPRK33Abstract"], indent=1))
