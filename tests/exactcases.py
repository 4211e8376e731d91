"""Closed-form element tensors (tests/golden/exact_element_tensors.npz, generated with sympy by
tests/golden/make_exact_tensors.py) against an assembly backend: the mathematical pin for what no passing
reference test covers (degree 2, stiffness, Nitsche with quadrature normals, CellDiameter)."""
import os

import numpy as np

import customquad_b200 as cq
from customquad_b200 import forms

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "exact_element_tensors.npz")
CASES = ["quad_q1", "quad_q2", "hex_q1", "hex_q2"]
RTOL = 1e-12


def load(name):
    z = np.load(GOLDEN)
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


def check(backend, name):
    """backend.matrix(form, qr_data) -> scipy CSR on the form's pattern."""
    c = load(name)
    tdim = len(c["x0"])
    mesh = (cq.create_rectangle if tdim == 2 else cq.create_box)([c["x0"], c["x0"] + c["h"]], [1] * tdim)
    V = cq.FunctionSpace(mesh, ("Lagrange", int(c["degree"])))
    d = V.dofmap.list[0]
    bulk = ([0], [c["bulk_pts"].reshape(-1)], [c["bulk_w"]])
    surf = ([0], [c["surf_pts"].reshape(-1)], [c["surf_w"]], [c["surf_n"].reshape(-1)])
    for kind, form, qr in (("mass", forms.mass(V), bulk), ("stiffness", forms.stiffness(V), bulk),
                           ("nitsche", forms.nitsche(V, float(c["betaN"])), surf)):  # fmt: skip
        A = backend.matrix(form, [qr]).toarray()[np.ix_(d, d)]  # global -> local dof order
        ref = c[kind]
        err = np.max(np.abs(A - ref)) / np.max(np.abs(ref))
        assert err <= RTOL, f"{name} {kind}: max rel error {err:.3e}"
