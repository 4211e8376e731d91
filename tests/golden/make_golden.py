"""Generates tests/golden/reference_known_answers.npz -- the known answers the reference's own
test-suite holds for the runtime-quadrature path, restated as closed forms.

The reference's tests compare customquad against dolfinx's standard assembly, which cannot
run here (dolfinx/basix/ffcx/PETSc absent, no network), so the right-hand sides are computed
from what that standard assembly evaluates exactly on axis-aligned Q1 meshes:

  * test/test_linear_basis.py:13-66 (+ test/common.py:39-108): integral f_h^2, integral f_h v,
    integral u v with f_h the nodal Q1 interpolant of fcn1..fcn4 -- all exact polynomial
    integrals: M = kron of 1-D Q1 mass matrices, b = M f, m = f^T M f.
  * test/test_linear_basis.py:69-115 (test_edges): (2x+y)(midpoint) * edge length.
  * test/test_linear_basis.py:118-212 (test_edge_integral): integral v ds = L/2 on the two dofs
    of the side; :215-333 (test_face_integral): area/4 on the four dofs of the face.
  * test/test_volume_area.py:10-74 (test_volume): count * cell volume for the 28 "cut" and
    36 "outside" cells of the 10x10 mesh of test/common.py:111-149.
  * test/test_volume_area.py:92-146 (test_full_area): perimeter 2*(Lx+Ly).

This script does NOT use the oracle or the CUDA path: it is an independent statement of the
expected values.  Run:  python tests/golden/make_golden.py
"""
import itertools
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from customquad_b200 import fem, mesh as cqmesh  # host data model only  # noqa: E402


def fcn1(x):
    return x[0] ** 0


def fcn2(x):
    return x[0]


def fcn3(x):
    return np.sin(np.pi * x[0]) * np.sin(np.pi * x[1])


def fcn4(x):
    return np.exp(x[0] * x[1])


FCNS = [fcn1, fcn2, fcn3, fcn4]


def mass_1d(n, length):
    h = length / n
    M = np.zeros((n + 1, n + 1))
    for e in range(n):
        M[e : e + 2, e : e + 2] += h * np.array([[1 / 3, 1 / 6], [1 / 6, 1 / 3]])
    return M


def global_mass(N, xmin, xmax):
    L = np.asarray(xmax, float) - np.asarray(xmin, float)
    Ms = [mass_1d(n, l) for n, l in zip(N, L)]
    if len(N) == 2:  # dof id = I*(ny+1) + J
        return np.kron(Ms[0], Ms[1])
    return np.kron(Ms[2], np.kron(Ms[1], Ms[0]))  # dof id = (K*ly + J)*lx + I


def main():
    out = {}
    # ---- test_quads_assembly / test_hexes_assembly -------------------------------------
    quad_cases = itertools.product([[1, 1], [3, 1], [1, 3], [3, 4]], [[0, 0], [-0.25, -10.25]], [[1, 1], [1.25, 17.5]])
    hex_cases = itertools.product(
        [[1, 1, 1], [3, 1, 1], [1, 3, 1], [1, 1, 3], [3, 4, 5]], [[0, 0, 0], [-0.25, -10.25, 0.25]], [[1, 1, 1], [1.25, 17.5, 4.4]]
    )
    idx = 0
    for N, xmin, xmax in list(quad_cases) + list(hex_cases):
        m = (cqmesh.create_rectangle if len(N) == 2 else cqmesh.create_box)([xmin, xmax], N)
        V = fem.FunctionSpaceQ(m, ("Lagrange", 1))
        X = V.tabulate_dof_coordinates().T
        M = global_mass(N, xmin, xmax)
        key = f"assembly/{idx}"
        out[key + "/N"] = np.array(N)
        out[key + "/xmin"] = np.array(xmin, float)
        out[key + "/xmax"] = np.array(xmax, float)
        out[key + "/matrix"] = M
        for k, f in enumerate(FCNS):
            fv = np.asarray(f(X), float) * np.ones(X.shape[1])
            out[key + f"/fcn{k + 1}/vector"] = M @ fv
            out[key + f"/fcn{k + 1}/scalar"] = np.array(fv @ M @ fv)
        idx += 1
    out["assembly/count"] = np.array(idx)

    # ---- test_edges: integral of 2x+y over the four edges of [-1,4]x[-1,1] ----------------
    xmin, xmax = np.array([-1.0, -1.0]), np.array([4.0, 1.0])
    mids = np.array([[1.5, -1.0], [1.5, 1.0], [-1.0, 0.0], [4.0, 0.0]])  # bottom, top, left, right
    lens = np.array([5.0, 5.0, 2.0, 2.0])
    out["edges/xmin"], out["edges/xmax"] = xmin, xmax
    out["edges/midpoints"], out["edges/lengths"] = mids, lens
    out["edges/exact"] = (2 * mids[:, 0] + mids[:, 1]) * lens

    # ---- test_edge_integral: integral v ds on one quad -----------------------------------------
    xmin, xmax = np.array([-0.25, -10.25]), np.array([1.25, 17.5])
    m = cqmesh.create_rectangle([xmin, xmax], [1, 1])
    V = fem.FunctionSpaceQ(m, ("Lagrange", 1))
    X = V.tabulate_dof_coordinates()
    d = xmax - xmin
    pts_ref = np.array([[0.5, 0.0], [0.5, 1.0], [0.0, 0.5], [1.0, 0.5]])  # test_linear_basis.py:181
    sides = [(1, xmin[1], d[0]), (1, xmax[1], d[0]), (0, xmin[0], d[1]), (0, xmax[0], d[1])]
    b = np.zeros((4, 4))
    for k, (axis, val, length) in enumerate(sides):
        b[k, np.isclose(X[:, axis], val)] = length / 2
    out["edge_integral/xmin"], out["edge_integral/xmax"] = xmin, xmax
    out["edge_integral/qr_pts_ref"] = pts_ref
    out["edge_integral/qr_w"] = np.array([d[0], d[0], d[1], d[1]]) / np.prod(d)
    out["edge_integral/exact"] = b

    # ---- test_face_integral: integral v ds on one hex ---------------------------------------------
    bbmin, bbmax = np.array([-0.25, -10.25, -4.4]), np.array([1.25, 17.5, 0.1])
    m = cqmesh.create_box([bbmin, bbmax], [1, 1, 1])
    V = fem.FunctionSpaceQ(m, ("Lagrange", 1))
    X = V.tabulate_dof_coordinates()
    d = bbmax - bbmin
    pts_ref = np.array([[0, 0.5, 0.5], [1, 0.5, 0.5], [0.5, 0, 0.5], [0.5, 1, 0.5], [0.5, 0.5, 0], [0.5, 0.5, 1.0]])  # :290-299
    fsize = np.array([d[1] * d[2]] * 2 + [d[0] * d[2]] * 2 + [d[0] * d[1]] * 2)
    faces = [(0, bbmin[0]), (0, bbmax[0]), (1, bbmin[1]), (1, bbmax[1]), (2, bbmin[2]), (2, bbmax[2])]
    b = np.zeros((6, 8))
    for k, (axis, val) in enumerate(faces):
        b[k, np.isclose(X[:, axis], val)] = fsize[k] / 4
    out["face_integral/bbmin"], out["face_integral/bbmax"] = bbmin, bbmax
    out["face_integral/qr_pts_ref"] = pts_ref
    out["face_integral/qr_w"] = fsize / np.prod(d)
    out["face_integral/exact"] = b

    # ---- test_volume / test_full_area: 10x10 mesh of test/common.py:111-149 -------------------
    N = 10
    xmin, xmax = np.array([-1.23, -11.11]), np.array([3.33, 99.99])
    h = (xmax - xmin) / N
    out["volume/N"], out["volume/xmin"], out["volume/xmax"] = np.array(N), xmin, xmax
    out["volume/num_outside"], out["volume/num_cut"], out["volume/num_uncut"] = np.array(4 * (N - 1)), np.array(4 * (N - 3)), np.array((N - 4) ** 2)
    out["volume/cut_exact"] = np.array(4 * (N - 3) * h[0] * h[1])
    out["volume/outside_exact"] = np.array(4 * (N - 1) * h[0] * h[1])
    out["volume/total_exact"] = np.array(np.prod(xmax - xmin))
    out["area/exact"] = np.array(2 * np.sum(xmax - xmin))

    path = os.path.join(os.path.dirname(__file__), "reference_known_answers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
