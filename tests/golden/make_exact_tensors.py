"""Closed-form element tensors for the items NO passing reference test pins (degree 2,
gradients / stiffness, the Nitsche form with quadrature normals and CellDiameter): exact symbolic
integration with sympy, independent of oracle/ and of the CUDA kernels.

    python tests/golden/make_exact_tensors.py   ->  tests/golden/exact_element_tensors.npz

Each case is one axis-aligned cell [x0, x0 + h] (what create_rectangle / create_box produce) with
  mass       A_ij = int_K psi_i psi_j dx                                  (test/common.py:91)
  stiffness  A_ij = int_K grad psi_i . grad psi_j dx                      (demo/poisson.py:154)
  nitsche    A_ij = int_G -(n.grad psi_j) psi_i - psi_j (n.grad psi_i) + betaN / h_K psi_i psi_j dS   (demo/poisson.py:156-158)
where psi_j = prod_d l_{lat[j, d]}((x_d - x0_d) / h_d) is the tensor Lagrange basis in the local dof order
the reference's call_basix permutation produces (customquad_b200.fem.point_frame_lattice; quad degree 2 with
the corrected permutation), G is a straight segment (2-D) or a planar parallelogram (3-D) inside the cell with
unit normal n, h_K the cell diagonal (ufl.CellDiameter).  The file also stores runtime rules that integrate
these polynomials exactly, in the reference's conventions (unit-cell points, cell-relative weights).
"""
import os
import sys

import numpy as np
import sympy as sp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from customquad_b200 import fem  # noqa: E402
from customquad_b200.mesh import CellType  # noqa: E402

BETA = sp.Rational(10)


def lagrange(deg, t, xi):
    nodes = [sp.Rational(k, deg) for k in range(deg + 1)]
    e = sp.Integer(1)
    for k, nk in enumerate(nodes):
        if k != t:
            e *= (xi - nk) / (nodes[t] - nk)
    return sp.expand(e)


def basis(tdim, deg, X, x0, h):
    ct = CellType.quadrilateral if tdim == 2 else CellType.hexahedron
    lat = fem.point_frame_lattice(ct, deg)
    psi = []
    for j in range(lat.shape[0]):
        e = sp.Integer(1)
        for d in range(tdim):
            e *= lagrange(deg, int(lat[j, d]), (X[d] - x0[d]) / h[d])
        psi.append(e)
    return psi


def gauss01(m):
    x, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (x + 1.0), 0.5 * w


def case(tdim, deg, x0, h, surf):
    X = sp.symbols("x y z")[:tdim]
    x0s, hs = [sp.nsimplify(v) for v in x0], [sp.nsimplify(v) for v in h]
    psi = basis(tdim, deg, X, x0s, hs)
    n = len(psi)
    lim = [(X[d], x0s[d], x0s[d] + hs[d]) for d in range(tdim)]
    mass = sp.zeros(n, n)
    stiff = sp.zeros(n, n)
    for i in range(n):
        for j in range(i, n):
            mass[i, j] = mass[j, i] = sp.integrate(psi[i] * psi[j], *lim)
            g = sum(sp.diff(psi[i], X[d]) * sp.diff(psi[j], X[d]) for d in range(tdim))
            stiff[i, j] = stiff[j, i] = sp.integrate(sp.expand(g), *lim)
    # surface patch: P(s[, t]) = P0 + s a (+ t b), fractions of the cell
    P0 = [x0s[d] + hs[d] * sp.nsimplify(surf["p0"][d]) for d in range(tdim)]
    a = [hs[d] * sp.nsimplify(surf["a"][d]) for d in range(tdim)]
    s, t = sp.symbols("s t")
    if tdim == 2:
        nrm = sp.Matrix([a[1], -a[0]])
        area = sp.sqrt(a[0] ** 2 + a[1] ** 2)
        par = {X[d]: P0[d] + s * a[d] for d in range(2)}
        plim = [(s, 0, 1)]
    else:
        b = [hs[d] * sp.nsimplify(surf["b"][d]) for d in range(3)]
        nrm = sp.Matrix(a).cross(sp.Matrix(b))
        area = sp.sqrt(sum(c**2 for c in nrm))
        par = {X[d]: P0[d] + s * a[d] + t * b[d] for d in range(3)}
        plim = [(s, 0, 1), (t, 0, 1)]
    nrm = nrm / area
    hK = sp.sqrt(sum(v**2 for v in hs))
    nit = sp.zeros(n, n)
    dn = [sum(nrm[d] * sp.diff(psi[i], X[d]) for d in range(tdim)) for i in range(n)]
    for i in range(n):
        for j in range(i, n):
            e = -dn[j] * psi[i] - psi[j] * dn[i] + BETA / hK * psi[i] * psi[j]
            nit[i, j] = nit[j, i] = sp.integrate(sp.expand(e.subs(par)), *plim) * area
    # runtime rules (reference conventions): bulk tensor Gauss, 3 points per direction for degree <= 2 (exact to degree 5)
    m = deg + 1
    g, gw = gauss01(m)
    grids = np.meshgrid(*([g] * tdim), indexing="ij")
    pts = np.stack([q.reshape(-1) for q in grids], axis=1)
    w = np.ones(len(pts))
    for q in np.meshgrid(*([gw] * tdim), indexing="ij"):
        w = w * q.reshape(-1)
    gs, gsw = gauss01(tdim * deg + 1)  # psi_i psi_j along the patch: degree <= 2 tdim deg per parameter
    p0, av = np.asarray(surf["p0"], float), np.asarray(surf["a"], float)
    cellvol = float(np.prod(h))
    if tdim == 2:
        spts = p0[None, :] + gs[:, None] * av[None, :]
        sw = gsw * float(area) / cellvol
    else:
        bv = np.asarray(surf["b"], float)
        S, T = np.meshgrid(gs, gs, indexing="ij")
        spts = p0[None, :] + S.reshape(-1, 1) * av[None, :] + T.reshape(-1, 1) * bv[None, :]
        sw = (gsw[:, None] * gsw[None, :]).reshape(-1) * float(area) / cellvol
    nn = np.tile(np.array([float(c) for c in nrm])[None, :], (len(spts), 1))
    f = lambda M: np.array(M.evalf(30).tolist(), dtype=np.float64)  # noqa: E731
    return dict(x0=np.asarray(x0, float), h=np.asarray(h, float), degree=deg, mass=f(mass), stiffness=f(stiff), nitsche=f(nit),
                bulk_pts=pts, bulk_w=w, surf_pts=spts, surf_w=sw, surf_n=nn, betaN=float(BETA))


def main():
    cases = {
        "quad_q1": case(2, 1, [-1.23, 0.4], [0.37, 0.21], dict(p0=[0.1, 0.2], a=[0.7, 0.5])),
        "quad_q2": case(2, 2, [0.3, -2.0], [0.25, 0.6], dict(p0=[0.9, 0.1], a=[-0.6, 0.8])),
        "hex_q1": case(3, 1, [0.1, -0.7, 1.2], [0.3, 0.2, 0.45], dict(p0=[0.1, 0.2, 0.3], a=[0.6, 0.1, 0.2], b=[0.1, 0.5, -0.2])),
        "hex_q2": case(3, 2, [-0.4, 0.5, 0.0], [0.5, 0.25, 0.4], dict(p0=[0.2, 0.1, 0.6], a=[0.5, 0.2, -0.3], b=[-0.1, 0.6, 0.2])),
    }
    flat = {f"{k}/{kk}": v for k, c in cases.items() for kk, v in c.items()}
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "exact_element_tensors.npz"), **flat)
    for k, c in cases.items():
        print(k, c["mass"].shape, "sum(mass) =", c["mass"].sum(), "vol =", np.prod(c["h"]))


if __name__ == "__main__":
    main()
