#!/usr/bin/env python
"""CutFEM Poisson on the unit circle (Q1 quadrilaterals) or the unit sphere (Q1 hexahedra): the
reference's demo/poisson.py (BASELINE.json configs[0]; -domain sphere is the small version of configs[2])
on this package.

Same steps as the demo: cut-cell quadrature for the level set (here customquad_b200.levelset instead
of Algoim, demo/poisson.py:88-113), bulk stiffness + load on the uncut cells with a standard Gauss
rule and ghost penalty on the cut-cell facets (:168-176; dolfinx in the demo), runtime-quadrature
assembly of the cut bulk part and of the Nitsche boundary terms with quadrature normals (:180-215) -- all of it,
ghost penalty included, added into one device CSR --,
inactive-dof locking (:229-232), a direct solve (:239-256; scipy here), L2 / H10 errors and the
area / volume functionals through assemble_scalar (:266-290).

    python examples/poisson.py -N 32        (needs a B200; prints one JSON line)
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import customquad_b200 as cq  # noqa: E402
from customquad_b200 import cutfem, forms, utils  # noqa: E402
from customquad_b200.levelset import (  # noqa: E402
    generate_qr_circle,
    generate_qr_circle_device,
    generate_qr_sphere,
    generate_qr_sphere_device,
)


def solve(N, betaN=10.0, betas=1.0, order=4, domain="circle", qr="host"):
    """qr = "device": the rules come from the GPU producer (cq_circle_* / cq_sphere_*) and stay in HBM."""
    import scipy.sparse.linalg as spla

    device_rules = None
    if domain == "circle":  # demo/poisson.py:39-50
        xmin, xmax = np.array([-1.11, -1.51]), np.array([1.55, 1.22])
        volume_exact, area_exact = np.pi, 2 * np.pi
        mesh = cq.create_rectangle([xmin, xmax], [N, N])
        if qr == "device":
            cut, uncut, outside, qb, qs = generate_qr_circle_device(mesh, 1.0, (0.0, 0.0), order)
            device_rules = (qb, qs)
            rules = (cut, uncut, outside, None, np.split(qb.host().w, qb.host().offsets[1:-1]), None, np.split(qs.host().w, qs.host().offsets[1:-1]), None)
        else:
            rules = generate_qr_circle(mesh, 1.0, (0.0, 0.0), order)
    elif domain in ("square", "flower"):
        # the reference's third level set (demo/square.hpp) and a smooth non-convex one given as callables: rules from
        # customquad_b200/levelset_general.py (host), the same hot path behind them
        from customquad_b200.levelset_general import generate_qr_box, generate_qr_levelset

        assert qr == "host", "device producers: circle and sphere"
        xmin, xmax = np.array([-1.11, -1.51]), np.array([1.55, 1.22])
        mesh = cq.create_rectangle([xmin, xmax], [N, N])
        if domain == "square":
            volume_exact, area_exact = 4.0, 8.0
            rules = generate_qr_box(mesh, order=order)
        else:
            th = np.linspace(0, 2 * np.pi, 400001)
            volume_exact = np.pi * (0.8**2 + 0.15**2 / 2)
            area_exact = float(np.trapezoid(np.sqrt((0.8 + 0.15 * np.cos(5 * th)) ** 2 + (0.75 * np.sin(5 * th)) ** 2), th))

            def phi(x):
                return np.hypot(x[:, 0], x[:, 1]) - (0.8 + 0.15 * np.cos(5 * np.arctan2(x[:, 1], x[:, 0])))

            def grad(x):
                r, t = np.hypot(x[:, 0], x[:, 1]), np.arctan2(x[:, 1], x[:, 0])
                dr = 0.75 * np.sin(5 * t)
                return np.stack([x[:, 0] / r - dr * x[:, 1] / r**2, x[:, 1] / r + dr * x[:, 0] / r**2], axis=1)

            rules = generate_qr_levelset(mesh, phi, grad, order=order)
    else:
        xmin, xmax = np.array([-1.11, -1.51, -1.23]), np.array([1.55, 1.22, 1.11])
        volume_exact, area_exact = 4 * np.pi / 3, 4 * np.pi
        mesh = cq.create_box([xmin, xmax], [N, N, N])
        if qr == "device":
            cut, uncut, outside, qb, qs = generate_qr_sphere_device(mesh, 1.0, (0.0, 0.0, 0.0), order)
            device_rules = (qb, qs)
            rules = (cut, uncut, outside, None, np.split(qb.host().w, qb.host().offsets[1:-1]), None, np.split(qs.host().w, qs.host().offsets[1:-1]), None)
        else:
            rules = generate_qr_sphere(mesh, 1.0, (0.0, 0.0, 0.0), order)
    gdim = len(xmin)
    cut_cells, uncut_cells, outside_cells, qr_pts, qr_w, qr_pts_bdry, qr_w_bdry, qr_n = rules
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    u_ex = cq.Analytic.sinprod(1.0)  # prod_d sin(pi x_d)  (:59-67)
    f = cq.Analytic.sinprod(gdim * np.pi**2)  # -div grad u_ex  (:147)
    g = cq.Function(V).interpolate(lambda x: np.prod(np.sin(np.pi * x[:gdim]), axis=0))  # (:146)

    qr_uncut = cutfem.standard_rule(mesh, uncut_cells, 2)  # dx_uncut: dolfinx's default degree for Q1 x Q1 gradients
    qr_bulk = [(cut_cells, qr_pts, qr_w)] if device_rules is None else [device_rules[0]]
    qr_bdry = [(cut_cells, qr_pts_bdry, qr_w_bdry, qr_n)] if device_rules is None else [device_rules[1]]

    # matrices: uncut bulk + ghost penalty (Ax, :168-170) + cut bulk + Nitsche boundary (A = Ax + Ac1 + Ac2, :203-205),
    # all into ONE device CSR whose pattern holds the cell couplings and the couplings across the ghost-penalty facets
    from customquad_b200 import assemble as asm

    facets = cutfem.ghost_penalty_facets(mesh, cut_cells, outside_cells)
    A = asm.create_matrix(forms.stiffness(V), facets=facets)
    cq.assemble_matrix(forms.stiffness(V), [qr_uncut], A=A)
    cutfem.ghost_penalty(V, facets, betas, A)  # betas * avg(h) * [n.grad u][n.grad v] dS(ghost_penalty_tag)
    cq.assemble_matrix(forms.stiffness(V), qr_bulk, A=A)
    cq.assemble_matrix(forms.nitsche(V, betaN), qr_bdry, A=A)
    A.assemble()
    # right-hand side: bx + bc1 + bc2 (:208-215); load with a degree-4 rule for the sine source
    b = cq.assemble_vector(forms.load(V, f), [cutfem.standard_rule(mesh, uncut_cells, 4)])
    b += cq.assemble_vector(forms.load(V, f), qr_bulk)
    b += cq.assemble_vector(forms.nitsche_rhs(V, g, betaN), qr_bdry)
    assert np.isfinite(b.array).all() and np.isfinite(A.norm())

    inactive = cutfem.get_inactive_dofs(V, cut_cells, uncut_cells)
    A = cutfem.lock_inactive_dofs(inactive, A)
    # the complete operator leaves the GPU once, for the direct solve (the solver is not part of the path)
    uh = cq.Function(V, spla.spsolve(A.to_scipy().tocsc(), b.array))
    assert np.isfinite(uh.array).all()

    def assemble(make_form):
        m_cut = cq.assemble_scalar(make_form(), qr_bulk)
        m_uncut = cq.assemble_scalar(make_form(), [cutfem.standard_rule(mesh, uncut_cells, 6)])
        return m_cut + m_uncut

    L2 = np.sqrt(assemble(lambda: forms.l2_error(uh, u_ex)))
    H10 = np.sqrt(assemble(lambda: forms.h10_error(uh, u_ex)))
    area_func = cq.assemble_scalar(forms.integral(mesh, 1.0), qr_bdry)
    volume_func = assemble(lambda: forms.integral(mesh, 1.0))
    return {
        "domain": domain, "N": N, "h": float(np.max(mesh.h())), "cut_cells": int(len(cut_cells)), "uncut_cells": int(len(uncut_cells)),
        "L2_err": float(L2), "H10_err": float(H10),
        "functional_volume_err": abs(volume_func - volume_exact) / volume_exact, "functional_area_err": abs(area_func - area_exact) / area_exact,
        "qr_volume_err": abs(utils.volume(xmin, xmax, [N] * gdim, uncut_cells, qr_w) - volume_exact) / volume_exact,
        "qr_area_err": abs(utils.area(xmin, xmax, [N] * gdim, qr_w_bdry) - area_exact) / area_exact,
    }  # fmt: skip


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-N", type=int, default=32)  # demo/poisson.py:13
    ap.add_argument("-betaN", type=float, default=10.0)
    ap.add_argument("-betas", type=float, default=1.0)
    ap.add_argument("-domain", default="circle", choices=["circle", "sphere", "square", "flower"])  # demo/poisson.py:17 (+ demo/square.hpp, a smooth callable level set)
    ap.add_argument("-qr", default="host", choices=["host", "device"], help="where the cut-cell rules are produced")
    a = ap.parse_args()
    print(json.dumps(solve(a.N, a.betaN, a.betas, domain=a.domain, qr=a.qr)))
