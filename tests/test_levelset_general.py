"""Producers of customquad_b200/levelset_general.py: the reference's third demo level set (demo/square.hpp) as an
axis-aligned box in 2-D / 3-D, and smooth 2-D level sets given as callables (what the reference gets from Algoim through
demo/algoim_utils.py:92-180).  Checked through integrals with known values and against the circle producer."""
import numpy as np
import pytest
from scipy.special import ellipe

import customquad_b200 as cq
from customquad_b200 import forms
from customquad_b200.levelset import generate_qr_circle
from customquad_b200.levelset_general import generate_qr_box, generate_qr_levelset


def _cells_lo(mesh):
    from customquad_b200.levelset_general import _cells_lo as f

    return f(mesh)


def _integrate(mesh, rules, f_bulk, f_surf, full_order=3):
    """sum over the cut cells' rules + a tensor Gauss rule on the uncut cells (physical coordinates)."""
    cut, uncut, _out, qp, qw, qpb, qwb, _qn = rules
    d = mesh.topology.dim
    h, lo = mesh.h(), _cells_lo(mesh)
    vol = float(np.prod(h))
    bulk = surf = 0.0
    for k, c in enumerate(cut):
        bulk += float(np.sum(qw[k] * vol * f_bulk(lo[c] + qp[k].reshape(-1, d) * h)))
        surf += float(np.sum(qwb[k] * vol * f_surf(lo[c] + qpb[k].reshape(-1, d) * h)))
    g, gw = np.polynomial.legendre.leggauss(full_order)
    g, gw = 0.5 * (g + 1), 0.5 * gw
    G = np.stack([x.reshape(-1) for x in np.meshgrid(*([g] * d), indexing="ij")], axis=1)
    GW = np.prod(np.stack([x.reshape(-1) for x in np.meshgrid(*([gw] * d), indexing="ij")], axis=1), axis=1)
    for c in uncut:
        bulk += float(np.sum(GW * vol * f_bulk(lo[c] + G * h)))
    return bulk, surf


one = lambda x: np.ones(x.shape[0])  # noqa: E731


def test_square_of_the_reference_demo_is_exact():
    mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [23, 19])  # demo/poisson.py:42-43
    rules = generate_qr_box(mesh)  # [-1, 1]^2: demo/square.hpp:19-23
    cut, uncut, outside = rules[:3]
    assert len(cut) + len(uncut) + len(outside) == mesh.num_cells and len(set(cut) & set(uncut)) == 0
    area, per = _integrate(mesh, rules, one, one)
    assert abs(area - 4.0) < 1e-13 and abs(per - 8.0) < 1e-13
    m, s = _integrate(mesh, rules, lambda x: x[:, 0] ** 2 * x[:, 1] ** 2 + x[:, 0], lambda x: x[:, 0] ** 2 + x[:, 1])
    assert abs(m - 4.0 / 9.0) < 1e-13  # int x^2 y^2 = (2/3)^2, int x = 0
    assert abs(s - (2 * 2.0 / 3.0 + 2 * 2.0)) < 1e-13  # x^2: 2/3 on each horizontal edge, 2 on each vertical; y: odd
    for k in range(len(cut)):  # unit axis normals (square.hpp:36-46)
        n = rules[7][k].reshape(-1, 2)
        assert np.all(np.abs(np.abs(n).sum(axis=1) - 1.0) < 1e-15) and np.all((n == 0.0).sum(axis=1) == 1)


def test_cube_is_exact():
    mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [9, 8, 7])
    rules = generate_qr_box(mesh, [-1.0, -0.9, -0.8], [1.0, 0.7, 0.6])
    vol, area = _integrate(mesh, rules, one, one)
    assert abs(vol - 2.0 * 1.6 * 1.4) < 1e-13
    assert abs(area - 2 * (2.0 * 1.6 + 2.0 * 1.4 + 1.6 * 1.4)) < 1e-13
    m, _ = _integrate(mesh, rules, lambda x: x[:, 0] ** 2 * x[:, 2], one)
    assert abs(m - (2.0 / 3.0) * 1.6 * (0.6**2 - 0.8**2) / 2) < 1e-13


def test_box_face_on_a_mesh_plane_is_refused():
    mesh = cq.create_rectangle([[-2.0, -2.0], [2.0, 2.0]], [8, 8])
    with pytest.raises(AssertionError):
        generate_qr_box(mesh)


def _circle(R):
    return (lambda x: x[:, 0] ** 2 + x[:, 1] ** 2 - R * R), (lambda x: 2.0 * x)


@pytest.mark.parametrize("N", [16, 32])
def test_smooth_level_set_reproduces_the_circle(N):
    mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [N, N])
    phi, grad = _circle(1.0)
    rules = generate_qr_levelset(mesh, phi, grad, order=4)
    ref = generate_qr_circle(mesh, 1.0, (0.0, 0.0), 4)
    assert np.array_equal(rules[0], ref[0]) and np.array_equal(rules[1], ref[1]) and np.array_equal(rules[2], ref[2])
    area, per = _integrate(mesh, rules, one, one)
    tol = 2e-6 if N == 16 else 2e-7
    assert abs(area - np.pi) < tol * np.pi and abs(per - 2 * np.pi) < tol * 2 * np.pi
    # cell by cell against the closed-form producer
    for k in range(len(ref[0])):
        assert abs(rules[4][k].sum() - ref[4][k].sum()) < 1e-5 and abs(rules[6][k].sum() - ref[6][k].sum()) < 1e-4
        n = rules[7][k].reshape(-1, 2)
        assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-14)


def test_ellipse_and_flower_converge():
    a, b = 0.9, 0.6
    phi_e = lambda x: (x[:, 0] / a) ** 2 + (x[:, 1] / b) ** 2 - 1.0  # noqa: E731
    grad_e = lambda x: np.stack([2 * x[:, 0] / a**2, 2 * x[:, 1] / b**2], axis=1)  # noqa: E731
    per_e = 4 * a * ellipe(1 - (b / a) ** 2)

    def phi_f(x):
        return np.hypot(x[:, 0], x[:, 1]) - (0.8 + 0.15 * np.cos(5 * np.arctan2(x[:, 1], x[:, 0])))

    def grad_f(x):
        r, th = np.hypot(x[:, 0], x[:, 1]), np.arctan2(x[:, 1], x[:, 0])
        dr = 0.75 * np.sin(5 * th)  # d/dtheta of -(0.8 + 0.15 cos 5 theta)
        return np.stack([x[:, 0] / r - dr * x[:, 1] / r**2, x[:, 1] / r + dr * x[:, 0] / r**2], axis=1)

    th = np.linspace(0, 2 * np.pi, 400001)
    rr, drr = 0.8 + 0.15 * np.cos(5 * th), -0.75 * np.sin(5 * th)
    per_f = float(np.trapezoid(np.sqrt(rr**2 + drr**2), th))
    area_f = np.pi * (0.8**2 + 0.15**2 / 2)
    err = {}
    for N in (24, 48):
        mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [N, N])
        A, P = _integrate(mesh, generate_qr_levelset(mesh, phi_e, grad_e, order=4), one, one)
        Af, Pf = _integrate(mesh, generate_qr_levelset(mesh, phi_f, grad_f, order=4), one, one)
        err[N] = (abs(A - np.pi * a * b), abs(P - per_e), abs(Af - area_f), abs(Pf - per_f))
    assert max(err[48][:2]) < 1e-6 and max(err[48][2:]) < 2e-5
    assert all(e48 < 0.25 * e24 + 1e-9 for e24, e48 in zip(err[24], err[48]))  # at least second order; in practice ~h^4+


def test_rules_feed_the_assembly_path():
    """The tuples go straight into assemble_scalar / assemble_matrix (here: the CPU oracle; the CUDA path takes the same)."""
    from oracle import oracle

    mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [12, 12])
    phi, grad = _circle(0.9)
    cut, uncut, _o, qp, qw, qpb, qwb, qn = generate_qr_levelset(mesh, phi, grad)
    volc = oracle.assemble_scalar(forms.integral(mesh, 1.0), [(cut, qp, qw)])
    area = oracle.assemble_scalar(forms.integral(mesh, 1.0), [(cut, qpb, qwb, qn)])
    assert abs(volc + len(uncut) * float(np.prod(mesh.h())) - np.pi * 0.81) < 1e-5
    assert abs(area - 2 * np.pi * 0.9) < 1e-5
