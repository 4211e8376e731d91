"""Host pieces around the path: the circle quadrature producer against the checks
the reference's demo prints (volume / area of the rules, demo/poisson.py:293-298;
utils.volume / utils.area, customquad/utils.py:228-243) and the ghost-penalty / inactive-dof helpers."""
import numpy as np

import customquad_b200 as cq
from customquad_b200 import cutfem, utils
from customquad_b200.levelset import generate_qr_circle

XMIN, XMAX = [-1.11, -1.51], [1.55, 1.22]  # demo/poisson.py:40-41


def test_circle_rules_volume_area_and_conventions():
    errs = []
    for N in (8, 16, 32):
        mesh = cq.create_rectangle([XMIN, XMAX], [N, N])
        cut, uncut, outside, qp, qw, qpb, qwb, qn = generate_qr_circle(mesh, 1.0, (0.0, 0.0), order=4)
        assert len(cut) + len(uncut) + len(outside) == mesh.num_cells
        assert len(set(cut) & set(uncut)) == 0 and len(set(cut) & set(outside)) == 0
        assert len(qp) == len(qw) == len(qpb) == len(qwb) == len(qn) == len(cut)
        for p, w, pb, wb, n in zip(qp, qw, qpb, qwb, qn):
            assert p.size == 2 * w.size and pb.size == 2 * wb.size == n.size
            assert np.all(w > 0) and w.sum() < 1.0 + 1e-12  # cell-relative weights (algoim_utils.hpp:127-132)
            assert np.all((p > -1e-12) & (p < 1 + 1e-12)) and np.all((pb > -1e-12) & (pb < 1 + 1e-12))  # unit cell (:134-143)
            nn = n.reshape(-1, 2)
            assert np.allclose(np.sum(nn * nn, axis=1), 1.0)
            # surface points lie on the circle and the normal is radial
        vol = utils.volume(np.asarray(XMIN), np.asarray(XMAX), [N, N], uncut, qw)
        area = utils.area(np.asarray(XMIN), np.asarray(XMAX), [N, N], qwb)
        errs.append(abs(vol - np.pi))
        assert abs(area - 2 * np.pi) < 1e-12
    assert errs[0] < 1e-4 and errs[1] < errs[0] / 20 and errs[2] < errs[1] / 20  # high-order convergence of the bulk rules


def test_sphere_rules_volume_area_and_conventions():
    from customquad_b200.levelset import generate_qr_sphere

    xmin, xmax = np.array([-1.11, -1.51, -1.23]), np.array([1.55, 1.22, 1.11])  # demo/poisson.py:46-47
    errs = []
    for N in (6, 12):
        mesh = cq.create_box([xmin, xmax], [N, N, N])
        cut, uncut, outside, qp, qw, qpb, qwb, qn = generate_qr_sphere(mesh, 1.0, (0.0, 0.0, 0.0), order=4)
        assert len(cut) + len(uncut) + len(outside) == mesh.num_cells
        h = mesh.h()
        for cell, p, w, pb, wb, n in zip(cut, qp, qw, qpb, qwb, qn):
            assert p.size == 3 * w.size and pb.size == 3 * wb.size == n.size
            assert np.all(w > 0) and w.sum() < 1.0 + 1e-10 and np.all(wb > 0)
            assert np.all((p > -1e-12) & (p < 1 + 1e-12)) and np.all((pb > -1e-12) & (pb < 1 + 1e-12))
            iz, rem = divmod(int(cell), N * N)
            iy, ix = divmod(rem, N)
            x = xmin + np.array([ix, iy, iz]) * h + pb.reshape(-1, 3) * h
            assert np.allclose(np.linalg.norm(x, axis=1), 1.0) and np.allclose(x, n.reshape(-1, 3))  # on the sphere, radial unit normal
        vol = utils.volume(xmin, xmax, [N, N, N], uncut, qw)
        area = utils.area(xmin, xmax, [N, N, N], qwb)
        errs.append((abs(vol - 4 * np.pi / 3) / (4 * np.pi / 3), abs(area - 4 * np.pi) / (4 * np.pi)))
    assert errs[0][0] < 1e-4 and errs[0][1] < 1e-4 and errs[1][0] < 1e-5 and errs[1][1] < 1e-5, errs


def test_surface_points_on_circle():
    N = 12
    mesh = cq.create_rectangle([XMIN, XMAX], [N, N])
    cut, _, _, _, _, qpb, _, qn = generate_qr_circle(mesh, 0.9, (0.1, -0.2), order=4)
    h = mesh.h()
    for cell, pb, n in zip(cut, qpb, qn):
        ix, iy = divmod(int(cell), N)
        lo = np.asarray(XMIN) + np.array([ix, iy]) * h
        x = lo + pb.reshape(-1, 2) * h
        r = x - np.array([0.1, -0.2])
        assert np.allclose(np.linalg.norm(r, axis=1), 0.9)
        assert np.allclose(r / 0.9, n.reshape(-1, 2))


def test_ghost_penalty_and_inactive_dofs():
    mesh = cq.create_rectangle([XMIN, XMAX], [16, 16])
    cut, uncut, outside, *_ = generate_qr_circle(mesh)
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    F = cutfem.ghost_penalty_facets(mesh, cut, outside)
    cutset, outset = set(cut.tolist()), set(outside.tolist())
    for lo, hi, _ in F:  # customquad/utils.py:166-175
        assert (lo in cutset and hi not in outset) or (hi in cutset and lo not in outset)
    S = cutfem.ghost_penalty_matrix(V, F, 1.0)
    X = V.tabulate_dof_coordinates()
    assert abs(S - S.T).max() < 1e-14
    for u in (np.ones(V.num_dofs), X[:, 0], X[:, 1], X[:, 0] * X[:, 1]):  # no jump of the normal derivative
        assert np.abs(S @ u).max() < 1e-12
    assert (X[:, 0] ** 2) @ (S @ (X[:, 0] ** 2)) > 1e-3
    inactive = cutfem.get_inactive_dofs(V, cut, uncut)
    d = V.dofmap.list
    active = np.unique(np.concatenate([d[cut].reshape(-1), d[uncut].reshape(-1)]))
    assert len(inactive) + len(active) == V.num_dofs and len(np.intersect1d(inactive, active)) == 0
    pts, w = cutfem.standard_rule(mesh, uncut, 2)[1:]
    assert pts.shape == (len(uncut), 8) and np.allclose(w.sum(axis=1), 1.0)
