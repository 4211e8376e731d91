"""The reference's own test cases (reference test/*.py), restated without dolfinx so they can be
run against any backend (the CPU oracle or the CUDA path).  Each function mirrors one test of
the reference and checks against the committed known answers (tests/golden/).  Tolerances are
the reference's: relative 1e-10 (test/test_linear_basis.py:38,66,115; test_volume_area.py:60,146),
absolute 1e-10 per dof (test_linear_basis.py:211-212,332-333).
"""
import numpy as np

import customquad_b200 as cq
from customquad_b200 import forms
from customquad_b200.quadrature import gauss_tensor_rule

TOL = 1e-10


def fcn1(x):
    return x[0] ** 0


def fcn2(x):
    return x[0]


def fcn3(x):
    return np.sin(np.pi * x[0]) * np.sin(np.pi * x[1])


def fcn4(x):
    return np.exp(x[0] * x[1])


FCNS = [fcn1, fcn2, fcn3, fcn4]


def assembly_case(backend, golden, idx):
    """test_quads_assembly / test_hexes_assembly (test_linear_basis.py:13-66) through the harness
    of test/common.py:39-108: all cells, FIAT degree-2 rule tiled as 2-D arrays."""
    key = f"assembly/{idx}"
    N, xmin, xmax = golden[key + "/N"], golden[key + "/xmin"], golden[key + "/xmax"]
    tdim = len(N)
    mesh = (cq.create_rectangle if tdim == 2 else cq.create_box)(np.array([xmin, xmax]), N)
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    num_cells = mesh.topology.num_cells
    cells = np.arange(num_cells)
    pts, w = gauss_tensor_rule(tdim, 2)
    qr_pts = np.tile(pts.flatten(), [num_cells, 1])
    qr_w = np.tile(w.flatten(), [num_cells, 1])
    qr_data = [(cells, qr_pts, qr_w)]

    A = backend.matrix(forms.mass(V), qr_data)
    A_ref = golden[key + "/matrix"]
    assert np.linalg.norm(A.toarray() - A_ref) / np.linalg.norm(A_ref) < TOL
    for k, fcn in enumerate(FCNS):
        f = cq.Function(V).interpolate(fcn)
        b = backend.vector(forms.load(V, f), qr_data)
        b_ref = golden[key + f"/fcn{k + 1}/vector"]
        assert np.linalg.norm(b - b_ref) / np.linalg.norm(b_ref) < TOL
        m = backend.scalar(forms.integral_sq(f), qr_data)
        m_ref = float(golden[key + f"/fcn{k + 1}/scalar"])
        assert abs(m - m_ref) / abs(m_ref) < TOL


def edges_case(backend, golden):
    """test_edges (test_linear_basis.py:69-115): integral of 2x+y over each edge of one quad via a
    1-point rule at the edge midpoint with w = edge_length / cell_area."""
    xmin, xmax = golden["edges/xmin"], golden["edges/xmax"]
    mesh = cq.create_rectangle(np.array([xmin, xmax]), [1, 1])
    form = forms.integral(mesh, cq.Analytic.affine(0.0, [2.0, 1.0]))
    cells = np.arange(mesh.topology.num_cells)
    cell_volume = np.prod(xmax - xmin)
    for k in range(4):
        midpoint = golden["edges/midpoints"][k]
        qr_pts = np.expand_dims((midpoint - xmin) / (xmax - xmin), axis=0)
        qr_w = np.array([[golden["edges/lengths"][k] / cell_volume]])
        m = backend.scalar(form, [(cells, qr_pts, qr_w)])
        m_exact = golden["edges/exact"][k]
        assert abs(m - m_exact) / abs(m_exact) < TOL


def edge_integral_case(backend, golden):
    """test_edge_integral (test_linear_basis.py:118-212): integral v ds per side of one quad equals
    the runtime cell integral with the midpoint rule; pins local dof <-> reference point."""
    xmin, xmax = golden["edge_integral/xmin"], golden["edge_integral/xmax"]
    mesh = cq.create_rectangle(np.array([xmin, xmax]), [1, 1])
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    cells = np.arange(1)
    for k in range(4):
        qr_pts_local = np.expand_dims(golden["edge_integral/qr_pts_ref"][k], axis=0)
        qr_w_local = np.expand_dims(golden["edge_integral/qr_w"][k], axis=(0, 1))
        b = backend.vector(forms.load(V, 1.0), [(cells, qr_pts_local, qr_w_local)])
        assert np.all(np.abs(b - golden["edge_integral/exact"][k]) < TOL)


def face_integral_case(backend, golden):
    """test_face_integral (test_linear_basis.py:215-333)."""
    bbmin, bbmax = golden["face_integral/bbmin"], golden["face_integral/bbmax"]
    mesh = cq.create_box(np.array([bbmin, bbmax]), [1, 1, 1])
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    cells = np.arange(1)
    for k in range(6):
        qr_pts_local = np.expand_dims(golden["face_integral/qr_pts_ref"][k], axis=0)
        qr_w_local = np.expand_dims(golden["face_integral/qr_w"][k], axis=(0, 1))
        b = backend.vector(forms.load(V, 1.0), [(cells, qr_pts_local, qr_w_local)])
        assert np.all(np.abs(b - golden["face_integral/exact"][k]) < TOL)


def get_mesh(golden):
    """test/common.py:111-149: 10x10 quads, rings of 'outside' and 'cut' cells by centroid."""
    N = int(golden["volume/N"])
    xmin, xmax = golden["volume/xmin"], golden["volume/xmax"]
    mesh = cq.create_rectangle(np.array([xmin, xmax]), [N, N])
    h = (xmax - xmin) / N
    centroids = np.mean(mesh.geometry.x[mesh.geometry.dofmap], axis=1)
    xc, yc = centroids[:, 0], centroids[:, 1]
    all_cells = np.arange(mesh.topology.num_cells, dtype=np.int32)

    def ring(k):
        return np.where((xc < xmin[0] + k * h[0]) | (xc > xmax[0] - k * h[0]) | (yc < xmin[1] + k * h[1]) | (yc > xmax[1] - k * h[1]))[0]

    outside_cells = ring(1)
    cut_cells = np.setdiff1d(ring(2), outside_cells)
    uncut_cells = np.setdiff1d(np.setdiff1d(all_cells, outside_cells), cut_cells)
    assert len(outside_cells) == int(golden["volume/num_outside"])
    assert len(cut_cells) == int(golden["volume/num_cut"])
    assert len(uncut_cells) == int(golden["volume/num_uncut"])
    return mesh, h, cut_cells, uncut_cells, outside_cells


def volume_case(backend, golden):
    """test_volume (test_volume_area.py:10-74): one-point rule xi = 0.5, w = 1 on cell subsets."""
    mesh, h, cut_cells, uncut_cells, outside_cells = get_mesh(golden)
    dim = 2
    cell_vol = h[0] * h[1]
    form = forms.integral(mesh, 1.0)
    qr_pts = np.tile([0.5] * dim, [len(cut_cells), 1])
    qr_w = np.tile(1.0, [len(cut_cells), 1])
    vol_cut = backend.scalar(form, [(cut_cells, qr_pts, qr_w)])
    assert abs(vol_cut - len(cut_cells) * cell_vol) / abs(vol_cut) < TOL
    assert abs(vol_cut - float(golden["volume/cut_exact"])) / abs(vol_cut) < TOL
    qr_pts = np.tile([0.5] * dim, [len(outside_cells), 1])
    qr_w = np.tile(1.0, [len(outside_cells), 1])
    vol_outside = backend.scalar(form, [(outside_cells, qr_pts, qr_w)])
    assert abs(vol_outside - float(golden["volume/outside_exact"])) / abs(vol_outside) < TOL
    vol_bulk = len(uncut_cells) * cell_vol  # dolfinx standard assembly of 1*dx(uncut) in the reference
    exact_vol = float(golden["volume/total_exact"])
    assert abs(vol_bulk + vol_cut + vol_outside - exact_vol) / exact_vol < TOL


def full_area_case(backend, golden):
    """test_full_area (test_volume_area.py:92-146): perimeter through 1-point rules with
    w = facet_area / cell_volume; `cells` is a Python list with the corner cells repeated."""
    mesh, h, *_ = get_mesh(golden)
    N = int(golden["volume/N"])
    exterior_cells, facet_area = [], []
    for ix in range(N):
        for iy in range(N):
            c = int(mesh.cell_index(ix, iy))
            if ix == 0:
                exterior_cells.append(c), facet_area.append(h[1])
            if ix == N - 1:
                exterior_cells.append(c), facet_area.append(h[1])
            if iy == 0:
                exterior_cells.append(c), facet_area.append(h[0])
            if iy == N - 1:
                exterior_cells.append(c), facet_area.append(h[0])
    num_ext_facets = len(exterior_cells)
    assert num_ext_facets == 4 * N and len(set(exterior_cells)) == 4 * N - 4  # corners twice
    qr_pts = np.tile([0.0] * 2, [num_ext_facets, 1])
    qr_w = np.empty([num_ext_facets, 1])
    for k in range(num_ext_facets):
        qr_w[k] = facet_area[k] / (h[0] * h[1])
    area = backend.scalar(forms.integral(mesh, 1.0), [(exterior_cells, qr_pts, qr_w)])
    exact_area = float(golden["area/exact"])
    assert abs(area - exact_area) / exact_area < TOL
