"""The GPU quadrature producer (cq_circle_classify + cq_circle_rules, csrc/cq_levelset.cuh) against the host
producer customquad_b200/levelset.py:generate_qr_circle, and its device-resident rules through the assembly
path: volume / perimeter of the disk (the reference's check, demo/poisson.py:283-298) and a stiffness matrix
equal to the one assembled from the host rules."""
import numpy as np
import pytest

import customquad_b200 as cq
from customquad_b200 import forms, levelset
from customquad_b200.quadrature import RaggedQuadrature

pytestmark = pytest.mark.gpu

CASES = [
    ([[-1.11, -1.51], [1.55, 1.22]], [32, 32], 1.0, (0.0, 0.0), 4),
    ([[-1.11, -1.51], [1.55, 1.22]], [17, 29], 0.83, (0.1, -0.2), 6),
    ([[-1.0, -1.0], [1.0, 1.0]], [16, 16], 0.5, (0.0, 0.0), 4),
    ([[-2.0, -2.0], [2.0, 2.0]], [3, 3], 0.4, (0.05, 0.1), 4),
    ([[-1.11, -1.51], [1.55, 1.22]], [256, 192], 1.3, (0.3, 0.1), 14),
]


@pytest.mark.parametrize("box,n,radius,center,order", CASES)
def test_device_producer_matches_host_producer(cuda_backend, box, n, radius, center, order):
    mesh = cq.create_rectangle(box, n)
    cut, uncut, outside, qp, qw, qpb, qwb, qn = levelset.generate_qr_circle(mesh, radius, center, order)
    dcut, duncut, doutside, qb, qs = levelset.generate_qr_circle_device(mesh, radius, center, order)
    assert qb.cells is None and qs.cells is None  # born on the device, no host copy until asked for
    qb, qs = qb.host(), qs.host()
    assert np.array_equal(dcut, cut) and np.array_equal(duncut, uncut) and np.array_equal(doutside, outside)
    assert np.array_equal(np.diff(qb.offsets), [w.size for w in qw])
    assert np.array_equal(np.diff(qs.offsets), [w.size for w in qwb])
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)  # noqa: E731
    np.testing.assert_allclose(qb.pts, cat(qp), rtol=0, atol=1e-12)
    # thin slices: (th - tl) is a difference of nearly equal numbers, and the GPU contracts a*b+c into FMAs
    np.testing.assert_allclose(qb.w, cat(qw), rtol=1e-9, atol=1e-14)
    # an arc end that rounds to the neighbouring 1e-14 moves a point by R * 1e-14 / h unit-cell lengths
    np.testing.assert_allclose(qs.pts, cat(qpb), rtol=0, atol=1e-12 + 2e-14 * radius / float(np.min(mesh.h())))
    np.testing.assert_allclose(qs.w, cat(qwb), rtol=1e-9, atol=1e-12)  # arc ends are angles rounded to 14 decimals
    np.testing.assert_allclose(qs.n, cat(qn), rtol=0, atol=1e-12)


def test_device_rules_feed_the_assembly_path(cuda_backend):
    """demo/poisson.py:283-298: volume and perimeter from the rules; and the cut-cell stiffness matrix."""
    N, R = 128, 1.0
    mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [N, N])
    cut, uncut, outside, qb, qs = levelset.generate_qr_circle_device(mesh, R, (0.0, 0.0), 4)
    one = forms.integral(mesh, 1.0)
    cellvol = float(np.prod(mesh.h()))
    vol = float(cq.assemble_scalar(one, [qb])) + cellvol * len(uncut)
    per = float(cq.assemble_scalar(one, [qs]))
    assert abs(vol - np.pi * R * R) < 1e-8
    assert abs(per - 2 * np.pi * R) < 1e-10
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    hcut, _, _, qp, qw, qpb, qwb, qn = levelset.generate_qr_circle(mesh, R, (0.0, 0.0), 4)
    A = cq.assemble_matrix(forms.stiffness(V), [qb])
    B = cq.assemble_matrix(forms.stiffness(V), [(hcut, qp, qw)])
    a, b = A.data.cpu().numpy(), B.data.cpu().numpy()
    assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b))
    An = cq.assemble_matrix(forms.nitsche(V, 10.0), [qs])
    Bn = cq.assemble_matrix(forms.nitsche(V, 10.0), [(hcut, qpb, qwb, qn)])
    a, b = An.data.cpu().numpy(), Bn.data.cpu().numpy()
    assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b))


SPHERE_CASES = [
    ([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [8, 8, 8], 1.0, (0.0, 0.0, 0.0), 4, 2),
    ([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [9, 7, 8], 0.83, (0.1, -0.2, 0.05), 4, 0),
    ([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], [4, 4, 4], 1.0, (0.0, 0.0, 0.0), 2, 2),
]


@pytest.mark.parametrize("box,n,radius,center,order,axis", SPHERE_CASES)
def test_device_sphere_producer_matches_host_producer(cuda_backend, box, n, radius, center, order, axis):
    mesh = cq.create_box(box, n)
    cut, uncut, outside, qp, qw, qpb, qwb, qn = levelset.generate_qr_sphere(mesh, radius, center, order, axis)
    dcut, duncut, doutside, qb, qs = levelset.generate_qr_sphere_device(mesh, radius, center, order, axis)
    qb, qs = qb.host(), qs.host()
    assert np.array_equal(dcut, cut) and np.array_equal(duncut, uncut) and np.array_equal(doutside, outside)
    assert np.array_equal(np.diff(qb.offsets), [w.size for w in qw])
    assert np.array_equal(np.diff(qs.offsets), [w.size for w in qwb])
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)  # noqa: E731
    hmin = float(np.min(mesh.h()))
    np.testing.assert_allclose(qb.pts, cat(qp), rtol=0, atol=1e-12)
    np.testing.assert_allclose(qb.w, cat(qw), rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(qs.pts, cat(qpb), rtol=0, atol=1e-12 + 2e-14 * radius / hmin)
    np.testing.assert_allclose(qs.w, cat(qwb), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(qs.n, cat(qn), rtol=0, atol=1e-12)


def test_device_sphere_rules_at_config3_resolution(cuda_backend):
    """Real cut-cell rules for the sphere on 96^3 hexahedra (config 3 is this at 256^3; the numpy producer needs minutes
    there): volume and area through assemble_scalar, and the Nitsche matrix on the surface rules is symmetric."""
    import torch

    N, R = 96, 1.0
    mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [N, N, N])
    cut, uncut, outside, qb, qs = levelset.generate_qr_sphere_device(mesh, R, (0.0, 0.0, 0.0), 4)
    assert len(cut) + len(uncut) + len(outside) == mesh.num_cells and len(cut) > 20000
    one = forms.integral(mesh, 1.0)
    cellvol = float(np.prod(mesh.h()))
    vol = float(cq.assemble_scalar(one, [qb])) + cellvol * len(uncut)
    area = float(cq.assemble_scalar(one, [qs]))
    assert abs(vol - 4 * np.pi / 3 * R**3) < 1e-6 * vol
    assert abs(area - 4 * np.pi * R * R) < 1e-5 * area
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    A = cq.assemble_matrix(forms.nitsche(V, 10.0), [qs])
    A.assemble()
    M = torch.sparse_csr_tensor(A.indptr, A.indices.to(torch.int64), A.data, size=A.shape)
    x = torch.rand(A.shape[0], dtype=torch.float64, device=A.data.device)
    y = torch.rand(A.shape[0], dtype=torch.float64, device=A.data.device)
    assert abs(float(y @ (M @ x)) - float(x @ (M @ y))) <= 1e-10 * float(A.data.abs().sum())
