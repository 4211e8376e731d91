"""csrc/cq_levelset.cuh (the per-cell functions of the GPU quadrature producer) against
customquad_b200/levelset.py:generate_qr_circle: the header is compiled for the host with g++
(tests/levelset_harness.cpp) -- same code as in the CUDA kernels, no GPU needed -- and must give the
same classification, the same number of points per cell and the same points / weights / normals.
Under -m gpu the kernels themselves are compared (tests/test_gpu_levelset_device.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import customquad_b200 as cq
from customquad_b200 import levelset

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ls") / "libls_harness.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "levelset_harness.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    L.ls_sizeof_grid.restype = C.c_int
    L.ls_sizeof_grid3.restype = C.c_int
    return L


CASES = [
    # (box, n, radius, center, order)
    ([[-1.11, -1.51], [1.55, 1.22]], [8, 8], 1.0, (0.0, 0.0), 4),      # demo/poisson.py:42-43 extents
    ([[-1.11, -1.51], [1.55, 1.22]], [32, 32], 1.0, (0.0, 0.0), 4),
    ([[-1.11, -1.51], [1.55, 1.22]], [17, 29], 0.83, (0.1, -0.2), 6),
    ([[-1.0, -1.0], [1.0, 1.0]], [4, 4], 1.0, (0.0, 0.0), 2),          # circle through cell corners and tangent to the box
    ([[-1.0, -1.0], [1.0, 1.0]], [16, 16], 0.5, (0.0, 0.0), 4),        # tangent to grid lines
    ([[-2.0, -2.0], [2.0, 2.0]], [3, 3], 0.4, (0.05, 0.1), 4),         # whole circle inside one cell
    ([[-1.11, -1.51], [1.55, 1.22]], [64, 48], 1.3, (0.3, 0.1), 14),   # leaves the mesh; 8 Gauss points
]


@pytest.mark.parametrize("box,n,radius,center,order", CASES)
def test_device_logic_matches_host_producer(harness, box, n, radius, center, order):
    from customquad_b200 import _lib

    mesh = cq.create_rectangle(box, n)
    D = levelset.circle_desc(mesh, radius, center, order)
    assert harness.ls_sizeof_grid() == C.sizeof(_lib.CircleDesc)
    nc = mesh.num_cells
    cls, nb, ns = (np.zeros(nc, dtype=np.int32) for _ in range(3))
    ip = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    harness.ls_classify_all(C.byref(D), ip(cls), ip(nb), ip(ns))
    cut, uncut, outside, qp, qw, qpb, qwb, qn = levelset.generate_qr_circle(mesh, radius, center, order)
    assert np.array_equal(np.nonzero(cls == 1)[0], cut)
    assert np.array_equal(np.nonzero(cls == 2)[0], uncut)
    assert np.array_equal(np.nonzero(cls == 0)[0], outside)
    assert nb[cls != 1].sum() == 0 and ns[cls != 1].sum() == 0
    for k, cell in enumerate(cut):
        assert nb[cell] == qw[k].size and ns[cell] == qwb[k].size, cell
        pts, w = np.zeros(2 * nb[cell]), np.zeros(nb[cell])
        pb, wb, nn = np.zeros(2 * ns[cell] + 1), np.zeros(ns[cell] + 1), np.zeros(2 * ns[cell] + 1)
        harness.ls_rules_cell(C.byref(D), C.c_int64(int(cell)), ip(pts), ip(w), ip(pb), ip(wb), ip(nn))
        np.testing.assert_allclose(pts, qp[k], rtol=0, atol=1e-12)
        np.testing.assert_allclose(w, qw[k], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(pb[:-1], qpb[k], rtol=0, atol=1e-12)
        # the arc ends are angles rounded to 14 decimals (numpy.round there, rint here): 1e-14 / (arc length) relative
        np.testing.assert_allclose(wb[:-1], qwb[k], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(nn[:-1], qn[k], rtol=0, atol=1e-12)


SPHERE_CASES = [
    # (box, n, radius, center, order, axis)
    ([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [6, 6, 6], 1.0, (0.0, 0.0, 0.0), 4, 2),   # demo/poisson.py:48-49 extents
    ([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [9, 7, 8], 0.83, (0.1, -0.2, 0.05), 4, 2),
    ([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], [4, 4, 4], 1.0, (0.0, 0.0, 0.0), 2, 2),         # through cell corners, tangent to the box
    ([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [5, 6, 7], 0.9, (0.0, 0.1, 0.0), 6, 0),   # sliced along x
    ([[-2.0, -2.0, -2.0], [2.0, 2.0, 2.0]], [3, 3, 3], 0.4, (0.05, 0.1, -0.02), 4, 1),      # whole ball inside one cell, sliced along y
]


@pytest.mark.parametrize("box,n,radius,center,order,axis", SPHERE_CASES)
def test_device_logic_matches_host_producer_sphere(harness, box, n, radius, center, order, axis):
    from customquad_b200 import _lib

    mesh = cq.create_box(box, n)
    D = levelset.sphere_desc(mesh, radius, center, order, axis)
    assert harness.ls_sizeof_grid3() == C.sizeof(_lib.SphereDesc)
    nc = mesh.num_cells
    cls, nb, ns = (np.zeros(nc, dtype=np.int32) for _ in range(3))
    ip = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    harness.ls_classify_all3(C.byref(D), ip(cls), ip(nb), ip(ns))
    cut, uncut, outside, qp, qw, qpb, qwb, qn = levelset.generate_qr_sphere(mesh, radius, center, order, axis)
    assert np.array_equal(np.nonzero(cls == 1)[0], cut)
    assert np.array_equal(np.nonzero(cls == 2)[0], uncut)
    assert np.array_equal(np.nonzero(cls == 0)[0], outside)
    hmin = float(np.min(mesh.h()))
    for k, cell in enumerate(cut):
        assert nb[cell] == qw[k].size and ns[cell] == qwb[k].size, cell
        pts, w = np.zeros(3 * nb[cell]), np.zeros(nb[cell])
        pb, wb, nn = np.zeros(3 * ns[cell] + 1), np.zeros(ns[cell] + 1), np.zeros(3 * ns[cell] + 1)
        harness.ls_rules_cell3(C.byref(D), C.c_int64(int(cell)), ip(pts), ip(w), ip(pb), ip(wb), ip(nn))
        np.testing.assert_allclose(pts, qp[k], rtol=0, atol=1e-12)
        np.testing.assert_allclose(w, qw[k], rtol=1e-11, atol=1e-15)
        np.testing.assert_allclose(pb[:-1], qpb[k], rtol=0, atol=1e-12 + 2e-14 * radius / hmin)
        np.testing.assert_allclose(wb[:-1], qwb[k], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(nn[:-1], qn[k], rtol=0, atol=1e-12)
