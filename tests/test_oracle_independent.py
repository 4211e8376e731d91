"""The oracle stands on its own feet: its sparsity pattern (oracle/cq_oracle.c, insertion-sorted rows like a PETSc
preallocation) and its qr_data normalisation (oracle/oracle.py:_rule, the reference's loop order) are separate
implementations from the product's (customquad_b200/sparsity.py, quadrature.py).  Here the two are held against each
other -- and against brute force -- so that "bit-exact sparsity" in the GPU parity tests compares two builders."""
import numpy as np
import pytest

import customquad_b200 as cq
from customquad_b200.quadrature import normalise_qr
from customquad_b200.sparsity import build_pattern
from oracle import oracle

MESHES = [
    ("quad", [[-1.0, -2.0], [3.0, 1.0]], [5, 4]),
    ("hex", [[0.0, 0.0, 0.0], [1.0, 2.0, 3.0]], [3, 2, 4]),
]


def _mesh(kind, pts, n):
    return (cq.create_rectangle if kind == "quad" else cq.create_box)(pts, n)


def _brute(dofmap, num_dofs, bs, cells):
    rows = [set() for _ in range(num_dofs)]
    for c in range(dofmap.shape[0]) if cells is None else cells:
        for i in dofmap[c]:
            rows[i].update(int(j) for j in dofmap[c])
    indptr, indices = [0], []
    for r in range(num_dofs):
        for _a in range(bs):
            for col in sorted(rows[r]):
                indices += [col * bs + b for b in range(bs)]
            indptr.append(len(indices))
    return np.asarray(indptr, dtype=np.int64), np.asarray(indices, dtype=np.int32)


@pytest.mark.parametrize("kind,pts,n", MESHES)
@pytest.mark.parametrize("degree", [1, 2])
@pytest.mark.parametrize("bs", [1, 2, 3])
@pytest.mark.parametrize("subset", [False, True])
def test_oracle_pattern_equals_product_pattern_and_brute_force(kind, pts, n, degree, bs, subset):
    mesh = _mesh(kind, pts, n)
    V = cq.FunctionSpace(mesh, ("Lagrange", degree), bs)
    cells = np.array([0, mesh.num_cells - 1, 3, 3, 1], dtype=np.int32) if subset else None
    o = oracle.build_pattern(V.dofmap.list, V.dofmap.num_dofs, bs, cells)
    p = build_pattern(V.dofmap.list, V.dofmap.num_dofs, bs, cells)
    ip, ix = _brute(V.dofmap.list, V.dofmap.num_dofs, bs, cells)
    assert np.array_equal(o.indptr, ip) and np.array_equal(o.indices, ix)
    assert np.array_equal(p.indptr, ip) and np.array_equal(p.indices, ix)
    assert o.nnz == p.nnz == len(ix)


@pytest.mark.parametrize("gdim", [2, 3])
@pytest.mark.parametrize("with_normals", [False, True])
@pytest.mark.parametrize("shape", ["lists", "rows", "python_cells"])
def test_oracle_rule_equals_product_normalise_qr(gdim, with_normals, shape):
    rng = np.random.default_rng(gdim * 10 + with_normals)
    C = 9
    cells = rng.integers(0, 50, C)
    if shape == "rows":  # 2-D arrays, one row per cell (test/common.py:51-52)
        nq = np.full(C, 4)
    else:  # lists of flat arrays with different lengths (demo/poisson.py:109-113); an empty rule in the middle
        nq = rng.integers(1, 7, C)
        nq[4] = 0
    pts = [rng.random(gdim * k) for k in nq]
    w = [rng.random(k) for k in nq]
    nrm = [rng.standard_normal(gdim * k) for k in nq]
    if shape == "rows":
        pts, w, nrm = np.stack(pts), np.stack(w), np.stack(nrm)
    if shape == "python_cells":
        cells = [int(c) for c in cells]  # test_volume_area.py:115-116
    qr = (cells, pts, w, nrm) if with_normals else (cells, pts, w)
    a, b = oracle._rule(qr, gdim), normalise_qr(qr, gdim)
    assert np.array_equal(a.cells, b.cells) and a.cells.dtype == b.cells.dtype == np.int32
    assert np.array_equal(a.offsets, b.offsets)
    assert np.array_equal(a.pts, b.pts) and np.array_equal(a.w, b.w)
    assert (a.n is None) == (b.n is None) and (a.n is None or np.array_equal(a.n, b.n))


def test_oracle_rule_asserts_like_the_reference():
    with pytest.raises(AssertionError):  # assemble_matrix.py:70-73
        oracle._rule(([0, 1], [np.zeros(2)], [np.zeros(1), np.zeros(1)]), 2)
    with pytest.raises(AssertionError):
        oracle._rule(([0], [np.zeros(3)], [np.zeros(1)]), 2)
