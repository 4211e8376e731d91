"""Owner-computes path of blocked Q2-hexahedron matrices (csrc/cq_owner.cuh; config 4: reference
demo/vector_poisson.py:95,121-127): every blocked row is stored once by the warp that owns its dof instead of being
RED-scattered into a zeroed matrix.  Held against the CPU oracle and against the RED path of the same library."""
import numpy as np
import pytest

import customquad_b200 as cq
from customquad_b200 import forms
from customquad_b200.assemble import set_variant
from customquad_b200.quadrature import synthetic_ragged
from tests.test_gpu_parity import close, make_mesh

pytestmark = pytest.mark.gpu


def _space(mesh, bs):
    return cq.VectorFunctionSpace(mesh, ("Lagrange", 2)) if bs == 3 else cq.FunctionSpace(mesh, ("Lagrange", 2), bs)


def _rule(mesh, rng, normals=True):
    """cell subset (rows nobody touches must still come out zero), repeats, a few zero-point entries"""
    cells = np.concatenate([rng.permutation(mesh.num_cells)[: mesh.num_cells * 2 // 3], rng.integers(0, mesh.num_cells, 7)])
    while True:
        qr = synthetic_ragged(cells, 3, 0, 31, seed=int(rng.integers(1 << 30)), normals=normals)
        if (np.diff(qr.offsets) == 0).any():
            return qr


@pytest.mark.parametrize("bs", [2, 3])
@pytest.mark.parametrize("distort", [False, True])
def test_overwrite_untouched_rows_and_accumulate(cuda_backend, oracle_backend, bs, distort):
    rng = np.random.default_rng(31 + bs)
    mesh = make_mesh(3, distort)
    V = _space(mesh, bs)
    q1, q2 = _rule(mesh, rng), _rule(mesh, rng)
    for form in (forms.stiffness(V), forms.mass(V), forms.nitsche(V, 10.0)):
        A = cq.assemble.create_matrix(form)
        A.data.fill_(7.0)  # zero=True has to overwrite every slot, touched or not
        cq.assemble_matrix(form, [q1], A=A, zero=True)
        A.assemble()
        R1 = oracle_backend.matrix(form, [q1])
        S = A.to_scipy()
        assert np.array_equal(S.indptr, R1.indptr) and np.array_equal(S.indices, R1.indices)
        assert close(S.data, R1.data), form.kind.name
        assert np.all(S.data[R1.data == 0.0] == 0.0)
        cq.assemble_matrix(form, [q2], A=A)  # second integral: the owner adds to its row
        A.assemble()
        R2 = oracle_backend.matrix(form, [q2])
        assert close(A.to_scipy().data, R1.data + R2.data), form.kind.name


def test_matches_red_path_and_is_reproducible(cuda_backend):
    rng = np.random.default_rng(5)
    mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [9, 7, 6])
    V = cq.VectorFunctionSpace(mesh, ("Lagrange", 2))
    qr = _rule(mesh, rng, normals=False)
    form = forms.stiffness(V)
    a = cq.assemble_matrix(form, [qr]).data.clone()
    b = cq.assemble_matrix(form, [qr]).data.clone()
    assert bool((a == b).all())  # fixed summation order (ascending list position per row)
    try:
        set_variant(2)  # staged kernels only: the RED scatter
        c = cq.assemble_matrix(form, [qr]).data.clone()
    finally:
        set_variant(0)
    assert float((a - c).abs().max()) <= 1e-13 * float(c.abs().max())
