"""Pins the CPU oracle against every known answer the reference's test-suite holds for the
runtime-quadrature path: the oracle may only be trusted as the checker of the
CUDA path after these pass."""
import numpy as np
import pytest

from tests import refcases

NUM_ASSEMBLY = 16 + 20  # test_quads_assembly meshes + test_hexes_assembly meshes


@pytest.mark.parametrize("idx", range(NUM_ASSEMBLY))
def test_quads_hexes_assembly(oracle_backend, golden, idx):
    assert int(golden["assembly/count"]) == NUM_ASSEMBLY
    refcases.assembly_case(oracle_backend, golden, idx)


def test_edges(oracle_backend, golden):
    refcases.edges_case(oracle_backend, golden)


def test_edge_integral(oracle_backend, golden):
    refcases.edge_integral_case(oracle_backend, golden)


def test_face_integral(oracle_backend, golden):
    refcases.face_integral_case(oracle_backend, golden)


def test_volume(oracle_backend, golden):
    refcases.volume_case(oracle_backend, golden)


def test_full_area(oracle_backend, golden):
    refcases.full_area_case(oracle_backend, golden)


def test_faithful_tables_equal_flat_tables():
    """The 'structure-faithful' flavour (one call_basix with heap tables per table per cell,
    call_basix.h:49-57) and the allocation-free flavour give identical element tensors."""
    import customquad_b200 as cq
    from customquad_b200 import forms
    from customquad_b200.quadrature import synthetic_ragged
    from oracle import oracle

    for mk, n in ((cq.create_rectangle, [3, 2]), (cq.create_box, [2, 2, 2])):
        lo = [-0.3, 0.1, 0.2][: len(n)]
        hi = [1.1, 2.3, 1.7][: len(n)]
        mesh = mk([lo, hi], n)
        for deg in (1, 2):
            V = cq.FunctionSpace(mesh, ("Lagrange", deg))
            qr = synthetic_ragged(np.arange(mesh.num_cells), mesh.geometry.dim, 3, 9, normals=True)
            for form in (forms.mass(V), forms.stiffness(V), forms.nitsche(V, 10.0)):
                a = oracle.element_tensors(form, qr, faithful=True)
                b = oracle.element_tensors(form, qr, faithful=False)
                assert np.array_equal(a, b)


def test_oracle_threads_match_serial():
    import customquad_b200 as cq
    from customquad_b200 import forms
    from customquad_b200.quadrature import synthetic_ragged
    from oracle import oracle

    mesh = cq.create_rectangle([[0, 0], [1, 2]], [9, 7])
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    qr = synthetic_ragged(np.arange(mesh.num_cells), 2, 4, 12)
    A1 = oracle.assemble_matrix(forms.stiffness(V), [qr], nthreads=1)
    A4 = oracle.assemble_matrix(forms.stiffness(V), [qr], nthreads=4)
    assert np.array_equal(A1.indices, A4.indices) and np.array_equal(A1.indptr, A4.indptr)
    assert abs(A1 - A4).max() <= 1e-13 * abs(A1).max()
