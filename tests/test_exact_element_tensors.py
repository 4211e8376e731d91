"""Oracle (CPU) and CUDA path against closed-form element tensors computed symbolically
(tests/golden/make_exact_tensors.py): degree 2, stiffness, Nitsche with quadrature normals and CellDiameter --
the items the reference's own (all-xfail or absent) tests leave unpinned."""
import pytest

from . import exactcases


@pytest.mark.parametrize("name", exactcases.CASES)
def test_oracle_matches_closed_forms(oracle_backend, name):
    exactcases.check(oracle_backend, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", exactcases.CASES)
def test_cuda_matches_closed_forms(cuda_backend, name):
    exactcases.check(cuda_backend, name)
