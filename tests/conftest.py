import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Known answers of the reference's own tests (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_known_answers.npz"))


class Backend:
    """Uniform view of an assembler implementation: results as host scipy / numpy / float."""

    def __init__(self, name, mod):
        self.name, self.mod = name, mod

    def matrix(self, form, qr_data):
        A = self.mod.assemble_matrix(form, qr_data)
        if hasattr(A, "assemble"):
            A.assemble()  # README.md:94-95
        return A.to_scipy() if hasattr(A, "to_scipy") else A

    def vector(self, form, qr_data):
        b = self.mod.assemble_vector(form, qr_data)
        return b.array if hasattr(b, "array") else b

    def scalar(self, form, qr_data):
        return float(self.mod.assemble_scalar(form, qr_data))


@pytest.fixture(scope="session")
def oracle_backend():
    from oracle import oracle

    return Backend("oracle", oracle)


@pytest.fixture(scope="session")
def cuda_backend():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import customquad_b200 as cq

    return Backend("cuda", cq)
