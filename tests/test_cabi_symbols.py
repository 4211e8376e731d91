"""The C-ABI library loads and exports every symbol include/cq_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cq_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("cq_assemble_matrix", "cq_assemble_vector", "cq_assemble_scalar", "cq_ctx_create", "cq_ctx_destroy", "cq_last_error",
              "cq_assemble_matrix_host", "cq_csr_accumulate_rows", "cq_tabulate", "cq_element_tensors", "cq_exchange_offrank",
              "cq_assemble_matrices_host", "cq_comm_init", "cq_peer_fence"):  # fmt: skip
        assert s in syms


def test_library_exports_every_declared_symbol():
    from customquad_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/cq_b200.h but not exported"
    L.cq_abi_version.restype = ctypes.c_int
    # 2: cq_csr.peer_row0 / peer_rowstart, cq_ipc_*; 3: cq_assemble_matrix_zero, cq_ctx_set_fused_zero, cq_circle_*, cq_sphere_*;
    # 4: cq_csr.flags, cq_assemble_matrices_host, cq_comm_*, cq_exchange_offrank*, cq_allreduce_sum, cq_peer_fence*;
    # 5: cq_csr.zero_keep_lo / zero_keep_hi, cq_ctx_set_fused_zero(-1)
    assert L.cq_abi_version() == 5
    assert set(_lib._SIGNATURES) == set(declared_symbols())


def test_ctx_create_fails_loudly_without_gpu():
    import torch

    from customquad_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.CqError):
        _lib.ctx(0)
