"""ctypes binding of libcq_b200.so (include/cq_b200.h).  No CPU fallback: if the library is not
built, or there is no CUDA device, the first assemble_* call raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CQ_LIB", os.path.join(HERE, "libcq_b200.so"))  # CQ_LIB: experiment builds

CQ_OK = 0
CQ_ERR_SPARSITY = -4


class FormDesc(C.Structure):
    _fields_ = [("form", C.c_int32), ("cell_type", C.c_int32), ("degree", C.c_int32), ("bs", C.c_int32),
                ("fn_kind", C.c_int32), ("nconsts", C.c_int32), ("consts", C.POINTER(C.c_double))]  # fmt: skip


class MeshDesc(C.Structure):
    _fields_ = [("num_cells", C.c_int64), ("num_nodes", C.c_int64), ("geom_dofmap", C.c_void_p), ("coords", C.c_void_p), ("flags", C.c_int32)]


class SpaceDesc(C.Structure):
    _fields_ = [("num_dofs", C.c_int64), ("dofmap", C.c_void_p), ("coeffs", C.c_void_p), ("elem_pos", C.c_void_p)]


class QuadratureDesc(C.Structure):
    _fields_ = [("num_entries", C.c_int64), ("num_points", C.c_int64), ("cells", C.c_void_p), ("offsets", C.c_void_p),
                ("pts", C.c_void_p), ("weights", C.c_void_p), ("normals", C.c_void_p),
                ("plane_entries", C.c_int64), ("plane_tiles", C.c_int32), ("pad_", C.c_int32)]  # fmt: skip


class CsrDesc(C.Structure):
    _fields_ = [("num_rows", C.c_int64), ("nnz", C.c_int64), ("indptr", C.c_void_p), ("indices", C.c_void_p), ("values", C.c_void_p),
                ("peer_row0", C.c_int64), ("peer_rowstart", C.c_void_p), ("flags", C.c_int32), ("pad_", C.c_int32),
                ("zero_keep_lo", C.c_int64), ("zero_keep_hi", C.c_int64)]  # fmt: skip

    def __init__(self, num_rows=0, nnz=0, indptr=None, indices=None, values=None, peer_row0=None, peer_rowstart=None, flags=0, zero_keep=(0, 0)):
        super().__init__(num_rows, nnz, indptr, indices, values, num_rows if peer_row0 is None else peer_row0, peer_rowstart, flags, 0,
                         int(zero_keep[0]), int(zero_keep[1]))


CQ_CSR_PEER_TARGET = 1


class MatrixJob(C.Structure):
    _fields_ = [("form", C.POINTER(FormDesc)), ("space", C.POINTER(SpaceDesc)), ("A", C.POINTER(CsrDesc)), ("values_host", C.c_void_p)]


class CircleDesc(C.Structure):
    _fields_ = [("xmin", C.c_double * 2), ("h", C.c_double * 2), ("n", C.c_int64 * 2), ("center", C.c_double * 2), ("radius", C.c_double),
                ("n1", C.c_int32), ("pad_", C.c_int32), ("g", C.c_double * 8), ("gw", C.c_double * 8)]  # fmt: skip


class SphereDesc(C.Structure):
    _fields_ = [("xmin", C.c_double * 3), ("h", C.c_double * 3), ("n", C.c_int64 * 3), ("center", C.c_double * 3), ("radius", C.c_double),
                ("n1", C.c_int32), ("axis", C.c_int32), ("g", C.c_double * 8), ("gw", C.c_double * 8), ("gu", C.c_double * 10), ("gwu", C.c_double * 10)]  # fmt: skip


class CqError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libcq_b200 error {code}: {msg}")
        self.code = code


_lib = None
_ctx = {}
_lock = threading.Lock()

# every symbol include/cq_b200.h declares (tests check the exports against the header)
_SIGNATURES = {
    "cq_abi_version": (C.c_int, []),
    "cq_last_error": (C.c_char_p, []),
    "cq_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "cq_ctx_destroy": (C.c_int, [C.c_void_p]),
    "cq_ctx_check": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cq_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "cq_ctx_set_variant": (C.c_int, [C.c_void_p, C.c_int]),
    "cq_ctx_set_fused_zero": (C.c_int, [C.c_void_p, C.c_int]),
    "cq_assemble_matrix": (C.c_int, [C.c_void_p, C.POINTER(FormDesc), C.POINTER(MeshDesc), C.POINTER(SpaceDesc), C.POINTER(QuadratureDesc), C.POINTER(CsrDesc), C.c_void_p]),
    "cq_assemble_matrix_zero": (C.c_int, [C.c_void_p, C.POINTER(FormDesc), C.POINTER(MeshDesc), C.POINTER(SpaceDesc), C.POINTER(QuadratureDesc), C.POINTER(CsrDesc), C.c_void_p]),
    "cq_assemble_vector": (C.c_int, [C.c_void_p, C.POINTER(FormDesc), C.POINTER(MeshDesc), C.POINTER(SpaceDesc), C.POINTER(QuadratureDesc), C.c_void_p, C.c_void_p]),
    "cq_assemble_scalar": (C.c_int, [C.c_void_p, C.POINTER(FormDesc), C.POINTER(MeshDesc), C.POINTER(SpaceDesc), C.POINTER(QuadratureDesc), C.c_void_p, C.c_void_p]),
    "cq_mesh_classify": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(MeshDesc), C.POINTER(C.c_int32), C.c_void_p]),
    "cq_scatter_map_stride": (C.c_int, [C.c_int, C.c_int]),
    "cq_build_scatter_map": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.POINTER(CsrDesc), C.c_void_p, C.c_void_p]),
    "cq_csr_expand_blocks": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cq_assemble_ghost_penalty": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(MeshDesc), C.POINTER(SpaceDesc), C.c_int64, C.c_void_p, C.POINTER(CsrDesc), C.c_void_p]),
    "cq_tabulate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cq_element_tensors": (C.c_int, [C.c_void_p, C.POINTER(FormDesc), C.POINTER(MeshDesc), C.POINTER(SpaceDesc), C.POINTER(QuadratureDesc), C.c_void_p, C.c_void_p]),
    "cq_csr_accumulate_rows": (C.c_int, [C.c_void_p, C.POINTER(CsrDesc), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cq_axpy": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cq_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "cq_ipc_open": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "cq_ipc_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cq_circle_classify": (C.c_int, [C.c_void_p, C.POINTER(CircleDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cq_circle_rules": (C.c_int, [C.c_void_p, C.POINTER(CircleDesc), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cq_sphere_classify": (C.c_int, [C.c_void_p, C.POINTER(SphereDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cq_sphere_rules": (C.c_int, [C.c_void_p, C.POINTER(SphereDesc), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cq_assemble_matrix_host": (C.c_int, [C.c_void_p, C.POINTER(FormDesc), C.POINTER(MeshDesc), C.POINTER(SpaceDesc), C.POINTER(QuadratureDesc), C.POINTER(CsrDesc), C.c_void_p, C.c_void_p]),
    "cq_assemble_vector_host": (C.c_int, [C.c_void_p, C.POINTER(FormDesc), C.POINTER(MeshDesc), C.POINTER(SpaceDesc), C.POINTER(QuadratureDesc), C.c_void_p, C.c_void_p, C.c_void_p]),
    "cq_assemble_scalar_host": (C.c_int, [C.c_void_p, C.POINTER(FormDesc), C.POINTER(MeshDesc), C.POINTER(SpaceDesc), C.POINTER(QuadratureDesc), C.c_void_p, C.c_void_p]),
    "cq_assemble_matrices_host": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(MatrixJob), C.POINTER(MeshDesc), C.POINTER(QuadratureDesc), C.c_void_p]),
    "cq_comm_unique_id": (C.c_int, [C.c_void_p]),
    "cq_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "cq_comm_destroy": (C.c_int, [C.c_void_p]),
    "cq_exchange_offrank": (C.c_int, [C.c_void_p, C.POINTER(CsrDesc), C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]),
    "cq_exchange_offrank_vector": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int64, C.c_void_p]),
    "cq_allreduce_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cq_peer_flags": (C.c_void_p, [C.c_void_p]),
    "cq_peer_fence_setup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cq_peer_fence": (C.c_int, [C.c_void_p, C.c_void_p]),
}  # fmt: skip


def load():
    """dlopen the in-tree library and declare the signatures.  Raises if it is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `python customquad_b200/build.py`); customquad_b200 has no CPU fallback"
                    )
                L = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(L, name)
                    fn.restype, fn.argtypes = res, args
                _lib = L
    return _lib


def last_error() -> str:
    return load().cq_last_error().decode()


def check(rc: int):
    if rc != CQ_OK:
        raise CqError(rc, last_error())


def ctx(device_index: int):
    """One cq_ctx per device per process (one ctx per device, not re-entrant)."""
    L = load()
    if device_index not in _ctx:
        h = C.c_void_p()
        check(L.cq_ctx_create(int(device_index), C.byref(h)))
        _ctx[device_index] = h
    return _ctx[device_index]


def launch_count(device_index: int = 0) -> int:
    return int(load().cq_ctx_launch_count(ctx(device_index)))
