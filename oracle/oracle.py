"""ctypes binding of the CPU oracle (oracle/cq_oracle.c).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (customquad_b200/) never does.

The three entry points have the reference's signatures
(reference customquad/assemble_matrix.py:10, assemble_vector.py:8, assemble_scalar.py:8):
``assemble_*(form, qr_data)`` with ``qr_data`` a list with one tuple per cell integral, and
follow the same host steps: get_dofs / get_vertices, pack coefficients and constants, create
and zero the matrix, fetch the kernel, run assemble_cells.  Results are host objects
(scipy CSR, numpy vector, float).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "cq_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "cq_b200.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        tmp = f"{so}.{os.getpid()}.tmp"  # several test processes may rebuild at once: write aside, rename atomically
        base = ["gcc", "-O2", "-g", "-fPIC", "-Wall", "-std=c11", "-D_GNU_SOURCE", "-shared", "-o", tmp, src, "-lm"]
        r = subprocess.run(base[:1] + ["-fopenmp"] + base[1:], capture_output=True, text=True)
        if r.returncode != 0:  # toolchain without libgomp: serial oracle
            subprocess.run(base, check=True)
        os.replace(tmp, so)
    return so


class _Args(C.Structure):
    _fields_ = [
        ("form", C.c_int), ("cell_type", C.c_int), ("degree", C.c_int), ("bs", C.c_int),
        ("fn_kind", C.c_int), ("faithful", C.c_int), ("nthreads", C.c_int),
        ("vertices", C.c_void_p), ("coords", C.c_void_p), ("dofmap", C.c_void_p),
        ("coeffs", C.c_void_p), ("ncoef", C.c_int), ("consts", C.c_void_p),
        ("num_entries", C.c_int64), ("cells", C.c_void_p), ("offsets", C.c_void_p),
        ("pts", C.c_void_p), ("w", C.c_void_p), ("normals", C.c_void_p),
    ]  # fmt: skip


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        # argtypes are mandatory: without them ctypes passes Python ints as 32-bit C ints and
        # truncates the addresses of large (mmap'd) numpy arrays.
        P = C.c_void_p
        L.cq_oracle_assemble_matrix.restype = C.c_int64
        L.cq_oracle_assemble_matrix.argtypes = [C.POINTER(_Args), P, P, P]
        L.cq_oracle_assemble_vector.restype = C.c_int64
        L.cq_oracle_assemble_vector.argtypes = [C.POINTER(_Args), P]
        L.cq_oracle_assemble_scalar.restype = C.c_int64
        L.cq_oracle_assemble_scalar.argtypes = [C.POINTER(_Args), P]
        L.cq_oracle_element_tensors.restype = C.c_int64
        L.cq_oracle_element_tensors.argtypes = [C.POINTER(_Args), C.c_int, P]
        L.cq_oracle_max_threads.restype = C.c_int
        L.cq_oracle_max_threads.argtypes = []
        L.cq_oracle_tabulate.restype = C.c_int
        L.cq_oracle_tabulate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.cq_oracle_pattern_build.restype = C.c_void_p
        L.cq_oracle_pattern_build.argtypes = [P, C.c_int64, C.c_int, C.c_int64, P, C.c_int64]
        L.cq_oracle_pattern_nnz.restype = C.c_int64
        L.cq_oracle_pattern_nnz.argtypes = [P, C.c_int]
        L.cq_oracle_pattern_export.restype = None
        L.cq_oracle_pattern_export.argtypes = [P, C.c_int, P, P]
        L.cq_oracle_pattern_free.restype = None
        L.cq_oracle_pattern_free.argtypes = [P]
        if os.environ.get("CQ_ORACLE_BACKTRACE"):
            L.cq_oracle_install_segv_handler()
        _LIB = L
    return _LIB


def max_threads() -> int:
    return int(lib().cq_oracle_max_threads())


# ---- the oracle's own input normalisation and sparsity (independent of customquad_b200/quadrature.py, sparsity.py) ----
class _Rule:
    """One qr_data tuple as the C loop wants it: what ``for k, cell in enumerate(cells): ... qr_pts[k], qr_w[k], qr_n[k]``
    (assemble_matrix.py:65-73,84-101) walks over, laid end to end."""

    def __init__(self, cells, offsets, pts, w, n):
        self.cells, self.offsets, self.pts, self.w, self.n = cells, offsets, pts, w, n
        self.num_entries = int(cells.shape[0])


def _rule(qr, gdim: int) -> _Rule:
    if all(hasattr(qr, a) for a in ("cells", "offsets", "pts", "w", "n")):  # already flat arrays (data, no logic involved)
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
        return _Rule(np.ascontiguousarray(qr.cells, dtype=np.int32), np.ascontiguousarray(qr.offsets, dtype=np.int64), f64(qr.pts), f64(qr.w), f64(qr.n))
    assert len(qr) in (3, 4)  # assemble_matrix.py:65-69
    cells = [int(c) for c in qr[0]]
    qr_pts, qr_w = qr[1], qr[2]
    qr_n = qr[3] if len(qr) == 4 else None
    assert len(cells) == len(qr_pts) and len(cells) == len(qr_w)  # :70-73
    if qr_n is not None:
        assert len(cells) == len(qr_n)
    offsets = np.zeros(len(cells) + 1, dtype=np.int64)
    pts, w, n = [], [], []
    for k in range(len(cells)):  # the reference's loop order (:84)
        wk = np.asarray(qr_w[k], dtype=np.float64).reshape(-1)
        pk = np.asarray(qr_pts[k], dtype=np.float64).reshape(-1)
        assert pk.size == gdim * wk.size, "qr_pts[k] must hold gdim coordinates per weight"
        offsets[k + 1] = offsets[k] + wk.size  # nq = len(qr_w[k]) (:87)
        pts.append(pk)
        w.append(wk)
        if qr_n is not None:
            nk = np.asarray(qr_n[k], dtype=np.float64).reshape(-1)
            assert nk.size == pk.size
            n.append(nk)
    cat = lambda parts: np.concatenate(parts) if parts else np.zeros(0)  # noqa: E731
    return _Rule(np.asarray(cells, dtype=np.int32), offsets, cat(pts), cat(w), cat(n) if qr_n is not None else None)


class Pattern:
    def __init__(self, indptr, indices, num_rows):
        self.indptr, self.indices, self.num_rows = indptr, indices, num_rows

    @property
    def nnz(self) -> int:
        return int(self.indptr[-1])


def build_pattern(dofmap, num_dofs: int, bs: int = 1, cells=None) -> Pattern:
    """dolfinx.cpp.fem.petsc.create_matrix(form) restated in oracle/cq_oracle.c (insertion-sorted rows)."""
    L = lib()
    d = np.ascontiguousarray(dofmap, dtype=np.int32)
    c = None if cells is None else np.ascontiguousarray(cells, dtype=np.int32)
    h = L.cq_oracle_pattern_build(_ptr(d), d.shape[0], d.shape[1], int(num_dofs), _ptr(c), 0 if c is None else c.shape[0])
    try:
        nnz = int(L.cq_oracle_pattern_nnz(h, int(bs)))
        indptr = np.zeros(int(num_dofs) * bs + 1, dtype=np.int64)
        indices = np.zeros(nnz, dtype=np.int32)
        L.cq_oracle_pattern_export(h, int(bs), _ptr(indptr), _ptr(indices))
    finally:
        L.cq_oracle_pattern_free(h)
    return Pattern(indptr, indices, int(num_dofs) * bs)


def _ptr(a):
    return None if a is None else a.ctypes.data


def _prepare(form, qr, nthreads, faithful):
    """Host steps of the reference's assemble_* drivers (assemble_matrix.py:11-17)."""
    mesh = form.mesh
    gdim = mesh.geometry.dim
    q = _rule(qr, gdim)
    keep = {}
    keep["vertices"] = np.ascontiguousarray(mesh.geometry.dofmap, dtype=np.int32)  # utils.get_vertices
    keep["coords"] = np.ascontiguousarray(mesh.geometry.x, dtype=np.float64)
    V = form.function_spaces[0] if form.function_spaces else (form.coefficient.function_space if form.coefficient is not None else None)
    keep["dofmap"] = None if V is None else np.ascontiguousarray(V.dofmap.list, dtype=np.int32)  # utils.get_dofs
    coeffs = None if form.coefficient is None else np.ascontiguousarray(form.coefficient.packed())  # pack_coefficients
    keep["coeffs"] = coeffs
    consts = form.consts()  # pack_constants
    keep["consts"] = consts if consts.size else np.zeros(1)
    keep["q"] = q
    a = _Args(
        int(form.kind), int(mesh.cell_type), int(form.degree), int(form.bs), int(form.fn_kind), int(faithful), int(nthreads),
        _ptr(keep["vertices"]), _ptr(keep["coords"]), _ptr(keep["dofmap"]),
        _ptr(coeffs), 0 if coeffs is None else coeffs.shape[1], _ptr(keep["consts"]),
        q.num_entries, _ptr(q.cells), _ptr(q.offsets), _ptr(q.pts), _ptr(q.w), _ptr(q.n),
    )  # fmt: skip
    if q.num_entries:
        assert q.cells.min() >= 0 and q.cells.max() < mesh.num_cells, "cell index out of range"
    return a, keep, V


def assemble_matrix(form, qr_data, nthreads: int = 1, faithful: bool = False, pattern=None):
    """reference customquad/assemble_matrix.py:10-47.  Returns scipy.sparse.csr_matrix whose
    structure is the full form pattern (create_matrix(form), :19-20)."""
    import scipy.sparse as sp

    assert form.rank == 2
    V = form.function_spaces[0]
    if pattern is None:
        pattern = build_pattern(V.dofmap.list, V.dofmap.num_dofs, V.bs)
    values = np.zeros(pattern.nnz)  # A.zeroEntries()
    for i, _id in enumerate(form.integral_ids()):  # one qr tuple per integral, by position (:26,42)
        a, keep, _ = _prepare(form, qr_data[i], nthreads, faithful)
        missing = lib().cq_oracle_assemble_matrix(C.byref(a), _ptr(pattern.indptr), _ptr(pattern.indices), _ptr(values))
        assert missing == 0, f"{missing} entries outside the sparsity pattern"
    idx_t = np.int32 if pattern.nnz < 2**31 else np.int64
    return sp.csr_matrix((values, pattern.indices, pattern.indptr.astype(idx_t)), shape=(pattern.num_rows, pattern.num_rows))


def assemble_vector(form, qr_data, nthreads: int = 1, faithful: bool = False) -> np.ndarray:
    """reference customquad/assemble_vector.py:8-49."""
    assert form.rank == 1
    V = form.function_spaces[0]
    b = np.zeros(V.num_dofs)  # create_vector(index_map, bs) (:27)
    for i, _id in enumerate(form.integral_ids()):
        a, keep, _ = _prepare(form, qr_data[i], nthreads, faithful)
        lib().cq_oracle_assemble_vector(C.byref(a), _ptr(b))
    return b


def assemble_scalar(form, qr_data, nthreads: int = 1, faithful: bool = False) -> float:
    """reference customquad/assemble_scalar.py:8-34."""
    assert form.rank == 0
    m = np.zeros(1)
    for i, _id in enumerate(form.integral_ids()):
        a, keep, _ = _prepare(form, qr_data[i], nthreads, faithful)
        lib().cq_oracle_assemble_scalar(C.byref(a), _ptr(m))
    return float(m[0])


def element_tensors(form, qr, nthreads: int = 1, faithful: bool = False) -> np.ndarray:
    """Un-scattered element tensors [num_entries, len] (one generated-kernel call per entry)."""
    a, keep, V = _prepare(form, qr, nthreads, faithful)
    n = (form.degree + 1) ** form.mesh.geometry.dim * form.bs
    ln = {2: n * n, 1: n, 0: 1}[form.rank]
    out = np.zeros((keep["q"].num_entries, ln))
    lib().cq_oracle_element_tensors(C.byref(a), form.rank, _ptr(out))
    return out


def tabulate(cell_type: int, degree: int, basix_derivative: int, pts: np.ndarray) -> np.ndarray:
    """call_basix (reference customquad/call_basix.h:14-81) on [nq, gdim] points -> [nq, nbf]."""
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    nq, gdim = pts.shape
    nbf = (degree + 1) ** gdim
    out = np.zeros((nq, nbf))
    r = lib().cq_oracle_tabulate(int(cell_type), degree, basix_derivative, nq, _ptr(pts), _ptr(out))
    assert r == nbf
    return out
