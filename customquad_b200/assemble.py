"""``assemble_matrix / assemble_vector / assemble_scalar(form, qr_data)``: the reference's three
entry points (reference customquad/assemble_matrix.py:10-47, assemble_vector.py:8-49,
assemble_scalar.py:8-34) with the cell loop, basix tabulation, generated kernel and
MatSetValuesLocal replaced by one CUDA launch per cell integral through the C-ABI
(include/cq_b200.h).  Host code only stages arrays; PyTorch is used for device buffers."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .forms import Form
from .la import CSRMatrix, Vector
from .quadrature import RaggedQuadrature, normalise_qr
from .sparsity import build_pattern, build_pattern_torch


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("customquad_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def _dev():
    torch = _torch()
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# ---- device staging of the read-only inputs (cached on the host objects) ---------------------
def mesh_on_device(mesh, dev):
    """utils.get_vertices (utils.py:66-71) on the device."""
    torch = _torch()
    if dev not in mesh._dev:
        gd = torch.from_numpy(np.ascontiguousarray(mesh.geometry.dofmap, dtype=np.int32)).to(dev)
        xc = torch.from_numpy(np.ascontiguousarray(mesh.geometry.x, dtype=np.float64)).to(dev)
        # setup-time classification: are all cells axis-aligned (constant diagonal Jacobian)?
        flags = C.c_int32(0)
        md = _lib.MeshDesc(mesh.num_cells, xc.shape[0], _ptr(gd), _ptr(xc), 0)
        _lib.check(_lib.load().cq_mesh_classify(_lib.ctx(dev.index), int(mesh.cell_type), C.byref(md), C.byref(flags), _stream_ptr()))
        mesh._dev[dev] = (gd, xc, int(flags.value))
    return mesh._dev[dev]


def space_on_device(V, dev):
    """utils.get_dofs (utils.py:33-63), kept un-blocked."""
    torch = _torch()
    if dev not in V._dev:
        V._dev[dev] = torch.from_numpy(np.ascontiguousarray(V.dofmap.list, dtype=np.int32)).to(dev)
    return V._dev[dev]


def pattern_on_device(V, dev, cells=None):
    """dolfinx.cpp.fem.petsc.create_matrix(form) (assemble_matrix.py:19): the CSR structure."""
    torch = _torch()
    key = ("pattern", dev, None if cells is None else hash(np.asarray(cells).tobytes()))
    if key not in V._dev:
        d = V.dofmap.list
        if cells is None and d.shape[0] * d.shape[1] ** 2 > 5e7:  # big mesh: build on the device
            indptr, indices = build_pattern_torch(space_on_device(V, dev), V.dofmap.num_dofs, V.bs)
        else:
            p = build_pattern(d, V.dofmap.num_dofs, V.bs, cells)
            indptr, indices = torch.from_numpy(p.indptr).to(dev), torch.from_numpy(p.indices).to(dev)
        V._dev[key] = (indptr, indices)
    return V._dev[key]


def scatter_map_on_device(V, dev, A):
    """cq_build_scatter_map: per-cell positions of the element tensor entries inside their CSR
    rows, built once per (space, pattern) -- the stand-in for the local-to-global mapping and
    preallocation a PETSc Mat carries.  Returns None when a row is too long for the uint8 map
    (the kernels then search the rows like MatSetValuesLocal does).

    The map of the form's full pattern is shared by every matrix of the space (cached on the space, next to the
    pattern); a matrix on its own pattern (facet couplings) keeps its map itself, so that a map can never outlive
    the pattern tensors it was built for."""
    torch = _torch()
    shared = V._dev.get(("pattern", dev, None))
    if shared is not None and shared[0] is A.indptr:
        cache, key = V._dev, ("scatter_map", dev)
    else:
        cache, key = A.__dict__.setdefault("_cache", {}), ("scatter_map", dev, id(V))
    if key not in cache:
        L = _lib.load()
        stride = L.cq_scatter_map_stride(int(V.mesh.cell_type), int(V.degree))
        dm = space_on_device(V, dev)
        out = torch.empty((dm.shape[0], stride), dtype=torch.uint8, device=dev)
        desc = _lib.CsrDesc(V.num_dofs, int(A.indices.shape[0]), _ptr(A.indptr), _ptr(A.indices), None)
        rc = L.cq_build_scatter_map(_lib.ctx(dev.index), int(V.mesh.cell_type), int(V.degree), int(V.bs), dm.shape[0], _ptr(dm), C.byref(desc), _ptr(out), _stream_ptr())
        if rc == -3:  # CQ_ERR_UNSUPPORTED: rows longer than 255
            out = None
        else:
            _lib.check(rc)
        cache[key] = out
    return cache[key]


def coefficient_on_device(f, dev):
    """dolfinx.cpp.fem.pack_coefficients (assemble_matrix.py:16): [num_cells, n*bs]."""
    torch = _torch()
    V = f.function_space
    d = space_on_device(V, dev)  # int32 indices gather directly (no int64 copy of the dofmap per call)
    arr = getattr(f, "device_array", None)  # a Function whose values live on the device (fem.Function.to_device): no upload
    if arr is None or arr.device != dev:
        arr = torch.from_numpy(f.array).to(dev)
    if V.bs == 1:
        return arr[d].contiguous()
    idx = (d.to(torch.int64)[:, :, None] * V.bs + torch.arange(V.bs, device=dev)[None, None, :]).reshape(d.shape[0], -1)
    return arr[idx].contiguous()


def qr_on_device(q: RaggedQuadrature, dev):
    """The device copy of a rule, made once and kept on the rule object: every later assemble_* call over the same
    rule (another form, another space, the right-hand side) reuses it -- the device-resident handle of the qr tuple."""
    torch = _torch()
    if dev not in q._dev:
        assert q.cells is not None, f"this rule lives on {q.device_arrays()[0]}, not on {dev}"
        up = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        q._dev[dev] = (up(q.cells), up(q.offsets), up(q.pts), up(q.w), up(q.n))
    return q._dev[dev]


L2_TILE_BYTES = int(os.environ.get("CQ_L2_TILE_BYTES", 24 << 20))  # matrix rows one traversal tile may keep live between two cell layers (B200 L2: 126 MB, shared with the rule stream)


def _traversal_hint(mesh, q, cells_dev, A):
    """cq_quadrature.plane_entries / plane_tiles for a matrix assembly over a structured hexahedral mesh whose rule lists
    whole cell layers in order (cells c0, c0 + 1, ... with c0 at the start of a layer): when the rows of the dof planes
    one cell layer scatters into do not fit the L2, walk the list tile by tile (include/cq_b200.h).  (0, 0) otherwise."""
    if A is None or mesh.topology.dim != 3 or len(getattr(mesh, "shape", ())) != 3 or q.num_entries == 0:
        return 0, 0
    per_layer = int(mesh.shape[0]) * int(mesh.shape[1])
    if q.num_entries < 2 * per_layer:
        return 0, 0
    key = ("layers", per_layer)
    if key not in q._dev:  # once per rule: consecutive cells starting at a layer boundary?
        c0 = int(cells_dev[0])
        ok = c0 % per_layer == 0 and int(cells_dev[-1]) - c0 + 1 == q.num_entries and bool((cells_dev[1:] - cells_dev[:-1] == 1).all())
        q._dev[key] = ok
    if not q._dev[key]:
        return 0, 0
    dof_planes = A.shape[0] / max(1, int(mesh.shape[2]))  # rows per cell layer (k dof planes, + 1 shared)
    live = 8.0 * A.nnz / A.shape[0] * dof_planes * 2.0    # bytes of matrix rows one cell layer touches, roughly
    tiles = int(np.ceil(live / L2_TILE_BYTES))
    if tiles <= 1:
        return 0, 0
    tiles = min(tiles, max(1, int(mesh.shape[1]) // 2))  # at least two cell rows per tile
    # whole cell rows per tile when the mesh allows it: the next divisor of the row count
    ny = int(mesh.shape[1])
    while tiles < ny // 2 and ny % tiles:
        tiles += 1
    return per_layer, tiles


# ---- descriptors --------------------------------------------------------------------------------
class _Call:
    """Everything one C-ABI call needs, kept alive for the duration of the call."""

    def __init__(self, form: Form, qr, dev, host_qr: bool = False, A: CSRMatrix | None = None):
        mesh = form.mesh
        gdim = mesh.geometry.dim
        self.q = normalise_qr(qr, gdim)
        q = self.q
        if q.num_entries and q._dev.get("checked_cells") != mesh.num_cells:  # validated once per rule
            cells = q.cells if q.cells is not None else q.device_arrays()[1][0]
            assert int(cells.min()) >= 0 and int(cells.max()) < mesh.num_cells, "cell index out of range"
            q._dev["checked_cells"] = mesh.num_cells
        if form.kind.name.startswith("NITSCHE"):
            assert q.has_normals, "this form needs normals: pass (cells, qr_pts, qr_w, qr_n)"
        self.keep = []
        gd, xc, mflags = mesh_on_device(mesh, dev)
        self.mesh = _lib.MeshDesc(mesh.num_cells, mesh.geometry.x.shape[0], _ptr(gd), _ptr(xc), mflags)
        self.V = form.function_spaces[0] if form.function_spaces else (form.coefficient.function_space if form.coefficient is not None else None)
        self.space = None
        if self.V is not None:
            dm = space_on_device(self.V, dev)
            co = coefficient_on_device(form.coefficient, dev) if form.coefficient is not None else None
            # the map belongs to the full-form pattern; matrices on a "cells" pattern search their rows
            pos = scatter_map_on_device(self.V, dev, A) if (A is not None and A.full_pattern) else None
            self.keep += [dm, co, pos]
            self.space = _lib.SpaceDesc(self.V.dofmap.num_dofs, _ptr(dm), _ptr(co), _ptr(pos))
        consts = form.consts()
        self._consts = (C.c_double * max(1, consts.size))(*consts.tolist())
        self.form = _lib.FormDesc(int(form.kind), int(mesh.cell_type), int(form.degree), int(form.bs), int(form.fn_kind), int(consts.size), self._consts)
        if host_qr:
            q = q.host()
            arrs = (q.cells, q.offsets, q.pts, q.w, q.n)
            hp = lambda a: None if a is None else C.c_void_p(a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr())  # noqa: E731
            self.keep += list(arrs)
            self.qr = _lib.QuadratureDesc(q.num_entries, q.num_points, *[hp(a) for a in arrs])
        else:
            t = qr_on_device(q, dev)
            self.keep += list(t)
            pe, pt = _traversal_hint(mesh, q, t[0], A)
            self.qr = _lib.QuadratureDesc(q.num_entries, q.num_points, *[_ptr(a) for a in t], pe, pt, 0)

    def space_ref(self):
        return None if self.space is None else C.byref(self.space)


def set_fused_zero(on) -> None:
    """The in-kernel zeroEntries of ``assemble_matrix(..., zero=True)`` on this device (cq_ctx_set_fused_zero):
    True = wherever a kernel can (mass / stiffness on quadrilaterals and Q1 hexahedra), False = never (memset first),
    None = the library's default (where it was measured to pay: the Q2 quadrilateral kernels, see include/cq_b200.h)."""
    _lib.check(_lib.load().cq_ctx_set_fused_zero(_lib.ctx(_dev().index), -1 if on is None else int(bool(on))))


def create_matrix(form: Form, pattern_cells=None, zero: bool = True, facets=None) -> CSRMatrix:
    """Zeroed matrix on the form's pattern (create_matrix + zeroEntries, assemble_matrix.py:19-20).
    ``zero=False`` leaves the values uninitialised for ``assemble_matrix(..., zero=True)``.
    ``facets`` [nf, 3] (low cell, high cell, axis; cutfem.ghost_penalty_facets): also the couplings across these
    interior facets, as dolfinx's create_matrix adds them for a form with a dS integral (demo/poisson.py:168-170)."""
    torch = _torch()
    dev = _dev()
    V = form.function_spaces[0]
    indptr, indices = pattern_on_device(V, dev, pattern_cells)
    if facets is not None and len(facets):
        from .sparsity import add_macro_couplings

        f = np.asarray(facets, dtype=np.int64)
        macro = np.concatenate([V.dofmap.list[f[:, 0]], V.dofmap.list[f[:, 1]]], axis=1)
        indptr, indices = add_macro_couplings(indptr, indices, V.dofmap.num_dofs, V.bs, macro)
    data = (torch.zeros if zero else torch.empty)(indices.shape[0], dtype=torch.float64, device=dev)
    n = V.num_dofs
    A = CSRMatrix(indptr, indices, data, (n, n), dev.index, _stream_ptr())
    A.full_pattern = pattern_cells is None
    return A


def assemble_ghost_penalty(V, facets, betas: float, A: CSRMatrix) -> CSRMatrix:
    """A += betas * avg(h) * inner(jump(n, grad u), jump(n, grad v)) * dS over ``facets`` (cq_assemble_ghost_penalty;
    demo/poisson.py:160,168-170, which the reference assembles with dolfinx).  ``A`` must come from
    ``create_matrix(form, facets=facets)``.  Axis-aligned meshes, degree 1 or 2, any block size."""
    torch = _torch()
    dev = _dev()
    gd, xc, mflags = mesh_on_device(V.mesh, dev)
    md = _lib.MeshDesc(V.mesh.num_cells, V.mesh.geometry.x.shape[0], _ptr(gd), _ptr(xc), mflags)
    dm = space_on_device(V, dev)
    sd = _lib.SpaceDesc(V.dofmap.num_dofs, _ptr(dm), None, None)
    f = torch.as_tensor(np.ascontiguousarray(facets, dtype=np.int32), device=dev)
    csr = _csr_desc(A)
    _lib.check(_lib.load().cq_assemble_ghost_penalty(_lib.ctx(dev.index), int(V.mesh.cell_type), int(V.degree), int(V.bs), float(betas), C.byref(md), C.byref(sd),
                                                     int(f.shape[0]), _ptr(f), C.byref(csr), _stream_ptr()))  # fmt: skip
    A.assembled = False
    return A


def peer_zero_plan(keep, peer_target: bool, has_integrals: bool):
    """zeroEntries of a matrix that takes part in a peer scatter: (what to zero BEFORE the fence, does the first assembly
    call zero the rest).  "keep": only the slots a neighbour adds into (``A.zero_keep``), the call zeroes everything else
    (cq_csr.zero_keep_*); "all": the whole matrix, the call only adds; None: nothing -- no neighbour writes into this
    matrix (the lowest rank of the slab partition), so the call may zero all of it after the fence."""
    if keep is not None and has_integrals:
        return "keep", True
    if peer_target or not has_integrals:
        return "all", False
    return None, True


def _csr_desc(A: CSRMatrix):
    return _lib.CsrDesc(A.shape[0], A.nnz, _ptr(A.indptr), _ptr(A.indices), _ptr(A.data),
                        getattr(A, "peer_row0", None), _ptr(getattr(A, "peer_rowstart", None)),
                        _lib.CQ_CSR_PEER_TARGET if getattr(A, "peer_target", False) else 0,
                        getattr(A, "zero_keep", None) or (0, 0))


def assemble_matrix(form: Form, qr_data, A: CSRMatrix | None = None, pattern: str = "form", zero: bool | None = None) -> CSRMatrix:
    """reference customquad/assemble_matrix.py:10-47.  Returns the (device) CSR matrix; like the
    reference's PETSc Mat it is "unassembled" until ``A.assemble()``.

    ``zero``: A.zeroEntries() (:20) as part of the call -- the first integral runs cq_assemble_matrix_zero (memset +
    kernel, or zeroing inside the kernel with ``set_fused_zero``).  Default: True for a matrix created here, False
    (accumulate) for a given A.

    Multi-GPU: a matrix whose rows receive a neighbour rank's peer scatter (``SlabExchange.enable_peer_scatter``; the
    neighbour's kernels add into its interface rows over NVLink at any time between two fences) is zeroed with the
    protocol zero the interface rows -> fence -> zero the rest and assemble (one call, ``cq_csr.zero_keep_*``); the
    caller completes it with ``exchange_matrix`` (the second fence) before reading.
    Every rank has to make the same calls in the same order."""
    assert form.rank == 2
    L = _lib.load()
    dev = _dev()
    ctx = _lib.ctx(dev.index)
    ids = form.integral_ids()
    if A is None:
        cells = None
        if pattern == "cells":
            cells = np.concatenate([np.asarray(normalise_qr(qr_data[i], form.mesh.geometry.dim).cells) for i in range(len(ids))])
        zero = True if zero is None else zero
        A = create_matrix(form, cells, zero=not zero)
    csr = _csr_desc(A)
    if zero and getattr(A, "peer_exchange", None) is not None:
        # peer scatter (either side): the slots a neighbour's kernels add into -- the interface rows, A.zero_keep -- must be
        # zero on every rank before any rank's kernel adds: zero them, fence; the rest of the matrix only this rank's
        # kernels touch, and the first call zeroes it (cq_assemble_matrix_zero skips the kept range)
        keep = getattr(A, "zero_keep", None)
        before, zero = peer_zero_plan(keep, A.peer_target, bool(ids))
        if before == "keep":
            A.data[keep[0] : keep[1]].zero_()
        elif before == "all":
            A.data.zero_()
        A.peer_exchange.fence(dev)
    if zero and not ids:
        A.data.zero_()
    for i, _id in enumerate(ids):  # qr tuples are matched to integrals by position (:26,42)
        c = _Call(form, qr_data[i], dev, A=A)
        fn = L.cq_assemble_matrix_zero if (zero and i == 0) else L.cq_assemble_matrix
        _lib.check(fn(ctx, C.byref(c.form), C.byref(c.mesh), c.space_ref(), C.byref(c.qr), C.byref(csr), _stream_ptr()))
    A.assembled = False
    return A


def assemble_vector(form: Form, qr_data, b: Vector | None = None) -> Vector:
    """reference customquad/assemble_vector.py:8-49."""
    assert form.rank == 1
    torch = _torch()
    L = _lib.load()
    dev = _dev()
    ctx = _lib.ctx(dev.index)
    V = form.function_spaces[0]
    if b is None:
        b = Vector(torch.zeros(V.num_dofs, dtype=torch.float64, device=dev), dev.index)  # create_vector (:27)
    for i, _id in enumerate(form.integral_ids()):
        c = _Call(form, qr_data[i], dev)
        _lib.check(L.cq_assemble_vector(ctx, C.byref(c.form), C.byref(c.mesh), c.space_ref(), C.byref(c.qr), _ptr(b.data), _stream_ptr()))
    return b


def assemble_scalar(form: Form, qr_data) -> float:
    """reference customquad/assemble_scalar.py:8-34: returns a Python float (no MPI reduction)."""
    assert form.rank == 0
    torch = _torch()
    L = _lib.load()
    dev = _dev()
    ctx = _lib.ctx(dev.index)
    m = torch.zeros(1, dtype=torch.float64, device=dev)
    for i, _id in enumerate(form.integral_ids()):
        c = _Call(form, qr_data[i], dev)
        _lib.check(L.cq_assemble_scalar(ctx, C.byref(c.form), C.byref(c.mesh), c.space_ref(), C.byref(c.qr), _ptr(m), _stream_ptr()))
    return float(m.item())


# ---- parity / debug hooks -------------------------------------------------------------------------
def element_tensors(form: Form, qr) -> np.ndarray:
    """Un-scattered element tensors [num_entries, len] (cq_element_tensors)."""
    torch = _torch()
    L = _lib.load()
    dev = _dev()
    c = _Call(form, qr, dev)
    n = (form.degree + 1) ** form.mesh.geometry.dim * form.bs
    ln = {2: n * n, 1: n, 0: 1}[form.rank]
    out = torch.zeros((c.q.num_entries, ln), dtype=torch.float64, device=dev)
    _lib.check(L.cq_element_tensors(_lib.ctx(dev.index), C.byref(c.form), C.byref(c.mesh), c.space_ref(), C.byref(c.qr), _ptr(out), _stream_ptr()))
    return out.cpu().numpy()


def tabulate(cell_type: int, degree: int, basix_derivative: int, pts: np.ndarray) -> np.ndarray:
    """call_basix on the device (cq_tabulate): [nq, gdim] points -> [nq, nbf]."""
    torch = _torch()
    L = _lib.load()
    dev = _dev()
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    nq, gdim = pts.shape
    p = torch.from_numpy(pts).to(dev)
    out = torch.zeros((nq, (degree + 1) ** gdim), dtype=torch.float64, device=dev)
    _lib.check(L.cq_tabulate(_lib.ctx(dev.index), int(cell_type), int(degree), int(basix_derivative), nq, _ptr(p), _ptr(out), _stream_ptr()))
    return out.cpu().numpy()


# ---- end-to-end (host buffer) calls -----------------------------------------------------------------
def assemble_matrix_host(form: Form, q: RaggedQuadrature, A: CSRMatrix, values_host) -> None:
    """cq_assemble_matrix_host: qr arrays and the result live in host (pinned) memory; the copies
    are part of the call.  ``q`` holds numpy arrays or pinned torch CPU tensors; ``values_host``
    is a float64 host buffer of A.nnz entries."""
    L = _lib.load()
    dev = _dev()
    c = _Call(form, q, dev, host_qr=True, A=A)
    csr = _csr_desc(A)
    vh = C.c_void_p(values_host.ctypes.data if isinstance(values_host, np.ndarray) else values_host.data_ptr())
    _lib.check(L.cq_assemble_matrix_host(_lib.ctx(dev.index), C.byref(c.form), C.byref(c.mesh), c.space_ref(), C.byref(c.qr), C.byref(csr), vh, _stream_ptr()))


def assemble_matrices_host(forms, q: RaggedQuadrature, mats, values_host) -> None:
    """cq_assemble_matrices_host: several bilinear forms over ONE qr tuple in host memory (the reference's demo
    integrates its bulk rule into a matrix and a vector, demo/poisson.py:193,209).  The rule crosses PCIe once; every
    matrix is zeroed, assembled and copied to its host buffer (``values_host[j]``, or None to keep it on the device)."""
    L = _lib.load()
    dev = _dev()
    calls = [_Call(f, q, dev, host_qr=True, A=A) for f, A in zip(forms, mats)]
    csrs = [_csr_desc(A) for A in mats]
    jobs = (_lib.MatrixJob * len(calls))()
    for j, (c, csr, vh) in enumerate(zip(calls, csrs, values_host)):
        jobs[j].form = C.pointer(c.form)
        jobs[j].space = C.pointer(c.space)
        jobs[j].A = C.pointer(csr)
        jobs[j].values_host = None if vh is None else (vh.ctypes.data if isinstance(vh, np.ndarray) else vh.data_ptr())
    _lib.check(L.cq_assemble_matrices_host(_lib.ctx(dev.index), len(calls), jobs, C.byref(calls[0].mesh), C.byref(calls[0].qr), _stream_ptr()))


def assemble_vector_host(form: Form, q: RaggedQuadrature, b: Vector, b_host) -> None:
    L = _lib.load()
    dev = _dev()
    c = _Call(form, q, dev, host_qr=True)
    bh = C.c_void_p(b_host.ctypes.data if isinstance(b_host, np.ndarray) else b_host.data_ptr())
    _lib.check(L.cq_assemble_vector_host(_lib.ctx(dev.index), C.byref(c.form), C.byref(c.mesh), c.space_ref(), C.byref(c.qr), _ptr(b.data), bh, _stream_ptr()))


def assemble_scalar_host(form: Form, q: RaggedQuadrature) -> float:
    L = _lib.load()
    dev = _dev()
    c = _Call(form, q, dev, host_qr=True)
    m = np.zeros(1)
    _lib.check(L.cq_assemble_scalar_host(_lib.ctx(dev.index), C.byref(c.form), C.byref(c.mesh), c.space_ref(), C.byref(c.qr), C.c_void_p(m.ctypes.data), _stream_ptr()))
    return float(m[0])


def set_variant(variant: int) -> None:
    """Kernel choice for the current device: 0 auto (staged bulk-copy kernel when applicable),
    1 direct-load kernel only, 2 staged kernel only (raises when not applicable)."""
    dev = _dev()
    _lib.check(_lib.load().cq_ctx_set_variant(_lib.ctx(dev.index), int(variant)))
