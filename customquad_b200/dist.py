"""Cell-partitioned assembly across the GPUs of one box.

The reference is serial; the only analogue of this step is PETSc's
MatAssemblyBegin/End stash that ``A.assemble()`` triggers (reference README.md:95,
demo/poisson.py:194).  Here: one process per GPU; the mesh is cut into R slabs along the slowest
cell axis (x for quads, z for hexes); rank r assembles the cells of slab r.  Every rank numbers
dofs on its slab *plus one ghost layer of cells below it* (so that its rows can hold the
couplings to the neighbour's dofs), owns the dof lattice lines above that ghost layer, and ships
the contributions it made to the one lattice line it does not own -- the top line of its slab,
owned by rank r+1 -- to that neighbour (NCCL send/recv through torch.distributed), which adds
them into its rows.  Rows are lattice-line-major, so both the ghost rows on the sender and the
interface rows on the receiver are contiguous, and a sender row is a prefix of the receiver's row
(columns are ordered line-major, the sender only knows the lines at or below the interface):
no packing kernel, no index exchange.

Scalars are all-reduced; vectors exchange the same interface line.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .fem import FunctionSpace, FunctionSpaceQ, VectorFunctionSpace  # noqa: F401
from .mesh import Mesh, create_box, create_rectangle


@dataclass
class Slab:
    """What rank `rank` of `world` works on."""

    mesh: Mesh            # slab + ghost layer (ghost cells first along the slab axis)
    cells: np.ndarray     # local ids of the cells this rank assembles (ghost layer excluded)
    first: int            # first global cell layer of the slab
    count: int            # cell layers in the slab
    ghost: int            # 1 if a ghost layer of cells was prepended, else 0
    rank: int
    world: int


def partition_layers(n_layers: int, world: int, weights=None) -> np.ndarray:
    """Slab boundaries [world+1] along the slab axis.  With `weights` (work per layer, e.g. the sum
    of nq*n^2 over the layer's cells) the split balances work, not cell count."""
    assert n_layers >= world, f"{world} ranks need at least {world} cell layers along the slab axis (got {n_layers})"
    if weights is None:
        return np.array([n_layers * r // world for r in range(world + 1)], dtype=np.int64)
    w = np.asarray(weights, dtype=np.float64)
    assert w.shape[0] == n_layers
    cum = np.concatenate([[0.0], np.cumsum(w)])
    b = np.searchsorted(cum, cum[-1] * np.arange(1, world) / world, side="left")
    b = np.concatenate([[0], b, [n_layers]]).astype(np.int64)
    for r in range(1, world + 1):  # every rank keeps at least one layer
        b[r] = max(b[r], b[r - 1] + 1)
    b[world] = n_layers
    for r in range(world - 1, 0, -1):
        b[r] = min(b[r], b[r + 1] - 1)
    return b


def make_slab(points, n, rank: int, world: int, bounds=None) -> Slab:
    """Slab `rank` of the global tensor mesh ``create_rectangle/create_box(points, n)``."""
    p = np.asarray(points, dtype=np.float64)
    n = [int(v) for v in n]
    tdim = len(n)
    axis = 0 if tdim == 2 else 2  # slowest cell axis of the mesh generators
    b = partition_layers(n[axis], world) if bounds is None else np.asarray(bounds)
    first, last = int(b[rank]), int(b[rank + 1])
    assert last > first, f"rank {rank} owns no cell layer (bounds {list(map(int, b))})"
    ghost = 1 if rank > 0 else 0
    lo = first - ghost
    h = (p[1, axis] - p[0, axis]) / n[axis]
    q = p.copy()
    q[0, axis] = p[0, axis] + lo * h
    q[1, axis] = p[0, axis] + last * h
    nn = list(n)
    nn[axis] = last - lo
    mesh = (create_rectangle if tdim == 2 else create_box)(q, nn)
    # The slab's vertices carry the GLOBAL mesh's coordinates bit for bit (same expression as create_rectangle /
    # create_box with the global index).  Generated from the slab's own corner they differ from the neighbour's copy
    # of the same vertices by an ulp, i.e. by eps * |x| / h relative in the cell width: 3e-12 on the interface rows of
    # the 8-GPU weak-scaling mesh (x ~ 20, h = 1.3e-3), seen by the in-bench parity check.
    gx = p[0, axis] + (p[1, axis] - p[0, axis]) * (np.arange(lo, last + 1) / n[axis])
    if tdim == 2:
        mesh.geometry.x[:, 0] = np.repeat(gx, nn[1] + 1)  # vertex id = ix * (ny + 1) + iy
    else:
        mesh.geometry.x.reshape(nn[2] + 1, nn[1] + 1, nn[0] + 1, 3)[..., 2] = gx[:, None, None]  # vertex id = (iz * nvy + iy) * nvx + ix
    mesh.xmin[axis], mesh.xmax[axis] = gx[0], gx[-1]
    per_layer = mesh.num_cells // nn[axis]
    cells = np.arange(ghost * per_layer, mesh.num_cells, dtype=np.int32)  # layers are the slowest cell index
    return Slab(mesh, cells, first, last - first, ghost, rank, world)


class SlabExchange:
    """Interface-row exchange for one matrix (or vector) of a slab-partitioned assembly.

    ``accumulate_rows(values, indptr, row0, nrows, recv_indptr, recv)`` and ``axpy(n, x, y)`` are
    the receive-side adds; by default they are the CUDA kernels behind cq_csr_accumulate_rows /
    cq_axpy (there is no CPU fallback; the gloo tests inject numpy doubles for the host logic).
    """

    def __init__(self, V: FunctionSpace, slab: Slab, indptr=None, accumulate_rows=None, axpy=None, group=None):
        self.V, self.slab, self.group = V, slab, group
        k, bs = V.degree, V.bs
        lshape = V.lattice_shape
        tdim = len(lshape)
        # dofs per lattice line/plane orthogonal to the slab axis (line-major numbering, fem.py)
        self.line = (lshape[1] if tdim == 2 else lshape[0] * lshape[1]) * bs
        nlines = lshape[0] if tdim == 2 else lshape[2]
        self.nrows = nlines * self.line
        self.has_up = slab.rank < slab.world - 1      # sends its top line to rank+1
        self.has_down = slab.rank > 0                 # receives rank-1's top line
        self.send_row0 = (nlines - 1) * self.line     # ghost rows (not owned): the slab's top line
        self.recv_row0 = k * slab.ghost * self.line   # the same line in this rank's numbering
        # rows this rank owns: everything above the ghost layer except the top line (owned by rank+1)
        self.owned_row0 = self.recv_row0 if self.has_down else 0
        self.owned_row1 = self.send_row0 if self.has_up else self.nrows
        self._acc_rows, self._axpy = accumulate_rows, axpy
        self.recv_indptr = None
        if indptr is not None:
            self.set_pattern(indptr)

    # -- matrices ---------------------------------------------------------------------------
    def set_pattern(self, indptr, indices=None):
        """Derive, from this rank's own pattern, the lengths of the rows rank-1 will send: the
        columns of an interface row that lie on lattice lines at or below the interface."""
        self.indptr = indptr
        if self.has_down:
            assert indices is not None, "receiver needs its column indices"
            r0, r1 = self.recv_row0, self.recv_row0 + self.line
            ip = _to_numpy(indptr[r0 : r1 + 1])
            cols = _to_numpy(indices[int(ip[0]) : int(ip[-1])])
            below = cols < r1  # columns on lines <= interface line  (col id = line*line_size + ...)
            cnt = np.add.reduceat(below.astype(np.int64), (ip[:-1] - ip[0])) if len(cols) else np.zeros(self.line, np.int64)
            # prefix property: the `cnt[r]` columns below the interface are the first entries of the row
            # (checked on every interface row: one-time setup)
            starts = ip[:-1] - ip[0]
            pos_in_row = np.arange(len(cols)) - np.repeat(starts, np.diff(ip))
            assert np.array_equal(below, pos_in_row < np.repeat(cnt, np.diff(ip))), "sender rows are not prefixes of the owner's rows"
            self.recv_indptr_host = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        if self.has_up:
            ip = _to_numpy(indptr[self.send_row0 : self.nrows + 1])
            self.send_lo, self.send_hi = int(ip[0]), int(ip[-1])

    def exchange_matrix(self, values, force_nccl=False):
        """values: this rank's CSR data (torch tensor).  After the call the owned rows are complete."""
        import torch
        import torch.distributed as dist

        if getattr(self, "peer", False) and not force_nccl:  # the kernels already scattered into the owner's memory
            self.fence(values.device)
            return
        if self._acc_rows is None and values.is_cuda and ensure_comm(values.device, self.group):
            # the whole exchange inside the library: grouped ncclSend / ncclRecv + accumulate kernel (cq_exchange_offrank)
            import ctypes as C

            from . import _lib

            if self.has_down and (self.recv_indptr is None or self.recv_indptr.device != values.device):
                self.recv_indptr = torch.from_numpy(self.recv_indptr_host).to(values.device)
            A = _lib.CsrDesc(int(self.indptr.shape[0] - 1), int(values.shape[0]), C.c_void_p(self.indptr.data_ptr()), None, C.c_void_p(values.data_ptr()))
            _lib.check(_lib.load().cq_exchange_offrank(
                _lib.ctx(values.device.index), C.byref(A),
                self.slab.rank + 1 if self.has_up else -1, self.send_lo if self.has_up else 0, self.send_hi if self.has_up else 0,
                self.slab.rank - 1 if self.has_down else -1, self.recv_row0, self.line if self.has_down else 0,
                C.c_void_p(self.recv_indptr.data_ptr()) if self.has_down else None, int(self.recv_indptr_host[-1]) if self.has_down else 0,
                C.c_void_p(torch.cuda.current_stream().cuda_stream)))  # fmt: skip
            return

        ops, recv = [], None
        if self.has_up:
            ops.append(dist.P2POp(dist.isend, values[self.send_lo : self.send_hi], self.slab.rank + 1, self.group))
        if self.has_down:
            recv = torch.empty(int(self.recv_indptr_host[-1]), dtype=values.dtype, device=values.device)
            ops.append(dist.P2POp(dist.irecv, recv, self.slab.rank - 1, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if recv is not None:
            if self.recv_indptr is None or self.recv_indptr.device != values.device:
                self.recv_indptr = torch.from_numpy(self.recv_indptr_host).to(values.device)
            (self._acc_rows or _cuda_accumulate_rows)(values, self.indptr, self.recv_row0, self.line, self.recv_indptr, recv)

    # -- matrices, fused: the kernel scatters the interface rows into the owner's memory ---------
    def enable_peer_scatter(self, A):
        """One-time setup for matrix `A` (same pattern as set_pattern): map the value array of rank+1 into
        this process (CUDA IPC, include/cq_b200.h cq_ipc_*) and build the slot table that makes the assembly
        kernels add this rank's contributions to its top lattice line -- owned by rank+1 -- straight into
        rank+1's matrix with red.global.add.f64 over NVLink.  After this, ``exchange_matrix`` is only the
        completion barrier.  Returns False (and leaves the NCCL exchange in place) if peer memory cannot
        be mapped.  Collective over the group."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        L = _lib.load()
        dev = A.data.device
        ctx = _lib.ctx(dev.index)
        rank, world = self.slab.rank, self.slab.world
        # what the lower neighbour needs from me: IPC handle of my values + the starts of my interface rows
        mine = None
        ok = True
        try:
            if self.has_down:
                handle = (C.c_ubyte * 64)()
                off = C.c_int64(0)
                _lib.check(L.cq_ipc_export(ctx, C.c_void_p(A.data.data_ptr()), handle, C.byref(off)))
                rows = _to_numpy(A.indptr[self.recv_row0 : self.recv_row0 + self.line + 1]).astype(np.int64)
                mine = (bytes(handle), int(off.value), rows)
        except Exception:  # noqa: BLE001 -- any failure => everybody falls back
            ok = False
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, mine), group=self.group)
        if not all(g[0] for g in gathered):
            return False
        ok = True
        if self.has_up:
            try:
                handle, off, rows = gathered[rank + 1][1]
                ptr, base = C.c_void_p(), C.c_void_p()
                hbuf = (C.c_ubyte * 64).from_buffer_copy(handle)
                _lib.check(L.cq_ipc_open(ctx, hbuf, off, C.byref(ptr), C.byref(base)))
                delta = ptr.value - A.data.data_ptr()
                assert delta % 8 == 0
                _open_mappings.append((dev.index, base.value))
                # sender row j of the top line is a prefix of the owner's row: same position inside the row
                table = rows + delta // 8
                self._peer_base = base
                A.peer_row0 = int(self.send_row0)
                A.peer_rowstart = torch.from_numpy(table).to(dev)
            except Exception:  # noqa: BLE001
                ok = False
        flags = [None] * world
        dist.all_gather_object(flags, ok, group=self.group)
        if not all(flags):
            A.peer_row0, A.peer_rowstart = None, None
            A.zero_keep = None
            return False
        self.peer = True
        # the owner side: a neighbour's kernels add into this matrix at any time between two fences, so zeroing it
        # needs the zero -> fence -> assemble -> fence protocol (assemble_matrix(..., zero=True) follows it)
        A.peer_target = self.has_down
        # the value slots the lower neighbour adds into: this rank's interface rows (cq_csr.zero_keep_lo / hi)
        A.zero_keep = tuple(int(v) for v in _to_numpy(A.indptr[[self.recv_row0, self.recv_row0 + self.line]])) if self.has_down else None
        A.peer_exchange = self
        setup_peer_fence(dev, self.group)
        return True

    def fence(self, device):
        """Stream-ordered barrier over the group: every rank's work enqueued so far (the zeroing of its
        matrix, or its assembly kernels with their peer REDs) is complete on all ranks' streams afterwards."""
        import torch
        import torch.distributed as dist

        if _peer_fence_ready.get(device.index):
            # flag words in the neighbours' memory (cq_peer_fence): one single-thread kernel, no NCCL in the step
            import ctypes as C

            from . import _lib

            _lib.check(_lib.load().cq_peer_fence(_lib.ctx(device.index), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            return
        if not hasattr(self, "_fence_buf") or self._fence_buf.device != device:
            self._fence_buf = torch.zeros(1, dtype=torch.float32, device=device)
        dist.all_reduce(self._fence_buf, group=self.group)

    # -- vectors ----------------------------------------------------------------------------
    def exchange_vector(self, b):
        import torch
        import torch.distributed as dist

        if self._axpy is None and b.is_cuda and ensure_comm(b.device, self.group):
            import ctypes as C

            from . import _lib

            _lib.check(_lib.load().cq_exchange_offrank_vector(
                _lib.ctx(b.device.index), C.c_void_p(b.data_ptr()),
                self.slab.rank + 1 if self.has_up else -1, self.send_row0, self.nrows,
                self.slab.rank - 1 if self.has_down else -1, self.recv_row0, self.line if self.has_down else 0,
                C.c_void_p(torch.cuda.current_stream().cuda_stream)))  # fmt: skip
            return
        ops, recv = [], None
        if self.has_up:
            ops.append(dist.P2POp(dist.isend, b[self.send_row0 : self.nrows], self.slab.rank + 1, self.group))
        if self.has_down:
            recv = torch.empty(self.line, dtype=b.dtype, device=b.device)
            ops.append(dist.P2POp(dist.irecv, recv, self.slab.rank - 1, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if recv is not None:
            (self._axpy or _cuda_axpy)(self.line, recv, b[self.recv_row0 : self.recv_row0 + self.line])

    # -- scalars ----------------------------------------------------------------------------
    @staticmethod
    def reduce_scalar(m, device=None, group=None) -> float:
        import torch
        import torch.distributed as dist

        t = torch.tensor([m], dtype=torch.float64, device=device)
        if t.is_cuda and ensure_comm(t.device, group):
            import ctypes as C

            from . import _lib

            _lib.check(_lib.load().cq_allreduce_sum(_lib.ctx(t.device.index), C.c_void_p(t.data_ptr()), 1, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return float(t.item())


# ---- process-wide state of the library-side communication (one cq_ctx per device per process) -------------------
_comm_ready: dict = {}        # device index -> bool: cq_comm_init done (NCCL communicator inside the cq_ctx)
_peer_fence_ready: dict = {}  # device index -> bool: cq_peer_fence_setup done
_open_mappings: list = []     # (device index, mapped base) of every cq_ipc_open, closed by close_peer_mappings()


def ensure_comm(device, group=None) -> bool:
    """Collective, once per process: create the NCCL communicator inside this device's cq_ctx (cq_comm_unique_id on
    rank 0, broadcast through torch.distributed, cq_comm_init).  CQ_DIST_TORCH=1 keeps the torch.distributed
    send/recv path instead.  Returns whether the library-side exchange (cq_exchange_offrank) can be used."""
    import ctypes as C
    import os

    import torch.distributed as dist

    from . import _lib

    if device.index in _comm_ready:
        return _comm_ready[device.index]
    ok = os.environ.get("CQ_DIST_TORCH", "0") != "1"
    ident = [None]
    if ok and dist.get_rank(group) == 0:
        buf = (C.c_ubyte * 128)()
        ok = _lib.load().cq_comm_unique_id(buf) == 0
        ident = [bytes(buf) if ok else None]
    dist.broadcast_object_list(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    mine = False
    if ident[0] is not None:
        buf = (C.c_ubyte * 128).from_buffer_copy(ident[0])
        mine = _lib.load().cq_comm_init(_lib.ctx(device.index), dist.get_world_size(group), dist.get_rank(group), buf) == 0
    flags = [None] * dist.get_world_size(group)
    dist.all_gather_object(flags, mine, group=group)
    _comm_ready[device.index] = all(flags)
    return _comm_ready[device.index]


def setup_peer_fence(device, group=None) -> bool:
    """Collective, once per process: every rank maps the fence words of its two slab neighbours (CUDA IPC) so that
    SlabExchange.fence is one small kernel (cq_peer_fence) instead of an NCCL all-reduce.  CQ_PEER_FENCE=0 keeps NCCL."""
    import ctypes as C
    import os

    import torch.distributed as dist

    from . import _lib

    if device.index in _peer_fence_ready:
        return _peer_fence_ready[device.index]
    L = _lib.load()
    ctx = _lib.ctx(device.index)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = None
    if os.environ.get("CQ_PEER_FENCE", "1") != "0":
        try:
            p = L.cq_peer_flags(ctx)
            handle, off = (C.c_ubyte * 64)(), C.c_int64(0)
            _lib.check(L.cq_ipc_export(ctx, C.c_void_p(p), handle, C.byref(off)))
            mine = (bytes(handle), int(off.value))
        except Exception:  # noqa: BLE001 -- any failure => everybody keeps the NCCL fence
            mine = None
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    ok = all(g is not None for g in gathered)
    if ok:
        try:
            ptrs = [None, None]
            for k, nb in enumerate((rank - 1, rank + 1)):
                if 0 <= nb < world:
                    handle, off = gathered[nb]
                    ptr, base = C.c_void_p(), C.c_void_p()
                    _lib.check(L.cq_ipc_open(ctx, (C.c_ubyte * 64).from_buffer_copy(handle), off, C.byref(ptr), C.byref(base)))
                    _open_mappings.append((device.index, base.value))
                    # word [1] of the down neighbour is written by its up neighbour (this rank); word [0] of the up neighbour by its down neighbour
                    ptrs[k] = C.c_void_p(ptr.value + (4 if k == 0 else 0))
            _lib.check(L.cq_peer_fence_setup(ctx, ptrs[0], ptrs[1]))
        except Exception:  # noqa: BLE001
            ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok, group=group)
    _peer_fence_ready[device.index] = all(flags)
    dist.barrier(group=group)
    return _peer_fence_ready[device.index]


def close_peer_mappings() -> None:
    """Unmap every neighbour allocation this process opened (cq_ipc_close); call once at teardown, after a barrier."""
    import ctypes as C

    from . import _lib

    while _open_mappings:
        dev, base = _open_mappings.pop()
        _lib.load().cq_ipc_close(_lib.ctx(dev), C.c_void_p(base))
    _peer_fence_ready.clear()


def _to_numpy(t):
    return t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)


def _cuda_accumulate_rows(values, indptr, row0, nrows, recv_indptr, recv):
    import ctypes as C

    from . import _lib

    assert values.is_cuda, "no CPU fallback: the receive-side add is a CUDA kernel"
    import torch

    L = _lib.load()
    A = _lib.CsrDesc(int(indptr.shape[0] - 1), int(values.shape[0]), C.c_void_p(indptr.data_ptr()), None, C.c_void_p(values.data_ptr()))
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.cq_csr_accumulate_rows(_lib.ctx(values.device.index), C.byref(A), row0, nrows, C.c_void_p(recv_indptr.data_ptr()), C.c_void_p(recv.data_ptr()), s))


def _cuda_axpy(n, x, y):
    import ctypes as C

    from . import _lib

    assert y.is_cuda, "no CPU fallback: the receive-side add is a CUDA kernel"
    import torch

    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(_lib.load().cq_axpy(_lib.ctx(y.device.index), int(n), C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), s))
