"""Result containers: what the reference hands back as PETSc Mat / Vec
(reference customquad/assemble_matrix.py:19-20,47; assemble_vector.py:27,49)."""
from __future__ import annotations

import numpy as np


class CSRMatrix:
    """Device CSR with the layout of a dolfinx/PETSc AIJ matrix: ``indptr`` int64 [nrows+1],
    ``indices`` int32 sorted inside each row, ``data`` float64 -- all torch device tensors.
    ``assemble()`` is the drop-in for ``A.assemble()`` (README.md:94-95): it waits for the
    device work and reports sparsity violations (PETSc would raise on a new nonzero)."""

    def __init__(self, indptr, indices, data, shape, device_index, stream_ptr=None):
        self.indptr, self.indices, self.data = indptr, indices, data
        self.shape = shape
        self._device_index = device_index
        self._stream_ptr = stream_ptr
        self.assembled = False
        self.full_pattern = True  # pattern of every cell of the mesh (create_matrix(form))
        # multi-GPU peer scatter (dist.SlabExchange.enable_peer_scatter): rows >= peer_row0 are added straight
        # into the owner rank's values over NVLink; peer_rowstart is the device slot table (cq_csr)
        self.peer_row0 = None
        self.peer_rowstart = None
        self.peer_target = False    # a neighbour's kernels add into this matrix (cq_csr.flags & CQ_CSR_PEER_TARGET)
        self.zero_keep = None       # ... into these value slots (its interface rows): cq_csr.zero_keep_lo / hi
        self.peer_exchange = None   # the SlabExchange whose fence orders the zeroing against the neighbours

    @property
    def size(self):
        return self.shape

    @property
    def nnz(self) -> int:
        return int(self.data.shape[0])

    def assemble(self):
        from . import _lib

        _lib.check(_lib.load().cq_ctx_check(_lib.ctx(self._device_index), self._stream_ptr))
        self.assembled = True
        return self

    def to_scipy(self):
        import scipy.sparse as sp

        if not self.assembled:
            self.assemble()
        idx_t = np.int32 if self.nnz < 2**31 else np.int64
        return sp.csr_matrix(
            (self.data.cpu().numpy(), self.indices.cpu().numpy(), self.indptr.cpu().numpy().astype(idx_t)), shape=self.shape
        )

    def getValuesCSR(self):
        """petsc4py-style (indptr, indices, data) host copies."""
        m = self.to_scipy()
        return m.indptr, m.indices, m.data

    def norm(self) -> float:
        if not self.assembled:
            self.assemble()
        return float(self.data.norm())

    def __iadd__(self, other: "CSRMatrix"):
        """``A += Ac1`` of demo/poisson.py:203-205 for matrices on the same pattern."""
        assert self.data.shape == other.data.shape, "patterns differ"
        self.data += other.data
        return self


class Vector:
    """Device vector; ``.array`` gives the host copy like ``petsc4py.Vec.array``."""

    def __init__(self, data, device_index):
        self.data = data
        self._device_index = device_index

    @property
    def array(self) -> np.ndarray:
        return self.data.cpu().numpy()

    def __iadd__(self, other: "Vector"):
        self.data += other.data
        return self

    def __len__(self):
        return int(self.data.shape[0])
