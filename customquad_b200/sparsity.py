"""CSR sparsity pattern of a bilinear form: the stand-in for
``dolfinx.cpp.fem.petsc.create_matrix(form)`` (reference customquad/assemble_matrix.py:19).

Pattern = union over the cells of the form's cell-integral domain of dofs[c] x dofs[c],
expanded by the block size, columns ascending inside each row, ``indices`` int32, ``indptr``
int64.  ``cells=None`` is the reference default ("form" pattern: every cell
of the mesh, as for ``dx_cut`` without subdomain id, demo/poisson.py:180,189); passing the
listed cells gives the pattern an assembled PETSc matrix keeps after squeezing untouched
slots (demo/poisson.py:181-183,190).

Setup code, not the timed path: numpy on the host, or the same algorithm with torch ops on
the device for meshes whose pair list is too large for the host (the 2048x2048 Q2 bench mesh
has 3.4e8 pairs).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Pattern:
    indptr: np.ndarray  # [nrows+1] int64
    indices: np.ndarray  # [nnz] int32
    num_rows: int

    @property
    def nnz(self) -> int:
        return int(self.indptr[-1])


def _expand_blocks_np(indptr, indices, bs):
    if bs == 1:
        return indptr, indices
    m = np.diff(indptr)  # scalar row lengths
    E = (indices.astype(np.int64)[:, None] * bs + np.arange(bs)[None, :]).reshape(-1)  # row r: E[indptr[r]*bs : indptr[r+1]*bs]
    len_exp = np.repeat(m * bs, bs)  # expanded row R = r*bs + a
    indptr_exp = np.zeros(len_exp.shape[0] + 1, dtype=np.int64)
    np.cumsum(len_exp, out=indptr_exp[1:])
    start = np.repeat(indptr[:-1] * bs, bs)
    src = np.repeat(start - indptr_exp[:-1], len_exp) + np.arange(indptr_exp[-1])
    return indptr_exp, E[src].astype(np.int32)


def build_pattern(dofmap: np.ndarray, num_dofs: int, bs: int = 1, cells=None) -> Pattern:
    """Host (numpy) builder.  dofmap [num_cells, n] scalar dofs."""
    d = np.asarray(dofmap, dtype=np.int64)
    if cells is not None:
        d = d[np.unique(np.asarray(cells, dtype=np.int64))]
    n = d.shape[1]
    keys = (d[:, :, None] * num_dofs + d[:, None, :]).reshape(-1)
    keys = np.unique(keys)
    rows = keys // num_dofs
    cols = (keys - rows * num_dofs).astype(np.int32)
    indptr = np.zeros(num_dofs + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=num_dofs), out=indptr[1:])
    indptr, cols = _expand_blocks_np(indptr, cols, bs)
    return Pattern(indptr, np.ascontiguousarray(cols, dtype=np.int32), num_dofs * bs)


def build_pattern_torch(dofmap, num_dofs: int, bs: int = 1, chunk_pairs: int = 1 << 28):
    """Device builder (torch tensors in, torch tensors out: indptr int64, indices int32).

    Row-chunked: the (dof, cell) incidences are sorted by dof once; a chunk of consecutive rows then owns a
    contiguous slice of that list, its candidate columns are the dofs of the incident cells, and de-duplicating
    them gives the FINAL rows of the chunk -- nothing is merged across chunks, so the peak memory is a few
    ``chunk_pairs`` keys whatever the mesh size (config 5: 6.7e7 hexahedra per GPU, 1.8e9 nonzeros)."""
    import torch

    dev = dofmap.device
    C, n = dofmap.shape
    d = dofmap.reshape(-1).to(torch.int64)
    order = torch.argsort(d, stable=True)           # incidences sorted by dof; order // n = cell
    inc_cell = torch.div(order, n, rounding_mode="floor")
    per_row = torch.bincount(d, minlength=num_dofs)  # incident cells per row
    del d, order
    inc_ptr = torch.zeros(num_dofs + 1, dtype=torch.int64, device=dev)
    inc_ptr[1:] = torch.cumsum(per_row, 0)
    # row chunks holding at most ~chunk_pairs candidate (row, col) pairs
    cand_ptr = (inc_ptr * n).cpu().numpy()
    bounds = [0]
    while bounds[-1] < num_dofs:
        r1 = int(np.searchsorted(cand_ptr, cand_ptr[bounds[-1]] + chunk_pairs, side="right")) - 1
        bounds.append(min(num_dofs, max(r1, bounds[-1] + 1)))
    dm = dofmap.to(torch.int64)
    counts = torch.zeros(num_dofs, dtype=torch.int64, device=dev)
    cols_parts = []
    for r0, r1 in zip(bounds[:-1], bounds[1:]):
        i0, i1 = int(inc_ptr[r0]), int(inc_ptr[r1])
        if i1 == i0:
            continue
        rows = torch.repeat_interleave(torch.arange(r0, r1, device=dev), per_row[r0:r1])  # row of every incidence
        keys = (rows[:, None] * num_dofs + dm[inc_cell[i0:i1]]).reshape(-1)
        keys = torch.unique(keys)
        rr = torch.div(keys, num_dofs, rounding_mode="floor")
        counts[r0:r1] = torch.bincount(rr - r0, minlength=r1 - r0)
        cols_parts.append((keys - rr * num_dofs).to(torch.int32))
        del rows, keys, rr
    cols = torch.cat(cols_parts) if cols_parts else torch.zeros(0, dtype=torch.int32, device=dev)
    del cols_parts
    indptr = torch.zeros(num_dofs + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(counts, 0)
    if bs > 1 and cols.is_cuda:
        # one kernel, no temporaries (cq_csr_expand_blocks): config 4 expands 1.08e9 scalar couplings into 9.69e9
        import ctypes as C

        from . import _lib

        indptr_exp = torch.empty(num_dofs * bs + 1, dtype=torch.int64, device=dev)
        cols_exp = torch.empty(int(indptr[-1]) * bs * bs, dtype=torch.int32, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        _lib.check(_lib.load().cq_csr_expand_blocks(_lib.ctx(dev.index), num_dofs, p(indptr), p(cols), bs, p(indptr_exp), p(cols_exp),
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))  # fmt: skip
        return indptr_exp, cols_exp
    if bs > 1:
        m = indptr[1:] - indptr[:-1]
        E = (cols.to(torch.int64)[:, None] * bs + torch.arange(bs, device=dev)[None, :]).reshape(-1)
        len_exp = torch.repeat_interleave(m * bs, bs)
        indptr_exp = torch.zeros(len_exp.shape[0] + 1, dtype=torch.int64, device=dev)
        indptr_exp[1:] = torch.cumsum(len_exp, 0)
        start = torch.repeat_interleave(indptr[:-1] * bs, bs)
        src = torch.repeat_interleave(start - indptr_exp[:-1], len_exp) + torch.arange(int(indptr_exp[-1]), device=dev)
        indptr, cols = indptr_exp, E[src].to(torch.int32)
    return indptr, cols.contiguous()


def add_macro_couplings(indptr, indices, num_dofs: int, bs: int, macro_dofmap):
    """The pattern (indptr, indices: torch tensors, blocked by ``bs``) plus the couplings of every "macro element"
    of ``macro_dofmap`` [num_macro, m] (scalar dofs): what dolfinx's create_matrix adds for an interior-facet integral
    (the two cells of a ghost-penalty facet couple all their dofs, demo/poisson.py:160,168-170).  Returns new
    (indptr, indices).  Setup code; the union is formed on the tensors' device."""
    import torch

    dev = indices.device
    nrows = num_dofs * bs
    rows = torch.repeat_interleave(torch.arange(nrows, device=dev), indptr[1:] - indptr[:-1])
    keys = rows * nrows + indices.to(torch.int64)
    md = torch.as_tensor(macro_dofmap, device=dev).to(torch.int64)
    blk = torch.arange(bs, device=dev)
    mr = (md[:, :, None] * bs + blk[None, None, :]).reshape(md.shape[0], -1)  # blocked dofs of every macro element
    extra = (mr[:, :, None] * nrows + mr[:, None, :]).reshape(-1)
    keys = torch.unique(torch.cat([keys, extra]))
    r = torch.div(keys, nrows, rounding_mode="floor")
    new_indptr = torch.zeros(nrows + 1, dtype=torch.int64, device=dev)
    new_indptr[1:] = torch.cumsum(torch.bincount(r, minlength=nrows), 0)
    return new_indptr, (keys - r * nrows).to(torch.int32).contiguous()
