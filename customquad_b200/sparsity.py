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


def build_pattern_torch(dofmap, num_dofs: int, bs: int = 1, chunk_cells: int = 1 << 22):
    """Device builder (torch tensors in, torch tensors out: indptr int64, indices int32).
    Same algorithm; pairs are generated and de-duplicated in chunks of cells to bound memory."""
    import torch

    d = dofmap.to(torch.int64)
    uniq = None
    for s in range(0, d.shape[0], chunk_cells):
        dc = d[s : s + chunk_cells]
        k = torch.unique((dc[:, :, None] * num_dofs + dc[:, None, :]).reshape(-1))
        uniq = k if uniq is None else torch.unique(torch.cat([uniq, k]))
    rows = torch.div(uniq, num_dofs, rounding_mode="floor")
    cols = (uniq - rows * num_dofs).to(torch.int32)
    indptr = torch.zeros(num_dofs + 1, dtype=torch.int64, device=d.device)
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=num_dofs), 0)
    if bs > 1:
        m = indptr[1:] - indptr[:-1]
        E = (cols.to(torch.int64)[:, None] * bs + torch.arange(bs, device=d.device)[None, :]).reshape(-1)
        len_exp = torch.repeat_interleave(m * bs, bs)
        indptr_exp = torch.zeros(len_exp.shape[0] + 1, dtype=torch.int64, device=d.device)
        indptr_exp[1:] = torch.cumsum(len_exp, 0)
        start = torch.repeat_interleave(indptr[:-1] * bs, bs)
        src = torch.repeat_interleave(start - indptr_exp[:-1], len_exp) + torch.arange(int(indptr_exp[-1]), device=d.device)
        indptr, cols = indptr_exp, E[src].to(torch.int32)
    return indptr, cols.contiguous()
