"""Host-side staging helpers with the reference's names (reference customquad/utils.py)."""
from __future__ import annotations

import numpy as np


def get_num_cells(mesh) -> int:
    """reference utils.py:18-21 (owned + ghost cells; single process => all cells)."""
    return mesh.topology.num_cells


def get_dofs(V):
    """reference utils.py:33-63: dofs [num_cells, n*bs] int32 and num_loc_dofs = n*bs; for
    bs > 1 blocked interleaved: dofs[cell, i*bs + j] = dof_i*bs + j (:59-62)."""
    d = V.dofmap.list
    bs = V.dofmap.index_map_bs
    num_loc_dofs = V.dofmap.dof_layout.num_dofs * bs
    if bs == 1:
        return d, num_loc_dofs
    dofs = (d[:, :, None] * bs + np.arange(bs, dtype=np.int32)[None, None, :]).reshape(d.shape[0], num_loc_dofs)
    return dofs.astype(np.int32), num_loc_dofs


def get_vertices(mesh):
    """reference utils.py:66-71: (vertices [num_cells, nv], coords [npts, 3], gdim)."""
    return mesh.geometry.dofmap, mesh.geometry.x, mesh.geometry.dim


def volume(xmin, xmax, NN, uncut_cells, qr_w):
    """reference utils.py:228-235: volume represented by a bulk rule (weights are cell-relative)."""
    gdim = len(NN)
    cellvol = np.prod((np.asarray(xmax) - np.asarray(xmin))[0:gdim]) / np.prod(NN)
    return float(sum(float(np.sum(w)) for w in qr_w) * cellvol + cellvol * len(uncut_cells))


def area(xmin, xmax, NN, qr_w_bdry):
    """reference utils.py:238-243."""
    gdim = len(NN)
    cellvol = np.prod((np.asarray(xmax) - np.asarray(xmin))[0:gdim]) / np.prod(NN)
    return float(sum(float(np.sum(w)) for w in qr_w_bdry) * cellvol)
