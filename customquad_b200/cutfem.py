"""CutFEM helpers around the assembly path: what demo/poisson.py does
right before and right after ``customquad.assemble_*`` with dolfinx / PETSc.

* ``get_inactive_dofs`` / ``lock_inactive_dofs``: reference customquad/utils.py:74-105.
* ``standard_rule``: the tiled Gauss rule that turns the runtime-quadrature path into dolfinx's
  standard assembly over a cell subset (``a_bulk * dx_uncut(uncut_cell_tag)``,
  demo/poisson.py:168-176; the equivalence the reference's tests check, test/common.py:48-57).
* ``ghost_penalty_facets`` / ``ghost_penalty``: ``betas * avg(h) * inner(jump(n, grad(u)),
  jump(n, grad(v))) * dS(ghost_penalty_tag)`` (demo/poisson.py:160,168-170) on the facets
  ``utils.get_facetags`` tags (utils.py:155-182), for Q1 / Q2 (any block size) on axis-aligned tensor meshes,
  added into the device CSR by a CUDA kernel (cq_assemble_ghost_penalty) on a pattern that contains the facet
  couplings (``assemble.create_matrix(form, facets=...)``).  The reference leaves this term to dolfinx.
  ``ghost_penalty_matrix`` is the host numpy statement of the same term the kernel is tested against.
"""
from __future__ import annotations

import numpy as np

from . import fem
from .quadrature import gauss_tensor_rule


def get_inactive_dofs(V, cut_cells, uncut_cells) -> np.ndarray:
    """reference utils.py:74-82 (blocked dof indices)."""
    d = V.dofmap.list
    bs = V.bs
    active = np.zeros(V.dofmap.num_dofs, dtype=bool)
    for cells in (cut_cells, uncut_cells):
        if len(cells):
            active[d[np.asarray(cells, dtype=np.int64)].reshape(-1)] = True
    scalar = np.nonzero(~active)[0]
    return (scalar[:, None] * bs + np.arange(bs)[None, :]).reshape(-1).astype(np.int32)


def lock_inactive_dofs(inactive_dofs, A):
    """reference utils.py:85-105: put 1 on the diagonal of the inactive rows (on the device CSR)
    and raise if any diagonal entry is still zero."""
    import torch

    dev = A.data.device
    rows = torch.as_tensor(np.asarray(inactive_dofs, dtype=np.int64), device=dev)
    if rows.numel():
        start, end = A.indptr[rows], A.indptr[rows + 1]
        cnt = end - start
        row_of = torch.repeat_interleave(rows, cnt)
        slot = torch.repeat_interleave(start - torch.cumsum(cnt, 0) + cnt, cnt) + torch.arange(int(cnt.sum()), device=dev)
        diag = slot[A.indices[slot].to(torch.int64) == row_of]
        assert diag.numel() == rows.numel(), "the pattern has no diagonal entry for an inactive dof"
        A.data[diag] += 1.0
    n = A.shape[0]
    all_rows = torch.repeat_interleave(torch.arange(n, device=dev), A.indptr[1:] - A.indptr[:-1])
    is_diag = A.indices.to(torch.int64) == all_rows
    d = torch.zeros(n, dtype=A.data.dtype, device=dev)
    d[all_rows[is_diag]] = A.data[is_diag]
    if bool((d == 0).any()):
        raise RuntimeError("Zeros on the diagonal should not happen")
    return A


def standard_rule(mesh, cells, degree):
    """(cells, qr_pts, qr_w) with the same tensor Gauss rule in every listed cell (2-D arrays, one
    row per cell, as test/common.py:48-53 builds them)."""
    tdim = mesh.topology.dim
    pts, w = gauss_tensor_rule(tdim, degree)
    cells = np.asarray(cells, dtype=np.int32)
    return (cells, np.tile(pts.reshape(1, -1), (len(cells), 1)), np.tile(w.reshape(1, -1), (len(cells), 1)))


def ghost_penalty_facets(mesh, cut_cells, outside_cells):
    """Interior facets shared by a cut cell and a cell that is not outside (utils.py:166-175):
    array [nf, 3] of (cell on the low side, cell on the high side, axis)."""
    shape = mesh.shape
    tdim = mesh.topology.dim
    is_cut = np.zeros(mesh.num_cells, dtype=bool)
    is_cut[np.asarray(cut_cells, dtype=np.int64)] = True
    is_out = np.zeros(mesh.num_cells, dtype=bool)
    is_out[np.asarray(outside_cells, dtype=np.int64)] = True
    grids = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    out = []
    for ax in range(tdim):
        sel = grids[ax] < shape[ax] - 1
        lo_idx = [g[sel] for g in grids]
        hi_idx = [g[sel] + (1 if a == ax else 0) for a, g in enumerate(grids)]
        lo, hi = mesh.cell_index(*lo_idx), mesh.cell_index(*hi_idx)
        keep = (is_cut[lo] & ~is_out[hi]) | (is_cut[hi] & ~is_out[lo])
        out.append(np.stack([lo[keep], hi[keep], np.full(int(keep.sum()), ax)], axis=1))
    return np.concatenate(out).astype(np.int64)


def ghost_penalty(V, facets, betas, A):
    """The ghost-penalty term added into the device matrix ``A`` (created with ``assemble.create_matrix(form,
    facets=facets)``) by the CUDA kernel behind cq_assemble_ghost_penalty: A = Ax + Ac1 + Ac2 + stabilisation stays on
    the GPU.  Degree 1 or 2, any block size, axis-aligned cells."""
    from .assemble import assemble_ghost_penalty

    return assemble_ghost_penalty(V, facets, betas, A)


def ghost_penalty_matrix(V, facets, betas):
    """Host numpy version (scipy CSR) for scalar spaces (bs = 1; blocked spaces: kron with the identity) on uniform
    axis-aligned meshes; kept as the small independent statement the device kernel is tested against
    (tests/test_gpu_cutfem_poisson.py).
    betas * avg(h) * int_F [n.grad u][n.grad v] dS over `facets` (h = cell diagonal, the same for both cells)."""
    import scipy.sparse as sp

    assert V.bs == 1, "ghost penalty (host): scalar spaces"
    mesh = V.mesh
    tdim = mesh.topology.dim
    h = mesh.h()
    hK = float(np.sqrt(np.sum(h * h)))
    g, gw = gauss_tensor_rule(tdim - 1, 2 * V.degree)  # exact: the integrand has degree 2k per facet coordinate
    n = V.dofmap.list.shape[1]
    rows, cols, vals = [], [], []
    for ax in range(tdim):
        f = facets[facets[:, 2] == ax]
        if not len(f):
            continue
        area = float(np.prod(np.delete(h, ax)))
        # facet points in the unit cell of the low cell (xi_ax = 1) and of the high cell (xi_ax = 0)
        P_lo = np.insert(g, ax, 1.0, axis=1)
        P_hi = np.insert(g, ax, 0.0, axis=1)
        d_lo = fem.tabulate(mesh.cell_type, V.degree, P_lo)[1][:, :, ax] / h[ax]  # [nq, n] d/dx_ax psi_j
        d_hi = fem.tabulate(mesh.cell_type, V.degree, P_hi)[1][:, :, ax] / h[ax]
        J = np.concatenate([-d_lo, d_hi], axis=1)  # jump over the 2n local dofs
        M = betas * hK * area * np.einsum("q,qi,qj->ij", gw, J, J)
        dofs = np.concatenate([V.dofmap.list[f[:, 0]], V.dofmap.list[f[:, 1]]], axis=1)  # [nf, 2n]
        rows.append(np.repeat(dofs, 2 * n, axis=1).reshape(-1))
        cols.append(np.tile(dofs, (1, 2 * n)).reshape(-1))
        vals.append(np.tile(M.reshape(-1), len(f)))
    N = V.dofmap.num_dofs
    if not rows:
        return sp.csr_matrix((N, N))
    return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N)).tocsr()
