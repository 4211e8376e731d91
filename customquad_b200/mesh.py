"""Tensor-product quad / hex meshes with dolfinx 0.6's *local* orderings.

The reference never builds meshes itself: demos and tests call
``dolfinx.mesh.create_rectangle`` / ``create_box`` (reference demo/poisson.py:71-82,
test/test_linear_basis.py:29-34,57-62) and customquad only reads
``mesh.geometry.x``, ``mesh.geometry.dofmap`` and the cell count
(reference customquad/utils.py:7-21,66-71).  dolfinx is not available on the GPU box,
so this module produces the same arrays: same attribute names, same local vertex order
inside a cell (which is what the basis-function permutation in the reference's
call_basix.h:60-74 compensates for):

* quadrilaterals (``create_rectangle``): vertices are numbered y-fastest and a cell's
  local vertices are ``(ix,iy), (ix,iy+1), (ix+1,iy), (ix+1,iy+1)`` -- local vertex 1
  is the +y neighbour;
* hexahedra (``create_box``): x-fastest, local vertices in basix order.

dolfinx additionally reorders cells and vertices globally; that reordering is
irrelevant to the assembly path (it only renames rows) and is not reproduced.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field

import numpy as np


class CellType(enum.IntEnum):
    """basix::cell::type values the reference passes to call_basix (call_basix.h:61,68)."""

    quadrilateral = 4
    hexahedron = 5


@dataclass
class Geometry:
    """Stand-in for ``dolfinx.mesh.Mesh.geometry`` (fields read by utils.get_vertices)."""

    x: np.ndarray  # [num_nodes, 3] float64, z = 0 in 2-D
    dim: int
    dofmap: np.ndarray  # [num_cells, nv] int32

    @property
    def array(self):  # geometry.dofmap.array in dolfinx 0.6 (utils.py:70)
        return self.dofmap.reshape(-1)


@dataclass
class Topology:
    dim: int
    cell_type: CellType
    num_cells: int


@dataclass
class Mesh:
    geometry: Geometry
    topology: Topology
    # structured-grid metadata (not present in dolfinx; used by the dofmap builder)
    shape: tuple = ()
    xmin: np.ndarray = None
    xmax: np.ndarray = None
    _dev: dict = field(default_factory=dict, repr=False)

    @property
    def num_cells(self) -> int:
        return self.topology.num_cells

    @property
    def cell_type(self) -> CellType:
        return self.topology.cell_type

    def cell_index(self, *ijk) -> np.ndarray:
        """Cell id of structured index (ix, iy[, iz]) (arrays allowed)."""
        if self.topology.dim == 2:
            ix, iy = ijk
            return np.asarray(ix) * self.shape[1] + np.asarray(iy)
        ix, iy, iz = ijk
        return (np.asarray(iz) * self.shape[1] + np.asarray(iy)) * self.shape[0] + np.asarray(ix)

    def h(self) -> np.ndarray:
        return (self.xmax - self.xmin) / np.asarray(self.shape, dtype=np.float64)


def create_rectangle(points, n, cell_type=CellType.quadrilateral, dtype=np.float64) -> Mesh:
    """``dolfinx.mesh.create_rectangle(comm, [xmin, xmax], [nx, ny], quadrilateral)``."""
    assert CellType(cell_type) == CellType.quadrilateral, "only quadrilateral meshes (call_basix.h:61)"
    p = np.asarray(points, dtype=np.float64)
    nx, ny = int(n[0]), int(n[1])
    xs = p[0, 0] + (p[1, 0] - p[0, 0]) * (np.arange(nx + 1) / nx)
    ys = p[0, 1] + (p[1, 1] - p[0, 1]) * (np.arange(ny + 1) / ny)
    x = np.zeros(((nx + 1) * (ny + 1), 3), dtype=dtype)
    x[:, 0] = np.repeat(xs, ny + 1)  # vertex id = ix*(ny+1) + iy  (y fastest)
    x[:, 1] = np.tile(ys, nx + 1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    v0 = (ix * (ny + 1) + iy).reshape(-1)  # cell id = ix*ny + iy
    cells = np.stack([v0, v0 + 1, v0 + ny + 1, v0 + ny + 2], axis=1).astype(np.int32)
    return Mesh(
        Geometry(x, 2, cells),
        Topology(2, CellType.quadrilateral, nx * ny),
        shape=(nx, ny),
        xmin=p[0].copy(),
        xmax=p[1].copy(),
    )


def create_box(points, n, cell_type=CellType.hexahedron, dtype=np.float64) -> Mesh:
    """``dolfinx.mesh.create_box(comm, [xmin, xmax], [nx, ny, nz], hexahedron)``."""
    assert CellType(cell_type) == CellType.hexahedron, "only hexahedral meshes (call_basix.h:68)"
    p = np.asarray(points, dtype=np.float64)
    nx, ny, nz = int(n[0]), int(n[1]), int(n[2])
    xs = p[0, 0] + (p[1, 0] - p[0, 0]) * (np.arange(nx + 1) / nx)
    ys = p[0, 1] + (p[1, 1] - p[0, 1]) * (np.arange(ny + 1) / ny)
    zs = p[0, 2] + (p[1, 2] - p[0, 2]) * (np.arange(nz + 1) / nz)
    nvx, nvy = nx + 1, ny + 1
    x = np.empty((nvx * nvy * (nz + 1), 3), dtype=dtype)
    xv = x.reshape(nz + 1, nvy, nvx, 3)  # vertex id = (iz*nvy + iy)*nvx + ix  (x fastest); filled by broadcasting
    xv[..., 0] = xs[None, None, :]
    xv[..., 1] = ys[None, :, None]
    xv[..., 2] = zs[:, None, None]
    assert nvx * nvy * (nz + 1) < 2**31, "vertex count exceeds int32 (the reference's geometry dofmap is int32)"
    # cell id = (iz*ny + iy)*nx + ix; first vertex by broadcasting (no index grids: config 5 slabs have 6.7e7 cells)
    v0 = ((np.arange(nz, dtype=np.int32) * nvy)[:, None, None] + np.arange(ny, dtype=np.int32)[None, :, None]) * np.int32(nvx) + np.arange(nx, dtype=np.int32)[None, None, :]
    v0 = v0.reshape(-1)
    sy, sz = nvx, nvx * nvy
    cells = np.empty((nx * ny * nz, 8), dtype=np.int32)
    for k, off in enumerate([0, 1, sy, sy + 1, sz, sz + 1, sz + sy, sz + sy + 1]):
        cells[:, k] = v0 + np.int32(off)
    return Mesh(
        Geometry(x, 3, cells),
        Topology(3, CellType.hexahedron, nx * ny * nz),
        shape=(nx, ny, nz),
        xmin=p[0].copy(),
        xmax=p[1].copy(),
    )
