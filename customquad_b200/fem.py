"""Function spaces, dofmaps and finite-element functions (host side, numpy).

Stand-ins for the pieces of ``dolfinx.fem.FunctionSpace`` the reference touches
(reference customquad/utils.py:33-63): ``V.mesh``, ``V.dofmap.index_map_bs``,
``V.dofmap.dof_layout.num_dofs``, ``V.dofmap.list``, ``V.dofmap.cell_dofs(c)``.

Local dof order is basix's entity order (vertices, edges, faces, interior; SURVEY.md A.3,
pinned by the reference's own mesh builder test/create_high_order_mesh.py:46-64,153-217)
relative to the cell's local vertices.  Global dofs are numbered by lattice position in
the same sweep order as the mesh vertices, so for degree 1 the dofmap equals the geometry
dofmap.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace

import numpy as np

from .mesh import CellType, Mesh

# ---------------------------------------------------------------------------------------
# Reference lattice coordinates r_j (units of 1/degree) of local dof j, in the cell's own
# (basix) frame: axis a runs from local vertex 0 to local vertex 2^a.
# ---------------------------------------------------------------------------------------
_HEX_EDGES = [(0, 1), (0, 2), (0, 4), (1, 3), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 6), (5, 7), (6, 7)]
_HEX_FACES = [(0, 1, 2, 3), (0, 1, 4, 5), (0, 2, 4, 6), (1, 3, 5, 7), (2, 3, 6, 7), (4, 5, 6, 7)]
_QUAD_EDGES = [(0, 1), (0, 2), (1, 3), (2, 3)]


def reference_lattice(cell_type: CellType, degree: int) -> np.ndarray:
    """[n, tdim] integer lattice coordinates (0..degree) of the local dofs, basix order."""
    tdim = 2 if CellType(cell_type) == CellType.quadrilateral else 3
    verts = np.array([[(v >> a) & 1 for a in range(tdim)] for v in range(2**tdim)]) * degree
    if degree == 1:
        return verts
    if degree != 2:
        raise ValueError("degree must be 1 or 2 (reference call_basix.h:63-73)")
    pts = [v for v in verts]
    edges = _QUAD_EDGES if tdim == 2 else _HEX_EDGES
    for a, b in edges:
        pts.append((verts[a] + verts[b]) // 2)
    if tdim == 3:
        for f in _HEX_FACES:
            pts.append(sum(verts[i] for i in f) // 4)
    pts.append(np.full(tdim, 1))
    return np.array(pts)


def point_frame_lattice(cell_type: CellType, degree: int) -> np.ndarray:
    """Lattice coordinates of the local dofs in the frame the runtime quadrature points use.

    The reference evaluates basix's basis through a dof permutation (call_basix.h:60-74):
    ``{0,2,1,3}`` for Q1 quads makes psi_1 = (1-xi0) xi1, i.e. the points' axis 0 is the
    physical x axis although local vertex 1 is the +y neighbour (pinned by
    test/test_linear_basis.py:181,189-190).  That is a swap of the two axes; hexes are
    already aligned (identity permutation, call_basix.h:70-73).
    """
    r = reference_lattice(cell_type, degree)
    if CellType(cell_type) == CellType.quadrilateral:
        return r[:, ::-1].copy()
    return r


def lagrange_1d(degree: int, t: int, xi):
    """Value and derivative of the 1-D equispaced Lagrange basis function attached to lattice
    coordinate t (node xi = t/degree) -- lagrange_variant::unset == equispaced for degree <= 2
    (call_basix.h:30)."""
    xi = np.asarray(xi, dtype=np.float64)
    if degree == 1:
        return (1.0 - xi, -np.ones_like(xi)) if t == 0 else (xi, np.ones_like(xi))
    if t == 0:
        return (1.0 - xi) * (1.0 - 2.0 * xi), 4.0 * xi - 3.0
    if t == 1:
        return 4.0 * xi * (1.0 - xi), 4.0 - 8.0 * xi
    return xi * (2.0 * xi - 1.0), 4.0 * xi - 1.0


def tabulate(cell_type: CellType, degree: int, pts: np.ndarray):
    """(psi [nq, n], dpsi [nq, n, tdim]) in local-dof order and the points' frame."""
    lat = point_frame_lattice(cell_type, degree)
    pts = np.asarray(pts, dtype=np.float64)
    nq, tdim = pts.shape
    n = lat.shape[0]
    psi = np.ones((nq, n))
    dpsi = np.ones((nq, n, tdim))
    for j in range(n):
        for d in range(tdim):
            v, dv = lagrange_1d(degree, int(lat[j, d]), pts[:, d])
            psi[:, j] *= v
            for e in range(tdim):
                dpsi[:, j, e] *= dv if e == d else v
    return psi, dpsi


# ---------------------------------------------------------------------------------------
@dataclass
class DofMap:
    """``V.dofmap``: scalar (un-blocked) cell dofs + block size."""

    list: np.ndarray  # [num_cells, n] int32  (dolfinx: dofmap.list.array.reshape)
    index_map_bs: int
    num_dofs: int  # scalar dofs (index_map.size_local)

    @property
    def dof_layout(self):
        return SimpleNamespace(num_dofs=self.list.shape[1])

    @property
    def index_map(self):
        return SimpleNamespace(size_local=self.num_dofs, num_ghosts=0)

    def cell_dofs(self, c: int) -> np.ndarray:
        return self.list[c]


@dataclass
class FunctionSpace:
    mesh: Mesh
    degree: int
    bs: int
    dofmap: DofMap
    lattice_shape: tuple
    _dev: dict = field(default_factory=dict, repr=False)

    @property
    def num_dofs(self) -> int:
        return self.dofmap.num_dofs * self.bs

    def tabulate_dof_coordinates(self) -> np.ndarray:
        """[num_scalar_dofs, 3] physical coordinates of the dof lattice."""
        m = self.mesh
        k = self.degree
        axes = [m.xmin[d] + (m.xmax[d] - m.xmin[d]) * (np.arange(k * m.shape[d] + 1) / (k * m.shape[d])) for d in range(m.topology.dim)]
        x = np.zeros((self.dofmap.num_dofs, 3))
        if m.topology.dim == 2:
            X, Y = np.meshgrid(axes[0], axes[1], indexing="ij")  # id = I*(k ny + 1) + J
            x[:, 0], x[:, 1] = X.reshape(-1), Y.reshape(-1)
        else:
            Z, Y, X = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")  # id = (K*ly + J)*lx + I
            x[:, 0], x[:, 1], x[:, 2] = X.reshape(-1), Y.reshape(-1), Z.reshape(-1)
        return x


def _build_dofmap(mesh: Mesh, degree: int) -> tuple[np.ndarray, int, tuple]:
    k = degree
    r = reference_lattice(mesh.cell_type, k)  # cell frame
    if k == 1:
        # degree 1: dofs are numbered like the vertices, in the same sweep -> the dofmap IS the geometry dofmap
        # (pinned in tests/test_host_model.py); no index grids for the 6.7e7-cell slabs of config 5
        sh = mesh.shape
        lshape = (sh[0] + 1, sh[1] + 1) if mesh.topology.dim == 2 else (sh[0] + 1, sh[1] + 1, sh[2] + 1)
        return mesh.geometry.dofmap, int(np.prod(lshape)), lshape
    if mesh.topology.dim == 2:
        nx, ny = mesh.shape
        ly = k * ny + 1
        ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
        ix, iy = ix.reshape(-1, 1), iy.reshape(-1, 1)
        # quad cell frame: axis 0 = +y, axis 1 = +x (local vertex 1 is the +y neighbour)
        I = k * ix + r[None, :, 1]
        J = k * iy + r[None, :, 0]
        dofs = I * ly + J
        return dofs.astype(np.int32), (k * nx + 1) * ly, (k * nx + 1, ly)
    nx, ny, nz = mesh.shape
    lx, ly, lz = k * nx + 1, k * ny + 1, k * nz + 1
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ix, iy, iz = ix.reshape(-1, 1), iy.reshape(-1, 1), iz.reshape(-1, 1)
    I = k * ix + r[None, :, 0]
    J = k * iy + r[None, :, 1]
    K = k * iz + r[None, :, 2]
    dofs = (K * ly + J) * lx + I
    assert lx * ly * lz < 2**31, "scalar dof count exceeds int32 (reference dofs are int32, utils.py:56)"
    return dofs.astype(np.int32), lx * ly * lz, (lx, ly, lz)


def FunctionSpaceQ(mesh: Mesh, element, bs: int = 1) -> FunctionSpace:
    """``dolfinx.fem.FunctionSpace(mesh, ("Lagrange", k))`` (test/common.py:41, poisson.py:137)."""
    family, degree = element
    assert family in ("Lagrange", "P", "Q", "CG"), "only Lagrange (family 1, README.md:153)"
    dofs, ndofs, lshape = _build_dofmap(mesh, int(degree))
    return FunctionSpace(mesh, int(degree), int(bs), DofMap(dofs, int(bs), ndofs), lshape)


def VectorFunctionSpace(mesh: Mesh, element) -> FunctionSpace:
    """``dolfinx.fem.VectorFunctionSpace`` (demo/vector_poisson.py:95): bs = gdim."""
    return FunctionSpaceQ(mesh, element, bs=mesh.geometry.dim)


class Function:
    """A finite-element function: ``x.array`` holds num_dofs*bs values, blocked interleaved."""

    def __init__(self, V: FunctionSpace, values=None):
        self.function_space = V
        self.array = np.zeros(V.num_dofs) if values is None else np.ascontiguousarray(values, dtype=np.float64)
        assert self.array.shape == (V.num_dofs,)
        self.device_array = None
        self._dev = {}

    @property
    def x(self):
        return SimpleNamespace(array=self.array)

    def to_device(self, dev=None):
        """Keeps a copy of the values in HBM (``device_array``, a torch tensor): every later assemble_* call with this
        coefficient packs from it instead of uploading ``array`` (8 MB per call at 1 M dofs).  Call again after changing
        ``array``; assigning a CUDA tensor to ``device_array`` directly (a solution that never left the GPU) works too."""
        import torch

        dev = torch.device("cuda", torch.cuda.current_device()) if dev is None else dev
        self.device_array = torch.from_numpy(self.array).to(dev)
        return self

    def interpolate(self, fcn) -> "Function":
        """Nodal interpolation; ``fcn(x)`` with x of shape [3, npoints] as in dolfinx
        (test/common.py:11-24). Vector-valued functions return [bs, npoints]."""
        V = self.function_space
        X = V.tabulate_dof_coordinates().T
        vals = np.asarray(fcn(X), dtype=np.float64)
        if V.bs == 1:
            self.array[:] = np.broadcast_to(vals.reshape(-1) if vals.ndim > 1 else vals, (V.dofmap.num_dofs,))
        else:
            self.array[:] = np.broadcast_to(vals, (V.bs, V.dofmap.num_dofs)).T.reshape(-1)
        self._dev.clear()
        self.device_array = None
        return self

    def packed(self) -> np.ndarray:
        """[num_cells, n*bs]: what dolfinx.cpp.fem.pack_coefficients gives for a single
        coefficient (assemble_matrix.py:16): cell-local dof values, blocked interleaved."""
        V = self.function_space
        d = V.dofmap.list
        if V.bs == 1:
            return self.array[d]
        idx = d[:, :, None] * V.bs + np.arange(V.bs)[None, None, :]
        return self.array[idx.reshape(d.shape[0], -1)]
