"""dolfinx objects -> customquad_b200 forms (the drop-in when FEniCSx is installed).

The reference's assemble_* functions take a compiled ``dolfinx.fem.Form`` and read a handful of arrays
from it (reference customquad/utils.py:7-71, assemble_matrix.py:11-20, assemble_vector.py:27,
assemble_scalar.py:9).  This module reads exactly the same attributes -- nothing of dolfinx is imported,
the objects are used duck-typed -- and builds the ``Mesh`` / ``FunctionSpace`` / ``Function`` / ``Form``
the package's assemblers take:

    import customquad_b200 as cq
    from customquad_b200 import dolfinx_adapter as dx
    a = dx.form(V, "stiffness")                      # V: dolfinx.fem.FunctionSpace (or a dolfinx Form over it)
    A = cq.assemble_matrix(a, [(cells, qr_pts, qr_w)])

What cannot come from dolfinx is the integrand: the reference gets it as ffcx-custom generated code
(assemble_matrix.py:27-30), which does not exist here, so the integrand is named (``kind``, the menu of
SURVEY 8a row a-6).  dolfinx is absent from this container; ``tests/test_dolfinx_adapter.py`` drives the
adapter with stand-in objects that expose the dolfinx 0.6 attribute layout the reference relies on
(``AdjacencyList.array``, ``dofmap.list`` as attribute or method, ``topology.index_map(tdim)``).
"""
from __future__ import annotations

import numpy as np

from . import forms as _forms
from .fem import DofMap, Function, FunctionSpace
from .mesh import CellType, Geometry, Mesh, Topology

_CELL_BY_NV = {(2, 4): CellType.quadrilateral, (3, 8): CellType.hexahedron}


def _num_cells(dmesh) -> int:
    """reference utils.py:7-21: owned + ghost cells."""
    tdim = dmesh.topology.dim
    im = dmesh.topology.index_map(tdim)
    return int(im.size_local) + int(im.num_ghosts)


def _adjacency(adj, rows: int) -> np.ndarray:
    """``AdjacencyList`` (``.array``), a method returning one (``dofmap.list()``, utils.py:52-55), or a 2-D ndarray
    (dolfinx >= 0.7) -> [rows, -1] int32."""
    if callable(adj):
        adj = adj()
    a = adj.array if hasattr(adj, "array") else adj
    a = np.ascontiguousarray(np.asarray(a), dtype=np.int32)
    return a.reshape(rows, -1)


def mesh(dmesh) -> Mesh:
    """``utils.get_vertices`` (reference utils.py:66-71) as a ``Mesh``: geometry.x [npts, 3], geometry.dofmap
    [num_cells, nv], gdim.  Only quadrilaterals and hexahedra (call_basix.h:60-74)."""
    if isinstance(dmesh, Mesh):
        return dmesh
    ncells = _num_cells(dmesh)
    gdim = int(dmesh.geometry.dim)
    tdim = int(dmesh.topology.dim)
    x = np.ascontiguousarray(np.asarray(dmesh.geometry.x, dtype=np.float64).reshape(-1, 3))
    gd = _adjacency(dmesh.geometry.dofmap, ncells)
    key = (tdim, gd.shape[1])
    if key not in _CELL_BY_NV or gdim != tdim:
        raise NotImplementedError(
            f"cells with {gd.shape[1]} geometry nodes in {tdim}-D (gdim {gdim}): only affine-order quadrilaterals / hexahedra "
            "are supported, like the reference's permutation table (call_basix.h:60-74)"
        )
    return Mesh(Geometry(x, gdim, gd), Topology(tdim, _CELL_BY_NV[key], ncells))


def _degree(V) -> int:
    for get in (
        lambda: V.ufl_element().degree(),  # dolfinx 0.6 (the reference's pin)
        lambda: V.element.basix_element.degree,
        lambda: V.ufl_element().degree,
    ):
        try:
            d = get()
            if isinstance(d, (int, np.integer)):
                return int(d)
        except Exception:
            pass
    raise ValueError("cannot read the element degree of the function space")


def function_space(V, cq_mesh: Mesh | None = None) -> FunctionSpace:
    """``utils.get_dofs`` (reference utils.py:33-63) kept un-blocked: dofmap.list [num_cells, n], index_map_bs,
    scalar dof count = index_map.size_local + num_ghosts."""
    if isinstance(V, FunctionSpace):
        return V
    m = cq_mesh if cq_mesh is not None else mesh(V.mesh)
    bs = int(V.dofmap.index_map_bs)
    n = int(V.dofmap.dof_layout.num_dofs)
    d = _adjacency(V.dofmap.list, m.num_cells)
    if d.shape[1] != n:
        raise ValueError(f"dofmap has {d.shape[1]} dofs per cell, dof_layout says {n}")
    im = V.dofmap.index_map
    ndofs = int(im.size_local) + int(im.num_ghosts)
    deg = _degree(V)
    if n != (deg + 1) ** m.topology.dim or deg not in (1, 2):
        raise NotImplementedError(f"Lagrange degree {deg} with {n} dofs per cell: only Q1 / Q2 (call_basix.h:63-73)")
    return FunctionSpace(m, deg, bs, DofMap(d, bs, ndofs), ())


def function(u, V: FunctionSpace | None = None) -> Function:
    """A ``dolfinx.fem.Function``: ``u.x.array`` (blocked interleaved, owned + ghost dofs)."""
    if isinstance(u, Function):
        return u
    W = V if V is not None else function_space(u.function_space)
    return Function(W, np.asarray(u.x.array, dtype=np.float64)[: W.num_dofs])


_KINDS = {
    "mass": lambda V, **k: _forms.mass(V),
    "stiffness": lambda V, **k: _forms.stiffness(V),
    "nitsche": lambda V, betaN, **k: _forms.nitsche(V, betaN),
    "load": lambda V, f, **k: _forms.load(V, f),
    "nitsche_rhs": lambda V, f, betaN, **k: _forms.nitsche_rhs(V, f, betaN),
}


def form(obj, kind: str, f=None, betaN: float | None = None):
    """The package ``Form`` of integrand ``kind`` over a dolfinx function space.

    ``obj``: a ``dolfinx.fem.FunctionSpace`` or a compiled ``dolfinx.fem.Form`` (its ``function_spaces[0]`` is
    used, like assemble_matrix.py:11).  ``kind``: "mass" | "stiffness" | "nitsche" | "load" | "nitsche_rhs".
    ``f``: the data of a linear form -- a ``dolfinx.fem.Function`` on the same space or a ``cq.Analytic``.
    Functionals: ``scalar_form``."""
    V = obj.function_spaces[0] if hasattr(obj, "function_spaces") else obj
    W = function_space(V)
    if kind not in _KINDS:
        raise ValueError(f"unknown integrand {kind!r}; one of {sorted(_KINDS)}")
    if f is not None and not isinstance(f, (_forms.Analytic, Function)) and not np.isscalar(f):  # numbers pass through (forms._data)
        f = function(f, W)
    return _KINDS[kind](W, f=f, betaN=betaN)


def scalar_form(kind: str, dmesh=None, f=None, u_ex=None):
    """Functionals (assemble_scalar.py:9 reads ``form.mesh``): "integral" (``f``: float or Analytic, over ``dmesh``),
    "integral_sq" (``f``: Function), "l2_error" / "h10_error" (``f``: Function u_h, ``u_ex``: Analytic)."""
    if kind == "integral":
        return _forms.integral(mesh(dmesh), 1.0 if f is None else f)
    fh = function(f)
    if kind == "integral_sq":
        return _forms.integral_sq(fh)
    if kind == "l2_error":
        return _forms.l2_error(fh, u_ex)
    if kind == "h10_error":
        return _forms.h10_error(fh, u_ex)
    raise ValueError(f"unknown functional {kind!r}")
