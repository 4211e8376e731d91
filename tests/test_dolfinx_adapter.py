"""customquad_b200.dolfinx_adapter driven by stand-ins with the dolfinx 0.6 attribute layout the
reference reads (customquad/utils.py:7-71, assemble_matrix.py:11): AdjacencyList objects with
``.array``, ``dofmap.list`` as attribute or as method, ``topology.index_map(tdim)``,
``ufl_element().degree()``, ``Function.x.array``.  dolfinx itself is not installed here.

The converted forms must assemble (CPU oracle here, CUDA path under -m gpu) to the same numbers as
the package's own mesh / space objects the stand-ins were made from -- also after a renumbering of
cells and dofs of the kind dolfinx applies (which the package's structured builders do not)."""
from types import SimpleNamespace as NS

import numpy as np
import pytest

import customquad_b200 as cq
from customquad_b200 import dolfinx_adapter as dx
from customquad_b200 import forms
from customquad_b200.quadrature import synthetic_ragged


class Adj:  # dolfinx.cpp.graph.AdjacencyList_int32
    def __init__(self, a):
        self.array = np.ascontiguousarray(a).reshape(-1)


def fake_dolfinx(V, list_as_method=False, perm_cells=None, perm_dofs=None, ghosts=0):
    """dolfinx-shaped views of a package FunctionSpace (optionally with cells / dofs renumbered)."""
    m = V.mesh
    nc = m.num_cells
    pc = np.arange(nc) if perm_cells is None else perm_cells  # new cell k = old cell pc[k]
    pd = np.arange(V.dofmap.num_dofs) if perm_dofs is None else perm_dofs  # new id of old dof i = pd[i]
    gd = m.geometry.dofmap[pc]
    dl = pd[V.dofmap.list[pc]].astype(np.int32)
    imap_cells = NS(size_local=nc - ghosts, num_ghosts=ghosts)
    dmesh = NS(
        geometry=NS(x=m.geometry.x.copy(), dim=m.geometry.dim, dofmap=Adj(gd)),
        topology=NS(dim=m.topology.dim, index_map=lambda tdim: imap_cells, create_connectivity=lambda a, b: None),
    )
    adj = Adj(dl)
    dofmap = NS(
        index_map_bs=V.bs,
        dof_layout=NS(num_dofs=dl.shape[1]),
        list=(lambda: adj) if list_as_method else adj,
        index_map=NS(size_local=V.dofmap.num_dofs, num_ghosts=0),
        cell_dofs=lambda c: dl[c],
    )
    dV = NS(mesh=dmesh, dofmap=dofmap, ufl_element=lambda: NS(degree=lambda: V.degree))
    return dV, pc, pd


def blocked(pd, bs):
    return (pd[:, None] * bs + np.arange(bs)[None, :]).reshape(-1)


@pytest.mark.parametrize("tdim,degree,bs", [(2, 1, 1), (2, 2, 1), (3, 1, 1), (2, 1, 2), (3, 2, 3)])
@pytest.mark.parametrize("list_as_method", [False, True])
def test_adapter_reads_the_reference_layout(tdim, degree, bs, list_as_method):
    mesh = cq.create_rectangle([[-1.0, -1.0], [4.0, 1.0]], [3, 2]) if tdim == 2 else cq.create_box([[0, 0, 0], [1, 2, 3]], [2, 2, 2])
    V = cq.FunctionSpace(mesh, ("Lagrange", degree)) if bs == 1 else cq.VectorFunctionSpace(mesh, ("Lagrange", degree))
    dV, _, _ = fake_dolfinx(V, list_as_method)
    W = dx.function_space(dV)
    assert W.degree == degree and W.bs == bs and W.num_dofs == V.num_dofs
    assert W.mesh.cell_type == mesh.cell_type and W.mesh.num_cells == mesh.num_cells
    assert W.dofmap.list.dtype == np.int32 and np.array_equal(W.dofmap.list, V.dofmap.list)
    assert np.array_equal(W.mesh.geometry.dofmap, mesh.geometry.dofmap) and np.array_equal(W.mesh.geometry.x, mesh.geometry.x)
    # utils.get_dofs / get_vertices on the converted objects give the reference's arrays (utils.py:33-71)
    d0, n0 = cq.utils.get_dofs(V)
    d1, n1 = cq.utils.get_dofs(W)
    assert n0 == n1 and np.array_equal(d0, d1)


def _cases(tdim):
    rng = np.random.default_rng(4)
    mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [5, 4]) if tdim == 2 else cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [3, 2, 3])
    mesh.geometry.x[:, :tdim] += rng.uniform(-0.15, 0.15, size=(mesh.geometry.x.shape[0], tdim)) * mesh.h()
    return mesh, rng


def _compare(backend, tdim, degree, bs):
    mesh, rng = _cases(tdim)
    V = cq.FunctionSpace(mesh, ("Lagrange", degree)) if bs == 1 else cq.VectorFunctionSpace(mesh, ("Lagrange", degree))
    nc, nd = mesh.num_cells, V.dofmap.num_dofs
    pc, pd = rng.permutation(nc), rng.permutation(nd)
    dV, pc, pd = fake_dolfinx(V, perm_cells=pc, perm_dofs=pd)
    f = cq.Function(V)
    f.array[:] = rng.uniform(-1, 1, size=V.num_dofs)
    fperm = np.empty_like(f.array)
    fperm[blocked(pd, bs)] = f.array  # the same function in the renumbered space
    du = NS(function_space=dV, x=NS(array=fperm))
    cells = np.sort(rng.choice(nc, size=nc - 3, replace=False)).astype(np.int32)
    qr = synthetic_ragged(cells, tdim, 3, 9, seed=2, normals=True)
    inv = np.empty(nc, dtype=np.int64)
    inv[pc] = np.arange(nc)
    qr2 = synthetic_ragged(inv[cells].astype(np.int32), tdim, 3, 9, seed=2, normals=True)  # same rules, renumbered cells
    P = blocked(pd, bs)
    for kind in ("mass", "stiffness", "nitsche"):
        a0 = {"mass": forms.mass, "stiffness": forms.stiffness, "nitsche": lambda V: forms.nitsche(V, 10.0)}[kind](V)
        a1 = dx.form(NS(function_spaces=[dV]), kind, betaN=10.0)
        A0 = backend.matrix(a0, [qr]).toarray()
        A1 = backend.matrix(a1, [qr2]).toarray()
        ref = A0
        got = A1[np.ix_(P, P)]
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref)), kind
    for kind in ("load", "nitsche_rhs"):
        b0 = backend.vector(forms.load(V, f) if kind == "load" else forms.nitsche_rhs(V, f, 10.0), [qr])
        b1 = backend.vector(dx.form(dV, kind, f=du, betaN=10.0), [qr2])
        assert np.max(np.abs(b1[P] - b0)) <= 1e-12 * np.max(np.abs(b0)), kind
    m0 = backend.scalar(forms.integral_sq(f), [qr])
    m1 = backend.scalar(dx.scalar_form("integral_sq", f=du), [qr2])
    assert abs(m1 - m0) <= 1e-12 * abs(m0)
    v0 = backend.scalar(forms.integral(mesh, 1.0), [qr])
    v1 = backend.scalar(dx.scalar_form("integral", dV.mesh), [qr2])
    assert abs(v1 - v0) <= 1e-12 * abs(v0)


@pytest.mark.parametrize("tdim,degree,bs", [(2, 1, 1), (2, 2, 1), (3, 1, 1), (2, 2, 2), (3, 2, 1)])
def test_adapter_forms_assemble_like_native_oracle(oracle_backend, tdim, degree, bs):
    _compare(oracle_backend, tdim, degree, bs)


@pytest.mark.gpu
@pytest.mark.parametrize("tdim,degree,bs", [(2, 1, 1), (2, 2, 2), (3, 1, 1), (3, 2, 1)])
def test_adapter_forms_assemble_like_native_cuda(cuda_backend, tdim, degree, bs):
    _compare(cuda_backend, tdim, degree, bs)


def test_adapter_rejects_what_the_reference_cannot_do():
    mesh = cq.create_rectangle([[0, 0], [1, 1]], [2, 2])
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    dV, _, _ = fake_dolfinx(V)
    tri = NS(geometry=NS(x=np.zeros((3, 3)), dim=2, dofmap=Adj(np.arange(3))), topology=NS(dim=2, index_map=lambda t: NS(size_local=1, num_ghosts=0)))
    with pytest.raises(NotImplementedError):
        dx.mesh(tri)  # simplices: no permutation in call_basix.h:60-74
    dV.ufl_element = lambda: NS(degree=lambda: 3)
    with pytest.raises(NotImplementedError):
        dx.function_space(dV)
    with pytest.raises(ValueError):
        dx.form(fake_dolfinx(V)[0], "curlcurl")


# ---- UFL integrand -> form menu (dolfinx_adapter.classify_integrand) with stand-in nodes carrying UFL's class names ----
class _Node:
    def __init__(self, *ops):
        self.ufl_operands = ops


def _ufl(name):
    return type(name, (_Node,), {})


_Sum, _Product, _Inner, _Dot, _Grad, _NablaGrad, _Division, _Power, _CellDiameter, _FacetNormal, _SpatialCoordinate = (
    _ufl(n) for n in "Sum Product Inner Dot Grad NablaGrad Division Power CellDiameter FacetNormal SpatialCoordinate".split()
)


class Argument(_Node):
    def __init__(self, number):
        super().__init__()
        self._n = number

    def number(self):
        return self._n


class Coefficient(_Node):
    pass


def test_ufl_integrands_of_the_reference_are_recognised():
    from customquad_b200 import dolfinx_adapter as dx

    u, v, n, h, g = Argument(1), Argument(0), _FacetNormal(), _CellDiameter(), Coefficient()
    neg = lambda e: _Product(-1, e)  # noqa: E731  (UFL writes -x as a product with IntValue(-1))
    assert dx.classify_integrand(_Inner(u, v)) == ("mass", {})  # test/common.py:91
    assert dx.classify_integrand(_Product(u, v)) == ("mass", {})
    assert dx.classify_integrand(_Inner(_Grad(u), _Grad(v))) == ("stiffness", {})  # demo/poisson.py:154
    assert dx.classify_integrand(_Inner(_NablaGrad(u), _NablaGrad(v))) == ("stiffness", {})  # demo/vector_poisson.py:121
    a_bdry = _Sum(_Sum(neg(_Inner(_Dot(n, _Grad(u)), v)), neg(_Inner(u, _Dot(n, _Grad(v))))), _Inner(_Product(_Division(10.0, h), u), v))
    assert dx.classify_integrand(a_bdry) == ("nitsche", {"betaN": 10.0})  # demo/poisson.py:156-158
    L_bdry = _Sum(neg(_Inner(g, _Dot(n, _Grad(v)))), _Inner(_Product(_Division(7.5, h), g), v))
    kind, data = dx.classify_integrand(L_bdry)  # demo/poisson.py:159
    assert kind == "nitsche_rhs" and data["betaN"] == 7.5 and data["f"] is g
    kind, data = dx.classify_integrand(_Inner(g, v))  # test/common.py:66-68
    assert kind == "load" and data["f"] is g
    assert dx.classify_integrand(_Product(1.0, v)) == ("load", {"f": 1.0})  # test_linear_basis.py:129
    assert dx.classify_integrand(1.0) == ("integral", {"f": 1.0})  # test_volume_area.py:58
    for e in (_Product(g, g), _Inner(g, g), _Power(g, 2)):  # test/common.py:42-44
        kind, data = dx.classify_integrand(e)
        assert kind == "integral_sq" and data["f"] is g


def test_ufl_integrands_outside_the_menu_raise():
    from customquad_b200 import dolfinx_adapter as dx

    u, v, n, h = Argument(1), Argument(0), _FacetNormal(), _CellDiameter()
    for e in (_Inner(_Grad(u), v),                                  # convection-like
              _Inner(_Product(_SpatialCoordinate(), u), v),        # variable coefficient
              _Sum(_Inner(u, v), _Inner(_Grad(u), _Grad(v))),      # two menu forms in one integrand
              _Product(_Division(10.0, h), _Inner(u, v))):          # Nitsche penalty alone
        with pytest.raises(NotImplementedError):
            dx.classify_integrand(e)


def test_form_from_ufl_builds_the_same_form_as_the_named_constructor():
    from customquad_b200 import dolfinx_adapter as dx
    from customquad_b200.forms import FormKind

    mesh = cq.create_rectangle([[0, 0], [1, 1]], [3, 2])
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    u, v, n, h = Argument(1), Argument(0), _FacetNormal(), _CellDiameter()
    neg = lambda e: _Product(-1, e)  # noqa: E731
    assert dx.form_from_ufl(_Inner(_Grad(u), _Grad(v)), V).kind == FormKind.STIFFNESS
    a = dx.form_from_ufl(_Sum(_Sum(neg(_Inner(_Dot(n, _Grad(u)), v)), neg(_Inner(u, _Dot(n, _Grad(v))))), _Inner(_Product(_Division(10.0, h), u), v)), V)
    assert a.kind == FormKind.NITSCHE and a.betaN == 10.0
    f = cq.Function(V)
    assert dx.form_from_ufl(_Inner(f, v), V).coefficient is f
    assert dx.form_from_ufl(1.0, dmesh=mesh).kind == FormKind.INTEGRAL
