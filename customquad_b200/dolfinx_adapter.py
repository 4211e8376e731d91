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


# ---- UFL integrand -> form menu -------------------------------------------------------------------
# The reference gets its element kernel from the UFL form through ffcx-custom (assemble_matrix.py:27-30).  Here the
# integrand is one of the forms of the menu, so a UFL integrand is RECOGNISED rather than compiled: the expression tree
# is walked through the public UFL node interface (class name, ``ufl_operands``, ``Argument.number()``), nothing of
# UFL is imported.  Recognised (everything the reference's demos and tests assemble with runtime quadrature, modulo the
# data functions, which must be V_h functions or numbers -- closed-form ``SpatialCoordinate`` data goes through
# ``forms.load(V, cq.Analytic...)``):
#   inner(u, v) | u*v                                                    -> mass            (test/common.py:91)
#   inner(grad(u), grad(v)) | inner(nabla_grad(u), nabla_grad(v))        -> stiffness       (demo/poisson.py:154, vector_poisson.py:121)
#   -(n.grad u) v - u (n.grad v) + betaN/h u v                           -> nitsche         (demo/poisson.py:156-158)
#   inner(f, v), f a Function or a number                                -> load            (test/common.py:66-68)
#   -g (n.grad v) + betaN/h g v, g a Function                            -> nitsche_rhs     (demo/poisson.py:159)
#   a number | f*f, inner(f, f), f**2 with f a Function                  -> integral | integral_sq   (test_volume_area.py:58, common.py:42-44)
# dolfinx / UFL are absent here: tests/test_dolfinx_adapter.py drives the matcher with stand-in nodes that carry UFL's
# class names and operand layout.  Untested against a real UFL.
def _cls(e) -> str:
    return type(e).__name__


def _ops(e):
    return tuple(getattr(e, "ufl_operands", ()))


def _is_number(e) -> bool:
    return isinstance(e, (int, float, np.integer, np.floating)) or _cls(e) in ("IntValue", "FloatValue", "ScalarValue", "RealValue", "Zero")


def _number(e) -> float:
    if isinstance(e, (int, float, np.integer, np.floating)):
        return float(e)
    if _cls(e) == "Zero":
        return 0.0
    v = getattr(e, "_value", None)
    if v is None and callable(getattr(e, "value", None)):
        v = e.value()
    return float(v)


def _const_value(e):
    """A number, a UFL scalar value, or a dolfinx.fem.Constant (``.value``): float, else None."""
    if _is_number(e):
        return _number(e)
    if _cls(e) == "Constant" and hasattr(e, "value"):
        return float(np.asarray(e.value).reshape(-1)[0])
    return None


def _terms(e, sign=1.0):
    """Flatten sums / negations / scalar factors into [(coefficient, expression)]."""
    c = _cls(e)
    if c == "Sum":
        out = []
        for o in _ops(e):
            out += _terms(o, sign)
        return out
    if c == "Product":
        a, b = _ops(e)
        va, vb = _const_value(a), _const_value(b)
        if va is not None and _cls(b) != "Constant":
            return _terms(b, sign * va)
        if vb is not None and _cls(a) != "Constant":
            return _terms(a, sign * vb)
    return [(sign, e)]


def _arg_number(e):
    """0 (test) / 1 (trial) for an Argument, through grad-free wrappers; None otherwise."""
    if _cls(e) == "Argument":
        return int(e.number())
    return None


def _strip_grad(e):
    """(inner expression, True) if e = grad(x) / nabla_grad(x), else (e, False)."""
    if _cls(e) in ("Grad", "NablaGrad"):
        return _ops(e)[0], True
    return e, False


def _is_normal(e) -> bool:
    return _cls(e) in ("FacetNormal", "QuadratureNormal", "CellNormal", "RuntimeNormal") or getattr(e, "is_quadrature_normal", False)


def _normal_derivative_of(e):
    """x if e = dot(n, grad(x)) / inner(n, grad(x)) / dot(grad(x), n), else None."""
    if _cls(e) not in ("Dot", "Inner"):
        return None
    a, b = _ops(e)
    for n, g in ((a, b), (b, a)):
        x, isg = _strip_grad(g)
        if _is_normal(n) and isg:
            return x
    return None


def _factors(e):
    """Flatten products / inner / dot of two operands into a factor list; divisions as ('/', den)."""
    c = _cls(e)
    if c in ("Product", "Inner", "Dot") and _normal_derivative_of(e) is None:
        a, b = _ops(e)
        return _factors(a) + _factors(b)
    if c == "Division":
        a, b = _ops(e)
        return _factors(a) + [("/", b)]
    if c == "Power":
        a, b = _ops(e)
        if _const_value(b) == 2.0:
            return _factors(a) + _factors(a)
    return [e]


def _describe(term):
    """One product term -> dict: test / trial involvement (plain 'v', gradient 'gv', normal derivative 'nv'), data factors."""
    d = {"test": None, "trial": None, "coeffs": [], "consts": 1.0, "over_h": False, "other": []}
    for f in _factors(term):
        if isinstance(f, tuple):
            if _cls(f[1]) == "CellDiameter":
                d["over_h"] = True
            elif _const_value(f[1]) is not None:
                d["consts"] /= _const_value(f[1])
            else:
                d["other"].append(f[1])
            continue
        x = _normal_derivative_of(f)
        mode = "n" if x is not None else ""
        if x is None:
            x, isg = _strip_grad(f)
            mode = "g" if isg else ""
        num = _arg_number(x)
        if num is not None:
            d["test" if num == 0 else "trial"] = mode + ("v" if num == 0 else "u")
        elif _const_value(f) is not None:
            d["consts"] *= _const_value(f)
        elif _cls(x) in ("Coefficient", "Function") or hasattr(x, "x") and hasattr(x, "function_space"):
            d["coeffs"].append((mode, x))
        else:
            d["other"].append(f)
    return d


def classify_integrand(integrand):
    """UFL integrand -> ("mass" | "stiffness" | "nitsche" | "load" | "nitsche_rhs" | "integral" | "integral_sq", data)
    with data = {"betaN": float, "f": Function-like or float}.  Raises NotImplementedError outside the menu."""
    if _const_value(integrand) is not None:
        return "integral", {"f": _const_value(integrand)}
    ts = [(c, _describe(t)) for c, t in _terms(integrand)]
    if any(d["other"] for _, d in ts):
        raise NotImplementedError("integrand outside the runtime-quadrature form menu (SURVEY 8a a-6): unrecognised factor " + repr(ts[0][1]["other"][:1] or [d["other"] for _, d in ts if d["other"]][0][:1]))
    shapes = sorted((d["test"] or "-", d["trial"] or "-", d["over_h"]) for _, d in ts)
    if len(ts) == 1:
        c, d = ts[0]
        if (d["test"], d["trial"]) == ("v", "u") and not d["coeffs"] and not d["over_h"] and c * d["consts"] == 1.0:
            return "mass", {}
        if (d["test"], d["trial"]) == ("gv", "gu") and not d["coeffs"] and c * d["consts"] == 1.0:
            return "stiffness", {}
        if d["test"] == "v" and d["trial"] is None and not d["over_h"]:
            if not d["coeffs"]:
                return "load", {"f": c * d["consts"]}
            if len(d["coeffs"]) == 1 and d["coeffs"][0][0] == "" and c * d["consts"] == 1.0:
                return "load", {"f": d["coeffs"][0][1]}
        if d["test"] is None and d["trial"] is None and len(d["coeffs"]) == 2 and d["coeffs"][0][1] is d["coeffs"][1][1] and c * d["consts"] == 1.0:
            return "integral_sq", {"f": d["coeffs"][0][1]}
    if shapes == [("nv", "u", False), ("v", "nu", False), ("v", "u", True)]:
        by = {(d["test"], d["trial"]): (c, d) for c, d in ts}
        if by[("nv", "u")][0] * by[("nv", "u")][1]["consts"] == -1.0 and by[("v", "nu")][0] * by[("v", "nu")][1]["consts"] == -1.0:
            c, d = by[("v", "u")]
            return "nitsche", {"betaN": c * d["consts"]}
    if shapes == [("nv", "-", False), ("v", "-", True)]:
        by = {d["test"]: (c, d) for c, d in ts}
        (c1, d1), (c2, d2) = by["nv"], by["v"]
        if c1 * d1["consts"] == -1.0 and len(d1["coeffs"]) == 1 and len(d2["coeffs"]) == 1 and d1["coeffs"][0][1] is d2["coeffs"][0][1]:
            return "nitsche_rhs", {"betaN": c2 * d2["consts"], "f": d1["coeffs"][0][1]}
    raise NotImplementedError(
        "integrand outside the runtime-quadrature form menu (mass, stiffness, symmetric Nitsche, load, Nitsche right-hand side, "
        "integral, integral of a square; SURVEY 8a a-6): name it with dolfinx_adapter.form(V, kind, ...) or customquad_b200.forms"
    )


def form_from_ufl(integrand, V=None, dmesh=None):
    """The package ``Form`` for a UFL integrand (the expression multiplied by the runtime ``dx`` in the reference,
    README.md:86-89).  ``V``: the dolfinx function space of the arguments (rank >= 1); ``dmesh``: the mesh (functionals)."""
    kind, data = classify_integrand(integrand)
    if kind in ("mass", "stiffness"):
        return form(V, kind)
    if kind == "nitsche":
        return form(V, kind, betaN=data["betaN"])
    if kind in ("load", "nitsche_rhs"):
        return form(V, kind, f=data["f"], betaN=data.get("betaN"))
    if kind == "integral":
        return scalar_form("integral", dmesh, f=data["f"])
    return scalar_form("integral_sq", f=data["f"])
