"""Form descriptors: what ``dolfinx.fem.form(integrand * dx_runtime)`` is to the reference.

The reference receives a compiled UFL form and reads from it the function space, the mesh,
packed coefficients / constants and a generated kernel
(reference customquad/assemble_matrix.py:11-30).  Without UFL/ffcx the integrand is named
instead: one constructor per integrand the reference's demos and tests assemble
(SURVEY.md 8a row a-6).  A ``Form`` carries exactly the attributes the reference reads
(``function_spaces``, ``mesh``) plus the integrand kind and its data.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, field

import numpy as np

from .fem import Function, FunctionSpace
from .mesh import Mesh


class FormKind(enum.IntEnum):  # == cq_form in include/cq_b200.h
    MASS = 0
    STIFFNESS = 1
    NITSCHE = 2
    LOAD = 10
    NITSCHE_RHS = 11
    INTEGRAL = 20
    INTEGRAL_SQ = 21
    L2_ERR = 22
    H10_ERR = 23


class FnKind(enum.IntEnum):  # == cq_fn
    COEF = 0
    CONST = 1
    AFFINE = 2
    SINPROD = 3


@dataclass(frozen=True)
class Analytic:
    """Closed-form data function of the physical coordinate (ufl.SpatialCoordinate)."""

    kind: FnKind
    params: tuple  # flat, per component

    @staticmethod
    def const(*c):
        return Analytic(FnKind.CONST, tuple(float(v) for v in c))

    @staticmethod
    def affine(c0, grad):
        """c0 + grad . x ; for a blocked function pass sequences of per-component (c0, grad)."""
        g = list(grad) + [0.0] * (3 - len(grad))
        return Analytic(FnKind.AFFINE, (float(c0), float(g[0]), float(g[1]), float(g[2])))

    @staticmethod
    def affine_blocked(comps):
        p = []
        for c0, grad in comps:
            g = list(grad) + [0.0] * (3 - len(grad))
            p += [float(c0), float(g[0]), float(g[1]), float(g[2])]
        return Analytic(FnKind.AFFINE, tuple(p))

    @staticmethod
    def sinprod(*amp):
        """amp[a] * prod_d sin(pi x_d)  (u_exact of demo/poisson.py:59-67)."""
        return Analytic(FnKind.SINPROD, tuple(float(v) for v in amp))

    def ncomp(self) -> int:
        return len(self.params) // 4 if self.kind == FnKind.AFFINE else len(self.params)

    def __call__(self, x, comp: int = 0):
        """Evaluate on x [3, npts] (numpy), used by tests."""
        if self.kind == FnKind.CONST:
            return np.full(x.shape[1], self.params[comp])
        if self.kind == FnKind.AFFINE:
            p = self.params[4 * comp : 4 * comp + 4]
            return p[0] + p[1] * x[0] + p[2] * x[1] + p[3] * x[2]
        raise NotImplementedError


RANK = {
    FormKind.MASS: 2, FormKind.STIFFNESS: 2, FormKind.NITSCHE: 2,
    FormKind.LOAD: 1, FormKind.NITSCHE_RHS: 1,
    FormKind.INTEGRAL: 0, FormKind.INTEGRAL_SQ: 0, FormKind.L2_ERR: 0, FormKind.H10_ERR: 0,
}  # fmt: skip


@dataclass
class Form:
    kind: FormKind
    mesh: Mesh
    function_spaces: list  # [V] (rank >= 1), [] for functionals (assemble_scalar.py:9 uses form.mesh)
    coefficient: Function | None = None  # the V_h function of the integrand, if any
    fn: Analytic | None = None  # the closed-form function of the integrand, if any
    betaN: float | None = None
    # degree / bs of the element the kernel tabulates (coefficient's space for functionals)
    degree: int = 1
    bs: int = 1
    _dev: dict = field(default_factory=dict, repr=False)

    @property
    def rank(self) -> int:
        return RANK[self.kind]

    @property
    def fn_kind(self) -> FnKind:
        if self.kind in (FormKind.L2_ERR, FormKind.H10_ERR):
            return self.fn.kind
        if self.fn is not None:
            return self.fn.kind
        return FnKind.COEF

    def consts(self) -> np.ndarray:
        """Packed constants, the analogue of dolfinx.cpp.fem.pack_constants (assemble_matrix.py:17):
        [betaN] for Nitsche forms, then the closed-form function's parameters."""
        c = []
        if self.kind in (FormKind.NITSCHE, FormKind.NITSCHE_RHS):
            c.append(float(self.betaN))
        if self.fn is not None:
            c += list(self.fn.params)
        return np.asarray(c, dtype=np.float64)

    def integral_ids(self):
        return [-1]  # one runtime cell integral, id -1 (assemble_vector.py:14-15)


def _data(V_or_mesh, f, bs):
    if isinstance(f, Function):
        # the kernels index the packed coefficient with the FORM's degree and block size (cq_space.coeffs [cells, n*bs])
        if isinstance(V_or_mesh, FunctionSpace):
            W = f.function_space
            assert W is V_or_mesh or (W.degree == V_or_mesh.degree and W.bs == V_or_mesh.bs and W.mesh is V_or_mesh.mesh), \
                "the coefficient must live in the form's space (same mesh, degree and block size)"
        return f, None
    if isinstance(f, Analytic):
        assert f.ncomp() == bs, f"function has {f.ncomp()} components, space has bs={bs}"
        return None, f
    if np.isscalar(f):
        return None, Analytic.const(*([float(f)] * bs))
    raise TypeError("data function must be a Function, an Analytic or a number")


def mass(V: FunctionSpace) -> Form:
    """inner(u, v) * dx_runtime (test/common.py:91)."""
    return Form(FormKind.MASS, V.mesh, [V, V], degree=V.degree, bs=V.bs)


def stiffness(V: FunctionSpace) -> Form:
    """inner(grad(u), grad(v)) (demo/poisson.py:154); nabla_grad version for bs>1
    (demo/vector_poisson.py:121)."""
    return Form(FormKind.STIFFNESS, V.mesh, [V, V], degree=V.degree, bs=V.bs)


def nitsche(V: FunctionSpace, betaN: float) -> Form:
    """-(n.grad u) v - u (n.grad v) + betaN/h u v with quadrature normals (demo/poisson.py:156-158)."""
    return Form(FormKind.NITSCHE, V.mesh, [V, V], betaN=betaN, degree=V.degree, bs=V.bs)


def load(V: FunctionSpace, f) -> Form:
    """inner(f, v): f in V_h (test/common.py:66-68), closed-form (poisson.py:147,155) or constant
    (test_linear_basis.py:129)."""
    c, fn = _data(V, f, V.bs)
    if c is not None:
        assert c.function_space is V or (c.function_space.degree == V.degree and c.function_space.bs == V.bs)
    return Form(FormKind.LOAD, V.mesh, [V], coefficient=c, fn=fn, degree=V.degree, bs=V.bs)


def nitsche_rhs(V: FunctionSpace, g, betaN: float) -> Form:
    """-g (n.grad v) + betaN/h g v (demo/poisson.py:159; vector_poisson.py:126)."""
    c, fn = _data(V, g, V.bs)
    return Form(FormKind.NITSCHE_RHS, V.mesh, [V], coefficient=c, fn=fn, betaN=betaN, degree=V.degree, bs=V.bs)


def integral(mesh: Mesh, f=1.0) -> Form:
    """f * dx(domain=mesh), f closed-form: volume/area (test_volume_area.py:58,141),
    2x+y (test_linear_basis.py:84-86)."""
    _, fn = _data(mesh, f, 1)
    assert fn is not None, "use integral_sq / l2_error for V_h functions"
    return Form(FormKind.INTEGRAL, mesh, [], fn=fn, degree=1, bs=1)


def integral_sq(f: Function) -> Form:
    """inner(f, f) * dx, f in V_h (test/common.py:42-44)."""
    V = f.function_space
    return Form(FormKind.INTEGRAL_SQ, V.mesh, [], coefficient=f, degree=V.degree, bs=V.bs)


def l2_error(uh: Function, u_ex: Analytic) -> Form:
    """(uh - u_ex)**2 (demo/poisson.py:275)."""
    V = uh.function_space
    assert u_ex.ncomp() == V.bs
    return Form(FormKind.L2_ERR, V.mesh, [], coefficient=uh, fn=u_ex, degree=V.degree, bs=V.bs)


def h10_error(uh: Function, u_ex: Analytic) -> Form:
    """(grad(uh) - grad(u_ex))**2 (demo/poisson.py:279)."""
    V = uh.function_space
    assert u_ex.ncomp() == V.bs
    return Form(FormKind.H10_ERR, V.mesh, [], coefficient=uh, fn=u_ex, degree=V.degree, bs=V.bs)
