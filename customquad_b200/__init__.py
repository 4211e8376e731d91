"""customquad_b200 -- B200-native runtime-quadrature assembly behind customquad's API.

Drop-in surface (reference customquad/__init__.py:1-3):

    from customquad_b200 import assemble_matrix, assemble_vector, assemble_scalar
    A = assemble_matrix(form, [(cells, qr_pts, qr_w)])        # or (cells, qr_pts, qr_w, qr_n)
    A.assemble()

The compute path is hand-written sm_100a CUDA behind the C-ABI in include/cq_b200.h
(customquad_b200/csrc); there is no CPU fallback: calling an assemble_* function without the
built library or without a GPU raises.
"""
from . import cutfem, dolfinx_adapter, fem, forms, la, levelset, mesh, quadrature, sparsity, utils  # noqa: F401
from .assemble import assemble_matrix, assemble_scalar, assemble_vector  # noqa: F401
from .fem import Function, FunctionSpaceQ as FunctionSpace, VectorFunctionSpace  # noqa: F401
from .forms import Analytic  # noqa: F401
from .mesh import CellType, create_box, create_rectangle  # noqa: F401

__all__ = [
    "assemble_matrix", "assemble_vector", "assemble_scalar",
    "create_rectangle", "create_box", "CellType", "FunctionSpace", "VectorFunctionSpace", "Function", "Analytic",
    "forms", "fem", "mesh", "quadrature", "sparsity", "utils", "dolfinx_adapter",
]
