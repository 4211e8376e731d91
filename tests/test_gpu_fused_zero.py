"""cq_assemble_matrix_zero (zeroEntries fused into the assembly kernel, csrc/cq_zero.cuh) against
zero-then-assemble (assemble_matrix.py:20 followed by :26-45), on matrices pre-filled with NaN:
every slot must come out zeroed-and-accumulated whatever the order of the cell list, for one wave
of stages (small lists) and for many (the zeroing runs `zero_ahead` waves ahead of the scatter)."""
import numpy as np
import pytest

import customquad_b200 as cq
from customquad_b200 import assemble as asm
from customquad_b200 import forms
from customquad_b200.quadrature import synthetic_ragged

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(autouse=True)
def fused_zero_on():
    """The in-kernel zeroing is opt-in (cq_ctx_set_fused_zero); everything here runs with it on."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    asm.set_fused_zero(True)
    yield
    asm.set_fused_zero(None)


def _mesh(tdim, n):
    if tdim == 2:
        return cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [n, n + 3])
    return cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [n, n + 1, n + 2])


def _space(mesh, degree, bs):
    return cq.FunctionSpace(mesh, ("Lagrange", degree)) if bs == 1 else cq.VectorFunctionSpace(mesh, ("Lagrange", degree))


def _cells(kind, ncells, rng):
    if kind == "all":
        return np.arange(ncells, dtype=np.int32)
    if kind == "reversed":
        return np.arange(ncells, dtype=np.int32)[::-1].copy()
    if kind == "shuffled":
        return rng.permutation(ncells).astype(np.int32)
    if kind == "subset":  # ascending subset: the last rows of the matrix are never touched
        c = np.sort(rng.choice(ncells, size=max(1, ncells // 3), replace=False)).astype(np.int32)
        return c[c < (ncells * 3) // 4]
    if kind == "repeats":
        return np.sort(rng.integers(0, ncells, size=ncells, dtype=np.int32))
    raise ValueError(kind)


def _check(form, qr):
    import torch

    A = asm.create_matrix(form)
    cq.assemble_matrix(form, [qr], A=A)  # zeroed by create_matrix, accumulate
    A.assemble()
    B = asm.create_matrix(form, zero=False)
    for _ in range(2):  # twice: the second call also has the counters of the first to reset
        B.data.fill_(float("nan"))
        cq.assemble_matrix(form, [qr], A=B, zero=True)
        B.assemble()
        torch.cuda.synchronize()
        a, b = A.data.cpu().numpy(), B.data.cpu().numpy()
        assert not np.isnan(b).any(), f"{int(np.isnan(b).sum())} of {b.size} slots were never zeroed"
        assert np.array_equal(a == 0.0, b == 0.0)
        assert np.linalg.norm(a - b) <= RTOL * np.linalg.norm(a)
        assert np.max(np.abs(a - b)) <= RTOL * np.max(np.abs(a))


# (tdim, degree, bs, form, cells per direction): n = 12 -> one wave; n = 200 (2-D) / 34 (3-D) -> 9-10 waves
CASES = [
    (2, 1, 1, "stiffness", 12), (2, 2, 1, "mass", 12), (3, 1, 1, "stiffness", 6),
    (2, 1, 1, "mass", 200), (2, 2, 1, "stiffness", 200), (2, 2, 2, "mass", 120), (2, 1, 2, "stiffness", 200),
    (3, 1, 1, "mass", 34), (3, 1, 1, "nitsche", 34), (3, 1, 3, "stiffness", 30),
]  # fmt: skip


@pytest.mark.parametrize("tdim,degree,bs,kind,n", CASES)
@pytest.mark.parametrize("order", ["all", "subset", "shuffled", "reversed", "repeats"])
def test_fused_zero_matches_zero_then_assemble(tdim, degree, bs, kind, n, order):
    rng = np.random.default_rng(5)
    mesh = _mesh(tdim, n)
    V = _space(mesh, degree, bs)
    form = {"mass": forms.mass, "stiffness": forms.stiffness, "nitsche": lambda V: forms.nitsche(V, 10.0)}[kind](V)
    cells = _cells(order, mesh.num_cells, rng)
    qr = synthetic_ragged(cells, tdim, 3, 11, seed=11, normals=kind == "nitsche")
    _check(form, qr)


def test_fused_zero_other_kernels_fall_back_to_memset():
    """Distorted mesh (point-by-point staged kernel) and Q2 hexahedra (tensor-core kernel): same contract.
    (So does the hexahedron Nitsche case above: the quadrilateral and the Q1 hexahedron mass / stiffness moment
    kernels carry the zeroing warp.)"""
    rng = np.random.default_rng(9)
    mesh = _mesh(2, 40)
    mesh.geometry.x[:, :2] += rng.uniform(-0.2, 0.2, size=(mesh.geometry.x.shape[0], 2)) * mesh.h()
    V = _space(mesh, 2, 1)
    _check(forms.stiffness(V), synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), 2, 3, 11, seed=3))
    mesh = _mesh(3, 5)
    V = _space(mesh, 2, 1)
    _check(forms.mass(V), synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), 3, 3, 11, seed=3))


@pytest.mark.parametrize("mode", [False, None], ids=["off", "default"])
def test_zero_in_call_without_fused_zeroing(mode):
    """Fused zeroing off / the library's default policy (in the kernel for Q2 quadrilateral matrices, memset + kernel
    otherwise)."""
    asm.set_fused_zero(mode)
    mesh = _mesh(2, 60)
    V = _space(mesh, 2, 1)
    _check(forms.mass(V), synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), 2, 3, 11, seed=3))
    mesh = _mesh(3, 34)
    V = _space(mesh, 1, 1)
    _check(forms.stiffness(V), synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), 3, 8, 8, seed=3))  # short-rule shape
    _check(forms.stiffness(V), synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), 3, 8, 27, seed=3))
    _check(forms.stiffness(V), synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), 3, 40, 120, seed=3))  # long rules: memset


def test_fused_zero_empty_list_zeroes():
    mesh = _mesh(2, 6)
    V = _space(mesh, 1, 1)
    B = asm.create_matrix(forms.mass(V), zero=False)
    B.data.fill_(float("nan"))
    cq.assemble_matrix(forms.mass(V), [(np.zeros(0, dtype=np.int32), [], [])], A=B, zero=True)
    assert float(B.data.abs().sum()) == 0.0


@pytest.mark.parametrize("mode", [False, None, True], ids=["memset", "default", "in-kernel"])
@pytest.mark.parametrize("tdim,degree", [(2, 1), (2, 2), (3, 1)])
def test_zero_keep_range_is_left_alone(tdim, degree, mode):
    """cq_csr.zero_keep_lo / hi (ABI 5): the zeroing call leaves that range of value slots alone -- the interface rows of a
    peer scatter, which the caller zeroed before its fence and a neighbour may already have added into -- and zeroes
    everything else, by memset or inside the kernel."""
    import torch

    asm.set_fused_zero(mode)
    mesh = _mesh(tdim, 40 if tdim == 2 else 12)
    V = _space(mesh, degree, 1)
    form = forms.mass(V)
    qr = synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), tdim, 3, 11, seed=4)
    A = asm.create_matrix(form)
    cq.assemble_matrix(form, [qr], A=A)
    ref = A.data.cpu().numpy()
    B = asm.create_matrix(form, zero=False)
    ip = B.indptr.cpu().numpy()
    r0 = B.shape[0] // 3
    for lo, hi in ((int(ip[r0]), int(ip[r0 + 37])), (int(ip[5]) + 1, int(ip[9]) + 3), (0, int(ip[3]))):  # aligned or not
        B.zero_keep = (lo, hi)
        B.data.fill_(float("nan"))
        B.data[lo:hi] = 7.0  # "already added into by the neighbour"
        cq.assemble_matrix(form, [qr], A=B, zero=True)
        torch.cuda.synchronize()
        b = B.data.cpu().numpy()
        assert not np.isnan(b).any()
        want = ref.copy()
        want[lo:hi] += 7.0
        assert np.max(np.abs(b - want)) <= RTOL * np.max(np.abs(want))
        assert np.array_equal(b[:lo] == 0.0, ref[:lo] == 0.0) and np.array_equal(b[hi:] == 0.0, ref[hi:] == 0.0)
    B.zero_keep = None
