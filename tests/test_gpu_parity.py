"""CUDA path vs the CPU oracle on identical seeded inputs, through the public API -> C-ABI.

Tolerance (BASELINE.json north_star: values within 1e-12 relative in FP64 given atomic
reordering): ||D||_F <= 1e-12 ||ref||_F and max|D| <= 1e-12 max|ref|.
Sparsity pattern (indptr, indices) bit-exact."""
import numpy as np
import pytest

import customquad_b200 as cq
from customquad_b200 import forms
from customquad_b200.quadrature import synthetic_ragged
from tests import refcases

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def close(a, ref, rtol=RTOL):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.max(np.abs(ref)) if ref.size else 0.0
    if scale == 0.0:
        return bool(np.all(a == 0.0))
    return np.linalg.norm(a - ref) <= rtol * np.linalg.norm(ref) and np.max(np.abs(a - ref)) <= rtol * scale


def make_mesh(tdim, distort, seed=7):
    if tdim == 2:
        mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [7, 5])  # demo/poisson.py:42-43
    else:
        mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [4, 3, 5])  # :48-49
    if distort:  # non-affine cells: J varies inside the cell
        rng = np.random.default_rng(seed)
        h = mesh.h()
        mesh.geometry.x[:, :tdim] += rng.uniform(-0.2, 0.2, size=(mesh.geometry.x.shape[0], tdim)) * h
    return mesh


def all_forms(V, rng):
    """Every integrand of SURVEY.md 8a a-6 on the space V."""
    mesh, bs, g = V.mesh, V.bs, V.mesh.geometry.dim
    f = cq.Function(V)
    f.array[:] = rng.uniform(-1, 1, size=V.num_dofs)
    amp = [g * np.pi**2 * (1.0 + 0.5 * a) for a in range(bs)]
    sinp = cq.Analytic.sinprod(*amp)
    aff = cq.Analytic.affine_blocked([(0.3 + a, [2.0, 1.0, -0.7][:g]) for a in range(bs)])
    const = cq.Analytic.const(*[1.0 + a for a in range(bs)])
    out = {
        "mass": forms.mass(V),
        "stiffness": forms.stiffness(V),
        "nitsche": forms.nitsche(V, 10.0),
        "load_coef": forms.load(V, f),
        "load_sinprod": forms.load(V, sinp),
        "load_affine": forms.load(V, aff),
        "load_const": forms.load(V, const),
        "nitsche_rhs_coef": forms.nitsche_rhs(V, f, 10.0),
        "nitsche_rhs_sinprod": forms.nitsche_rhs(V, sinp, 10.0),
        "integral_sq_coef": forms.integral_sq(f),
        "l2_err_sinprod": forms.l2_error(f, sinp),
        "l2_err_affine": forms.l2_error(f, aff),
        "h10_err_sinprod": forms.h10_error(f, sinp),
        "h10_err_affine": forms.h10_error(f, aff),
    }
    if bs == 1:
        out["integral_one"] = forms.integral(mesh, 1.0)
        out["integral_affine"] = forms.integral(mesh, cq.Analytic.affine(0.0, [2.0, 1.0, 0.5][:g]))
        out["integral_sinprod"] = forms.integral(mesh, cq.Analytic.sinprod(1.0))
    return out


SPACES = [(2, 1, 1), (2, 2, 1), (3, 1, 1), (3, 2, 1), (2, 1, 2), (2, 2, 2), (3, 1, 3), (3, 2, 3)]


@pytest.mark.parametrize("cell_type,degree", [(4, 1), (4, 2), (5, 1), (5, 2)])
def test_tabulate_matches_call_basix(cuda_backend, cell_type, degree):
    from customquad_b200.assemble import tabulate
    from oracle import oracle

    gdim = 2 if cell_type == 4 else 3
    pts = np.random.default_rng(3).random((257, gdim))
    for deriv in range(gdim + 1):
        dev = tabulate(cell_type, degree, deriv, pts)
        ref = oracle.tabulate(cell_type, degree, deriv, pts)
        assert np.max(np.abs(dev - ref)) <= 1e-14 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("tdim,degree,bs", SPACES)
@pytest.mark.parametrize("distort", [False, True])
def test_element_tensors_match_oracle(cuda_backend, tdim, degree, bs, distort):
    """One generated-kernel call per cell (README.md:133-158), before any scatter."""
    from customquad_b200.assemble import element_tensors
    from oracle import oracle

    rng = np.random.default_rng(11)
    mesh = make_mesh(tdim, distort)
    V = cq.FunctionSpace(mesh, ("Lagrange", degree)) if bs == 1 else cq.VectorFunctionSpace(mesh, ("Lagrange", degree))
    cells = rng.permutation(mesh.num_cells)[: max(3, mesh.num_cells // 2)]
    qr = synthetic_ragged(cells, tdim, 1, 19, seed=5, normals=True)
    for name, form in all_forms(V, rng).items():
        dev = element_tensors(form, qr)
        ref = oracle.element_tensors(form, qr)
        assert dev.shape == ref.shape
        assert close(dev, ref), f"{name}: max rel {np.max(np.abs(dev - ref)) / np.max(np.abs(ref)):.3e}"


@pytest.fixture(params=[1, 0], ids=["direct", "auto-staged"])
def variant(request, cuda_backend):
    """Both kernel families: 1 = direct-load kernel, 0 = auto (bulk-copy staged kernel wherever
    it applies)."""
    from customquad_b200.assemble import set_variant

    set_variant(request.param)
    yield request.param
    set_variant(0)


@pytest.mark.parametrize("tdim,degree,bs", SPACES)
def test_assembly_matches_oracle(cuda_backend, oracle_backend, variant, tdim, degree, bs):
    """Full path: element kernels + scatter into CSR / vector / scalar, ragged rules, cell subset
    with repeats."""
    rng = np.random.default_rng(23)
    mesh = make_mesh(tdim, True)
    V = cq.FunctionSpace(mesh, ("Lagrange", degree)) if bs == 1 else cq.VectorFunctionSpace(mesh, ("Lagrange", degree))
    cells = np.concatenate([rng.permutation(mesh.num_cells), rng.integers(0, mesh.num_cells, 9)])
    qr = synthetic_ragged(cells, tdim, 2, 33, seed=9, normals=True)
    for name, form in all_forms(V, rng).items():
        if form.rank == 2:
            A, R = cuda_backend.matrix(form, [qr]), oracle_backend.matrix(form, [qr])
            assert np.array_equal(A.indptr, R.indptr) and np.array_equal(A.indices, R.indices), name
            assert close(A.data, R.data), f"{name}: {np.max(np.abs(A.data - R.data)) / np.max(np.abs(R.data)):.3e}"
        elif form.rank == 1:
            assert close(cuda_backend.vector(form, [qr]), oracle_backend.vector(form, [qr])), name
        else:
            m, r = cuda_backend.scalar(form, [qr]), oracle_backend.scalar(form, [qr])
            assert abs(m - r) <= RTOL * abs(r), f"{name}: {m} vs {r}"


# ---- the reference's own tests on the CUDA path -----------------------------------------------
@pytest.mark.parametrize("idx", range(36))
def test_reference_quads_hexes_assembly(cuda_backend, golden, idx):
    refcases.assembly_case(cuda_backend, golden, idx)


def test_reference_edges(cuda_backend, golden):
    refcases.edges_case(cuda_backend, golden)


def test_reference_edge_integral(cuda_backend, golden):
    refcases.edge_integral_case(cuda_backend, golden)


def test_reference_face_integral(cuda_backend, golden):
    refcases.face_integral_case(cuda_backend, golden)


def test_reference_volume(cuda_backend, golden):
    refcases.volume_case(cuda_backend, golden)


def test_reference_full_area(cuda_backend, golden):
    refcases.full_area_case(cuda_backend, golden)


# ---- input conventions and edge cases -----------------------------------------------------------
def test_input_conventions(cuda_backend, oracle_backend):
    """cells as list / int32 / int64, qr as list of flat arrays or 2-D arrays."""
    mesh = make_mesh(2, False)
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    rng = np.random.default_rng(1)
    cells = [3, 0, 7, 7, 12]
    pts = [rng.random(2 * n) for n in (4, 1, 9, 2, 6)]
    w = [rng.random(n) for n in (4, 1, 9, 2, 6)]
    ref = oracle_backend.matrix(forms.stiffness(V), [(cells, pts, w)])
    for c in (cells, np.array(cells, dtype=np.int32), np.array(cells, dtype=np.int64)):
        A = cuda_backend.matrix(forms.stiffness(V), [(c, pts, w)])
        assert close(A.data, ref.data)
    P2 = rng.random((5, 6))
    W2 = rng.random((5, 3))
    A = cuda_backend.matrix(forms.mass(V), [(cells, P2, W2)])
    assert close(A.data, oracle_backend.matrix(forms.mass(V), [(cells, P2, W2)]).data)


@pytest.mark.parametrize("nq_lo,nq_hi", [(0, 70), (40, 200)], ids=["short", "long-rules"])
@pytest.mark.parametrize("degree,bs", [(1, 1), (2, 1), (1, 2), (2, 2)])
def test_moment_kernel_matches_oracle(cuda_backend, oracle_backend, degree, bs, nq_lo, nq_hi):
    """Axis-aligned quadrilateral meshes take the monomial-moment kernel for mass / stiffness
    (cq_moment.cuh); the result must equal both the oracle and the point-by-point kernels.  A mean of more than 64
    points per cell takes the long-rule shape (more moment warps, four points in flight per lane; entries that
    overflow a stage are read from global memory)."""
    from customquad_b200.assemble import mesh_on_device, set_variant
    import torch

    mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [23, 17])
    assert mesh_on_device(mesh, torch.device("cuda", torch.cuda.current_device()))[2] == 1  # CQ_MESH_AXIS_ALIGNED
    V = cq.FunctionSpace(mesh, ("Lagrange", degree)) if bs == 1 else cq.VectorFunctionSpace(mesh, ("Lagrange", degree))
    rng = np.random.default_rng(31)
    cells = np.concatenate([rng.permutation(mesh.num_cells), rng.integers(0, mesh.num_cells, 40)])
    qr = synthetic_ragged(cells, 2, nq_lo, nq_hi, seed=13, normals=True)
    for form in (forms.mass(V), forms.stiffness(V)):
        R = oracle_backend.matrix(form, [qr])
        set_variant(0)
        A = cuda_backend.matrix(form, [qr])
        set_variant(3)  # same staged kernel family without moments
        B = cuda_backend.matrix(form, [qr])
        set_variant(0)
        assert close(A.data, R.data), f"{form.kind.name}: {np.max(np.abs(A.data - R.data)) / np.max(np.abs(R.data)):.3e}"
        assert close(B.data, R.data)
    dist = make_mesh(2, True)
    assert mesh_on_device(dist, torch.device("cuda", torch.cuda.current_device()))[2] == 0


@pytest.mark.parametrize("nq_lo,nq_hi", [(0, 40), (36, 52), (40, 200)], ids=["short", "long-rules-3-slots", "long-rules"])
@pytest.mark.parametrize("bs", [1, 3])
def test_moment_kernel_hexahedra(cuda_backend, oracle_backend, bs, nq_lo, nq_hi):
    """Axis-aligned Q1 hexahedra (create_box) take the 27-moment kernel for mass / stiffness; with rules of 40..200
    points per cell (real cut cells: config 3) most entries of a stage overflow its shared-memory slice and are read
    from global memory; means of 33..56 points per cell take the three-slot shape of the long-rule kernel
    (cq_launch.cuh:launch_auto)."""
    from customquad_b200.assemble import mesh_on_device, set_variant
    import torch

    mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [7, 5, 6])
    assert mesh_on_device(mesh, torch.device("cuda", torch.cuda.current_device()))[2] == 1  # CQ_MESH_AXIS_ALIGNED
    V = cq.FunctionSpace(mesh, ("Lagrange", 1)) if bs == 1 else cq.VectorFunctionSpace(mesh, ("Lagrange", 1))
    rng = np.random.default_rng(37)
    cells = np.concatenate([rng.permutation(mesh.num_cells), rng.integers(0, mesh.num_cells, 40), np.arange(mesh.num_cells)])
    qr = synthetic_ragged(cells, 3, nq_lo, nq_hi, seed=17, normals=True)
    for form in (forms.mass(V), forms.stiffness(V), forms.nitsche(V, 10.0)):
        R = oracle_backend.matrix(form, [qr])
        set_variant(0)
        A = cuda_backend.matrix(form, [qr])
        set_variant(3)  # same staged kernel family without moments
        B = cuda_backend.matrix(form, [qr])
        set_variant(0)
        assert close(A.data, R.data), f"{form.kind.name}: {np.max(np.abs(A.data - R.data)) / np.max(np.abs(R.data)):.3e}"
        assert close(B.data, R.data)


def test_staged_kernel_stage_overflow_and_tails(cuda_backend, oracle_backend):
    """Staged kernel corner cases: a rule far longer than a shared-memory stage (read from global
    memory), an odd total point count (manual tail copy), many stages, zero-point entries."""
    from customquad_b200.assemble import set_variant

    set_variant(2)
    try:
        rng = np.random.default_rng(5)
        for tdim in (2, 3):
            mesh = make_mesh(tdim, True)
            V = cq.FunctionSpace(mesh, ("Lagrange", 1))
            nq = rng.integers(0, 40, size=300)
            nq[17] = 9001  # overflows any stage
            nq[-1] = nq[-1] | 1
            if nq.sum() % 2 == 0:
                nq[-1] += 2
                nq[5] += 1
            assert nq.sum() % 2 == 1
            cells = rng.integers(0, mesh.num_cells, size=300)
            pts = [rng.random(tdim * n) for n in nq]
            w = [rng.random(n) / max(n, 1) for n in nq]
            nr = [rng.standard_normal(tdim * n) for n in nq]
            for form in (forms.stiffness(V), forms.nitsche(V, 10.0)):
                A, R = cuda_backend.matrix(form, [(cells, pts, w, nr)]), oracle_backend.matrix(form, [(cells, pts, w, nr)])
                assert close(A.data, R.data)
            lf = forms.load(V, cq.Analytic.sinprod(3.0))
            assert close(cuda_backend.vector(lf, [(cells, pts, w)]), oracle_backend.vector(lf, [(cells, pts, w)]))
            m, r = cuda_backend.scalar(forms.integral(mesh, 1.0), [(cells, pts, w)]), oracle_backend.scalar(forms.integral(mesh, 1.0), [(cells, pts, w)])
            assert abs(m - r) <= RTOL * abs(r)
    finally:
        set_variant(0)


def test_empty_and_degenerate(cuda_backend, oracle_backend):
    mesh = make_mesh(2, False)
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    empty = (np.zeros(0, dtype=np.int32), [], [])
    assert cuda_backend.matrix(forms.mass(V), [empty]).data.sum() == 0.0
    assert np.all(cuda_backend.vector(forms.load(V, 1.0), [empty]) == 0.0)
    assert cuda_backend.scalar(forms.integral(mesh, 1.0), [empty]) == 0.0
    # entries with zero points among non-empty ones; one entry with a very long rule
    rng = np.random.default_rng(2)
    nq = [0, 5, 0, 3001, 1, 0]
    cells = [1, 2, 3, 4, 5, 6]
    pts = [rng.random(2 * n) for n in nq]
    w = [rng.random(n) / max(n, 1) for n in nq]
    for form in (forms.stiffness(V), forms.mass(V)):
        A, R = cuda_backend.matrix(form, [(cells, pts, w)]), oracle_backend.matrix(form, [(cells, pts, w)])
        assert close(A.data, R.data)
    m = cuda_backend.scalar(forms.integral(mesh, 1.0), [(cells, pts, w)])
    assert abs(m - oracle_backend.scalar(forms.integral(mesh, 1.0), [(cells, pts, w)])) <= RTOL * abs(m)


def test_error_behaviour(cuda_backend):
    """Same asserts as the reference (assemble_matrix.py:70-73) + loud failures."""
    from customquad_b200._lib import CqError

    mesh = make_mesh(2, False)
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    with pytest.raises(AssertionError):
        cq.assemble_matrix(forms.mass(V), [([0, 1], [np.zeros(2)], [np.zeros(1), np.zeros(1)])])
    with pytest.raises(AssertionError):
        cq.assemble_matrix(forms.mass(V), [([0, 1], [np.zeros(2), np.zeros(2)], [np.zeros(1)])])
    with pytest.raises(AssertionError):
        cq.assemble_matrix(forms.nitsche(V, 10.0), [([0], [np.zeros(2)], [np.ones(1)])])  # no normals
    # a cell outside the "cells" pattern is reported at assemble() (PETSc: new nonzero error)
    A = cq.assemble.create_matrix(forms.mass(V), pattern_cells=[0, 1])
    cq.assemble_matrix(forms.mass(V), [([mesh.num_cells - 1], [np.full(2, 0.5)], [np.ones(1)])], A=A)
    with pytest.raises(CqError):
        A.assemble()


def test_pattern_builders_agree(cuda_backend):
    import torch

    from customquad_b200.sparsity import build_pattern, build_pattern_torch

    for tdim, degree, bs in [(2, 2, 1), (3, 1, 3), (2, 1, 2)]:
        mesh = make_mesh(tdim, False)
        V = cq.FunctionSpace(mesh, ("Lagrange", degree), bs) if bs > 1 else cq.FunctionSpace(mesh, ("Lagrange", degree))
        p = build_pattern(V.dofmap.list, V.dofmap.num_dofs, bs)
        ip, ix = build_pattern_torch(torch.from_numpy(V.dofmap.list).cuda(), V.dofmap.num_dofs, bs, chunk_pairs=200)
        assert np.array_equal(ip.cpu().numpy(), p.indptr) and np.array_equal(ix.cpu().numpy(), p.indices)


def test_host_buffer_entry_points(cuda_backend, oracle_backend):
    """cq_assemble_*_host: qr and result in host memory, chunked copies inside the call."""
    import torch

    from customquad_b200 import assemble as asm

    mesh = cq.create_rectangle([[0, 0], [1, 1]], [300, 300])  # > 2^16 entries => chunked path
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    qr = synthetic_ragged(np.arange(mesh.num_cells), 2, 1, 6, seed=4)
    form = forms.stiffness(V)
    A = asm.create_matrix(form)
    vals = np.zeros(A.nnz)
    asm.assemble_matrix_host(form, qr, A, vals)
    R = oracle_backend.matrix(form, [qr])
    assert close(vals, R.data)
    # rows go back to the host as soon as no later chunk touches them: reversed and shuffled cell
    # orders (nothing / little is final early) and a degree-2 space must give the same values
    rng = np.random.default_rng(8)
    V2 = cq.FunctionSpace(mesh, ("Lagrange", 2))
    for cells, space in ((np.arange(mesh.num_cells)[::-1].copy(), V), (rng.permutation(mesh.num_cells), V), (np.arange(mesh.num_cells), V2)):
        q2 = synthetic_ragged(cells, 2, 1, 6, seed=5)
        f2 = forms.mass(space)
        A2 = asm.create_matrix(f2)
        v2 = np.full(A2.nnz, np.nan)
        asm.assemble_matrix_host(f2, q2, A2, v2)
        assert close(v2, oracle_backend.matrix(f2, [q2]).data)
    b = cq.la.Vector(torch.zeros(V.num_dofs, dtype=torch.float64, device="cuda"), 0)
    bh = np.zeros(V.num_dofs)
    lf = forms.load(V, cq.Analytic.sinprod(2 * np.pi**2))
    asm.assemble_vector_host(lf, qr, b, bh)
    assert close(bh, oracle_backend.vector(lf, [qr]))
    m = asm.assemble_scalar_host(forms.integral(mesh, 1.0), qr)
    assert abs(m - oracle_backend.scalar(forms.integral(mesh, 1.0), [qr])) <= RTOL * abs(m)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,degree", [((6, 24, 5), 1), ((13, 19, 7), 1), ((4, 16, 3), 2)])
def test_plane_tile_traversal_hint(cuda_backend, oracle_backend, monkeypatch, shape, degree):
    """cq_quadrature.plane_entries / plane_tiles: the persistent kernels walk a layer-by-layer list tile by tile (a
    bijection of the stages, kernels.cuh:stage_of).  Same matrices as the oracle, for exact and drifting planes
    (per-layer cell counts that are / are not multiples of the 32-entry stage) and with a partial last stage."""
    from customquad_b200 import assemble as asm

    mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], list(shape))
    V = cq.FunctionSpace(mesh, ("Lagrange", degree))
    qr = synthetic_ragged(np.arange(mesh.num_cells), 3, 1, 12, seed=23, normals=True)
    monkeypatch.setattr(asm, "L2_TILE_BYTES", 1)  # every matrix "exceeds the L2": as many tiles as the hint allows
    form = forms.stiffness(V)
    A = asm.create_matrix(form)
    c = asm._Call(form, qr, asm._dev(), A=A)
    assert c.qr.plane_entries == shape[0] * shape[1] and 1 < c.qr.plane_tiles <= shape[1] // 2
    for form in (forms.stiffness(V), forms.mass(V), forms.nitsche(V, 10.0)):
        R = oracle_backend.matrix(form, [qr])
        S = cuda_backend.matrix(form, [qr])
        assert np.array_equal(S.indptr, R.indptr) and np.array_equal(S.indices, R.indices)
        assert close(S.data, R.data), form.kind.name
    # a list that does not start at a layer boundary, or is not consecutive, gets no hint
    sub = synthetic_ragged(np.arange(3, mesh.num_cells), 3, 1, 4, seed=1)
    assert asm._Call(forms.mass(V), sub, asm._dev(), A=A).qr.plane_tiles == 0


def test_far_from_origin_mesh_matches_recentred_oracle(cuda_backend, oracle_backend):
    """Rank 7's slab of the N = 8 weak-scaling mesh (x ~ 20, h = 1.3e-3: kappa = max|x| / h = 1.6e4).  The reference's
    J = sum_v x_v grad psi_v carries a rounding error of a few eps * kappa of its own, so on identical inputs the two
    sides agree to max(1e-12, 4 eps kappa) (bench.py:parity_tolerance); against the oracle on the SAME cells with the
    geometry translated to the origin (these integrands do not depend on x) the CUDA path agrees to 1e-12: its cell
    constants come from vertex differences, which are exact.  profiles/r2s2/far_mesh_parity.jsonl."""
    import bench
    import customquad_b200 as cq
    from customquad_b200 import dist as cqdist, forms
    from customquad_b200.quadrature import synthetic_ragged

    world, ny = 8, 64
    hx = (bench.XMAX[0] - bench.XMIN[0]) / bench.MESH_N
    slab = cqdist.make_slab([[bench.XMIN[0], bench.XMIN[1]], [bench.XMIN[0] + bench.MESH_N * world * hx, bench.XMIN[1] + ny * hx]],
                            [bench.MESH_N * world, ny], world - 1, world)  # fmt: skip
    mesh = slab.mesh
    tol, kappa = bench.parity_tolerance(mesh)
    assert kappa > 1e4 and tol < 2e-11
    cells = np.arange(8 * ny, dtype=np.int32)
    qr = synthetic_ragged(cells, 2, 16, 64, seed=11)
    x = mesh.geometry.x
    for deg in (1, 2):
        V = cq.FunctionSpace(mesh, ("Lagrange", deg))
        for name in ("stiffness", "mass"):
            f = getattr(forms, name)(V)
            S = cuda_backend.matrix(f, [qr])
            R = oracle_backend.matrix(f, [qr])
            try:
                mesh.geometry.x = x - np.array([x[:, 0].min(), 0.0, 0.0])
                R0 = oracle_backend.matrix(f, [qr])
            finally:
                mesh.geometry.x = x
            assert np.array_equal(S.indices, R.indices) and np.array_equal(S.indptr, R.indptr)
            sc = np.abs(R0.data).max()
            assert np.abs(S.data - R.data).max() <= tol * sc
            assert np.abs(S.data - R0.data).max() <= 1e-12 * sc


def test_device_resident_coefficient(cuda_backend, oracle_backend):
    """A coefficient whose values live in HBM (Function.to_device / device_array) packs without an upload and gives the
    same vectors and functionals as the host array (reference: pack_coefficients, assemble_vector.py:16)."""
    import torch

    rng = np.random.default_rng(3)
    for tdim, degree, bs in ((2, 2, 1), (3, 1, 3)):
        mesh = make_mesh(tdim, True)
        V = cq.FunctionSpace(mesh, ("Lagrange", degree)) if bs == 1 else cq.VectorFunctionSpace(mesh, ("Lagrange", degree))
        f = cq.Function(V)
        f.array[:] = rng.uniform(-1, 1, V.num_dofs)
        qr = synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), tdim, 1, 9, seed=3, normals=True)
        lf = forms.load(V, f)
        ref = oracle_backend.vector(lf, [qr])
        host = cuda_backend.vector(lf, [qr])
        f.to_device()
        assert f.device_array.is_cuda
        f.array[:] = 0.0  # the device copy is what counts now
        dev = cuda_backend.vector(lf, [qr])
        assert close(dev, host) and close(dev, ref)  # (the adds are floating-point atomics: equal up to summation order)
        f.device_array = torch.zeros_like(f.device_array)
        assert np.all(cuda_backend.vector(lf, [qr]) == 0.0)
