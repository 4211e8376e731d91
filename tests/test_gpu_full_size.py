"""BASELINE.json-size checks of the CUDA path (parity on the full problem through
size-independent properties + the oracle on a fixed random sample of >= 65,536 cells assembled into the
same sparsity).  Properties used: constants are in the kernel of the stiffness matrix (row sums 0), the
mass matrix integrates 1 (sum of entries = sum_c |K_c| sum_q w_q), symmetry (x'Ay = y'Ax), linearity
in the weights, and for scalars sum w = measure."""
import os

import numpy as np
import pytest
import torch

import customquad_b200 as cq
from customquad_b200 import assemble as asm
from customquad_b200 import forms
from customquad_b200.quadrature import synthetic_ragged

pytestmark = pytest.mark.gpu
RTOL = 1e-12
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _row_sums(A):
    rows = torch.repeat_interleave(torch.arange(A.shape[0], device=A.data.device), A.indptr[1:] - A.indptr[:-1])
    return torch.zeros(A.shape[0], dtype=torch.float64, device=A.data.device).index_add_(0, rows, A.data)


def _spmv(A, x):
    M = torch.sparse_csr_tensor(A.indptr, A.indices.to(torch.int64), A.data, size=A.shape)
    return M @ x


def _sample_parity(form, qr, A, sample, oracle, bs=1):
    """Oracle on `sample` entries of the rule vs the CUDA assembly of the same entries into the full
    pattern: compare every (row, col) the oracle produced."""
    sub = qr.select(sample)
    B = asm.create_matrix(form)
    cq.assemble_matrix(form, [sub], A=B)
    B.assemble()
    R = oracle.assemble_matrix(form, [sub], nthreads=oracle.max_threads(), pattern=_cells_pattern(form, sub)).tocoo()
    keep = R.data != 0.0
    rr, cc, vv = R.row[keep].astype(np.int64), R.col[keep].astype(np.int64), R.data[keep]
    # look the oracle's (row, col) up in the device CSR
    ip = B.indptr.cpu().numpy()
    ix = B.indices.cpu().numpy()
    data = B.data.cpu().numpy()
    pos = np.empty(len(rr), dtype=np.int64)
    for k in range(0, len(rr), 1 << 20):  # vectorised row-wise search in chunks
        r, c = rr[k : k + (1 << 20)], cc[k : k + (1 << 20)]
        lo, hi = ip[r].copy(), ip[r + 1].copy()
        while np.any(lo < hi):
            mid = (lo + hi) // 2
            act = lo < hi
            less = np.zeros_like(act)
            less[act] = ix[mid[act]] < c[act]
            lo = np.where(act & less, mid + 1, lo)
            hi = np.where(act & ~less, mid, hi)
        assert np.array_equal(ix[lo], c), "sparsity mismatch"
        pos[k : k + (1 << 20)] = lo
    got = data[pos]
    assert np.linalg.norm(got - vv) <= RTOL * np.linalg.norm(vv)
    assert np.max(np.abs(got - vv)) <= RTOL * np.max(np.abs(vv))
    # and nothing else was touched
    assert abs(np.abs(data).sum() - np.abs(got).sum()) <= 1e-9 * np.abs(got).sum()


def _cells_pattern(form, sub):
    from oracle.oracle import build_pattern  # the oracle's own builder

    V = form.function_spaces[0]
    return build_pattern(V.dofmap.list, V.dofmap.num_dofs, V.bs, cells=sub.cells)


@pytest.mark.parametrize("degree", [1, 2])
def test_config2_full_size(cuda_backend, degree):
    """BASELINE.json configs[1]: 2048 x 2048 quads, ragged U{16..64} rules, all cells."""
    from oracle import oracle

    N = 2048
    mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [N, N])
    V = cq.FunctionSpace(mesh, ("Lagrange", degree))
    qr = synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), 2, 16, 64, seed=20231)
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.rand(V.num_dofs, dtype=torch.float64, device=dev, generator=g)
    y = torch.rand(V.num_dofs, dtype=torch.float64, device=dev, generator=g)
    cell_vol = float(np.prod(mesh.h()))
    sw = np.add.reduceat(qr.w, qr.offsets[:-1])  # sum of weights per cell

    K = cq.assemble_matrix(forms.stiffness(V), [qr])
    K.assemble()
    assert float(_row_sums(K).abs().max()) <= RTOL * float(K.data.abs().max()) * 25  # constants in the kernel
    assert abs(float(x @ _spmv(K, y)) - float(y @ _spmv(K, x))) <= RTOL * abs(float(x @ _spmv(K, y)))
    M = cq.assemble_matrix(forms.mass(V), [qr])
    M.assemble()
    total = float(M.data.sum())
    assert abs(total - cell_vol * sw.sum()) <= RTOL * abs(total)  # integrates 1
    assert abs(float(x @ _spmv(M, y)) - float(y @ _spmv(M, x))) <= RTOL * abs(float(x @ _spmv(M, y)))
    # linearity in the weights: doubling them doubles the matrix
    q2 = synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), 2, 16, 64, seed=20231)
    q2.w *= 2.0
    K2 = cq.assemble_matrix(forms.stiffness(V), [q2])
    K2.assemble()
    assert float((K2.data - 2.0 * K.data).abs().max()) <= RTOL * float(K.data.abs().max())
    del K2, q2
    # oracle on a fixed random sample of 65,536 cells (seeded), both forms
    sample = np.sort(np.random.default_rng(7).choice(mesh.num_cells, 65536, replace=False))
    _sample_parity(forms.stiffness(V), qr, K, sample, oracle)
    _sample_parity(forms.mass(V), qr, M, sample, oracle)


def test_config3_shape_hex_q1(cuda_backend):
    """BASELINE.json configs[2] shape: Q1 hexahedra, cut-cell volume rules U{8..27} and surface rules
    with normals; 128^3 here (the 256^3 pattern needs 5.5 GB of setup; same kernels)."""
    from oracle import oracle

    mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [128, 128, 128])
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    rng = np.random.default_rng(3)
    cut = np.sort(rng.choice(mesh.num_cells, 200000, replace=False)).astype(np.int32)  # ~2e5 cut cells (cfg 3)
    qb = synthetic_ragged(cut, 3, 8, 27, seed=11)
    qs = synthetic_ragged(cut, 3, 4, 16, seed=12, normals=True)
    vol = cq.assemble_scalar(forms.integral(mesh, 1.0), [qb])
    sw = np.add.reduceat(qb.w, qb.offsets[:-1]).sum()
    assert abs(vol - float(np.prod(mesh.h())) * sw) <= RTOL * abs(vol)
    K = cq.assemble_matrix(forms.stiffness(V), [qb])
    K.assemble()
    assert float(_row_sums(K).abs().max()) <= RTOL * float(K.data.abs().max()) * 27
    Nm = cq.assemble_matrix(forms.nitsche(V, 10.0), [qs])
    Nm.assemble()
    dev = K.data.device
    x = torch.rand(V.num_dofs, dtype=torch.float64, device=dev)
    y = torch.rand(V.num_dofs, dtype=torch.float64, device=dev)
    assert abs(float(x @ _spmv(Nm, y)) - float(y @ _spmv(Nm, x))) <= 1e-11 * abs(float(x @ _spmv(Nm, y)))  # symmetric Nitsche form
    sample = np.sort(rng.choice(len(cut), 65536, replace=False))
    _sample_parity(forms.stiffness(V), qb, K, sample, oracle)
    _sample_parity(forms.nitsche(V, 10.0), qs, Nm, sample, oracle)
    b = cq.assemble_vector(forms.nitsche_rhs(V, cq.Analytic.sinprod(1.0), 10.0), [qs]).array
    r = oracle.assemble_vector(forms.nitsche_rhs(V, cq.Analytic.sinprod(1.0), 10.0), [qs], nthreads=oracle.max_threads())
    assert np.max(np.abs(b - r)) <= RTOL * np.max(np.abs(r))


def test_config4_shape_hex_q2_vector(cuda_backend):
    """BASELINE.json configs[3] shape: vector Poisson on Q2 hexahedra (bs = 3), 3^3 Gauss points per
    cell; 24^3 here (same kernels: FP64 tensor-core path + blocked scatter)."""
    from customquad_b200.quadrature import gauss_tensor_rule
    from oracle import oracle

    mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [24, 24, 24])
    V = cq.VectorFunctionSpace(mesh, ("Lagrange", 2))
    pts, w = gauss_tensor_rule(3, 4)
    C = mesh.num_cells
    qr = cq.quadrature.normalise_qr((np.arange(C), np.tile(pts.reshape(-1), [C, 1]), np.tile(w, [C, 1])), 3)
    K = cq.assemble_matrix(forms.stiffness(V), [qr])
    K.assemble()
    assert float(_row_sums(K).abs().max()) <= 1e-11 * float(K.data.abs().max())  # rigid translations
    M = cq.assemble_matrix(forms.mass(V), [qr])
    M.assemble()
    vol = float(np.prod(np.asarray(mesh.xmax) - np.asarray(mesh.xmin)))
    assert abs(float(M.data.sum()) - 3 * vol) <= 1e-11 * 3 * vol  # exact rule: each component integrates 1
    sample = np.sort(np.random.default_rng(2).choice(C, 4096, replace=False))
    _sample_parity(forms.stiffness(V), qr, K, sample, oracle, bs=3)
    f = cq.Analytic.sinprod(2 * np.pi**2, 0.0, 0.0)  # demo/vector_poisson.py f
    b = cq.assemble_vector(forms.load(V, f), [qr]).array
    r = oracle.assemble_vector(forms.load(V, f), [qr], nthreads=oracle.max_threads())
    assert np.max(np.abs(b - r)) <= RTOL * np.max(np.abs(r))


def test_long_list_takes_the_ticket_path(cuda_backend):
    """Lists longer than 512 rounds of grid x slots stages draw their stages from a ticket counter (cq_launch.cuh:
    choose_dynamic, the config-5 regime) instead of the static interleaving.  1.2e7 one-point entries on a 3500^2 Q1
    mesh: linearity, total mass, and the oracle on a sample of cells."""
    from oracle import oracle

    N = 3500
    mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [N, N])
    V = cq.FunctionSpace(mesh, ("Lagrange", 1))
    qr = synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), 2, 1, 1, seed=3)
    assert qr.num_entries > 512 * 148 * 4 * 32
    M = cq.assemble_matrix(forms.mass(V), [qr])
    M.assemble()
    cell_vol = float(np.prod(mesh.h()))
    assert abs(float(M.data.sum()) - cell_vol * qr.w.sum()) <= RTOL * abs(float(M.data.sum()))
    K = cq.assemble_matrix(forms.stiffness(V), [qr])
    K.assemble()
    assert float(_row_sums(K).abs().max()) <= RTOL * float(K.data.abs().max()) * 25
    # the full-list matrices themselves against the oracle: rows of the dofs strictly inside a block of cells get all
    # their contributions from that block, so the oracle on the block's entries must reproduce those rows of K and M
    for a, b0, m in ((100, 2000, 160), (3300, 37, 160)):
        ix, iy = np.meshgrid(np.arange(a, a + m), np.arange(b0, b0 + m), indexing="ij")
        block = (ix * N + iy).reshape(-1)  # cell id = ix * ny + iy
        sub = qr.select(block)
        di, dj = np.meshgrid(np.arange(a + 1, a + m), np.arange(b0 + 1, b0 + m), indexing="ij")
        rows = (di * (N + 1) + dj).reshape(-1)  # Q1 dof id = i * (ny + 1) + j
        for form, A in ((forms.stiffness(V), K), (forms.mass(V), M)):
            R = oracle.assemble_matrix(form, [sub], nthreads=oracle.max_threads(), pattern=_cells_pattern(form, sub)).tocsr()
            ip = A.indptr.cpu().numpy()
            sel = torch.as_tensor((ip[rows][:, None] + np.arange(9)[None, :]).reshape(-1), device=A.data.device)  # interior rows hold 9 entries
            got_v, got_c = A.data[sel].cpu().numpy(), A.indices[sel].cpu().numpy()
            ref = R[rows]
            assert np.array_equal(np.diff(ref.indptr), np.full(len(rows), 9)) and np.array_equal(ref.indices, got_c)
            assert np.max(np.abs(got_v - ref.data)) <= RTOL * np.max(np.abs(ref.data))
    b = cq.assemble_vector(forms.load(V, 1.0), [qr]).array  # the vector kernels share the producer
    assert abs(b.sum() - cell_vol * qr.w.sum()) <= RTOL * abs(b.sum())


def _bench_leg(config, timeout=900):
    """One extra leg of bench.py in a process of its own (its allocations: config 4 holds 133 GB)."""
    import json
    import subprocess
    import sys

    torch.cuda.empty_cache()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--only-config", str(config)], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_config3_full_size_256_on_device_rules(cuda_backend):
    """BASELINE.json configs[2] at its full size: sphere on Q1 hexahedra 256^3, rules from the device producer; volume
    and area of the unit sphere, and a 32768-cut-cell sample of the stiffness and Nitsche matrices against the oracle."""
    d = _bench_leg(3)
    assert d["cells"] == 256**3 and d["cut_cells"] > 150000
    assert d["volume_rel_err"] < 1e-8 and d["area_rel_err"] < 1e-6
    assert d["parity"]["ok"] and d["parity"]["stiffness_cut"]["sparsity_ok"] and d["parity"]["nitsche_cut"]["max_rel"] <= RTOL


def test_config4_full_size_128_q2_vector(cuda_backend):
    """BASELINE.json configs[3] at its full size: vector Poisson on Q2 hexahedra 128^3 (9.69e9 non-zeros, 116 GB of CSR).
    The sum of all entries of the stiffness matrix vanishes, a 2048-cell sample matches the oracle to 1e-12."""
    d = _bench_leg(4)
    if d.get("skipped"):
        pytest.skip(d["skipped"])
    assert d["nnz"] == 9692015625
    assert d["sum_of_all_entries_rel"] < 1e-13
    assert d["parity"]["ok"] and d["parity"]["sparsity_ok"] and d["parity"]["max_rel"] <= RTOL
