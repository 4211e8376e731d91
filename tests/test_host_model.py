"""Host-side logic: mesh orderings, dofmaps, sparsity, qr normalisation (no GPU)."""
import numpy as np
import pytest

import customquad_b200 as cq
from customquad_b200 import fem, forms
from customquad_b200.quadrature import gauss_tensor_rule, normalise_qr, synthetic_ragged
from customquad_b200.sparsity import build_pattern


def test_quad_local_vertex_order_is_dolfinx():
    """local vertex 1 is the +y neighbour (reference test_linear_basis.py:181)."""
    m = cq.create_rectangle([[0, 0], [2, 3]], [2, 3])
    x = m.geometry.x[m.geometry.dofmap]  # [cells, 4, 3]
    assert np.allclose(x[:, 1] - x[:, 0], [0, 1, 0])
    assert np.allclose(x[:, 2] - x[:, 0], [1, 0, 0])
    assert np.allclose(x[:, 3] - x[:, 0], [1, 1, 0])


def test_hex_local_vertex_order_is_basix():
    m = cq.create_box([[0, 0, 0], [2, 3, 4]], [2, 3, 4])
    x = m.geometry.x[m.geometry.dofmap]
    for v in range(8):
        assert np.allclose(x[:, v] - x[:, 0], [(v >> 0) & 1, (v >> 1) & 1, (v >> 2) & 1])


@pytest.mark.parametrize("tdim,degree", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_dofmap_positions_match_lattice(tdim, degree):
    """Global dof coordinates of a cell's local dofs sit at the element's lattice points, in the
    frame the quadrature points use."""
    m = cq.create_rectangle([[-1, 0.5], [2, 1.5]], [3, 2]) if tdim == 2 else cq.create_box([[0, 0, 0], [1, 2, 3]], [2, 2, 3])
    V = cq.FunctionSpace(m, ("Lagrange", degree))
    X = V.tabulate_dof_coordinates()
    lat = fem.point_frame_lattice(m.cell_type, degree)
    h = m.h()
    x0 = m.geometry.x[m.geometry.dofmap[:, 0], :tdim]  # lower corner of each cell
    for j in range(lat.shape[0]):
        expect = x0 + lat[j] / degree * h
        assert np.allclose(X[V.dofmap.list[:, j], :tdim], expect)
    assert len(np.unique(V.dofmap.list)) == V.dofmap.num_dofs


def test_q1_dofmap_equals_geometry_dofmap():
    for m in (cq.create_rectangle([[0, 0], [1, 1]], [4, 3]), cq.create_box([[0, 0, 0], [1, 1, 1]], [2, 3, 4])):
        V = cq.FunctionSpace(m, ("Lagrange", 1))
        assert np.array_equal(V.dofmap.list, m.geometry.dofmap)


def test_partition_of_unity_and_interpolation():
    rng = np.random.default_rng(0)
    for ct, tdim in ((cq.CellType.quadrilateral, 2), (cq.CellType.hexahedron, 3)):
        for degree in (1, 2):
            pts = rng.random((20, tdim))
            psi, dpsi = fem.tabulate(ct, degree, pts)
            assert np.allclose(psi.sum(1), 1) and np.allclose(dpsi.sum(1), 0)
            lat = fem.point_frame_lattice(ct, degree) / degree
            assert np.allclose(fem.tabulate(ct, degree, lat)[0], np.eye(lat.shape[0]))


def test_pattern_q1_quad_counts():
    m = cq.create_rectangle([[0, 0], [1, 1]], [8, 8])
    V = cq.FunctionSpace(m, ("Lagrange", 1))
    p = build_pattern(V.dofmap.list, V.dofmap.num_dofs)
    assert p.nnz == (3 * 8 + 1) ** 2  # (3N+1)^2 for Q1 on an NxN grid; 2048 -> 37,761,025
    for r in range(p.num_rows):
        row = p.indices[p.indptr[r] : p.indptr[r + 1]]
        assert np.all(np.diff(row) > 0)
    pb = build_pattern(V.dofmap.list, V.dofmap.num_dofs, bs=2)
    assert pb.nnz == 4 * p.nnz and pb.num_rows == 2 * p.num_rows
    dense = np.zeros((p.num_rows, p.num_rows), bool)
    for r in range(p.num_rows):
        dense[r, p.indices[p.indptr[r] : p.indptr[r + 1]]] = True
    db = np.zeros((pb.num_rows, pb.num_rows), bool)
    for r in range(pb.num_rows):
        row = pb.indices[pb.indptr[r] : pb.indptr[r + 1]]
        assert np.all(np.diff(row) > 0)
        db[r, row] = True
    assert np.array_equal(db, np.kron(dense, np.ones((2, 2), bool)))
    sub = build_pattern(V.dofmap.list, V.dofmap.num_dofs, cells=[0, 0, 9])
    assert sub.nnz == 31  # cells (0,0) and (1,1) share one vertex


def test_normalise_qr_conventions():
    rng = np.random.default_rng(1)
    pts = [rng.random(2 * n) for n in (3, 1, 4)]
    w = [rng.random(n) for n in (3, 1, 4)]
    q = normalise_qr(([5, 5, 2], pts, w), 2)
    assert q.offsets.tolist() == [0, 3, 4, 8] and q.cells.dtype == np.int32 and q.n is None
    assert np.array_equal(q.pts, np.concatenate(pts))
    q2 = normalise_qr((np.arange(4), rng.random((4, 6)), rng.random((4, 3)), rng.random((4, 6))), 2)
    assert q2.offsets.tolist() == [0, 3, 6, 9, 12] and q2.n is not None
    with pytest.raises(AssertionError):
        normalise_qr(([0, 1], pts, w), 2)  # len(cells) != len(qr_pts)  (assemble_matrix.py:72)
    with pytest.raises(AssertionError):
        normalise_qr(([0, 1, 2], pts, w, w[:2]), 2)  # len(cells) != len(qr_n)  (:70)
    with pytest.raises(AssertionError):
        normalise_qr(([0], [np.zeros(3)], [np.zeros(1)]), 2)  # 3 coordinates for 1 weight in 2-D
    sel = q.select([2, 0])
    assert sel.cells.tolist() == [2, 5] and np.array_equal(sel.w, np.concatenate([w[2], w[0]]))


def test_gauss_rule_is_fiat_like():
    pts, w = gauss_tensor_rule(3, 2)
    assert pts.shape == (8, 3) and abs(w.sum() - 1) < 1e-15
    assert np.allclose(sorted(set(np.round(pts[:, 0], 12))), [0.5 - 0.5 / np.sqrt(3), 0.5 + 0.5 / np.sqrt(3)])


def test_synthetic_generator_is_seeded():
    a = synthetic_ragged(np.arange(10), 2, 16, 64, normals=True)
    b = synthetic_ragged(np.arange(10), 2, 16, 64, normals=True)
    assert np.array_equal(a.pts, b.pts) and np.array_equal(a.offsets, b.offsets)
    assert a.nq().min() >= 16 and a.nq().max() <= 64
    assert np.allclose(np.linalg.norm(a.n.reshape(-1, 2), axis=1), 1)


def test_forms_pack_constants():
    m = cq.create_rectangle([[0, 0], [1, 1]], [2, 2])
    V = cq.FunctionSpace(m, ("Lagrange", 1))
    f = forms.nitsche_rhs(V, cq.Analytic.sinprod(3.0), 10.0)
    assert f.consts().tolist() == [10.0, 3.0] and f.rank == 1
    assert forms.mass(V).consts().size == 0
    VV = cq.VectorFunctionSpace(m, ("Lagrange", 2))
    assert VV.bs == 2 and VV.num_dofs == 2 * 25
    d, nl = cq.utils.get_dofs(VV)
    assert nl == 18 and np.array_equal(d[:, 1], VV.dofmap.list[:, 0] * 2 + 1)  # utils.py:59-62


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = cq.create_rectangle([[0, 0], [1, 1]], [2, 2])
    V = cq.FunctionSpace(m, ("Lagrange", 1))
    with pytest.raises(RuntimeError):
        cq.assemble_matrix(forms.mass(V), [([0], [np.full(2, 0.5)], [np.ones(1)])])


def test_degree1_dofmap_is_the_geometry_dofmap_and_box_mesh_matches_index_grids():
    """The degree-1 shortcut of fem._build_dofmap and the broadcast construction of create_box (no index grids)."""
    import customquad_b200 as cq
    from customquad_b200.fem import reference_lattice

    for mesh in (cq.create_rectangle([[0, 0], [1, 2]], [3, 5]), cq.create_box([[0, 0, 0], [1, 2, 3]], [3, 4, 2])):
        V = cq.FunctionSpace(mesh, ("Lagrange", 1))
        assert np.array_equal(V.dofmap.list, mesh.geometry.dofmap)
        # independent statement: dof of local vertex v of cell (ix, iy[, iz]) by lattice position
        r = reference_lattice(mesh.cell_type, 1)
        if mesh.topology.dim == 2:
            nx, ny = mesh.shape
            for ix in range(nx):
                for iy in range(ny):
                    want = [(ix + r[v, 1]) * (ny + 1) + iy + r[v, 0] for v in range(4)]
                    assert list(V.dofmap.list[ix * ny + iy]) == want
        else:
            nx, ny, nz = mesh.shape
            for iz in range(nz):
                for iy in range(ny):
                    for ix in range(nx):
                        c = (iz * ny + iy) * nx + ix
                        want = [((iz + r[v, 2]) * (ny + 1) + iy + r[v, 1]) * (nx + 1) + ix + r[v, 0] for v in range(8)]
                        assert list(V.dofmap.list[c]) == want
                        x0 = mesh.geometry.x[mesh.geometry.dofmap[c, 0]]
                        assert np.allclose(x0, [ix / nx, 2 * iy / ny, 3 * iz / nz])


def test_row_chunked_torch_pattern_builder_on_cpu_tensors():
    import torch

    import customquad_b200 as cq
    from customquad_b200.sparsity import build_pattern, build_pattern_torch

    for mesh, degree, bs in ((cq.create_rectangle([[0, 0], [1, 1]], [6, 5]), 2, 2), (cq.create_box([[0, 0, 0], [1, 1, 1]], [3, 4, 2]), 1, 1),
                             (cq.create_box([[0, 0, 0], [1, 1, 1]], [2, 2, 3]), 2, 3)):
        V = cq.FunctionSpace(mesh, ("Lagrange", degree), bs)
        p = build_pattern(V.dofmap.list, V.dofmap.num_dofs, bs)
        for chunk in (50, 1000, 1 << 28):
            ip, ix = build_pattern_torch(torch.from_numpy(np.ascontiguousarray(V.dofmap.list)), V.dofmap.num_dofs, bs, chunk_pairs=chunk)
            assert np.array_equal(ip.numpy(), p.indptr) and np.array_equal(ix.numpy(), p.indices)


def test_slab_meshes_carry_the_global_coordinates_bit_for_bit():
    """dist.make_slab: a vertex has the same float coordinates on every rank that holds it (and as in the serial mesh).
    Generated from the slab's own corner they would differ by an ulp -- eps * |x| / h relative in the cell width, which
    the in-bench interface parity check saw as 2e-12 on the 8-GPU weak-scaling mesh."""
    import customquad_b200 as cq
    from customquad_b200 import dist as cqdist

    P, n = [[-1.11, -1.51], [20.17, 1.22]], [64, 8]
    g = cq.create_rectangle(P, n)
    for r in range(4):
        s = cqdist.make_slab(P, n, r, 4)
        lo = s.first - s.ghost
        assert np.array_equal(g.geometry.x[:, 0].reshape(65, 9)[lo : lo + s.mesh.shape[0] + 1].reshape(-1), s.mesh.geometry.x[:, 0])
    P3, n3 = [[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [4, 3, 16]
    g = cq.create_box(P3, n3)
    for r in range(3):
        s = cqdist.make_slab(P3, n3, r, 3)
        lo = s.first - s.ghost
        assert np.array_equal(g.geometry.x[:, 2].reshape(17, 4, 5)[lo : lo + s.mesh.shape[2] + 1].reshape(-1), s.mesh.geometry.x[:, 2])


def test_peer_zero_plan():
    """assemble_matrix(..., zero=True) on a matrix of a peer scatter: what is zeroed before the fence (customquad_b200/
    assemble.py; the reference zeroes the whole Mat and lets PETSc's stash exchange order the rest, assemble_matrix.py:20)."""
    from customquad_b200.assemble import peer_zero_plan

    assert peer_zero_plan((10, 50), True, True) == ("keep", True)      # owner side: interface rows now, the rest in the call
    assert peer_zero_plan(None, False, True) == (None, True)           # lowest rank: nobody writes into it
    assert peer_zero_plan(None, True, True) == ("all", False)          # a target without a known range: the old protocol
    assert peer_zero_plan((10, 50), True, False) == ("all", False)     # no integral to carry the zeroing
    assert peer_zero_plan(None, False, False) == ("all", False)
