"""Multi-process (gloo, world_size 2 and 3) test of the slab partition + interface-row exchange
(customquad_b200/dist.py) on CPU.  The local assembler here is the oracle (this is a
test of the host-side partition / ownership / exchange logic, which is device independent); the
receive-side adds are numpy doubles injected in place of the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import customquad_b200 as cq
from customquad_b200 import dist as cqdist
from customquad_b200 import forms
from customquad_b200.quadrature import synthetic_ragged
from customquad_b200.sparsity import build_pattern


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _np_accumulate_rows(values, indptr, row0, nrows, recv_indptr, recv):
    v, ip, rp, r = values.numpy(), np.asarray(indptr), recv_indptr.numpy(), recv.numpy()
    for k in range(nrows):
        n = rp[k + 1] - rp[k]
        v[ip[row0 + k] : ip[row0 + k] + n] += r[rp[k] : rp[k + 1]]


def _np_axpy(n, x, y):
    y += x


def _global_problem(tdim, degree, bs):
    if tdim == 2:
        pts, n = [[-1.11, -1.51], [1.55, 1.22]], [7, 4]
    else:
        pts, n = [[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [3, 2, 7]
    return pts, n


def _worker(rank, world, port, tdim, degree, bs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle

        pts, n = _global_problem(tdim, degree, bs)
        slab = cqdist.make_slab(pts, n, rank, world)
        V = cq.FunctionSpace(slab.mesh, ("Lagrange", degree), bs)
        # every rank draws the rule of the GLOBAL cell list and keeps its own cells
        gmesh = (cq.create_rectangle if tdim == 2 else cq.create_box)(pts, n)
        gq = synthetic_ragged(np.arange(gmesh.num_cells), tdim, 1, 9, seed=3, normals=True)
        per_layer = gmesh.num_cells // n[0 if tdim == 2 else 2]
        mine = np.arange(slab.first * per_layer, (slab.first + slab.count) * per_layer)
        q = gq.select(mine)
        q.cells = (q.cells - (slab.first - slab.ghost) * per_layer).astype(np.int32)  # global -> local cell ids
        assert np.array_equal(np.sort(q.cells), slab.cells)

        pat = build_pattern(V.dofmap.list, V.dofmap.num_dofs, bs)
        ex = cqdist.SlabExchange(V, slab, accumulate_rows=_np_accumulate_rows, axpy=_np_axpy)
        ex.set_pattern(pat.indptr, pat.indices)
        A = oracle.assemble_matrix(forms.stiffness(V), [q], pattern=pat)
        vals = torch.from_numpy(A.data.copy())
        ex.exchange_matrix(vals)
        lf = forms.load(V, cq.Analytic.sinprod(*([2.0] * bs)))
        b = torch.from_numpy(oracle.assemble_vector(lf, [q]))
        ex.exchange_vector(b)
        m = oracle.assemble_scalar(forms.integral(slab.mesh, 1.0), [q])
        m = cqdist.SlabExchange.reduce_scalar(m)
        np.savez(os.path.join(out, f"r{rank}.npz"), vals=vals.numpy(), b=b.numpy(), m=m, indptr=pat.indptr, indices=pat.indices,
                 owned=np.array([ex.owned_row0, ex.owned_row1]), line=ex.line, ghost=slab.ghost, first=slab.first)  # fmt: skip
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,tdim,degree,bs", [(2, 2, 1, 1), (2, 2, 2, 2), (3, 2, 1, 1), (2, 3, 1, 1), (2, 3, 2, 3)])
def test_slab_exchange_reproduces_serial_assembly(tmp_path, world, tdim, degree, bs):
    from oracle import oracle

    port = _free_port()
    mp.spawn(_worker, args=(world, port, tdim, degree, bs, str(tmp_path)), nprocs=world, join=True)
    # serial reference on the global mesh
    pts, n = _global_problem(tdim, degree, bs)
    gmesh = (cq.create_rectangle if tdim == 2 else cq.create_box)(pts, n)
    V = cq.FunctionSpace(gmesh, ("Lagrange", degree), bs)
    gq = synthetic_ragged(np.arange(gmesh.num_cells), tdim, 1, 9, seed=3, normals=True)
    A = oracle.assemble_matrix(forms.stiffness(V), [gq]).toarray()
    b = oracle.assemble_vector(forms.load(V, cq.Analytic.sinprod(*([2.0] * bs))), [gq])
    m = oracle.assemble_scalar(forms.integral(gmesh, 1.0), [gq])
    k = degree
    covered = 0
    for r in range(world):
        d = np.load(tmp_path / f"r{r}.npz")
        line = int(d["line"])
        r0, r1 = d["owned"]
        # local row R sits `(first - ghost) * k` lattice lines above the global origin
        shift = (int(d["first"]) - int(d["ghost"])) * k * line
        ip, ix, vals = d["indptr"], d["indices"], d["vals"]
        for R in range(r0, r1):
            cols = ix[ip[R] : ip[R + 1]].astype(np.int64) + shift
            ref = A[R + shift]
            assert np.allclose(vals[ip[R] : ip[R + 1]], ref[cols], rtol=1e-12, atol=1e-13 * np.abs(A).max())
            mask = np.ones(A.shape[1], bool)
            mask[cols] = False
            assert np.all(ref[mask] == 0.0)  # the owner's row holds every coupling of the global row
        assert np.allclose(d["b"][r0:r1], b[r0 + shift : r1 + shift], rtol=1e-12, atol=1e-13 * np.abs(b).max())
        assert abs(float(d["m"]) - m) <= 1e-12 * abs(m)
        covered += r1 - r0
    assert covered == A.shape[0]  # the owned row ranges tile the global rows exactly


def test_partition_balances_work():
    w = np.ones(64)
    w[:8] = 20.0  # a band of heavily cut layers
    b = cqdist.partition_layers(64, 4, w)
    loads = [w[b[i] : b[i + 1]].sum() for i in range(4)]
    assert b[0] == 0 and b[-1] == 64 and np.all(np.diff(b) >= 1)
    assert max(loads) <= 1.6 * (w.sum() / 4)
    assert np.array_equal(cqdist.partition_layers(10, 3), [0, 3, 6, 10])
