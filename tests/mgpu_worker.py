"""Run under torchrun (one rank per GPU): cell-partitioned CUDA assembly + NCCL interface exchange,
checked on every rank against the serial CPU oracle on the global mesh.  Exit code 0 = parity."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

import customquad_b200 as cq  # noqa: E402
from customquad_b200 import assemble as asm  # noqa: E402
from customquad_b200 import dist as cqdist  # noqa: E402
from customquad_b200 import forms  # noqa: E402
from customquad_b200.quadrature import synthetic_ragged  # noqa: E402


def close(a, ref, rtol=1e-12):
    return np.linalg.norm(a - ref) <= rtol * np.linalg.norm(ref) and np.max(np.abs(a - ref)) <= rtol * np.max(np.abs(ref))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import oracle

    for tdim, degree, bs, pts, n in [
        (2, 1, 1, [[-1.11, -1.51], [1.55, 1.22]], [64, 48]),
        (2, 2, 2, [[-1.11, -1.51], [1.55, 1.22]], [33, 20]),
        (3, 1, 1, [[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [6, 5, 17]),
    ]:
        axis = 0 if tdim == 2 else 2
        slab = cqdist.make_slab(pts, n, rank, world)
        V = cq.FunctionSpace(slab.mesh, ("Lagrange", degree), bs)
        gmesh = (cq.create_rectangle if tdim == 2 else cq.create_box)(pts, n)
        gq = synthetic_ragged(np.arange(gmesh.num_cells), tdim, 1, 40, seed=3, normals=True)
        per_layer = gmesh.num_cells // n[axis]
        q = gq.select(np.arange(slab.first * per_layer, (slab.first + slab.count) * per_layer))
        q.cells = (q.cells - (slab.first - slab.ghost) * per_layer).astype(np.int32)

        form = forms.stiffness(V)
        A = asm.create_matrix(form)
        ex = cqdist.SlabExchange(V, slab)
        ex.set_pattern(A.indptr, A.indices)
        cq.assemble_matrix(form, [q], A=A)
        ex.exchange_matrix(A.data)
        A.assemble()
        # the same matrix with the fused path: interface rows scattered into the owner's memory by the kernel
        A2 = asm.create_matrix(form)
        ex2 = cqdist.SlabExchange(V, slab)
        ex2.set_pattern(A2.indptr, A2.indices)
        fused = ex2.enable_peer_scatter(A2)
        assert fused, "peer memory could not be mapped"
        for _ in range(2):  # twice: zero -> fence -> assemble -> fence is the repeatable protocol
            A2.data.zero_()
            ex2.fence(A2.data.device)
            cq.assemble_matrix(form, [q], A=A2)
            ex2.exchange_matrix(A2.data)
        A2.assemble()
        o0, o1 = int(A.indptr[ex.owned_row0]), int(A.indptr[ex.owned_row1])
        a1, a2 = A.data[o0:o1].cpu().numpy(), A2.data[o0:o1].cpu().numpy()
        assert close(a2, a1), f"peer scatter differs from the NCCL exchange: {np.max(np.abs(a2 - a1))}"
        lf = forms.load(V, cq.Analytic.sinprod(*([2.0] * bs)))
        b = cq.assemble_vector(lf, [q])
        ex.exchange_vector(b.data)
        m = cqdist.SlabExchange.reduce_scalar(cq.assemble_scalar(forms.integral(slab.mesh, 1.0), [q]), device=torch.device("cuda", local))

        # serial reference (every rank computes it: small problems)
        GV = cq.FunctionSpace(gmesh, ("Lagrange", degree), bs)
        R = oracle.assemble_matrix(forms.stiffness(GV), [gq]).tocsr()
        rb = oracle.assemble_vector(forms.load(GV, cq.Analytic.sinprod(*([2.0] * bs))), [gq])
        rm = oracle.assemble_scalar(forms.integral(gmesh, 1.0), [gq])
        S = A.to_scipy()
        shift = (slab.first - slab.ghost) * degree * ex.line
        r0, r1 = ex.owned_row0, ex.owned_row1
        mine = S[r0:r1].tocoo()
        ref = R[r0 + shift : r1 + shift].tocoo()
        # compare as dense-by-key: (row, global col) -> value; the local pattern may hold extra zero slots
        loc = {(int(i), int(j) + shift): v for i, j, v in zip(mine.row, mine.col, mine.data)}
        refd = {(int(i), int(j)): v for i, j, v in zip(ref.row, ref.col, ref.data)}
        assert set(refd) <= set(loc), "owner row misses a coupling"
        keys = sorted(refd)
        assert close(np.array([loc[k] for k in keys]), np.array([refd[k] for k in keys])), (tdim, degree, bs)
        assert all(loc[k] == 0.0 for k in set(loc) - set(refd))
        assert close(b.array[r0:r1], rb[r0 + shift : r1 + shift])
        assert abs(m - rm) <= 1e-12 * abs(rm)
    dist.barrier()
    if rank == 0:
        print(f"mgpu parity ok on {world} ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
