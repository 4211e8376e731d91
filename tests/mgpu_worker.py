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
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle import oracle

    checked = 0
    for tdim, degree, bs, pts, n in [
        (2, 1, 1, [[-1.11, -1.51], [1.55, 1.22]], [64, 48]),
        (2, 2, 2, [[-1.11, -1.51], [1.55, 1.22]], [33, 20]),
        (3, 1, 1, [[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [6, 5, 17]),
        (3, 1, 3, [[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [4, 5, 9]),   # blocked hexahedra (vector_poisson-style)
        (3, 2, 1, [[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [3, 4, 8]),   # Q2 hexahedra: tensor-core kernel
    ]:
        axis = 0 if tdim == 2 else 2
        # work-balanced boundaries from a synthetic weight per layer, the same on every rank
        wts = 1.0 + np.arange(n[axis]) % 3
        bounds = cqdist.partition_layers(n[axis], world, weights=wts)
        slab = cqdist.make_slab(pts, n, rank, world, bounds=bounds)
        V = cq.FunctionSpace(slab.mesh, ("Lagrange", degree), bs)
        gmesh = (cq.create_rectangle if tdim == 2 else cq.create_box)(pts, n)
        gq = synthetic_ragged(np.arange(gmesh.num_cells), tdim, 1, 40, seed=3, normals=True)
        per_layer = gmesh.num_cells // n[axis]
        q = gq.select(np.arange(slab.first * per_layer, (slab.first + slab.count) * per_layer))
        q.cells = (q.cells - (slab.first - slab.ghost) * per_layer).astype(np.int32)
        GV = cq.FunctionSpace(gmesh, ("Lagrange", degree), bs)

        for kind in ("stiffness", "mass", "nitsche"):
            mk = (lambda W: forms.nitsche(W, 10.0)) if kind == "nitsche" else getattr(forms, kind)
            form = mk(V)
            # (1) separate exchange: ncclSend / ncclRecv + accumulate kernel inside the library (cq_exchange_offrank)
            A = asm.create_matrix(form)
            ex = cqdist.SlabExchange(V, slab)
            ex.set_pattern(A.indptr, A.indices)
            cq.assemble_matrix(form, [q], A=A)
            ex.exchange_matrix(A.data)
            A.assemble()
            # (2) fused: interface rows scattered into the owner's memory by the kernel, fences through peer-memory flags
            A2 = asm.create_matrix(form)
            ex2 = cqdist.SlabExchange(V, slab)
            ex2.set_pattern(A2.indptr, A2.indices)
            assert ex2.enable_peer_scatter(A2), "peer memory could not be mapped"
            for _ in range(2):  # twice: zero -> fence -> assemble -> fence is the repeatable protocol (zero=True follows it)
                cq.assemble_matrix(form, [q], A=A2, zero=True)
                ex2.exchange_matrix(A2.data)
            A2.assemble()
            o0, o1 = int(A.indptr[ex.owned_row0]), int(A.indptr[ex.owned_row1])
            a1, a2 = A.data[o0:o1].cpu().numpy(), A2.data[o0:o1].cpu().numpy()
            assert close(a2, a1), f"peer scatter differs from the NCCL exchange ({kind}): {np.max(np.abs(a2 - a1))}"
            # serial reference (every rank computes it: small problems), on the oracle's own pattern
            R = oracle.assemble_matrix(mk(GV), [gq]).tocsr()
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
            assert close(np.array([loc[k] for k in keys]), np.array([refd[k] for k in keys])), (tdim, degree, bs, kind)
            assert all(loc[k] == 0.0 for k in set(loc) - set(refd))
            checked += 1

        f = cq.Function(V)
        gf = cq.Function(GV)
        gf.array[:] = np.random.default_rng(11).uniform(-1, 1, GV.num_dofs)
        shift = (slab.first - slab.ghost) * degree * ex.line
        f.array[:] = gf.array[shift : shift + V.num_dofs]  # the slab's dofs are a contiguous range of the global numbering
        for lf, glf in ((forms.load(V, cq.Analytic.sinprod(*([2.0] * bs))), forms.load(GV, cq.Analytic.sinprod(*([2.0] * bs)))),
                        (forms.nitsche_rhs(V, f, 10.0), forms.nitsche_rhs(GV, gf, 10.0))):  # fmt: skip
            b = cq.assemble_vector(lf, [q])
            ex.exchange_vector(b.data)
            rb = oracle.assemble_vector(glf, [gq])
            assert close(b.array[r0:r1], rb[r0 + shift : r1 + shift]), (tdim, degree, bs, lf.kind)
        m = cqdist.SlabExchange.reduce_scalar(cq.assemble_scalar(forms.integral(slab.mesh, 1.0), [q]), device=dev)
        rm = oracle.assemble_scalar(forms.integral(gmesh, 1.0), [gq])
        assert abs(m - rm) <= 1e-12 * abs(rm)
    dist.barrier()
    if rank == 0:
        print(f"mgpu parity ok on {world} ranks ({checked} matrix cases x (library NCCL exchange, fused peer scatter) + vectors + scalars; "
              f"library comm {cqdist._comm_ready}, peer fence {cqdist._peer_fence_ready})")
    cqdist.close_peer_mappings()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
