#!/usr/bin/env python
"""Vector Poisson with a VectorFunctionSpace: the reference's demo/vector_poisson.py on this package, and the
BASELINE.json configs[3] workload ("demo/vector_poisson.py-style 3D vector Poisson on Q2 hex 128^3").

1. ``python examples/vector_poisson.py -N 32``  -- the demo itself (demo/vector_poisson.py is 2-D only): CutFEM on the
   unit circle, VectorFunctionSpace Q1 (bs = 2, :95), u_ex = (sin(pi x) sin(pi y), 0) (:99-106),
   a_bulk = inner(nabla_grad u, nabla_grad v) (:121), the blocked symmetric Nitsche terms with quadrature normals
   (:123-126), ghost penalty (:127, per component), runtime-quadrature assembly on the cut cells (:160-181), inactive-dof
   locking (:191-198), direct solve (scipy instead of MUMPS), L2 / H10 errors through assemble_scalar (:226-238).
2. ``python examples/vector_poisson.py -config4 N`` -- the 3-D Q2 version at N^3 hexahedra with the 3^3 Gauss rule in
   every cell: blocked stiffness matrix (FP64 tensor-core kernel, cq_dmma.cuh), mass matrix and load vector; prints
   timings and the properties the full-size test checks (no solve: 5e7 unknowns at N = 128).

Needs a B200; prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import customquad_b200 as cq  # noqa: E402
from customquad_b200 import cutfem, forms  # noqa: E402
from customquad_b200.levelset import generate_qr_circle  # noqa: E402


def solve(N, betaN=10.0, betas=1.0, order=4):
    import scipy.sparse.linalg as spla

    xmin, xmax = np.array([-1.11, -1.51]), np.array([1.55, 1.22])  # demo/vector_poisson.py:38-39
    mesh = cq.create_rectangle([xmin, xmax], [N, N])
    cut_cells, uncut_cells, outside_cells, qr_pts, qr_w, qr_pts_bdry, qr_w_bdry, qr_n = generate_qr_circle(mesh, 1.0, (0.0, 0.0), order)
    V = cq.VectorFunctionSpace(mesh, ("Lagrange", 1))  # :95
    u_ex = cq.Analytic.sinprod(1.0, 0.0)  # (sin(pi x) sin(pi y), 0)  (:99-106)
    f = cq.Analytic.sinprod(2 * np.pi**2, 0.0)  # (:107-110)
    g = cq.Function(V).interpolate(lambda x: np.stack([np.sin(np.pi * x[0]) * np.sin(np.pi * x[1]), 0.0 * x[0]]))  # g = u_ufl (:112)

    qr_uncut = cutfem.standard_rule(mesh, uncut_cells, 2)
    qr_bulk = [(cut_cells, qr_pts, qr_w)]  # :157
    qr_bdry = [(cut_cells, qr_pts_bdry, qr_w_bdry, qr_n)]  # :158

    from customquad_b200 import assemble as asm

    facets = cutfem.ghost_penalty_facets(mesh, cut_cells, outside_cells)
    A = asm.create_matrix(forms.stiffness(V), facets=facets)  # cell couplings + couplings across the ghost-penalty facets
    cq.assemble_matrix(forms.stiffness(V), [qr_uncut], A=A)  # Ax (:139-144): uncut bulk ...
    cutfem.ghost_penalty(V, facets, betas, A)  # ... + a_stab (:127), per component, on the device
    cq.assemble_matrix(forms.stiffness(V), qr_bulk, A=A)  # Ac1 (:160-165)
    cq.assemble_matrix(forms.nitsche(V, betaN), qr_bdry, A=A)  # Ac2 (:167-170)
    A.assemble()
    b = cq.assemble_vector(forms.load(V, f), [cutfem.standard_rule(mesh, uncut_cells, 4)])
    b += cq.assemble_vector(forms.load(V, f), qr_bulk)  # :176
    b += cq.assemble_vector(forms.nitsche_rhs(V, g, betaN), qr_bdry)  # :181
    assert np.isfinite(b.array).all() and np.isfinite(A.norm())  # :185-186

    inactive = cutfem.get_inactive_dofs(V, cut_cells, uncut_cells)  # :191
    A = cutfem.lock_inactive_dofs(inactive, A)
    uh = cq.Function(V, spla.spsolve(A.to_scipy().tocsc(), b.array))
    assert np.isfinite(uh.array).all()

    def assemble(make_form):  # :226-231
        return cq.assemble_scalar(make_form(), qr_bulk) + cq.assemble_scalar(make_form(), [cutfem.standard_rule(mesh, uncut_cells, 6)])

    L2 = np.sqrt(assemble(lambda: forms.l2_error(uh, u_ex)))
    H10 = np.sqrt(assemble(lambda: forms.h10_error(uh, u_ex)))
    return {"demo": "vector_poisson", "N": N, "h": float(np.max(mesh.h())), "cut_cells": int(len(cut_cells)), "dofs": int(V.num_dofs),
            "L2_err": float(L2), "H10_err": float(H10)}  # fmt: skip


def config4(N):
    import torch

    from customquad_b200 import assemble as asm
    from customquad_b200.quadrature import RaggedQuadrature, gauss_tensor_rule

    dev = torch.device("cuda", torch.cuda.current_device())
    t0 = time.time()
    mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [N, N, N])
    V = cq.VectorFunctionSpace(mesh, ("Lagrange", 2))
    pts, w = gauss_tensor_rule(3, 4)  # 3^3 points: exact for the Q2 mass matrix
    C, nq = mesh.num_cells, len(w)
    qr = RaggedQuadrature.on_device(torch.arange(C, dtype=torch.int32, device=dev), torch.arange(C + 1, dtype=torch.int64, device=dev) * nq,
                                    torch.from_numpy(pts.reshape(-1)).to(dev).repeat(C), torch.from_numpy(w).to(dev).repeat(C), None, 3)  # fmt: skip
    fK = forms.stiffness(V)
    K = asm.create_matrix(fK)
    torch.cuda.synchronize()
    t_setup = time.time() - t0

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    ms_zero = timed(lambda: K.data.zero_())
    ms_K = timed(lambda: asm.assemble_matrix(fK, [qr], A=K))  # adds to what is there
    ms_Kz = timed(lambda: asm.assemble_matrix(fK, [qr], A=K, zero=True))  # A.zeroEntries() + assembly in one call: what assemble_matrix(form, qr) does
    K.data.zero_()
    asm.assemble_matrix(fK, [qr], A=K)
    K.assemble()
    rows = torch.repeat_interleave(torch.arange(K.shape[0], device=dev), K.indptr[1:] - K.indptr[:-1]) if K.nnz < 2**31 else None
    rowsum = None
    if rows is not None:
        rs = torch.zeros(K.shape[0], dtype=torch.float64, device=dev).index_add_(0, rows, K.data)
        rowsum = float(rs.abs().max() / K.data.abs().max())  # rigid translations are in the kernel
    del rows
    chunks = [K.data[i : i + (1 << 28)] for i in range(0, K.nnz, 1 << 28)]  # no full-size temporaries (77.5 GB of values at N = 128)
    total_rel = abs(float(K.data.sum())) / sum(float(c.abs().sum()) for c in chunks)  # 1^T K 1 = 0
    ms_M = timed(lambda: asm.assemble_matrix(forms.mass(V), [qr], A=K))
    K.data.zero_()
    asm.assemble_matrix(forms.mass(V), [qr], A=K)
    vol = float(np.prod(np.asarray(mesh.xmax) - np.asarray(mesh.xmin)))
    mass_err = abs(float(K.data.sum()) - 3 * vol) / (3 * vol)
    b = cq.assemble_vector(forms.load(V, cq.Analytic.sinprod(2 * np.pi**2, 0.0, 0.0)), [qr])
    alg = C * (4 + 8 + 4 * 8 + 4 * 27) + C * nq * 32 + 24 * mesh.geometry.x.shape[0] + 8 * K.nnz
    return {"config": f"vector Poisson, Q2 hexahedra {N}^3, bs = 3, {nq} points per cell (BASELINE.json configs[3])", "cells": C, "dofs": int(V.num_dofs),
            "nnz": int(K.nnz), "csr_GB": K.nnz * 12 / 1e9, "setup_s": t_setup, "zero_ms": ms_zero, "stiffness_ms": ms_K, "stiffness_zero_fused_ms": ms_Kz, "mass_ms": ms_M,
            "algorithmic_GB": alg / 1e9, "stiffness_frac_hbm_6450": alg / ms_K / 1e6 / 6450.9, "stiffness_zero_fused_frac_hbm_6450": alg / ms_Kz / 1e6 / 6450.9, "stiffness_rowsum_rel": rowsum, "stiffness_total_sum_rel": total_rel,
            "mass_total_rel_err": mass_err, "load_norm": float(b.data.norm())}  # fmt: skip


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-N", type=int, default=32)  # demo/vector_poisson.py:13
    ap.add_argument("-betaN", type=float, default=10.0)
    ap.add_argument("-betas", type=float, default=1.0)
    ap.add_argument("-config4", type=int, default=0, help="N: time the 3-D Q2 blocked assembly on N^3 hexahedra instead of running the demo")
    a = ap.parse_args()
    print(json.dumps(config4(a.config4) if a.config4 else solve(a.N, a.betaN, a.betas)))
