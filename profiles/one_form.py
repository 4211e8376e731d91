"""Development: time one (mesh shape, degree, form) case repeatedly (for ncu captures and quick sweeps).
usage: python profiles/one_form.py quad|hex N degree form nq_lo nq_hi [reps]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import customquad_b200 as cq  # noqa: E402
from customquad_b200 import assemble as asm  # noqa: E402
from customquad_b200 import forms  # noqa: E402
from customquad_b200.quadrature import synthetic_ragged  # noqa: E402

shape, N, degree, kind, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
tdim = 2 if shape == "quad" else 3
mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [N, N]) if tdim == 2 else cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [N, N, N])
BS = int(os.environ.get("CQ_BS", "1"))  # block size of the space
ZERO = os.environ.get("CQ_ZERO", "0") == "1"  # matrices: A.zeroEntries() as part of the call
if "CQ_FUSED" in os.environ:  # in-kernel zeroing: 1 on, 0 off (default: the library's policy)
    asm.set_fused_zero(os.environ["CQ_FUSED"] == "1")
V = cq.FunctionSpace(mesh, ("Lagrange", degree), BS) if BS > 1 else cq.FunctionSpace(mesh, ("Lagrange", degree))
qr = synthetic_ragged(np.arange(mesh.num_cells, dtype=np.int32), tdim, lo, hi, seed=5, normals=True)
dev = torch.device("cuda", 0)
asm.qr_on_device(qr, dev)
if kind in ("mass", "stiffness", "nitsche"):
    form = forms.nitsche(V, 10.0) if kind == "nitsche" else getattr(forms, kind)(V)
    A = asm.create_matrix(form)
    fn = lambda: asm.assemble_matrix(form, [qr], A=A, zero=ZERO)  # noqa: E731
    out = 8 * A.nnz
elif kind in ("load", "load_coef", "nitsche_rhs"):
    if kind == "load":
        form = forms.load(V, cq.Analytic.sinprod(3.0))
    else:
        f = cq.Function(V)
        f.array[:] = np.random.default_rng(1).uniform(-1, 1, V.num_dofs)
        form = forms.load(V, f) if kind == "load_coef" else forms.nitsche_rhs(V, f, 10.0)
    b = cq.la.Vector(torch.zeros(V.num_dofs, dtype=torch.float64, device=dev), 0)
    fn = lambda: asm.assemble_vector(form, [qr], b=b)  # noqa: E731
    out = 8 * V.num_dofs
else:
    form = forms.integral(mesh, 1.0)
    fn = lambda: asm.assemble_scalar(form, [qr])  # noqa: E731
    out = 8
for _ in range(2):
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
ms = float(np.median(ts))
n, nv = (degree + 1) ** tdim, 2**tdim
normals = kind in ("nitsche", "nitsche_rhs")
ab = qr.num_entries * (4 + 8 + 4 * nv + 4 * n) + qr.num_points * 8 * (tdim + 1 + (tdim if normals else 0)) + 24 * mesh.geometry.x.shape[0] + out
peak, _ = bench.measured_peak()
print(json.dumps({"case": f"{shape} {N} Q{degree} {kind} U{{{lo}..{hi}}}", "ms": round(ms, 4), "cells": int(qr.num_entries), "points": int(qr.num_points),
                  "alg_GBs": round(ab / ms / 1e6, 1), "frac_hbm": round(ab / ms / 1e6 / peak, 3)}))
