"""Timing of the other BASELINE.json configs' kernel shapes (3-D Q1/Q2 hex, Nitsche with normals,
vector forms) at sizes that fit a quick run.  Prints one JSON line per case:
ms per launch, entries/s, algorithmic GB/s and fraction of the measured HBM peak."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import customquad_b200 as cq  # noqa: E402
from customquad_b200 import assemble as asm  # noqa: E402
from customquad_b200 import forms  # noqa: E402
from customquad_b200.quadrature import synthetic_ragged  # noqa: E402


def time_call(fn, reps=5):
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
    return float(np.median(ts))


def run(name, mesh, degree, bs, nq_lo, nq_hi, which, frac_cells=1.0):
    peak, _ = bench.measured_peak()
    tdim = mesh.geometry.dim
    V = cq.FunctionSpace(mesh, ("Lagrange", degree), bs)
    rng = np.random.default_rng(1)
    cells = np.arange(mesh.num_cells, dtype=np.int32)
    if frac_cells < 1.0:
        cells = np.sort(rng.choice(mesh.num_cells, int(mesh.num_cells * frac_cells), replace=False)).astype(np.int32)
    qr = synthetic_ragged(cells, tdim, nq_lo, nq_hi, seed=5, normals=True)
    dev = torch.device("cuda", 0)
    asm.qr_on_device(qr, dev)
    n = (degree + 1) ** tdim
    nv = 2**tdim
    for kind in which:
        if kind in ("mass", "stiffness", "nitsche"):
            form = forms.nitsche(V, 10.0) if kind == "nitsche" else getattr(forms, kind)(V)
            A = asm.create_matrix(form)
            fn = lambda: asm.assemble_matrix(form, [qr], A=A)  # noqa: E731
            fn()
            A.assemble()
            out_bytes = 8 * A.nnz
        elif kind == "load":
            form = forms.load(V, cq.Analytic.sinprod(*([3.0] * bs)))
            b = cq.la.Vector(torch.zeros(V.num_dofs, dtype=torch.float64, device=dev), 0)
            fn = lambda: asm.assemble_vector(form, [qr], b=b)  # noqa: E731
            out_bytes = 8 * V.num_dofs
        else:
            form = forms.integral(mesh, 1.0)
            fn = lambda: asm.assemble_scalar(form, [qr])  # noqa: E731
            out_bytes = 8
        ms = time_call(fn)
        normals = kind == "nitsche"
        ab = qr.num_entries * (4 + 8 + 4 * nv + 4 * n) + qr.num_points * 8 * (tdim + 1 + (tdim if normals else 0)) + 24 * mesh.geometry.x.shape[0] + out_bytes
        print(json.dumps({"case": name, "form": kind, "cells": int(qr.num_entries), "points": int(qr.num_points), "ms": round(ms, 4),
                          "entries_per_s": qr.num_entries / ms * 1e3, "alg_GBs": ab / ms / 1e6, "frac_hbm": ab / ms / 1e6 / peak}), flush=True)  # fmt: skip


if __name__ == "__main__":
    sel = sys.argv[1:] or ["hexq1", "hexq2", "quadnitsche"]
    if os.environ.get("CQ_VARIANT"):  # 1: the direct (un-staged) kernels
        asm.set_variant(int(os.environ["CQ_VARIANT"]))
    t0 = time.time()
    if "hexq1" in sel:
        mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [128, 128, 128])
        run("Q1 hex 128^3, 8 pts (cfg5 bulk)", mesh, 1, 1, 8, 8, ["stiffness", "mass", "load", "scalar"])
        run("Q1 hex 128^3, U{8..27} pts + normals (cfg3 cut)", mesh, 1, 1, 8, 27, ["stiffness", "nitsche", "scalar"])
    if "hexq2" in sel:
        mesh = cq.create_box([[-1.11, -1.51, -1.23], [1.55, 1.22, 1.11]], [48, 48, 48])
        run("Q2 hex 48^3 bs=3, 27 pts (cfg4)", mesh, 2, 3, 27, 27, ["stiffness", "mass", "load"])
        run("Q2 hex 48^3 bs=1, 27 pts", mesh, 2, 1, 27, 27, ["stiffness", "nitsche"])
    if "quadnitsche" in sel:
        mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [1024, 1024])
        run("Q1 quad 1024^2, U{4..16} pts + normals (surface rules)", mesh, 1, 1, 4, 16, ["nitsche", "load", "scalar"])
        run("Q1 quad vector 1024^2 bs=2, U{16..64}", mesh, 1, 2, 16, 64, ["stiffness", "load"])
    print("total s", time.time() - t0, file=sys.stderr)
