"""One rank's share of config 5 at N = 2 (Q1 hex 512 x 512 x 256, 6.7e7 cells, 8-point rules) on ONE GPU: per-launch times of the
pieces of the step (memset, stiffness, mass), with and without the traversal-tile hint.  usage: python profiles/config5_single.py [nz]"""
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
from customquad_b200.quadrature import RaggedQuadrature  # noqa: E402

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(os.environ.get("C5_N", "512"))
dev = torch.device("cuda", 0)
mesh = cq.create_box(bench.BOX3, [N, N, nz])
V = cq.FunctionSpace(mesh, ("Lagrange", 1))
C = mesh.num_cells
g = torch.Generator(device=dev).manual_seed(1)
off = torch.arange(C + 1, dtype=torch.int64, device=dev) * 8
q = RaggedQuadrature.on_device(torch.arange(C, dtype=torch.int32, device=dev), off, torch.rand(3 * 8 * C, dtype=torch.float64, device=dev, generator=g),
                               torch.rand(8 * C, dtype=torch.float64, device=dev, generator=g) / 8, None, 3)
peak, _ = bench.measured_peak()


def t(fn, reps=3):
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


A = asm.create_matrix(forms.stiffness(V))
ab = bench.algorithmic_bytes(C, 8 * C, 3, 8, 8, mesh.geometry.x.shape[0], A.nnz)
print(json.dumps({"cells": C, "nnz": A.nnz, "algorithmic_GB": ab / 1e9, "memset_ms": t(lambda: A.data.zero_())}), flush=True)
for tile in (24 << 20, 48 << 20, 12 << 20, 1 << 60):
    asm.L2_TILE_BYTES = tile
    for kind in ("stiffness", "mass"):
        form = getattr(forms, kind)(V)
        c = asm._Call(form, q, dev, A=A)
        ms = t(lambda: asm.assemble_matrix(form, [q], A=A))
        print(json.dumps({"tile_MB": tile >> 20, "tiles": int(c.qr.plane_tiles), "form": kind, "ms": ms, "frac_hbm": ab / ms / 1e6 / peak, "cells_per_s": C / ms * 1e3}), flush=True)
