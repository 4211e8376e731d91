"""Cut-cell quadrature producer for the circle on the bench-size mesh: host numpy (levelset.generate_qr_circle) vs
the GPU kernels (generate_qr_circle_device)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import customquad_b200 as cq  # noqa: E402
from customquad_b200 import levelset  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
mesh = cq.create_rectangle([[-1.11, -1.51], [1.55, 1.22]], [N, N])
levelset.generate_qr_circle_device(mesh, 1.0, (0.0, 0.0), 4)  # warm-up (context, allocator)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    cut, uncut, outside, qb, qs = levelset.generate_qr_circle_device(mesh, 1.0, (0.0, 0.0), 4)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
t0 = time.perf_counter()
h = levelset.generate_qr_circle(mesh, 1.0, (0.0, 0.0), 4)
th = time.perf_counter() - t0
print(json.dumps({"mesh": f"{N}x{N}", "cut_cells": int(len(cut)), "bulk_points": int(qb.num_points), "surface_points": int(qs.num_points),
                  "device_ms_incl_host_copies": round(1e3 * min(ts), 3), "host_numpy_ms": round(1e3 * th, 1)}))
