"""Builds libcq_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the tree)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcq_b200.so")
SOURCES = ["cq_api.cu", "cq_comm.cu", "cq_launch_matrix.cu", "cq_launch_vector.cu", "cq_launch_scalar.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--expt-relaxed-constexpr",
]  # fmt: skip


def _newest_header() -> float:
    t = os.path.getmtime(os.path.join(HERE, "..", "include", "cq_b200.h"))
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h")):
            t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    return t


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    hdr_t = _newest_header()
    objs, jobs = [], []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(sp), hdr_t):
            jobs.append([nvcc, *NVCC_FLAGS, *extra_flags, "-c", sp, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed: {cmd[-3]}")

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart", "-ldl"])
    build_flatten(force)
    return LIB


FLATTEN_SO = os.path.join(HERE, "_cq_flatten.so")


def build_flatten(force: bool = False) -> str:
    """The CPython helper that flattens the reference's lists of per-cell arrays (csrc/cq_pyflatten.c), gcc, in-tree."""
    import sysconfig

    import numpy

    src = os.path.join(CSRC, "cq_pyflatten.c")
    if force or not os.path.exists(FLATTEN_SO) or os.path.getmtime(FLATTEN_SO) < os.path.getmtime(src):
        tmp = f"{FLATTEN_SO}.{os.getpid()}.tmp"
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-Wall", f"-I{sysconfig.get_paths()['include']}", f"-I{numpy.get_include()}", src, "-o", tmp], check=True)
        os.replace(tmp, FLATTEN_SO)
    return FLATTEN_SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=["-Xptxas", "-v"] if "--ptxas" in sys.argv else ()))
