"""The ctypes mirrors in customquad_b200/_lib.py have the layout the C compiler gives the structs of
include/cq_b200.h (sizes and field offsets, checked with gcc: no GPU needed)."""
import ctypes
import os
import subprocess
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

STRUCTS = {
    "cq_form_desc": ("FormDesc", ["form", "cell_type", "degree", "bs", "fn_kind", "nconsts", "consts"]),
    "cq_mesh": ("MeshDesc", ["num_cells", "num_nodes", "geom_dofmap", "coords", "flags"]),
    "cq_space": ("SpaceDesc", ["num_dofs", "dofmap", "coeffs", "elem_pos"]),
    "cq_quadrature": ("QuadratureDesc", ["num_entries", "num_points", "cells", "offsets", "pts", "weights", "normals", "plane_entries", "plane_tiles", "pad_"]),
    "cq_csr": ("CsrDesc", ["num_rows", "nnz", "indptr", "indices", "values", "peer_row0", "peer_rowstart", "flags", "pad_", "zero_keep_lo", "zero_keep_hi"]),
    "cq_matrix_job": ("MatrixJob", ["form", "space", "A", "values_host"]),
    "cq_circle": ("CircleDesc", ["xmin", "h", "n", "center", "radius", "n1", "pad_", "g", "gw"]),
    "cq_sphere": ("SphereDesc", ["xmin", "h", "n", "center", "radius", "n1", "axis", "g", "gw", "gu", "gwu"]),
}


def test_struct_layouts_match_the_header():
    from customquad_b200 import _lib

    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "cq_b200.h"', "int main(void) {"]
    for cname, (_, fields) in STRUCTS.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf(" %zu", offsetof({cname}, {f}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "layout.c"), os.path.join(d, "layout")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        cname, size, *offs = line.split()
        pyname, fields = STRUCTS[cname]
        S = getattr(_lib, pyname)
        assert ctypes.sizeof(S) == int(size), cname
        assert [f for f, _ in S._fields_] == fields, cname
        assert [getattr(S, f).offset for f in fields] == [int(o) for o in offs], cname
