"""Runtime quadrature data: normalising the reference's ``qr_data`` tuples into ragged CSR form.

The reference accepts, per cell integral, a tuple ``(cells, qr_pts, qr_w)`` or
``(cells, qr_pts, qr_w, qr_n)`` (reference customquad/assemble_matrix.py:65-73) where
``cells`` is a list / int array (test/test_volume_area.py:115-116,136; test/common.py:49),
and ``qr_*`` are lists of flat float64 arrays, one per listed cell (demo/poisson.py:109-113),
or 2-D arrays with one row per cell (test/common.py:51-52).  Repeated cells and arbitrary
subsets are allowed (test_volume_area.py:121-136).  The device path wants one flat stream
per quantity plus a per-cell offset array -- that is ``RaggedQuadrature``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class RaggedQuadrature:
    cells: np.ndarray  # [C] int32 mesh cell ids
    offsets: np.ndarray  # [C+1] int64
    pts: np.ndarray  # [P*gdim] float64, point-major interleaved
    w: np.ndarray  # [P] float64
    n: np.ndarray | None  # [P*gdim] float64 or None
    gdim: int
    _dev: dict = field(default_factory=dict, repr=False)

    @property
    def num_entries(self) -> int:
        return int(self._dev["sizes"][0]) if self.cells is None else int(self.cells.shape[0])

    @property
    def num_points(self) -> int:
        return int(self._dev["sizes"][1]) if self.w is None else int(self.w.shape[0])

    def nq(self) -> np.ndarray:
        return np.diff(self.host().offsets)

    # ---- device-resident rules ---------------------------------------------------------------------------------
    # A rule is uploaded once (assemble.qr_on_device caches the device copy on this object) and then serves every
    # assemble_* call that integrates over it -- the reference's demo reuses its bulk rule for a matrix and a vector
    # (demo/poisson.py:193,209).  A rule can also be BORN on the device (a producer kernel, a synthetic generator):
    @classmethod
    def on_device(cls, cells, offsets, pts, w, n, gdim: int) -> "RaggedQuadrature":
        """Wrap torch CUDA tensors (cells int32 [C], offsets int64 [C+1], pts f64 [P*gdim], w f64 [P], n f64 [P*gdim]
        or None) as a rule without a host copy; ``host()`` makes one on demand."""
        q = cls(None, None, None, None, None, gdim)
        assert cells.is_cuda and cells.dtype.is_floating_point is False and offsets.shape[0] == cells.shape[0] + 1
        assert pts.shape[0] == gdim * w.shape[0] and (n is None or n.shape[0] == pts.shape[0])
        q._dev[cells.device] = (cells.contiguous(), offsets.contiguous(), pts.contiguous(), w.contiguous(), None if n is None else n.contiguous())
        q._dev["sizes"] = (int(cells.shape[0]), int(w.shape[0]))
        q._dev["has_normals"] = n is not None
        return q

    @property
    def has_normals(self) -> bool:
        return self._dev["has_normals"] if self.cells is None else self.n is not None

    def device_arrays(self):
        """(device, (cells, offsets, pts, w, n)) of a device-born rule."""
        for k, v in self._dev.items():
            if not isinstance(k, str):
                return k, v
        raise KeyError("no device copy")

    def host(self) -> "RaggedQuadrature":
        """This rule with numpy arrays (a device-born rule is copied back once)."""
        if self.cells is not None:
            return self
        if "host" not in self._dev:
            _, t = self.device_arrays()
            h = lambda a: None if a is None else a.cpu().numpy()  # noqa: E731
            self._dev["host"] = RaggedQuadrature(h(t[0]), h(t[1]), h(t[2]), h(t[3]), h(t[4]), self.gdim)
        return self._dev["host"]

    def select(self, idx) -> "RaggedQuadrature":
        """Sub-list (entries idx, in that order) -- used for cell partitioning."""
        if self.cells is None:
            return self.host().select(idx)
        idx = np.asarray(idx, dtype=np.int64)
        nq = self.nq()[idx]
        off = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(nq, out=off[1:])
        src = np.repeat(self.offsets[:-1][idx] - off[:-1], nq) + np.arange(off[-1])
        g = self.gdim
        pts = self.pts.reshape(-1, g)[src].reshape(-1)
        n = None if self.n is None else self.n.reshape(-1, g)[src].reshape(-1)
        return RaggedQuadrature(self.cells[idx].copy(), off, pts, self.w[src].copy(), n, g)


def _flatten(q, ncells: int, what: str, out=None, total_hint=None):
    """list-of-arrays or 2-D array -> (flat float64, per-entry lengths).  ``out(total)``: allocator of the flat
    array (e.g. pinned host memory), default numpy."""
    if isinstance(q, np.ndarray) and q.dtype != object:
        q = np.ascontiguousarray(q, dtype=np.float64)
        if q.ndim == 1:  # a single row for a single cell
            q = q.reshape(1, -1) if ncells == 1 else q.reshape(ncells, -1)
        assert q.ndim == 2
        assert q.shape[0] == ncells, f"len(cells) != len({what})"  # assemble_matrix.py:72-73
        return q.reshape(-1), np.full(ncells, q.shape[1], dtype=np.int64)
    assert len(q) == ncells, f"len(cells) != len({what})"
    if ncells == 0:
        return np.zeros(0), np.zeros(0, dtype=np.int64)
    fl = _flatten_helper()
    if fl is not None and isinstance(q, (list, tuple)):
        # one pass through the buffer protocol (csrc/cq_pyflatten.c): count, allocate, copy
        lens = np.empty(ncells, dtype=np.int64)
        try:
            if total_hint is not None:  # size known from the weights: one pass (a wrong size raises ValueError)
                flat = np.empty(total_hint, dtype=np.float64) if out is None else out(total_hint)
                try:
                    if fl.flatten(q, flat, lens) == total_hint:
                        return flat, lens
                except ValueError:
                    pass
            total = fl.flatten(q, None, lens)
            flat = np.empty(total, dtype=np.float64) if out is None else out(total)
            fl.flatten(q, flat, None)
            return flat, lens
        except TypeError:  # some item is not a contiguous float64 array: the general path converts
            pass
    lens = np.fromiter((np.size(a) for a in q), dtype=np.int64, count=ncells)
    flat = np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1) for a in q])
    return flat, lens


_FLATTEN = False


def _flatten_helper():
    """The in-tree CPython extension built by customquad_b200/build.py, or None when it has not been built (the
    numpy path gives the same arrays, seconds slower on millions of cells)."""
    global _FLATTEN
    if _FLATTEN is False:
        try:
            from . import _cq_flatten

            _FLATTEN = _cq_flatten
        except ImportError:
            _FLATTEN = None
    return _FLATTEN


def normalise_qr(qr, gdim: int, out=None) -> RaggedQuadrature:
    """One reference ``qr_data[i]`` tuple -> RaggedQuadrature.  Raises AssertionError on the
    same length mismatches the reference asserts (assemble_matrix.py:65-73).  ``out(total)`` allocates the flat float64
    arrays (pass a pinned-memory allocator to flatten straight into buffers the H2D copies can read)."""
    if isinstance(qr, RaggedQuadrature):
        assert qr.gdim == gdim
        return qr
    assert len(qr) in (3, 4), "qr tuple must be (cells, qr_pts, qr_w[, qr_n])"
    cells = np.asarray(qr[0])
    if cells.size == 0:
        cells = cells.astype(np.int32)
    assert cells.ndim == 1 and np.issubdtype(cells.dtype, np.integer), "cells must be a 1-D integer list/array"
    C = int(cells.shape[0])
    w, nw = _flatten(qr[2], C, "qr_w", out)
    pts, npts = _flatten(qr[1], C, "qr_pts", out, gdim * int(w.shape[0]))
    assert np.array_equal(npts, nw * gdim), "qr_pts must hold gdim coordinates per weight"
    n = None
    if len(qr) == 4:
        n, nn = _flatten(qr[3], C, "qr_n", out, gdim * int(w.shape[0]))
        assert np.array_equal(nn, nw * gdim), "qr_n must hold gdim components per weight"
    off = np.zeros(C + 1, dtype=np.int64)
    np.cumsum(nw, out=off[1:])
    return RaggedQuadrature(cells.astype(np.int32), off, pts, w, n, gdim)


# ---------------------------------------------------------------------------------------
def gauss_tensor_rule(tdim: int, degree: int):
    """The rule the reference's tests feed in: ``FIAT.create_quadrature(UFCQuadrilateral /
    UFCHexahedron, degree)`` (test/common.py:50) = tensor Gauss-Legendre on [0,1]^d with
    (degree+2)//2 points per direction, first coordinate slowest, weights summing to 1
    (SURVEY.md 8c (vi)).  Returns (pts [nq, tdim], w [nq])."""
    m = (degree + 2) // 2
    x, w = np.polynomial.legendre.leggauss(m)
    x, w = 0.5 * (x + 1.0), 0.5 * w
    grids = np.meshgrid(*([x] * tdim), indexing="ij")
    pts = np.stack([g.reshape(-1) for g in grids], axis=1)
    wg = np.meshgrid(*([w] * tdim), indexing="ij")
    wt = np.ones_like(wg[0])
    for g in wg:
        wt = wt * g
    return pts, wt.reshape(-1)


def synthetic_ragged(cells, gdim: int, nq_lo: int, nq_hi: int, seed: int = 20231, normals: bool = False) -> RaggedQuadrature:
    """Synthetic ragged rules of SURVEY.md 8d: nq_c ~ U{lo..hi}, points ~ U(0,1)^d,
    weights ~ U(0.5,1.5)/nq_c, normals = normalised N(0, I)."""
    rng = np.random.default_rng(seed)
    cells = np.asarray(cells, dtype=np.int32)
    C = cells.shape[0]
    nq = rng.integers(nq_lo, nq_hi + 1, size=C, dtype=np.int64)
    off = np.zeros(C + 1, dtype=np.int64)
    np.cumsum(nq, out=off[1:])
    P = int(off[-1])
    pts = rng.random(P * gdim)
    w = rng.uniform(0.5, 1.5, size=P) / np.repeat(nq, nq)
    n = None
    if normals:
        v = rng.standard_normal((P, gdim))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        n = v.reshape(-1)
    return RaggedQuadrature(cells, off, pts, w, n, gdim)
