"""Deterministic synthetic inputs for the five BASELINE.json configs.

All generators draw from the counter-based hash ``splitmix64`` with ``SEED = 0xDEADBEEF`` so the
same matrix is produced anywhere (numpy only, no RNG state).  Matrices are patterns; rows are sorted
and de-duplicated within each column; indices are 1-based Int64 (Julia's SparseMatrixCSC layout).
"""
from __future__ import annotations

import numpy as np

from .types import SparseMatrixCSC

SEED = np.uint64(0xDEADBEEF)
_U = np.uint64


def splitmix64(x):
    """uint64 -> uint64, vectorised (wraparound arithmetic)."""
    with np.errstate(over="ignore"):
        x = np.asarray(x, dtype=np.uint64) + _U(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> _U(30))) * _U(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U(27))) * _U(0x94D049BB133111EB)
        return z ^ (z >> _U(31))


def _unit(x):
    return splitmix64(x).astype(np.float64) / 18446744073709551616.0


def _from_pairs(m: int, n: int, rows0: np.ndarray, cols0: np.ndarray) -> SparseMatrixCSC:
    """0-based (row, col) pairs -> sorted, de-duplicated 1-based CSC."""
    key = cols0.astype(np.int64) * np.int64(m) + rows0.astype(np.int64)
    key = np.unique(key)
    cols = key // m
    rows = key - cols * m
    counts = np.bincount(cols, minlength=n)
    colptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int64) + 1
    return SparseMatrixCSC(m, n, colptr, rows.astype(np.int64) + 1)


def laplacian5(g: int = 256) -> SparseMatrixCSC:
    """C1: 5-point stencil on a g x g grid; vertex v = y*g + x."""
    v = np.arange(g * g, dtype=np.int64)
    x, y = v % g, v // g
    rows = [v]
    cols = [v]
    for ok, off in ((x > 0, -1), (x < g - 1, 1), (y > 0, -g), (y < g - 1, g)):
        rows.append(v[ok] + off)
        cols.append(v[ok])
    return _from_pairs(g * g, g * g, np.concatenate(rows), np.concatenate(cols))


def erdos_renyi(n: int = 1_000_000, d: int = 10, m: int | None = None) -> SparseMatrixCSC:
    """C2: column j holds rows ``h(SEED ^ (j*16 + t)) mod m`` for t = 0..d-1."""
    m = n if m is None else m
    j = np.repeat(np.arange(n, dtype=np.uint64), d)
    t = np.tile(np.arange(d, dtype=np.uint64), n)
    rows = splitmix64(SEED ^ (j * _U(16) + t)) % _U(m)
    return _from_pairs(m, n, rows.astype(np.int64), j.astype(np.int64))


def rmat(scale: int = 24, edges: int | None = None, abcd=(0.57, 0.19, 0.19, 0.05), chunk: int = 1 << 24) -> SparseMatrixCSC:
    """C3: R-MAT, 2^scale vertices, ``edges`` draws (default 16 per vertex), quadrant of edge e at
    level l from ``u(SEED ^ (e*32 + l))``; duplicates removed."""
    n = 1 << scale
    edges = 16 * n if edges is None else edges
    a, b, c, _ = abcd
    keys = []
    for e0 in range(0, edges, chunk):
        e = np.arange(e0, min(edges, e0 + chunk), dtype=np.uint64)
        r = np.zeros(len(e), dtype=np.int64)
        cc = np.zeros(len(e), dtype=np.int64)
        for l in range(scale):
            u = _unit(SEED ^ (e * _U(32) + _U(l)))
            right = ((u >= a) & (u < a + b)) | (u >= a + b + c)
            down = u >= a + b
            r = (r << 1) | down
            cc = (cc << 1) | right
        keys.append(np.unique(cc * np.int64(n) + r))
    key = np.unique(np.concatenate(keys))
    cols = key // n
    rows = key - cols * n
    counts = np.bincount(cols, minlength=n)
    colptr = np.concatenate(([0], np.cumsum(counts))).astype(np.int64) + 1
    return SparseMatrixCSC(n, n, colptr, rows + 1)


def banded(n: int = 1 << 22, bw: int = 64, keep_mod: int = 4) -> SparseMatrixCSC:
    """C4: entry (i,j), |i-j| <= bw, present iff i == j or ``h(SEED ^ (i*2^22 + j)) mod 4 == 0``."""
    rows, cols = [], []
    j = np.arange(n, dtype=np.int64)
    for off in range(-bw, bw + 1):
        i = j + off
        ok = (i >= 0) & (i < n)
        ii, jj = i[ok], j[ok]
        if off != 0:
            h = splitmix64(SEED ^ (ii.astype(np.uint64) * _U(1 << 22) + jj.astype(np.uint64)))
            keep = (h % _U(keep_mod)) == 0
            ii, jj = ii[keep], jj[keep]
        rows.append(ii)
        cols.append(jj)
    return _from_pairs(n, n, np.concatenate(rows), np.concatenate(cols))


def random_geometric(n: int = 1 << 23, mean_degree: float = 8.0) -> SparseMatrixCSC:
    """C5: n points in the unit square, radius sqrt(deg/(pi n)); vertices relabelled by
    (cell_y, cell_x, v); symmetric pattern without self-loops."""
    v = np.arange(n, dtype=np.uint64)
    px = _unit(SEED ^ (_U(2) * v))
    py = _unit(SEED ^ (_U(2) * v + _U(1)))
    r = np.sqrt(mean_degree / (np.pi * n))
    g = max(1, int(np.floor(1.0 / r)))
    cx = np.minimum((px * g).astype(np.int64), g - 1)
    cy = np.minimum((py * g).astype(np.int64), g - 1)
    order = np.lexsort((np.arange(n), cx, cy))
    px, py, cx, cy = px[order], py[order], cx[order], cy[order]
    cell = cy * g + cx
    start = np.searchsorted(cell, np.arange(g * g + 1))
    rows, cols = [], []
    idx = np.arange(n, dtype=np.int64)
    # half the neighbourhood (dx, dy) in {(0,0),(1,0),(-1,1),(0,1),(1,1)}; pairs mirrored afterwards
    cnt = start[1:] - start[:-1]
    maxc = int(cnt.max()) if n else 0
    for dx, dy in ((0, 0), (1, 0), (-1, 1), (0, 1), (1, 1)):
        nx, ny = cx + dx, cy + dy
        ok = (nx >= 0) & (nx < g) & (ny < g)
        ncell = np.where(ok, ny * g + nx, 0)
        s, c = start[ncell], np.where(ok, cnt[ncell], 0)
        for t in range(maxc):
            has = t < c
            a = idx[has]
            b = s[has] + t
            if dx == 0 and dy == 0:
                keep = b > a
                a, b = a[keep], b[keep]
            d2 = (px[a] - px[b]) ** 2 + (py[a] - py[b]) ** 2
            near = d2 <= r * r
            rows.append(a[near])
            cols.append(b[near])
    a = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
    b = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    return _from_pairs(n, n, np.concatenate([a, b]), np.concatenate([b, a]))


def sprand_pattern(m: int, n: int, density: float, seed: int = 0) -> SparseMatrixCSC:
    """Parity-scale stand-in for the reference tests' ``sprand(m, n, p)``: entry (i,j) present iff
    ``u(h(SEED ^ seed) ^ (i*n + j)) < p``."""
    s = splitmix64(SEED ^ _U(seed))
    i, j = np.divmod(np.arange(m * n, dtype=np.uint64), _U(max(n, 1)))
    keep = _unit(s ^ (i * _U(max(n, 1)) + j)) < density
    return _from_pairs(m, n, i[keep].astype(np.int64), j[keep].astype(np.int64))
