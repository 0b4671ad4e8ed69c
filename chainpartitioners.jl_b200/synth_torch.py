"""GPU versions of the synthetic generators in ``synth.py`` (harness code: torch is used for device
memory and sorting only).  They produce the SAME matrices as the numpy generators -- same
``splitmix64`` counter hash, same seeds -- but in seconds at the full BASELINE sizes (R-MAT scale 24,
banded 2^22, random-geometric 2^23).  ``tests/test_gpu_parity.py::test_torch_generators_match_numpy``
checks the equality at small sizes.
"""
from __future__ import annotations

import numpy as np
import torch

from .types import SparseMatrixCSC

SEED = 0xDEADBEEF
_M64 = (1 << 64) - 1


def _i64(x: int) -> int:
    """uint64 constant -> the int64 with the same bits"""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    """logical right shift of int64 bit patterns"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    """int64 bit patterns in -> int64 bit patterns out (wraparound arithmetic == uint64 arithmetic)"""
    z = x + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def _unit(h: torch.Tensor) -> torch.Tensor:
    """uint64 -> float64 in [0, 1) exactly as numpy's ``h.astype(float64) / 2**64``"""
    hi = _lsr(h, 1).to(torch.float64) * 2.0 + (h & 1).to(torch.float64)  # round-to-nearest-even like numpy's uint64 -> f64 cast
    return hi / 18446744073709551616.0


def _umod(h: torch.Tensor, m: int) -> torch.Tensor:
    """uint64(h) mod m for int64 bit patterns (m < 2^31)"""
    hi = _lsr(h, 32)
    lo = h & 0xFFFFFFFF
    return ((hi % m) * ((1 << 32) % m) + lo % m) % m


def _from_keys(m: int, n: int, key: torch.Tensor) -> SparseMatrixCSC:
    """key = col * m + row (0-based) -> sorted, de-duplicated 1-based CSC on the host"""
    key = torch.unique(key)
    cols = torch.div(key, m, rounding_mode="floor")
    rows = key - cols * m
    counts = torch.bincount(cols, minlength=n)
    colptr = torch.zeros(n + 1, dtype=torch.int64, device=key.device)
    colptr[1:] = torch.cumsum(counts, 0)
    return SparseMatrixCSC(m, n, (colptr + 1).cpu().numpy(), (rows + 1).cpu().numpy())


def erdos_renyi(n: int = 1_000_000, d: int = 10, device="cuda") -> SparseMatrixCSC:
    j = torch.arange(n, dtype=torch.int64, device=device).repeat_interleave(d)
    t = torch.arange(d, dtype=torch.int64, device=device).repeat(n)
    rows = _umod(splitmix64(_i64(SEED) ^ (j * 16 + t)), n)
    return _from_keys(n, n, j * n + rows)


def rmat(scale: int = 24, edges: int | None = None, abcd=(0.57, 0.19, 0.19, 0.05), chunk: int = 1 << 25, device="cuda") -> SparseMatrixCSC:
    n = 1 << scale
    edges = 16 * n if edges is None else edges
    a, b, c, _ = abcd
    keys = []
    for e0 in range(0, edges, chunk):
        e = torch.arange(e0, min(edges, e0 + chunk), dtype=torch.int64, device=device)
        r = torch.zeros_like(e)
        cc = torch.zeros_like(e)
        for l in range(scale):
            u = _unit(splitmix64(_i64(SEED) ^ (e * 32 + l)))
            right = ((u >= a) & (u < a + b)) | (u >= a + b + c)
            down = u >= a + b
            r = (r << 1) | down.to(torch.int64)
            cc = (cc << 1) | right.to(torch.int64)
        keys.append(torch.unique(cc * n + r))
    return _from_keys(n, n, torch.cat(keys))


def banded(n: int = 1 << 22, bw: int = 64, keep_mod: int = 4, device="cuda") -> SparseMatrixCSC:
    j = torch.arange(n, dtype=torch.int64, device=device)
    keys = []
    for off in range(-bw, bw + 1):
        i = j + off
        ok = (i >= 0) & (i < n)
        ii, jj = i[ok], j[ok]
        if off != 0:
            h = splitmix64(_i64(SEED) ^ (ii * (1 << 22) + jj))
            keep = (h & (keep_mod - 1)) == 0 if keep_mod & (keep_mod - 1) == 0 else _umod(h, keep_mod) == 0
            ii, jj = ii[keep], jj[keep]
        keys.append(jj * n + ii)
    return _from_keys(n, n, torch.cat(keys))


def random_geometric(n: int = 1 << 23, mean_degree: float = 8.0, device="cuda") -> SparseMatrixCSC:
    v = torch.arange(n, dtype=torch.int64, device=device)
    px = _unit(splitmix64(_i64(SEED) ^ (2 * v)))
    py = _unit(splitmix64(_i64(SEED) ^ (2 * v + 1)))
    r = float(np.sqrt(mean_degree / (np.pi * n)))
    g = max(1, int(np.floor(1.0 / r)))
    cx = torch.clamp((px * g).to(torch.int64), max=g - 1)
    cy = torch.clamp((py * g).to(torch.int64), max=g - 1)
    order = torch.argsort((cy * g + cx) * n + v)  # == lexsort((v, cx, cy))
    px, py, cx, cy = px[order], py[order], cx[order], cy[order]
    cell = cy * g + cx
    start = torch.searchsorted(cell, torch.arange(g * g + 1, dtype=torch.int64, device=device))
    cnt = start[1:] - start[:-1]
    maxc = int(cnt.max()) if n else 0
    idx = torch.arange(n, dtype=torch.int64, device=device)
    keys = []
    for dx, dy in ((0, 0), (1, 0), (-1, 1), (0, 1), (1, 1)):
        nx, ny = cx + dx, cy + dy
        ok = (nx >= 0) & (nx < g) & (ny < g)
        ncell = torch.where(ok, ny * g + nx, torch.zeros_like(nx))
        s = start[ncell]
        c = torch.where(ok, cnt[ncell], torch.zeros_like(nx))
        for t in range(maxc):
            has = t < c
            a = idx[has]
            b = s[has] + t
            if dx == 0 and dy == 0:
                keep = b > a
                a, b = a[keep], b[keep]
            d2 = (px[a] - px[b]) ** 2 + (py[a] - py[b]) ** 2
            near = d2 <= r * r
            a, b = a[near], b[near]
            keys.append(a * n + b)
            keys.append(b * n + a)
    key = torch.cat(keys) if keys else torch.zeros(0, dtype=torch.int64, device=device)
    return _from_keys(n, n, key)
