"""The five BASELINE.json configurations as data: generator, public call, comparison.

Shared by ``bench.py`` (measurement), ``tests/test_full_size.py`` (bit-exact parity at BASELINE sizes) and
``bench_configs.py``.  ``call(impl, M, extra)`` takes the implementation module as an argument -- the product
(``chainb200``) or, in the tests and the CPU-baseline legs only, the oracle binding -- so nothing in this package
imports the oracle.  Parameters follow SURVEY.md section 8(d).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import numpy as np

from . import types as T

AFF = T.AffineConnectivityModel(0, 10, 1, 100)  # the reference's own net_model (test/runbenchmarks.jl:17)


@dataclass
class Workload:
    key: str
    title: str
    make: Callable          # (scale, device) -> host SparseMatrixCSC the solve runs on + extra inputs
    call: Callable          # (impl, M, extra) -> result; M = host matrix or (product only) a DeviceMatrix
    same: Callable          # (a, b) -> bool, bit-exact comparison of two results
    describe: Callable      # result -> small dict for the bench line

    def sizes(self, A):
        return {"m": int(A.m), "n": int(A.n), "nnz": int(A.nnz)}


def _same_split(a, b):
    return a.K == b.K and np.array_equal(a.spl, b.spl)


def _same_pair(a, b):
    return _same_split(a[0], b[0]) and _same_split(a[1], b[1])


def _gen(device):
    if device:
        from . import synth_torch as g
    else:
        from . import synth as g
    return g


def _c1(scale, device):
    from . import synth

    g = max(8, int(round(256 * np.sqrt(scale))))
    return synth.laplacian5(g), None


def _c2(scale, device):
    n = max(1000, int(1_000_000 * scale))
    return _gen(device).erdos_renyi(n, 10), None


def _c3(scale, device):
    s = 24 if scale >= 1 else max(10, int(round(24 + np.log2(scale))))
    return _gen(device).rmat(s, 16 << s), None


def _c4(scale, device):
    from . import api

    n = max(1024, int((1 << 22) * scale))
    A = _gen(device).banded(n, 64)
    X = api.adjointpattern(A)  # the solve runs on X with the row partition Pi of A's columns (runbenchmarks.jl:40-45)
    Pi = T.SplitPartition(*_equi_chunks(A.n, 4))
    return X, Pi


def _equi_chunks(n, w):
    """pack_stripe(A, EquiChunker(w)) (EquiPartitioner.jl:15-21): chunks of w columns, closed form."""
    K = -(-n // w)
    spl = np.minimum(1 + w * np.arange(K + 1, dtype=np.int64), n + 1)
    return K, spl


def _c5(scale, device):
    n = max(1024, int((1 << 23) * scale))
    return _gen(device).random_geometric(n), None


def _k3(A):
    return min(1024, max(2, A.n // 64))


def _k5(A):
    return min(256, max(2, A.n // 64))


_BLK = T.BlockComponentCostModel(int, 1, 3, (1, T.identity), (1, T.identity))
_SYM = T.AffineMonotonizedSymmetricConnectivityModel(0, 0, 1, 100, 90)

WORKLOADS = {
    "C1": Workload("C1", "configs[0]: 5-point Laplacian 256x256, K=8, DynamicBottleneckSplitter(AffineConnectivityModel(0,10,1,100))", _c1,
                   lambda impl, M, extra: impl.partition_stripe(M, 8, T.DynamicBottleneckSplitter(AFF)), _same_split,
                   lambda r: {"K": int(r.K)}),
    "C2": Workload("C2", "configs[1]: Erdos-Renyi 1Mx1M, 10 nnz/col, K=64, BisectCostBottleneckSplitter(AffineConnectivityModel(0,10,1,100), 0.01)", _c2,
                   lambda impl, M, extra: impl.partition_stripe(M, 64, T.BisectCostBottleneckSplitter(AFF, 0.01)), _same_split,
                   lambda r: {"K": int(r.K)}),
    "C3": Workload("C3", "configs[2]: R-MAT scale 24 (16M x 16M, ~2.6e8 nnz), K=1024, LazyBisectCostBottleneckSplitter(AffineConnectivityModel(0,10,1,100), 0.01)", _c3,
                   lambda impl, M, extra: impl.partition_stripe(M, _k3(M), T.LazyBisectCostBottleneckSplitter(AFF, 0.01)), _same_split,
                   lambda r: {"K": int(r.K)}),
    "C4a": Workload("C4a", "configs[3]: banded 4M, bandwidth 64, pack_stripe(X, DynamicTotalChunker(BlockComponentCostModel{Int}(1,3,(1,identity),(1,identity)), 8), EquiChunker(4) rows)", _c4,
                    lambda impl, M, extra: impl.pack_stripe(M, T.DynamicTotalChunker(_BLK, 8), extra), _same_split,
                    lambda r: {"chunks": int(r.K)}),
    "C4b": Workload("C4b", "configs[3]: banded 4M, bandwidth 64, pack_stripe(X, ConvexTotalChunker(ConstrainedCost(AffineConnectivityModel(0,0,0,1), VertexCount(), 8)))", _c4,
                    lambda impl, M, extra: impl.pack_stripe(M, T.ConvexTotalChunker(T.ConstrainedCost(T.AffineConnectivityModel(0, 0, 0, 1), T.VertexCount(), 8))), _same_split,
                    lambda r: {"chunks": int(r.K)}),
    "C5": Workload("C5", "configs[4]: random-geometric graph 8M vertices, K=256, partition_plaid(AlternatingPartitioner(LazyBisect(MonotonizedSymmetric(0,0,1,100,90), 0.1) x2))", _c5,
                   lambda impl, M, extra: impl.partition_plaid(M, _k5(M), T.AlternatingPartitioner(T.LazyBisectCostBottleneckSplitter(_SYM, 0.1), T.LazyBisectCostBottleneckSplitter(_SYM, 0.1))),
                   _same_pair, lambda r: {"K": int(r[0].K)}),
}
