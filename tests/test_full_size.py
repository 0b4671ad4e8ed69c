"""Bit-exact parity at the FULL BASELINE.json sizes: all five configurations, split vectors of the CUDA path
(through the C ABI, host arrays in, split vector out) compared with the CPU oracle on the same input.

CPU oracle wall clock on the GPU box's host (one thread): C1 ~55 s, C2 0.3 s, C3 ~9 s, C4a ~2 s, C4b ~19 s, C5 ~5 s.
The matrices come from the GPU generators (``synth_torch``: same hash, same seeds as ``synth``;
``test_torch_generators_match_numpy`` pins the equality).  Set CPB_SKIP_FULL_SIZE=1 to skip (debugging only).
"""
import gc
import os

import numpy as np
import pytest

import chainb200 as cp
from chainb200.workloads import WORKLOADS

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("CPB_SKIP_FULL_SIZE") == "1", reason="CPB_SKIP_FULL_SIZE=1")]


def _release():
    gc.collect()
    try:
        import torch

        torch.cuda.empty_cache()
    except Exception:
        pass
    cp.trim_memory()


def _check(ref, key, also_resident=True):
    w = WORKLOADS[key]
    M, extra = w.make(1.0, True)
    _release()
    got = w.call(cp, M, extra)                      # public call, HOST arrays (the e2e path)
    exp = w.call(ref, M, extra)                     # CPU oracle, same input
    assert w.same(got, exp), f"{key}: the device result differs from the CPU oracle at full size"
    if also_resident:
        dM = cp.device_matrix(M)
        try:
            assert w.same(w.call(cp, dM, extra), exp), f"{key}: resident-matrix call differs"
        finally:
            dM.close()
    del M
    _release()
    return got


def test_config1_full_size(ref):
    """C1: 5-point Laplacian 256 x 256 (n = 65,536), K = 8, DynamicBottleneckSplitter (DynamicSplitter.jl:15-50)."""
    r = _check(ref, "C1")
    assert r.spl[0] == 1 and r.spl[-1] == 65537


def test_config2_full_size(ref):
    """C2: Erdos-Renyi 1M x 1M, K = 64, BisectCost eps = 0.01 (BisectCostBottleneckSplitter.jl:6-63)."""
    _check(ref, "C2")


def test_config3_full_size(ref):
    """C3: R-MAT scale 24 (2.6e8 nonzeros), K = 1024, LazyBisect eps = 0.01 (LazyBisectCost...:140-258)."""
    r = _check(ref, "C3")
    assert r.K == 1024 and r.spl[-1] == (1 << 24) + 1


def test_config4a_full_size(ref):
    """C4a: banded 4M, pack_stripe DynamicTotalChunker(block model, 8) with EquiChunker(4) rows (DynamicChunker.jl:20-56,
    BlockCosts.jl:46-142)."""
    _check(ref, "C4a")


def test_config4b_full_size(ref):
    """C4b: banded 4M, pack_stripe ConvexTotalChunker(ConstrainedCost(connectivity, VertexCount, 8)) (ConvexTotalChunker.jl:141-265)."""
    _check(ref, "C4b")


def test_config5_full_size(ref):
    """C5: random-geometric 8M, K = 256, AlternatingPartitioner(LazyBisect(sym) x 2) (AlternatingPartitioner.jl:18-32,
    LazyBisectCost...:260-388)."""
    _check(ref, "C5")
