"""Host-side logic of the Python mirror that needs no GPU: partition conversions (Partitions.jl:37-81), the reduction of
non-contiguous partitions to contiguous ones used by the objective evaluators (Costs.jl:34-39, 52-57), value marshalling for the
prefix matrices, method / model codes of the C ABI."""
import numpy as np
import pytest

import chainb200 as cp
from chainb200 import api, types as T
from helpers import sprand


def test_partition_conversions_round_trip():
    rng = np.random.default_rng(40)
    for _ in range(50):
        n, K = int(rng.integers(1, 40)), int(rng.integers(1, 8))
        P = cp.MapPartition(K, rng.integers(1, K + 1, n))
        D = cp.convert(cp.DomainPartition, P)
        assert D.spl[0] == 1 and D.spl[-1] == n + 1 and np.all(np.diff(D.spl) >= 0)
        assert sorted(D.prm.tolist()) == list(range(1, n + 1))
        for k in range(K):  # stable inside a part (Partitions.jl:56-66: counting sort)
            part = D.prm[D.spl[k] - 1 : D.spl[k + 1] - 1]
            assert np.all(np.diff(part) > 0) and np.all(P.asg[part - 1] == k + 1)
        assert cp.convert(cp.MapPartition, D) == P
        S = cp.SplitPartition(K, np.sort(np.concatenate([[1, n + 1], rng.integers(1, n + 2, K - 1)])))
        assert cp.convert(cp.MapPartition, cp.convert(cp.DomainPartition, S)) == cp.convert(cp.MapPartition, S)


def test_rows_by_part_matches_the_oracle_side(ref):
    """api._rows_by_part (device path) and pyoracle._contiguous (checker) rename the rows the same way, and the renamed
    partition is contiguous with the same part sizes."""
    rng = np.random.default_rng(41)
    for _ in range(30):
        m, n, K = int(rng.integers(1, 30)), int(rng.integers(1, 20)), int(rng.integers(1, 6))
        A = sprand(rng, m, n, 0.3)
        Pi = cp.MapPartition(K, rng.integers(1, K + 1, m))
        row_new, Pis = api._rows_by_part(Pi, m)
        assert sorted(row_new.tolist()) == list(range(1, m + 1))
        assert np.array_equal(np.diff(Pis.spl), np.bincount(Pi.asg - 1, minlength=K))
        for r in range(m):  # row r lands inside its part's range
            k = Pi.asg[r]
            assert Pis.spl[k - 1] <= row_new[r] < Pis.spl[k]
        B, _, Pis2 = ref._contiguous(A, cp.SplitPartition(1, [1, n + 1]), Pi)
        assert np.array_equal(Pis2.spl, Pis.spl)
        want = ref.permute(A, None, row_new)
        assert np.array_equal(B.colptr, want.colptr) and np.array_equal(B.rowval, want.rowval)


def test_prefix_value_marshalling():
    v, dt = api._as_i64_values(np.array([1, 2, 2**64 - 1], dtype=np.uint64))
    assert dt == np.uint64 and v.dtype == np.int64 and v.tolist() == [1, 2, -1]  # the same 64-bit words
    v, dt = api._as_i64_values(np.array([-3, 4], dtype=np.int32))
    assert dt == np.int64 and v.tolist() == [-3, 4]
    with pytest.raises(TypeError):
        api._as_i64_values(np.array([1.0, 2.0]))


def test_method_and_model_codes_match_the_header():
    import os
    import re

    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "chainb200.h")).read()
    enum = {k: int(v) for k, v in re.findall(r"(CPB_(?:MODEL|SPLIT|PACK)_[A-Z_0-9]+)\s*=\s*(-?\d+)", hdr)}
    f = cp.AffineConnectivityModel(0, 10, 1, 100)
    want = [(cp.DynamicBottleneckSplitter(f), "CPB_SPLIT_DYNAMIC_BOTTLENECK"), (cp.DynamicTotalSplitter(f), "CPB_SPLIT_DYNAMIC_TOTAL"),
            (cp.BisectCostBottleneckSplitter(f, 0.1), "CPB_SPLIT_BISECT_COST"), (cp.LazyBisectCostBottleneckSplitter(f, 0.1), "CPB_SPLIT_LAZY_BISECT_COST"),
            (cp.BisectIndexBottleneckSplitter(f), "CPB_SPLIT_BISECT_INDEX"), (cp.FlipBisectCostBottleneckSplitter(f, 0.1), "CPB_SPLIT_FLIP_BISECT_COST"),
            (cp.LazyFlipBisectCostBottleneckSplitter(f, 0.1), "CPB_SPLIT_LAZY_FLIP_BISECT_COST"), (cp.FlipBisectIndexBottleneckSplitter(f), "CPB_SPLIT_FLIP_BISECT_INDEX"),
            (cp.ConvexTotalSplitter(f), "CPB_SPLIT_CONVEX_TOTAL"), (cp.DynamicBottleneckChunker(f), "CPB_SPLIT_DYNAMIC_BOTTLENECK_CHUNKER"),
            (cp.DynamicTotalChunker(f), "CPB_SPLIT_DYNAMIC_TOTAL_CHUNKER"), (cp.EquiSplitter(), "CPB_SPLIT_EQUI")]
    for mtd, name in want:
        assert T.split_method_code(mtd)[0] == enum[name], name
    kinds = [(cp.AffineWorkModel(0, 1, 1), "CPB_MODEL_WORK"), (f, "CPB_MODEL_CONNECTIVITY"), (cp.AffinePrimaryConnectivityModel(0, 1, 1, 1, 1), "CPB_MODEL_PRIMCONN"),
             (cp.AffineSecondaryConnectivityModel(0, 1, 1, 1, 1), "CPB_MODEL_SECCONN"), (cp.AffinePrimaryEdgeCutModel(0, 1, 1, 1), "CPB_MODEL_PRIMEDGE"),
             (cp.AffineSecondaryEdgeCutModel(0, 1, 1, 1), "CPB_MODEL_SECEDGE"), (cp.AffineEnvelopeModel(0, 1, 1, 1), "CPB_MODEL_ENVELOPE")]
    for mdl, name in kinds:
        assert mdl.kind == enum[name], name
