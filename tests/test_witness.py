"""The two independent restatements of the reference's hot solvers -- oracle/cpo_*.hpp (C++, the checker of the GPU
tests) and oracle/pywitness.py (a literal Python transliteration with set-based counts) -- must return the same split
vectors on random small inputs, ties included.  A disagreement means one of them misreads the reference."""
import numpy as np
import pytest

import chainb200 as cp
import pywitness as wit


def random_matrix(rng, n_max=28, m_max=20):
    n = int(rng.integers(1, n_max + 1))
    m = int(rng.integers(1, m_max + 1))
    dens = rng.choice([0.05, 0.15, 0.4, 0.9])
    cols, colptr = [], [1]
    for _ in range(n):
        rows = np.zeros(0, dtype=np.int64) if rng.random() < 0.2 else np.flatnonzero(rng.random(m) < dens) + 1
        cols.append(rows)
        colptr.append(colptr[-1] + len(rows))
    rowval = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    return cp.SparseMatrixCSC(m, n, np.array(colptr, dtype=np.int64), rowval.astype(np.int64))


def random_model(rng):
    if rng.random() < 0.5:  # small integers: many exact ties
        coef = tuple(int(x) for x in rng.integers(0, 4, 4))
    else:
        coef = tuple(float(x) for x in rng.choice([0.0, 0.5, 1.0, 2.5, 10.0], 4))
    return coef, cp.AffineConnectivityModel(*coef)


@pytest.mark.parametrize("seed", range(6))
def test_splitters_agree(ref, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(60):
        A = random_matrix(rng)
        coef, mdl = random_model(rng)
        f = wit.Conn(A, coef)
        K = int(rng.integers(1, 7))
        eps = float(rng.choice([0.5, 0.1, 0.01]))
        assert wit.dynamic_splitter(f, K, False) == ref.partition_stripe(A, K, cp.DynamicBottleneckSplitter(mdl)).spl.tolist(), (A.colptr, A.rowval, coef, K)
        assert wit.dynamic_splitter(f, K, True) == ref.partition_stripe(A, K, cp.DynamicTotalSplitter(mdl)).spl.tolist(), (A.colptr, A.rowval, coef, K)
        assert wit.bisect_cost(f, K, eps) == ref.partition_stripe(A, K, cp.BisectCostBottleneckSplitter(mdl, eps)).spl.tolist(), (A.colptr, A.rowval, coef, K, eps)
        assert wit.lazy_bisect_connectivity(f, K, eps) == ref.partition_stripe(A, K, cp.LazyBisectCostBottleneckSplitter(mdl, eps)).spl.tolist(), (A.colptr, A.rowval, coef, K, eps)


@pytest.mark.parametrize("seed", range(6))
def test_chunkers_agree(ref, seed):
    rng = np.random.default_rng(2000 + seed)
    for _ in range(60):
        A = random_matrix(rng)
        coef, mdl = random_model(rng)
        f = wit.Conn(A, coef)
        w_max = int(rng.integers(1, 7))
        got = wit.dynamic_total_chunker(f, w_max)
        exp = ref.pack_stripe(A, cp.DynamicTotalChunker(cp.ConstrainedCost(mdl, cp.VertexCount(), w_max)))
        assert got == exp.spl.tolist(), ("dynamic", A.colptr, A.rowval, coef, w_max)
        # the convex chunker presumes the quadrangle inequality: connectivity-type costs with beta >= 0 obey it
        got = wit.convex_total_chunker_constrained(f, w_max)
        exp = ref.pack_stripe(A, cp.ConvexTotalChunker(cp.ConstrainedCost(mdl, cp.VertexCount(), w_max)))
        assert got == exp.spl.tolist(), ("convex", A.colptr, A.rowval, coef, w_max)
        rho = float(rng.choice([0.0, 0.3, 0.9, 1.0]))
        nn = []
        exp = ref.pack_stripe(A, cp.OverlapChunker(rho, w_max), n_nets=nn)
        spl, nets = wit.overlap_chunker(A, rho, w_max)
        assert spl == exp.spl.tolist() and nets == nn[0].tolist(), ("overlap", A.colptr, A.rowval, rho, w_max)


@pytest.mark.parametrize("seed", range(4))
def test_concave_forms_agree(ref, seed):
    """a23: the queue routine chunk_concave! (ConcaveTotalChunker.jl:57-114) and its two unconstrained callers, transliterated
    independently in Python, against the C++ oracle that checks the device -- on connectivity costs (not concave: the result is
    then whatever the queue leaves, which is exactly what has to be reproduced), ties included."""
    rng = np.random.default_rng(3000 + seed)
    for _ in range(60):
        A = random_matrix(rng)
        coef, mdl = random_model(rng)
        f = wit.Conn(A, coef)
        got = wit.concave_total_chunker(f)
        exp = ref.pack_stripe(A, cp.ConcaveTotalChunker(mdl))
        assert got == exp.spl.tolist(), ("concave chunker", A.colptr, A.rowval, coef)
        K = int(rng.integers(1, 7))
        got = wit.concave_total_splitter(f, K)
        exp = ref.partition_stripe(A, K, cp.ConcaveTotalSplitter(mdl))
        assert got == exp.spl.tolist(), ("concave splitter", A.colptr, A.rowval, coef, K)
