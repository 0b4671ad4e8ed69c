"""Pins the CPU oracle's count structures against the reference's own brute-force test definitions
(test/test_SparsePrefixMatrices.jl:14-92, test/test_SparseColorArrays.jl:1-92,
test/test_EnvelopeMatrices.jl:1-19, test/test_util.jl:1-11)."""
import numpy as np
import pytest

import chainb200 as cp
from helpers import (ref_dianetcount, ref_dominancecount, ref_envelope, ref_netcount, ref_selfnetcount,
                     ref_selfpincount, sprand)

HINTS = [cp.NoHint(), cp.RandomHint(), cp.SparseHint(), cp.StepHint()]
DIMS = [1, 2, 3, 7, 8, 9]  # test_SparsePrefixMatrices.jl:24


def probes(rng, trials, *ranges):
    """corner points + random points (test_SparsePrefixMatrices.jl:17-20)"""
    ends = [(r[0], r[-1]) for r in ranges]
    pts = [tuple(e[t] for e, t in zip(ends, idx)) for idx in np.ndindex(*(2,) * len(ends))]
    for _ in range(trials):
        pts.append(tuple(int(rng.integers(e[0], e[1] + 1)) for e in ends))
    return pts


def test_adjointpattern(ref):
    rng = np.random.default_rng(1)
    for m in range(1, 30):
        for n in range(1, 30, 3):
            A = sprand(rng, m, n, 0.5)
            B = ref.adjointpattern(A)
            S = A.to_scipy().T.tocsc()
            S.sort_indices()
            assert np.array_equal(B.colptr, S.indptr + 1) and np.array_equal(B.rowval, S.indices + 1)


@pytest.mark.parametrize("kw", [dict(), dict(H=1), dict(H=2), dict(H=3), dict(H=4), dict(b=1), dict(b=2), dict(b=3), dict(b=4)])
def test_dominancecount_jump(ref, kw):
    rng = np.random.default_rng(2)
    hints = [cp.SparseHint()] if kw else HINTS
    for hint in hints:
        for m in DIMS:
            for n in DIMS:
                A = sprand(rng, m, n, 0.5)
                pts = probes(rng, 10, (1, m + 1), (1, n + 1))
                i = [p[0] for p in pts]
                j = [p[1] for p in pts]
                got = ref.dominancecount(A, i, j, hint=hint, **kw)
                exp = [ref_dominancecount(A, a, b) for a, b in pts]
                assert got.tolist() == exp, (hint, m, n, kw)


def ref_prefix(rows, cols, vals, i, j):
    """test_SparsePrefixMatrices.jl:14-15 on coordinate lists: count and (wrap-around) sum of the points below (i, j)."""
    sel = (rows <= i - 1) & (cols <= j - 1)
    return int(sel.sum()), int(vals[sel].sum(dtype=np.uint64))


@pytest.mark.parametrize("kw", [dict(), dict(H=1), dict(H=2), dict(H=3), dict(H=4), dict(b=1), dict(b=2), dict(b=3), dict(b=4)])
def test_dominancesum_reference_structure(ref, kw):
    """The restated DominanceSum (SparsePrefixMatrices.jl:1-254: b-ary tree, permuted value copies, cached digit counts and
    cumulative sums) for the reference's own parameter grid (test_SparsePrefixMatrices.jl:27-41) with UInt values: every probed
    entry equals sum(A[1:i-1, 1:j-1]) with wrap-around, and the independent offline sweep."""
    rng = np.random.default_rng(32)
    for m in DIMS + [15, 16, 17, 31, 32, 33, 63, 64, 65]:
        for n in DIMS:
            A = sprand(rng, m, n, 0.5)
            val = rng.integers(0, 2**64, A.nnz, dtype=np.uint64)
            cols = np.repeat(np.arange(1, n + 1), np.diff(A.colptr))
            pts = probes(rng, 10, (1, m + 1), (1, n + 1))
            i, j = [p[0] for p in pts], [p[1] for p in pts]
            got = ref.dominancesum(A, val, i, j, **kw)
            assert got.tolist() == [ref_prefix(A.rowval, cols, val, a, b)[1] for a, b in pts], (m, n, kw)
            assert np.array_equal(got, ref.prefix_query(m, n, A.nnz, A.colptr, A.rowval, val, i, j))


def test_dominancesum_and_rook_structures(ref):
    """dominancesum, rookcount!, rooksum! (test_SparsePrefixMatrices.jl:43-69): UInt values with wrap-around, a random
    permutation with odd values; every entry equals the definition.  The sweep also reproduces dominancecount."""
    rng = np.random.default_rng(31)
    for m in DIMS + [31, 32, 33]:
        for n in DIMS:
            A = sprand(rng, m, n, 0.5)
            val = rng.integers(0, 2**64, A.nnz, dtype=np.uint64)
            cols = np.repeat(np.arange(1, n + 1), np.diff(A.colptr))
            pts = probes(rng, 10, (1, m + 1), (1, n + 1))
            i, j = [p[0] for p in pts], [p[1] for p in pts]
            exp = [ref_prefix(A.rowval, cols, val, a, b) for a, b in pts]
            assert ref.prefix_query(m, n, A.nnz, A.colptr, A.rowval, None, i, j).tolist() == [e[0] for e in exp]
            assert ref.prefix_query(m, n, A.nnz, A.colptr, A.rowval, None, i, j).tolist() == ref.dominancecount(A, i, j).tolist()
            assert ref.prefix_query(m, n, A.nnz, A.colptr, A.rowval, val, i, j).tolist() == [e[1] for e in exp]
        N = m
        idx = rng.permutation(N) + 1
        val = (rng.integers(0, 2**63, N, dtype=np.uint64) << np.uint64(1)) + np.uint64(1)
        pts = probes(rng, 100, (1, N + 1), (1, N + 1))
        i, j = [p[0] for p in pts], [p[1] for p in pts]
        exp = [ref_prefix(idx, np.arange(1, N + 1), val, a, b) for a, b in pts]
        assert ref.prefix_query(N, N, N, None, idx, None, i, j).tolist() == [e[0] for e in exp]
        assert ref.prefix_query(N, N, N, None, idx, val, i, j).tolist() == [e[1] for e in exp]


def test_dominancecount_larger(ref):
    rng = np.random.default_rng(3)
    for (m, n, p) in [(100, 80, 0.1), (300, 500, 0.02), (1000, 1000, 0.005)]:
        A = sprand(rng, m, n, p)
        i = rng.integers(1, m + 2, 400)
        j = rng.integers(1, n + 2, 400)
        exp = [ref_dominancecount(A, int(a), int(b)) for a, b in zip(i, j)]
        for hint in HINTS:
            assert ref.dominancecount(A, i, j, hint=hint).tolist() == exp


def test_dominancecount_step_walk(ref):
    rng = np.random.default_rng(4)
    for m in DIMS:
        for n in DIMS:
            A = sprand(rng, m, n, 0.5)
            i, j = [int(rng.integers(1, m + 2))], [int(rng.integers(1, n + 2))]
            for _ in range(60):
                mv = rng.integers(0, 6)
                a, b = i[-1], j[-1]
                if mv == 0 and a + 1 <= m + 1: a += 1
                elif mv == 1 and a - 1 >= 1: a -= 1
                elif mv == 2 and b + 1 <= n + 1: b += 1
                elif mv == 3 and b - 1 >= 1: b -= 1
                elif mv == 4: a, b = int(rng.integers(1, m + 2)), int(rng.integers(1, n + 2))
                i.append(a); j.append(b)
            got = ref.dominancecount_walk(A, i, j)
            assert got.tolist() == [ref_dominancecount(A, a, b) for a, b in zip(i, j)]


@pytest.mark.parametrize("hint", HINTS, ids=lambda h: type(h).__name__)
def test_color_arrays(ref, hint):
    rng = np.random.default_rng(5)
    for m in DIMS + [20]:
        for n in DIMS + [25]:
            A = sprand(rng, m, n, 0.5)
            pts = [tuple(sorted(p)) for p in probes(rng, 12, (1, n + 1), (1, n + 1))]
            j = [p[0] for p in pts]
            jp = [p[1] for p in pts]
            assert ref.netcount(A, j, jp, hint).tolist() == [ref_netcount(A, a, b) for a, b in pts]
            assert ref.selfnetcount(A, j, jp, hint).tolist() == [ref_selfnetcount(A, a, b) for a, b in pts]
            assert ref.pincount(A, j, jp, hint).tolist() == [int(A.colptr[b - 1] - A.colptr[a - 1]) for a, b in pts]
    for m in DIMS + [20, 40]:
        A = sprand(rng, m, m, 0.5)
        pts = [tuple(sorted(p)) for p in probes(rng, 12, (1, m + 1), (1, m + 1))]
        j = [p[0] for p in pts]
        jp = [p[1] for p in pts]
        assert ref.selfpincount(A, j, jp, hint).tolist() == [ref_selfpincount(A, a, b) for a, b in pts]
        assert ref.dianetcount(A, j, jp, hint).tolist() == [ref_dianetcount(A, a, b) for a, b in pts]


def test_color_arrays_fixtures(ref, fixtures):
    rng = np.random.default_rng(6)
    for name, A in fixtures.items():
        n = A.n
        j = rng.integers(1, n + 2, 200)
        jp = rng.integers(1, n + 2, 200)
        j, jp = np.minimum(j, jp), np.maximum(j, jp)
        exp = [ref_netcount(A, int(a), int(b)) for a, b in zip(j, jp)]
        exps = [ref_selfnetcount(A, int(a), int(b)) for a, b in zip(j, jp)]
        for hint in HINTS:
            assert ref.netcount(A, j, jp, hint).tolist() == exp, name
            assert ref.selfnetcount(A, j, jp, hint).tolist() == exps, name
        if A.m == A.n:
            for hint in HINTS:
                assert ref.dianetcount(A, j, jp, hint).tolist() == [ref_dianetcount(A, int(a), int(b)) for a, b in zip(j, jp)]
                assert ref.selfpincount(A, j, jp, hint).tolist() == [ref_selfpincount(A, int(a), int(b)) for a, b in zip(j, jp)]


def test_rowenvelope(ref):
    rng = np.random.default_rng(7)
    for m in list(range(1, 40)) + [100]:
        A = sprand(rng, m, m, 0.3)
        pts = [tuple(sorted(p)) for p in probes(rng, 20, (1, m + 1), (1, m + 1))]
        pts = [p for p in pts if p[0] < p[1]]
        lo, hi = ref.rowenvelope(A, [p[0] for p in pts], [p[1] for p in pts])
        assert list(zip(lo.tolist(), hi.tolist())) == [ref_envelope(A, a, b) for a, b in pts]
