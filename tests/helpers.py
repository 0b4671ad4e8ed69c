"""Shared test helpers: fixture loading, brute-force definitions (the reference's own test
definitions re-expressed), random inputs."""
import os

import numpy as np

import chainb200 as cp

HERE = os.path.dirname(os.path.abspath(__file__))


def load_fixtures():
    """The six matrices of /root/reference/test/matrices.jl (tests/golden/make_fixtures.py)."""
    z = np.load(os.path.join(HERE, "golden", "matrices.npz"))
    names = sorted({k.rsplit("__", 1)[0] for k in z.files})
    out = {}
    for key in names:
        out[key.replace("__", "/")] = cp.SparseMatrixCSC.from_coo(int(z[key + "__m"]), int(z[key + "__n"]), z[key + "__I"], z[key + "__J"])
    return out


def sprand(rng, m, n, p):
    """dropzeros!(sprand(m, n, p)) as a pattern."""
    mask = rng.random((m, n)) < p
    I, J = np.nonzero(mask)
    return cp.SparseMatrixCSC.from_coo(m, n, I + 1, J + 1)


def col_rows(A, j):
    return A.rowval[A.colptr[j - 1] - 1 : A.colptr[j] - 1]


def rows_in(A, j, jp):
    return A.rowval[A.colptr[j - 1] - 1 : A.colptr[jp - 1] - 1]


# ---- test/test_SparseColorArrays.jl:1-11 and test/test_SparsePrefixMatrices.jl:14 -------------
def ref_dominancecount(A, i, j):
    r = rows_in(A, 1, j)
    return int(np.sum(r <= i - 1))


def ref_netcount(A, j, jp):
    return len(set(rows_in(A, j, jp).tolist()))


def ref_selfnetcount(A, j, jp):
    inside = set(rows_in(A, j, jp).tolist())
    outside = set(rows_in(A, 1, j).tolist()) | set(rows_in(A, jp, A.n + 1).tolist())
    return len(inside - outside)


def ref_selfpincount(A, j, jp):
    r = rows_in(A, j, jp)
    return int(np.sum((r >= j) & (r <= jp - 1)))


def ref_dianetcount(A, j, jp):
    s = set(rows_in(A, j, jp).tolist()) | set(range(j, jp))
    return len(s)


def ref_envelope(A, j, jp):
    r = rows_in(A, j, jp)
    if len(r) == 0:
        return (A.m + 1, 0)
    return (int(r.min()), int(r.max()))


def ref_counts(A, j, jp):
    """dict of every count the cost models consume, from the set definitions above."""
    nv = jp - j
    np_ = int(A.colptr[jp - 1] - A.colptr[j - 1])
    return dict(nv=nv, np=np_, net=ref_netcount(A, j, jp))


def ref_cost(mdl, A, j, jp, Pi=None):
    """Cost models evaluated from the brute-force counts (left-to-right sums)."""
    nv = jp - j
    npin = int(A.colptr[jp - 1] - A.colptr[j - 1])
    c = mdl.coef if hasattr(mdl, "coef") else None
    k = mdl.kind
    if k == cp.MODEL_WORK:
        return c[0] + nv * c[1] + npin * c[2]
    if k in (cp.MODEL_CONNECTIVITY,):
        return c[0] + nv * c[1] + npin * c[2] + ref_netcount(A, j, jp) * c[3]
    if k == cp.MODEL_ENVELOPE:
        lo, hi = ref_envelope(A, j, jp)
        return c[0] + nv * c[1] + npin * c[2] + max(hi - lo, 0) * c[3]
    if k == cp.MODEL_MONOSYM:
        deg = np.diff(A.colptr)[j - 1 : jp - 1]
        over = int(np.sum(np.maximum(deg - c[4], 0)))
        return c[0] + nv * c[1] + over * c[2] + ref_dianetcount(A, j, jp) * c[3]
    if k == cp.MODEL_SYMCONN:
        d = ref_netcount(A, j, jp)
        r = ref_dianetcount(A, j, jp) - nv
        return c[0] + nv * c[1] + npin * c[2] + (d - r) * c[3] + r * c[4]
    if k == cp.MODEL_HYPEREDGE:
        d = ref_netcount(A, j, jp)
        l = ref_selfnetcount(A, j, jp)
        return c[0] + nv * c[1] + npin * c[2] + l * c[3] + (d - l) * c[4]
    if k == cp.MODEL_SYMEDGECUT:
        l = ref_selfpincount(A, j, jp)
        return c[0] + nv * c[1] + l * c[2] + (npin - l) * c[3]
    if k == cp.MODEL_COLBLOCK:
        return cp.block_component(mdl.alpha_col, nv) + ref_netcount(A, j, jp) * cp.block_component(mdl.beta_col, nv)
    if k == cp.MODEL_BLOCK:
        rows = set(rows_in(A, j, jp).tolist())
        parts = sorted({int(np.searchsorted(Pi.spl, r, side="right")) for r in rows})
        cost = cp.block_component(mdl.alpha_col, nv)
        for br, bc in zip(mdl.beta_row, mdl.beta_col):
            d = sum(cp.block_component(br, int(Pi.spl[kk] - Pi.spl[kk - 1])) for kk in parts)
            cost += d * cp.block_component(bc, nv)
        return cost
    raise ValueError(k)


def cost_matrix(mdl, A, Pi=None):
    """c[j, j'] for all 1 <= j <= j' <= n+1 (brute force; tiny inputs only)."""
    n = A.n
    C = np.full((n + 2, n + 2), np.inf)
    for j in range(1, n + 2):
        for jp in range(j, n + 2):
            C[j, jp] = ref_cost(mdl, A, j, jp, Pi)
    return C


def brute_optimum(C, n, K, total):
    """Optimal objective of splitting columns 1..n into exactly K contiguous parts."""
    best = np.full((K + 1, n + 2), np.inf)
    for jp in range(1, n + 2):
        best[1, jp] = C[1, jp]
    for k in range(2, K + 1):
        for jp in range(1, n + 2):
            vals = [(best[k - 1, j] + C[j, jp]) if total else max(best[k - 1, j], C[j, jp]) for j in range(1, jp + 1)]
            best[k, jp] = min(vals)
    return best[K, n + 1]


def objective(C, spl, total):
    vals = [C[int(spl[k]), int(spl[k + 1])] for k in range(len(spl) - 1)]
    return sum(vals) if total else max(vals)


def brute_chunk_optimum(C, n, w_max=None):
    """Optimal total cost over all chunkings into parts of width <= w_max."""
    best = np.full(n + 2, np.inf)
    best[1] = 0
    for jp in range(2, n + 2):
        lo = 1 if w_max is None else max(1, jp - w_max)
        best[jp] = min(best[j] + C[j, jp] for j in range(lo, jp))
    return best[n + 1]


def check_split(spl, n, K=None):
    spl = np.asarray(spl)
    assert spl[0] == 1 and spl[-1] == n + 1
    assert np.all(np.diff(spl) >= 0)
    if K is not None:
        assert len(spl) == K + 1
