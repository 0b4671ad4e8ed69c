"""A SECOND, independent restatement of the reference's six hot solvers -- TEST INFRASTRUCTURE ONLY.

Why: the reference (pure Julia) cannot run in this image and holds no golden outputs, so exact split vectors under ties
are pinned only by restatement.  ``oracle/cpo_*.hpp`` is one restatement (C++, with the reference's data structures, also
the timed CPU baseline).  This file is another one, written separately and as literally as Python allows from the Julia
sources -- same loops, same comparison operators, same update order -- with the cost oracle reduced to its DEFINITION
(distinct rows of a column range counted with a set), so it shares no code and no data structure with the C++ oracle.
``tests/test_witness.py`` fuzzes the two against each other; a disagreement means one of them misreads the reference.

Transliterated (file:line of /root/reference/src):
  DynamicSplitter.jl:15-50, 89-99            partition_stripe(::AbstractDynamicSplitter), unravel_splits
  BisectCostBottleneckSplitter.jl:6-63       partition_stripe(::BisectCostBottleneckSplitter)
  LazyBisectCostBottleneckSplitter.jl:140-258 partition_stripe(::LazyBisect...{<:AbstractConnectivityModel})
  DynamicChunker.jl:20-75                    pack_stripe(::DynamicTotalChunker{<:ConstrainedCost}), unravel_chunks!
  ConvexTotalChunker.jl:57-112, 141-165, 211-265  chunk_convex!, pack_stripe(::ConvexTotalChunker{<:ConstrainedCost}),
                                             chunk_convex_constrained!
  OverlapChunker.jl:6-75                     pack_stripe(::OverlapChunker)
  ConnectivityCosts.jl:20, 25-35             AffineConnectivityModel, bound_stripe
Arrays are 1-based like the source: index 0 of every list is unused padding.
"""
from __future__ import annotations

import math

TYPEMAX = (1 << 63) - 1


class Conn:
    """AffineConnectivityModel(alpha, beta_vertex, beta_pin, beta_net) (ConnectivityCosts.jl:7-20) and its oracle
    f(j, j') = mdl(j' - j, pos[j'] - pos[j], nets[j, j']) with nets = #distinct rows, by definition."""

    def __init__(self, A, coef):
        self.m, self.n = A.m, A.n
        self.pos = [0] + [int(x) for x in A.colptr]     # pos[1..n+1], 1-based values
        self.idx = [0] + [int(x) for x in A.rowval]     # idx[1..N]
        self.a, self.bv, self.bp, self.bn = coef
        self.is_float = any(isinstance(c, float) for c in coef)

    def model(self, n_vertices, n_pins, n_nets):
        return self.a + n_vertices * self.bv + n_pins * self.bp + n_nets * self.bn   # left to right, :20

    def nets(self, j, jp):
        rows = set()
        for q in range(self.pos[j], self.pos[jp]):
            rows.add(self.idx[q])
        return len(rows)

    def __call__(self, j, jp, k=None):
        return self.model(jp - j, self.pos[jp] - self.pos[j], self.nets(j, jp))


def fld(x, y):
    """Julia's fld: floor division on Int64; on Float64 round((x - mod(x, y)) / y) with mod built on the exact rem (fmod)."""
    if isinstance(x, float) or isinstance(y, float):
        x, y = float(x), float(y)
        r = math.fmod(x, y)
        if r == 0:
            r = math.copysign(r, y)
        elif (r > 0) != (y > 0):
            r = r + y
        return float(round((x - r) / y))
    return x // y


def bound_stripe(f: Conn, K):  # ConnectivityCosts.jl:25-35
    assert f.bv >= 0 and f.bp >= 0 and f.bn >= 0
    c_hi = f(1, f.n + 1)
    c_lo = f.a + fld(c_hi - f.a, K)
    return (c_lo, c_hi)


def unravel_splits(K, n, ptr):  # DynamicSplitter.jl:89-99, ptr[k][j']
    spl = [0] * (K + 2)
    spl[K + 1] = n + 1
    for k in range(K, 0, -1):
        spl[k] = ptr[k][spl[k + 1]]
    return spl[1:]


def dynamic_splitter(f: Conn, K, total: bool):  # DynamicSplitter.jl:15-50
    n = f.n
    g = (lambda a, b: a + b) if total else max
    ptr = [[0] * (n + 2) for _ in range(K + 1)]          # ptr[k][j']
    cst = [[math.inf] * (n + 2) for _ in range(K + 1)]   # typemax: never wins a `<=` against a real cost
    cst[1][1] = f(1, 1, 1)
    ptr[1][1] = 1
    for jp in range(2, n + 2):
        cst[1][jp] = f(1, jp, 1)
        ptr[1][jp] = 1
    for k in range(2, K + 1):
        jp0 = n + 1 if k == K else 1
        for jp in range(jp0, n + 2):
            cst[k][jp] = g(cst[k - 1][1], f(1, jp, k))
            ptr[k][jp] = 1
            for j in range(2, jp + 1):
                c2 = g(cst[k - 1][j], f(j, jp, k))
                if c2 <= cst[k][jp]:
                    cst[k][jp] = c2
                    ptr[k][jp] = j
    return unravel_splits(K, n, ptr)


def bisect_cost(f: Conn, K, eps):  # BisectCostBottleneckSplitter.jl:6-63
    n = f.n

    def search(j, jp_lo, jp_hi, k, c):
        jp_lo = max(j, jp_lo)
        while jp_lo <= jp_hi:
            jp = (jp_lo + jp_hi) >> 1
            if f(j, jp, k) <= c:
                jp_lo = jp + 1
            else:
                jp_hi = jp - 1
        return jp_hi

    spl_lo = [0] + [1] * (K + 1)
    spl_lo[K + 1] = n + 1
    spl_hi = [0] + [n + 1] * (K + 1)
    spl_hi[1] = 1
    spl = [0] * (K + 2)
    spl[1] = 1
    spl[K + 1] = n + 1
    c_lo, c_hi = bound_stripe(f, K)
    c_lo, c_hi = c_lo / 1, c_hi / 1
    while c_lo * (1 + eps) < c_hi:
        c = (c_lo + c_hi) / 2
        spl[1] = 1
        chk = True
        for k in range(1, K):
            spl[k + 1] = search(spl[k], spl_lo[k + 1], spl_hi[k + 1], k, c)
            if spl[k + 1] < spl[k]:
                chk = False
                for t in range(k + 1, K + 1):
                    spl[t] = spl[k]
                break
        if chk and f(spl[K], spl[K + 1], K) <= c:
            c_hi = c
            spl_hi = list(spl)
        else:
            c_lo = c
            spl_lo = list(spl)
    return spl_hi[1:]


def lazy_bisect_connectivity(f: Conn, K, eps):  # LazyBisectCostBottleneckSplitter.jl:140-258
    n, m = f.n, f.m
    pos, idx = f.pos, f.idx
    N = pos[n + 1] - 1
    spl = [0] * (K + 2)
    spl[1] = 1
    spl_hi = [0] + [n + 1] * (K + 1)
    spl_hi[1] = 1
    hst = [0] * (m + 1)
    cch = [0] * (N + 1)
    mdl = lambda nv, np_, nn, k: f.model(nv, np_, nn)

    def probe_init(c):
        spl[1] = 1
        j = 1
        k = 1
        n_vertices = n_pins = n_net = 0
        for jp in range(1, n + 1):
            n_vertices += 1
            n_pins += pos[jp + 1] - pos[jp]
            for q in range(pos[jp], pos[jp + 1]):
                i = idx[q]
                if hst[i] < j:
                    n_net += 1
                cch[q] = hst[i]
                hst[i] = jp
            while k < K and mdl(n_vertices, n_pins, n_net, k) > c:
                spl[k + 1] = jp
                j = jp
                k += 1
                n_vertices = 1
                n_pins = pos[jp + 1] - pos[jp]
                n_net = pos[jp + 1] - pos[jp]
        res = k < K or mdl(n_vertices, n_pins, n_net, K) <= c
        while k <= K:
            spl[k + 1] = n + 1
            k += 1
        return res

    def probe(c):
        spl[1] = 1
        j = 1
        k = 1
        n_vertices = n_pins = n_net = 0
        for jp in range(1, n + 1):
            n_vertices += 1
            n_pins += pos[jp + 1] - pos[jp]
            for q in range(pos[jp], pos[jp + 1]):
                if cch[q] < j:
                    n_net += 1
            while mdl(n_vertices, n_pins, n_net, k) > c:
                if k == K:
                    return False
                spl[k + 1] = jp
                j = jp
                k += 1
                n_vertices = 1
                n_pins = pos[jp + 1] - pos[jp]
                n_net = pos[jp + 1] - pos[jp]
        while k <= K:
            spl[k + 1] = n + 1
            k += 1
        return True

    c_lo, c_hi = bound_stripe(f, K)
    c_lo, c_hi = c_lo / 1, c_hi / 1
    for k in range(1, K + 1):
        c_lo = max(c_lo, mdl(0, 0, 0, k))
    if c_lo * (1 + eps) < c_hi:
        c = (c_lo + c_hi) / 2
        if probe_init(c):
            c_hi = c
            spl_hi = list(spl)
        else:
            c_lo = c
    while c_lo * (1 + eps) < c_hi:
        c = (c_lo + c_hi) / 2
        if probe(c):
            c_hi = c
            spl_hi = list(spl)
        else:
            c_lo = c
    return spl_hi[1:]


def unravel_chunks(spl, n):  # DynamicChunker.jl:58-75 (spl[1..n+1], in place in the source)
    K = 0
    jp = n + 1
    end = n + 1
    while jp != 1:
        j = spl[jp]
        spl[end - K] = jp
        K += 1
        jp = j
    spl[1] = 1
    for k in range(1, K + 1):
        spl[k + 1] = spl[end - K + k]
    return spl[1:K + 2]


def dynamic_total_chunker(f: Conn, w_max):  # DynamicChunker.jl:20-56, w = VertexCount(): w(j, j') = j' - j
    n = f.n
    cst = [0] * (n + 2)
    spl = [0] * (n + 2)
    j0 = 1
    for jp in range(2, n + 2):
        while jp - j0 > w_max:
            j0 += 1
        assert j0 < jp
        best_c = cst[j0] + f(j0, jp)
        best_j = j0
        for j in range(j0 + 1, jp):
            c = cst[j] + f(j, jp)
            if c < best_c:
                best_c = c
                best_j = j
        cst[jp] = best_c
        spl[jp] = best_j
    return unravel_chunks(spl, n)


def chunk_convex(cst, ptr, f, j0, jp1, ftr):  # ConvexTotalChunker.jl:57-112; ftr = list used as the deque's back
    ftr.clear()
    ftr.append((j0, jp1 + 1))
    for jp in range(j0 + 1, jp1 + 1):
        (j, h) = ftr[-1]
        c = f(j, jp)
        c2 = f(jp - 1, jp)
        if c <= c2:
            if c <= cst[jp]:
                cst[jp] = c
                ptr[jp] = j
            if h == jp + 1:
                ftr.pop()
        else:
            if c2 <= cst[jp]:
                cst[jp] = c2
                ptr[jp] = jp - 1
            while ftr:
                (j, h) = ftr[-1]
                if f(jp - 1, h - 1) < f(j, h - 1):
                    ftr.pop()
                else:
                    break
            if not ftr:
                ftr.append((jp - 1, jp1 + 1))
            else:
                (j, h) = ftr[-1]
                h_lo = jp + 1
                h_hi = h - 1
                while h_lo <= h_hi:
                    h = (h_lo + h_hi) >> 1
                    if f(jp - 1, h - 1) < f(j, h - 1):
                        h_lo = h + 1
                    else:
                        h_hi = h - 1
                h = h_hi
                if jp + 1 != h:
                    ftr.append((jp - 1, h))


def chunk_concave(cst, ptr, f, j0, jp1):  # ConcaveTotalChunker.jl:57-114; the CircularDeque as a collections.deque
    from collections import deque

    ftr = deque()
    ftr.append((j0, j0 + 1))
    for jp in range(j0 + 1, jp1 + 1):
        (j, h) = ftr[0]
        c = f(j, jp)
        c2 = f(jp - 1, jp)
        if c2 <= c:
            if c2 <= cst[jp]:
                cst[jp] = c2
                ptr[jp] = jp - 1
            ftr.clear()
            ftr.append((jp - 1, jp + 1))
        else:
            if c <= cst[jp]:
                cst[jp] = c
                ptr[jp] = j
            while True:
                (j, h) = ftr[-1]
                if f(jp - 1, h) <= f(j, h):
                    ftr.pop()
                else:
                    break
            (j, h) = ftr[-1]
            h_lo = h + 1
            h_hi = jp1
            while h_lo <= h_hi:
                h = (h_lo + h_hi) >> 1
                if f(jp - 1, h) > f(j, h):
                    h_lo = h + 1
                else:
                    h_hi = h - 1
            h = h_lo
            if h != jp1 + 1:
                ftr.append((jp - 1, h))
            (j, _) = ftr.popleft()
            if not ftr or ftr[0][1] != jp + 1:
                ftr.appendleft((j, jp + 1))


def concave_total_chunker(f: Conn):  # ConcaveTotalChunker.jl:9-24
    n = f.n
    INF = float("inf")  # typemax(cost_type): compares like the reference's sentinel, never enters a sum
    spl = [0] * (n + 2)
    cst = [INF] * (n + 2)
    cst[1] = 0
    chunk_concave(cst, spl, lambda j, jp: cst[j] + f(j, jp), 1, n + 1)
    return unravel_chunks(spl, n)


def concave_total_splitter(f: Conn, K):  # ConcaveTotalChunker.jl:26-55
    n = f.n
    if K == 1:
        return [1, n + 1]
    INF = float("inf")
    ptr = [[0] * (n + 2) for _ in range(K + 1)]
    cst = [[INF] * (n + 2) for _ in range(K + 1)]
    for jp in range(1, n + 2):
        cst[1][jp] = f(1, jp)
        ptr[1][jp] = 1
    for k in range(2, K + 1):
        fp = lambda j, jp, k=k: cst[k - 1][j] + f(j, jp)
        for jp in range(1, n + 2):
            cst[k][jp] = fp(jp, jp)
            ptr[k][jp] = jp
        chunk_concave(cst[k], ptr[k], fp, 1, n + 1)
    return unravel_splits(K, n, ptr)


def convex_total_chunker_constrained(f: Conn, w_max):  # ConvexTotalChunker.jl:141-165 + 211-265, w = VertexCount()
    n = f.n
    w = lambda j, jp: jp - j
    ftr = []
    size = 2 * n + 2
    s_j, s_jp, s_ptr, s_cst = [0] * size, [0] * size, [0] * size, [0] * size
    spl = [0] * (n + 2)
    cst = [TYPEMAX] * (n + 2)
    cst[1] = 0
    fp = lambda j, jp: cst[j] + f(j, jp)
    J0, Jp1 = 1, n + 1
    jp1 = J0 + 1
    while jp1 < Jp1 and w(J0, jp1 + 1) <= w_max:
        jp1 += 1
    j0 = J0
    while True:
        chunk_convex(cst, spl, fp, j0, jp1, ftr)
        if jp1 == Jp1:
            break
        jp = jp1
        I = 1
        for j in range(j0 + 1, jp1 + 1):
            if jp > jp1:
                s_jp[I] = jp
                I += 1
                s_j[I] = j
            while jp < Jp1 and w(j, jp + 1) <= w_max:
                jp += 1
                s_jp[I] = jp
                I += 1
                s_j[I] = j
        I += 1
        for i in range(2, I):
            s_cst[i] = TYPEMAX
        fp2 = lambda i, i2: fp(s_j[I - i], s_jp[I - i2])
        chunk_convex(s_cst, s_ptr, fp2, 1, I - 1, ftr)
        for i2 in range(2, I):
            cst[s_jp[I - i2]] = s_cst[i2]
            spl[s_jp[I - i2]] = s_j[I - s_ptr[i2]]
        j0 = jp1
        jp1 = s_jp[I - 2]
    return unravel_chunks(spl, n)


def overlap_chunker(A, rho, w_max):  # OverlapChunker.jl:6-75 -> (spl, n_nets)
    m, n = A.m, A.n
    pos = [0] + [int(x) for x in A.colptr]
    idx = [0] + [int(x) for x in A.rowval]
    hst = [0] * (m + 1)
    spl = [0] * (n + 2)
    n_nets = [0] * (n + 1)
    d = pos[2] - pos[1]
    c = pos[2] - pos[1]
    j = 1
    K = 0
    spl[1] = 1
    for q in range(pos[1], pos[2]):
        hst[idx[q]] = 1
    for jp in range(2, n + 1):
        c2 = pos[jp + 1] - pos[jp]
        d2 = d
        cc2 = 0
        for q in range(pos[jp], pos[jp + 1]):
            i = idx[q]
            h = hst[i]
            if abs(h) == j:
                cc2 += 1
                hst[i] = -jp
            elif j < h:
                hst[i] = jp
            elif h < -j:
                cc2 += 1
                hst[i] = -jp
            else:
                d2 += 1
                hst[i] = jp
        w = jp - j
        if w == w_max or cc2 < rho * min(c, c2):
            K += 1
            spl[K + 1] = jp
            n_nets[K] = d
            j = jp
            d = c2
        else:
            d = d2
    K += 1
    n_nets[K] = d
    spl[K + 1] = n + 1
    return spl[1:K + 2], n_nets[1:K + 1]
