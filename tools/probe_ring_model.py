#!/usr/bin/env python
"""Sequential numpy model of k_probe_ring's index arithmetic (chunk ownership, windows, boundary ownership, window
sizing, end conditions), checked against a plain greedy probe on random inputs.  No GPU: it exists because a GPU run is
expensive and the arithmetic is easy to get wrong.  Chunk size / ring depth are parameters so that small inputs
exercise many chunks per part and many parts per chunk.

    python tools/probe_ring_model.py            # random cases, prints the number checked
"""
import sys

import numpy as np

NCTA = 8


def links_of(colptr, rowval):
    """prev[q] = 1 + position of the previous nonzero of the same row (0 = none); colidx[q]; P[1..n+1] (0-based offsets)."""
    N = len(rowval)
    prev = np.zeros(N, dtype=np.int64)
    last = {}
    n = len(colptr) - 1
    colidx = np.zeros(N, dtype=np.int64)
    for c in range(n):
        for q in range(colptr[c], colptr[c + 1]):
            colidx[q] = c
            r = rowval[q]
            prev[q] = last.get(r, -1) + 1
            last[r] = q
    return prev, colidx


def greedy_probe(P, Wt, prev, coef, K, c):
    """The reference's streaming probe on the link array: returns (spl, feasible)."""
    n1 = len(P) - 1  # P indexed 1..n+1
    a, bv, bp, bn = coef
    spl = [1]
    j = 1
    for k in range(1, K + 1):
        if a > c:
            return spl, False
        e0 = P[j]
        r = j
        g = 0
        best = j
        # extend boundary by boundary
        for rr in range(j + 1, n1 + 1):
            g += int(np.sum(prev[P[rr - 1]:P[rr]] <= e0)) if P[rr] > P[rr - 1] else 0
            cost = a + (rr - j) * bv + (Wt[rr] - Wt[j]) * bp + g * bn
            if cost <= c:
                best = rr
            else:
                break
        if k == K:
            return spl + [n1], best == n1
        if best == n1:
            return spl + [n1] * (K + 1 - k), True
        spl.append(best)
        j = best
    return spl, False


def ring_probe(P, Wt, prev, colidx, coef, K, c, C, WMAX):
    n1 = len(P) - 1
    Ne = len(prev)
    a, bv, bp, bn = coef
    nchunks = (Ne + C - 1) // C
    chunk_col = [int(colidx[g * C]) if g * C < Ne else 0 for g in range(nchunks + 1)]
    spl = [1]
    j = 1
    pcur, wcur = P[1], Wt[1]
    W = w_est = WMAX
    steps = 0
    for k in range(1, K + 1):
        if a > c:
            return spl, False, steps
        e0, wj = pcur, wcur
        jlast, grun = j, 0
        first, missed = True, False
        g_win = e0 // C
        while True:
            steps += 1
            tot = {}      # window position -> chunk total
            per_cta = []
            for crank in range(NCTA):
                l0 = (g_win + 7 - crank) >> 3
                p0 = (crank - g_win) & 7
                chunks = []
                for v in range(W):
                    g = (l0 + v) * 8 + crank
                    assert g >= g_win and g - g_win == v * 8 + p0
                    x0 = g * C
                    head = first and g == g_win
                    if head:
                        ja = j + 1
                    else:
                        ja = chunk_col[g] + 2 if x0 < Ne else n1 + 1
                    if x0 >= Ne and not head:
                        jb = 0
                    else:
                        jb = n1 if x0 + C >= Ne else chunk_col[g + 1] + 1
                    nb = jb - ja + 1 if jb >= ja else 0
                    flags = np.zeros(C + 1, dtype=np.int64)  # prefix counts inside the chunk
                    lo, hi = max(x0, e0), min(x0 + C, Ne)
                    if hi > lo:
                        f = (prev[lo:hi] <= e0).astype(np.int64)
                        flags[lo - x0 + 1:hi - x0 + 1] = np.cumsum(f)
                        flags[hi - x0 + 1:] = flags[hi - x0]
                    tot[v * 8 + p0] = int(flags[C])
                    chunks.append((x0, ja, nb, flags, v * 8 + p0))
                per_cta.append(chunks)
            nwin = 8 * W
            base = np.concatenate(([0], np.cumsum([tot[q] for q in range(nwin)])))
            feas = nbs = 0
            best = None
            for crank in range(NCTA):
                cnt = 0
                lastp = lastw = 0
                seen_fail = False
                nb_cta = 0
                for (x0, ja, nb, flags, q) in per_cta[crank]:
                    nb_cta += nb
                    for r in range(ja, ja + nb):
                        x = P[r] - x0
                        assert 0 <= x <= C, (x, r, x0, C)
                        g = grun + int(base[q]) + int(flags[x])
                        ok = a + (r - j) * bv + (Wt[r] - wj) * bp + g * bn <= c
                        if ok:
                            assert not seen_fail, "feasible boundaries of a CTA must form a prefix"
                            cnt += 1
                            lastp, lastw = P[r], Wt[r]
                        else:
                            seen_fail = True
                feas += cnt
                nbs += nb_cta
                if cnt > 0 and (best is None or lastp >= best[0]):
                    best = (lastp, lastw)
            if best is not None:
                pcur, wcur = best
            first = False
            jlast += feas
            if feas < nbs:
                break
            if (g_win + 8 * W) * C >= Ne:
                break
            grun += int(base[nwin])
            g_win += 8 * W
            W = WMAX
            missed = True
        elems = pcur - e0
        want = (elems + (elems >> 3) + 9 * C - 1) // (C * NCTA)
        w_est = WMAX if missed else max(min(want, WMAX), w_est - 1 if w_est > 1 else 1)
        W = w_est
        if k == K:
            return spl + [n1], jlast == n1, steps
        if jlast == n1:
            return spl + [n1] * (K + 1 - k), True, steps
        assert pcur == P[jlast] and wcur == Wt[jlast], (pcur, P[jlast], jlast)
        spl.append(jlast)
        j = jlast
    return spl, False, steps


def main():
    rng = np.random.default_rng(7)
    checked = 0
    for case in range(400):
        n = int(rng.integers(1, 60))
        m = int(rng.integers(1, 40))
        dens = rng.choice([0.02, 0.1, 0.3, 0.8])
        cols = []
        colptr = [0]
        for c in range(n):
            if rng.random() < 0.3:
                rows = np.zeros(0, dtype=np.int64)  # empty columns
            else:
                rows = np.flatnonzero(rng.random(m) < dens)
            cols.append(rows)
            colptr.append(colptr[-1] + len(rows))
        rowval = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
        colptr = np.array(colptr)
        prev, colidx = links_of(colptr, rowval)
        P = np.concatenate(([0], colptr))  # P[x] = colptr[x-1], x = 1..n+1
        Wt = P
        coef = (int(rng.integers(0, 3)), int(rng.integers(0, 4)), int(rng.integers(0, 3)), int(rng.integers(0, 5)))
        K = int(rng.integers(1, 9))
        total = coef[0] + n * coef[1] + len(rowval) * coef[2] + m * coef[3]
        for c in sorted(set(rng.integers(0, max(total, 1) + 2, 6).tolist())):
            exp = greedy_probe(P, Wt, prev, coef, K, c)
            for C in (4, 8, 16):
                for WMAX in (1, 2, 6):
                    got = ring_probe(P, Wt, prev, colidx, coef, K, c, C, WMAX)
                    assert got[1] == exp[1] and (not exp[1] or got[0] == exp[0]), (case, c, C, WMAX, got, exp)
                    checked += 1
    print("ring model == greedy probe on", checked, "probes")


if __name__ == "__main__":
    sys.exit(main())
