"""Prototype: iterative rebalancing of the planner's bounding partition (numpy), compared with the optimal bottleneck."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import chainb200 as cp
from chainb200 import synth
import pyoracle as ref

def links(A):
    row = np.asarray(A.rowval) - 1
    order = np.argsort(row, kind="stable")
    prev = np.zeros(len(row), dtype=np.int64)
    same = row[order[1:]] == row[order[:-1]]
    prev[order[1:][same]] = order[:-1][same] + 1
    return prev

def part_costs(prev, pos, spl, coef):
    a, bv, bp, bn = coef
    K = len(spl) - 1
    c = np.zeros(K)
    for k in range(K):
        e0, e1 = pos[spl[k]], pos[spl[k + 1]]
        nets = int(np.count_nonzero(prev[e0:e1] <= e0))
        c[k] = a + bv * (spl[k + 1] - spl[k]) + bp * (e1 - e0) + bn * nets
    return c

def run(name, A, K, coef=(0, 10, 1, 100), iters=8):
    f = cp.AffineConnectivityModel(*coef)
    n = A.n
    pos = np.asarray(A.colptr) - 1  # 0-based offsets, pos[x] for x = 0..n (column x starts at pos[x])
    prev = links(A)
    t0 = time.time()
    Phi = ref.partition_stripe(A, K, cp.LazyBisectCostBottleneckSplitter(f, 0.01))
    spl_opt = np.asarray(Phi.spl) - 1
    copt = part_costs(prev, pos, spl_opt, coef).max()
    a, bv, bp, bn = coef
    U = bv * np.arange(n + 1) + (bp + bn) * pos  # surrogate weight at column boundary x (0-based)
    spl = np.searchsorted(U, U[0] + (U[-1] - U[0]) * np.arange(K + 1) / K).astype(np.int64)
    spl[0] = 0; spl[K] = n
    best = np.inf
    out = []
    for it in range(iters + 1):
        c = part_costs(prev, pos, spl, coef)
        best = min(best, c.max())
        out.append("%.3f" % (c.max() / copt))
        if it == iters: break
        w = np.maximum(c - a, 1e-9)
        # rescaled surrogate: inside old part k the surrogate increments are scaled so that the part weighs its exact cost
        scale = w / np.maximum(U[spl[1:]] - U[spl[:-1]], 1e-9)
        dU = np.diff(U)
        part_of_col = np.searchsorted(spl, np.arange(n), side="right") - 1
        W = np.concatenate([[0], np.cumsum(dU * scale[np.clip(part_of_col, 0, K - 1)])])
        def greedy(T):
            cuts = [0]; s_ = 0
            for k in range(K):
                r = int(np.searchsorted(W, W[s_] + T, side="right") - 1)
                if r <= s_: return None
                cuts.append(min(r, n)); s_ = cuts[-1]
                if s_ >= n: break
            if cuts[-1] < n: return None
            return cuts + [n] * (K + 1 - len(cuts))
        lo, hi = W[-1] / K, W[-1]
        for _ in range(40):
            mid = 0.5 * (lo + hi)
            if greedy(mid) is None: lo = mid
            else: hi = mid
        spl = np.array(greedy(hi), dtype=np.int64)
    print(name, "K", K, "bound/opt per iteration:", " ".join(out), " best %.4f" % (best / copt), flush=True)

run("rmat 16", synth.rmat(16, 16 << 16), 64)
run("rmat 18", synth.rmat(18, 16 << 18), 256)
run("rmat 20", synth.rmat(20, 16 << 20), 256)
run("rmat 20", synth.rmat(20, 16 << 20), 1024)
run("er 200k", synth.erdos_renyi(200000, 10), 64)
run("rgg 1M", synth.random_geometric(1 << 20, 8.0), 256, coef=(0, 0, 1, 100))
