"""Differential fuzzer: random matrices x cost models x solvers through the C ABI of libchainb200.so against the CPU oracle.

Not collected by pytest (a long-running checker, needs a GPU):

    python tests/fuzz_parity.py --seconds 240 --seed 1

Every case draws a matrix shape family (empty columns, empty rows, a dense row, banded, rectangular, 1 x n, m x 1 ...), a
cost model with random integer or Float64 coefficients, a solver and K, and compares split vectors bit for bit (and objective
values) between the device and the oracle.  Combinations the device declines with CPB_ERR_UNSUPPORTED, or the oracle rejects,
are counted and skipped.  Any mismatch is printed with everything needed to replay it, and the exit code is 1."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import chainb200 as cp  # noqa: E402
from chainb200 import synth  # noqa: E402
import pyoracle as ref  # noqa: E402
from helpers import sprand  # noqa: E402


def rand_large_matrix(rng):
    """mid-size inputs: several probe tiles, rows beyond the row-segment limit (radix-sort link path), many rank blocks"""
    fam = rng.integers(0, 4)
    if fam == 0:
        scale = int(rng.integers(12, 17))
        return synth.rmat(scale, int((1 << scale) * rng.integers(4, 24)))
    if fam == 1:
        return synth.erdos_renyi(int(rng.integers(20000, 150000)), int(rng.integers(2, 12)))
    if fam == 2:
        return synth.banded(int(rng.integers(5000, 60000)), int(rng.integers(1, 40)), int(rng.integers(2, 6)))
    return synth.random_geometric(int(rng.integers(10000, 80000)), float(rng.integers(3, 12)))


def rand_matrix(rng):
    fam = rng.integers(0, 8)
    if fam == 0:
        m, n = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        return sprand(rng, m, n, float(rng.choice([0.0, 0.1, 0.4, 0.9])))
    if fam == 1:  # square
        n = int(rng.integers(1, 70))
        return sprand(rng, n, n, float(rng.choice([0.02, 0.1, 0.3])))
    if fam == 2:  # rectangular
        m, n = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        return sprand(rng, m, n, float(rng.choice([0.02, 0.1, 0.3])))
    if fam == 3:  # a dense row and a dense column on top of a sparse pattern
        n = int(rng.integers(2, 60))
        mask = rng.random((n, n)) < 0.05
        mask[int(rng.integers(0, n)), :] = True
        mask[:, int(rng.integers(0, n))] = rng.random(n) < 0.8
        I, J = np.nonzero(mask)
        return cp.SparseMatrixCSC.from_coo(n, n, I + 1, J + 1)
    if fam == 4:  # banded, square
        n = int(rng.integers(2, 200))
        hb = int(rng.integers(0, 4))
        I, J = np.nonzero(np.abs(np.subtract.outer(np.arange(n), np.arange(n))) <= hb)
        keep = rng.random(len(I)) < 0.8
        return cp.SparseMatrixCSC.from_coo(n, n, I[keep] + 1, J[keep] + 1)
    if fam == 5:  # blocks of empty columns
        m, n = int(rng.integers(1, 40)), int(rng.integers(4, 80))
        mask = rng.random((m, n)) < 0.2
        a = int(rng.integers(0, n))
        mask[:, a : a + int(rng.integers(1, n))] = False
        I, J = np.nonzero(mask)
        return cp.SparseMatrixCSC.from_coo(m, n, I + 1, J + 1)
    if fam == 6:  # larger, sparse: several probe tiles / rank blocks
        n = int(rng.integers(300, 3000))
        return synth.erdos_renyi(n, int(rng.integers(1, 8)))
    m, n = (1, int(rng.integers(1, 30))) if rng.random() < 0.5 else (int(rng.integers(1, 30)), 1)
    return sprand(rng, m, n, 0.7)


def coefs(rng, k, allow_float=True):
    if allow_float and rng.random() < 0.35:
        return [float(x) for x in rng.choice([0.0, 0.25, 0.5, 1.0, 1.5, 3.0, 6.5, 10.0], k)]
    return [int(x) for x in rng.choice([0, 0, 1, 1, 2, 3, 5, 10, 100], k)]


def rand_model(rng, A):
    """(model, needs row partition, decreasing)"""
    sq = A.m == A.n
    kinds = ["work", "conn", "conn", "hyper", "primconn", "secconn", "primedge", "secedge", "envelope"] + (["monosym", "symconn", "symedge"] if sq else [])
    kind = kinds[int(rng.integers(0, len(kinds)))]
    if kind == "work":
        return cp.AffineWorkModel(*coefs(rng, 3)), False, False
    if kind == "conn":
        return cp.AffineConnectivityModel(*coefs(rng, 4)), False, False
    if kind == "hyper":
        return cp.AffineHyperedgeCutModel(*coefs(rng, 4)), False, False
    if kind == "envelope":
        return cp.AffineEnvelopeModel(*coefs(rng, 4)), False, False
    if kind == "monosym":
        c = coefs(rng, 4)
        return cp.AffineMonotonizedSymmetricConnectivityModel(*c, type(c[0])(rng.integers(0, 6))), False, False
    if kind == "symconn":
        return cp.AffineSymmetricConnectivityModel(*coefs(rng, 5)), False, False
    if kind == "symedge":
        return cp.AffineSymmetricEdgeCutModel(*coefs(rng, 4)), False, False
    if kind == "primconn":
        return cp.AffinePrimaryConnectivityModel(*coefs(rng, 5)), True, False
    if kind == "primedge":
        return cp.AffinePrimaryEdgeCutModel(*coefs(rng, 4)), True, False
    if kind == "secconn":
        c = coefs(rng, 5)
        c[3], c[4] = min(c[3], c[4]), max(c[3], c[4])
        return cp.AffineSecondaryConnectivityModel(*c), True, True
    c = coefs(rng, 4)
    c[2], c[3] = min(c[2], c[3]), max(c[2], c[3])
    return cp.AffineSecondaryEdgeCutModel(*c), True, True


def rand_splitter(rng, f, decreasing, constrained_ok, large=False):
    eps = float(rng.choice([0.3, 0.1, 0.01, 0.001]))
    if large:  # the O(n^2) reference DPs would take minutes here
        if decreasing:
            return rng.choice([cp.FlipBisectIndexBottleneckSplitter(f), cp.FlipBisectCostBottleneckSplitter(f, eps), cp.LazyFlipBisectCostBottleneckSplitter(f, eps)])
        return rng.choice([cp.BisectCostBottleneckSplitter(f, eps), cp.LazyBisectCostBottleneckSplitter(f, eps), cp.LazyBisectCostBottleneckSplitter(f, eps),
                           cp.BisectIndexBottleneckSplitter(f), cp.EquiSplitter()])
    spec = f
    if constrained_ok and rng.random() < 0.25:
        w = cp.VertexCount() if rng.random() < 0.6 else cp.AffineWorkModel(*[int(x) for x in rng.choice([0, 1, 2], 3)])
        spec = cp.ConstrainedCost(f, w, int(rng.integers(1, 40)))
    if decreasing:
        return rng.choice([cp.FlipBisectIndexBottleneckSplitter(f), cp.FlipBisectCostBottleneckSplitter(f, eps), cp.LazyFlipBisectCostBottleneckSplitter(f, eps),
                           cp.DynamicBottleneckSplitter(f)])
    return rng.choice([cp.DynamicBottleneckSplitter(spec), cp.DynamicTotalSplitter(spec), cp.DynamicBottleneckChunker(spec), cp.DynamicTotalChunker(spec),
                       cp.BisectCostBottleneckSplitter(f, eps), cp.LazyBisectCostBottleneckSplitter(f, eps), cp.BisectIndexBottleneckSplitter(f),
                       cp.ConvexTotalSplitter(spec), cp.ConcaveTotalSplitter(spec), cp.EquiSplitter()])


def rand_packer(rng, A):
    w_max = int(rng.integers(1, 12))
    r = rng.integers(0, 7)
    if r == 6:  # ConcaveTotalChunker.jl:9-24 (any random-access model: the device follows the queue routine step by step)
        f = rng.choice([cp.AffineConnectivityModel(*coefs(rng, 4)), cp.AffineWorkModel(*coefs(rng, 3)), cp.ColumnBlockComponentCostModel(int, 5, lambda w: int(9 * np.sqrt(w)))])
        return cp.ConcaveTotalChunker(f), False
    if r == 0:
        return cp.OverlapChunker(float(rng.choice([0.0, 0.3, 0.7, 0.9, 1.0])), w_max), False
    if r == 1:
        return cp.StrictChunker(w_max), False
    if r == 2:
        return cp.EquiChunker(w_max), False
    if r == 3:
        f = cp.AffineConnectivityModel(*coefs(rng, 4, allow_float=False))
        return cp.ConvexTotalChunker(cp.ConstrainedCost(f, cp.VertexCount(), w_max)), False
    if r == 4:
        f = rng.choice([cp.AffineConnectivityModel(*coefs(rng, 4)), cp.AffineWorkModel(*coefs(rng, 3)), cp.ColumnBlockComponentCostModel(int, 3, lambda w: 1 + w)])
        return cp.DynamicTotalChunker(cp.ConstrainedCost(f, cp.VertexCount(), w_max)), False
    a, b = int(rng.integers(0, 4)), int(rng.integers(0, 4))
    f = cp.BlockComponentCostModel(int, a, b, (int(rng.integers(0, 5)), cp.identity), (int(rng.integers(0, 5)), lambda x: 2 * x))
    return cp.DynamicTotalChunker(cp.ConstrainedCost(f, cp.VertexCount(), w_max)), True


def misc_case(rng, A, stats):
    """structures and drivers around the solvers: prefix matrices, colour counts, adjointpattern, objectives of
    non-contiguous partitions, the plaid driver.  Returns a description of the first difference, or None."""
    r = int(rng.integers(0, 5))
    if r == 0:
        val = rng.integers(0, 2**64, A.nnz, dtype=np.uint64) if rng.random() < 0.5 else rng.integers(-9, 10, A.nnz)
        i, j = rng.integers(1, A.m + 2, 24), rng.integers(1, A.n + 2, 24)
        C, S = cp.dominancecount(A), cp.dominancesum(A, val)
        gc, gs = C.query(i, j), S.query(i, j)
        C.close(); S.close()
        if not np.array_equal(gc, ref.dominancecount(A, i, j)) or not np.array_equal(gs, ref.prefix_query(A.m, A.n, A.nnz, A.colptr, A.rowval, val, i, j)):
            return "prefix structures"
    elif r == 1:
        j = rng.integers(1, A.n + 2, 24)
        jp = rng.integers(1, A.n + 2, 24)
        j, jp = np.minimum(j, jp), np.maximum(j, jp)
        names = ["pincount", "netcount", "selfnetcount"] + (["dianetcount", "selfpincount"] if A.m == A.n else [])
        for nm in names:
            if not np.array_equal(getattr(cp, nm)(A).query(j, jp), getattr(ref, nm)(A, j, jp)):
                return nm
    elif r == 2:
        g, e = cp.adjointpattern(A), ref.adjointpattern(A)
        if not (np.array_equal(g.colptr, e.colptr) and np.array_equal(g.rowval, e.rowval)):
            return "adjointpattern"
    elif r == 3:
        K = int(rng.integers(1, 6))
        Phi = cp.MapPartition(K, rng.integers(1, K + 1, A.n))
        Pi = cp.MapPartition(K, rng.integers(1, K + 1, A.m))
        f, needs_pi, _ = rand_model(rng, A)
        if isinstance(f, (cp.AffineSymmetricConnectivityModel, cp.AffineSymmetricEdgeCutModel, cp.AffineMonotonizedSymmetricConnectivityModel)):
            return None  # diagonal-aware models have no meaning after a column permutation alone
        pi = Pi if needs_pi else None
        for fn in ("bottleneck_value", "total_value"):
            if getattr(cp, fn)(A, Phi, f, pi) != getattr(ref, fn)(A, Phi, f, pi):
                return fn + " of a MapPartition, " + type(f).__name__
    else:
        K = int(rng.integers(1, 6))
        eps = float(rng.choice([0.1, 0.01]))
        f1, f2 = cp.AffineConnectivityModel(*coefs(rng, 4)), cp.AffineConnectivityModel(*coefs(rng, 4))
        meth = cp.AlternatingPartitioner(cp.LazyBisectCostBottleneckSplitter(f1, eps), cp.BisectCostBottleneckSplitter(f2, eps))
        (Pg, Fg), (Pr, Fr) = cp.partition_plaid(A, K, meth), ref.partition_plaid(A, K, meth)
        if not (np.array_equal(Pg.spl, Pr.spl) and np.array_equal(Fg.spl, Fr.spl)):
            return "partition_plaid"
    stats["misc"] += 1
    return None


def describe(A):
    return dict(m=A.m, n=A.n, colptr=A.colptr.tolist(), rowval=A.rowval.tolist()) if A.nnz <= 400 else dict(m=A.m, n=A.n, nnz=A.nnz)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-cases", type=int, default=10**9)
    ap.add_argument("--focus", default="", help="'total': only the unconstrained total-cost splitters on 65..600 columns -- the layers the device "
                                                  "solves by monotone divide & conquer (an empirical property of these costs, DESIGN.md 4)")
    ap.add_argument("--large", type=float, default=0.0, help="fraction of mid-size inputs (10^4..10^5 columns, up to ~10^6 nonzeros)")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    cp.init(0)
    t0 = time.time()
    stats = dict(cases=0, compared=0, queries=0, misc=0, unsupported=0, oracle_rejected=0, mismatches=0)
    by_method = {}
    while time.time() - t0 < args.seconds and stats["cases"] < args.max_cases:
        stats["cases"] += 1
        large = rng.random() < args.large
        A = rand_large_matrix(rng) if large else rand_matrix(rng)
        if args.focus == "total":
            n = int(rng.integers(65, 600))
            fam = int(rng.integers(0, 4))
            if fam == 0:
                A = sprand(rng, int(rng.integers(1, 400)), n, float(rng.choice([0.005, 0.02, 0.1, 0.3])))
            elif fam == 1:
                A = sprand(rng, n, n, float(rng.choice([0.005, 0.02, 0.1])))
            elif fam == 2:
                hb = int(rng.integers(0, 6))
                I, J = np.nonzero(np.abs(np.subtract.outer(np.arange(n), np.arange(n))) <= hb)
                keep = rng.random(len(I)) < float(rng.choice([0.3, 0.8, 1.0]))
                A = cp.SparseMatrixCSC.from_coo(n, n, I[keep] + 1, J[keep] + 1)
            else:  # a few dense rows / columns over a sparse pattern
                mask = rng.random((n, n)) < 0.01
                for _ in range(int(rng.integers(1, 4))):
                    mask[int(rng.integers(0, n)), :] |= rng.random(n) < 0.7
                    mask[:, int(rng.integers(0, n))] |= rng.random(n) < 0.5
                I, J = np.nonzero(mask)
                A = cp.SparseMatrixCSC.from_coo(n, n, I + 1, J + 1)
            try:
                fs = [cp.AffineConnectivityModel(*coefs(rng, 4)), cp.AffineWorkModel(*coefs(rng, 3))]
                if A.m == A.n:
                    c = coefs(rng, 4)
                    fs.append(cp.AffineMonotonizedSymmetricConnectivityModel(*c, type(c[0])(rng.integers(0, 6))))
                f = fs[int(rng.integers(0, len(fs)))]
                K = int(rng.integers(2, 12))
                mtd = cp.DynamicTotalSplitter(f) if rng.random() < 0.6 else cp.ConvexTotalSplitter(f)
                r = ref.partition_stripe(A, K, mtd)
                g = cp.partition_stripe(A, K, mtd)
                stats["compared"] += 1
                if not np.array_equal(g.spl, r.spl):
                    stats["mismatches"] += 1
                    print("MISMATCH", type(mtd).__name__, f.__dict__, "K =", K, "A =", describe(A), "\n  device:", g.spl.tolist(), "\n  oracle:", r.spl.tolist(),
                          "totals", cp.total_value(A, g, f), ref.total_value(A, r, f), flush=True)
            except cp.CpbError as e:
                stats["unsupported"] += 1
            continue
        packing = rng.random() < 0.25
        Pi = None
        if rng.random() < 0.12:
            mtd = None
            try:
                what = misc_case(rng, A, stats)
            except Exception as e:
                what = "ERROR " + type(e).__name__ + " " + str(e)[:200]
            if what:
                stats["mismatches"] += 1
                print("MISC MISMATCH", what, "A =", describe(A), flush=True)
            continue
        try:
            if packing:
                mtd, needs_pi = rand_packer(rng, A)
                if needs_pi:
                    Pi = ref.pack_stripe(ref.adjointpattern(A), cp.EquiChunker(int(rng.integers(1, 5))))
                run = lambda mod: mod.pack_stripe(A, mtd, Pi) if Pi is not None else mod.pack_stripe(A, mtd)
                K = None
            else:
                f, needs_pi, decreasing = rand_model(rng, A)
                K = int(rng.integers(1, min(A.n + 3, 14))) if not large else int(rng.choice([2, 7, 64, 300, 1500]))
                if needs_pi:
                    Pi = ref.partition_stripe(ref.adjointpattern(A), K, cp.EquiSplitter())
                mtd = rand_splitter(rng, f, decreasing, constrained_ok=not needs_pi, large=large)
                if rng.random() < 0.3:  # the oracle itself: c(j, j', k) on random ranges, and bound_stripe where the reference has one
                    j = rng.integers(1, A.n + 2, 24)
                    jp = rng.integers(1, A.n + 2, 24)
                    j, jp = np.minimum(j, jp), np.maximum(j, jp)
                    kk = rng.integers(1, K + 1, 24)
                    ocl = cp.oracle_stripe(f, A, Pi) if Pi is not None else cp.oracle_stripe(f, A)
                    got = ocl.query(j, jp, kk)
                    ocl.close()
                    exp = ref.oracle_query(f, A, j, jp, kk, Pi=Pi)
                    stats["queries"] += 1
                    if not np.array_equal(got, exp):
                        stats["mismatches"] += 1
                        print("QUERY MISMATCH", type(f).__name__, f.__dict__, "Pi =", None if Pi is None else Pi.spl.tolist(), "\n  A =", describe(A),
                              "\n  j, jp, k =", j.tolist(), jp.tolist(), kk.tolist(), "\n  device:", got.tolist(), "\n  oracle:", exp.tolist(), flush=True)
                    try:
                        eb = ref.bound_stripe(A, K, f) if Pi is None else None
                    except Exception:
                        eb = None
                    if eb is not None:
                        gb = cp.bound_stripe(A, K, f)
                        if tuple(map(float, gb)) != tuple(map(float, eb)):
                            stats["mismatches"] += 1
                            print("BOUND MISMATCH", type(f).__name__, f.__dict__, "K =", K, "A =", describe(A), gb, eb, flush=True)
                run = lambda mod: mod.partition_stripe(A, K, mtd, Pi) if Pi is not None else mod.partition_stripe(A, K, mtd)
            try:
                r = run(ref)
            except Exception:
                stats["oracle_rejected"] += 1
                continue
            try:
                g = run(cp)
            except cp.CpbError as e:
                if e.code == -2:
                    stats["unsupported"] += 1
                    continue
                raise
            name = type(mtd).__name__ + ("/" + type(getattr(mtd, "f", None)).__name__ if hasattr(mtd, "f") else "")
            by_method[name] = by_method.get(name, 0) + 1
            stats["compared"] += 1
            stats["large"] = stats.get("large", 0) + int(large)
            if g.K != r.K or not np.array_equal(g.spl, r.spl):
                stats["mismatches"] += 1
                print("MISMATCH", name, "K =", K, "model =", getattr(getattr(mtd, "f", None), "__dict__", None), "Pi =", None if Pi is None else Pi.spl.tolist(),
                      "\n  A =", describe(A), "\n  device:", g.spl.tolist()[:40], "\n  oracle:", r.spl.tolist()[:40], flush=True)
        except Exception as e:  # an error on either side that is not a declared refusal is a finding too
            stats["mismatches"] += 1
            print("ERROR", type(e).__name__, str(e)[:300], "method =", type(mtd).__name__ if "mtd" in dir() else None, "A =", describe(A), flush=True)
    print("fuzz:", stats, "seed", args.seed, "seconds %.0f" % (time.time() - t0))
    print("compared by method:", dict(sorted(by_method.items())))
    sys.exit(1 if stats["mismatches"] else 0)


if __name__ == "__main__":
    main()
