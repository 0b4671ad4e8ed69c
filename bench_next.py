#!/usr/bin/env python
"""bench_next.py -- the widened rows on one B200, each beside the CPU oracle on the same input.

Not the driver's bench (bench.py, config 2).  One JSON line per case to stdout and gpurun_out/next_<tag>.jsonl: the public
call with the pattern resident in HBM and with host arrays, the single-threaded CPU restatement of the reference, and whether
the results are identical.  Cases (Erdos-Renyi n = m = 10^6, ~10 nonzeros per column unless noted):

  bisect_index      BisectIndexBottleneckSplitter(connectivity), K = 64                     (BisectIndexBottleneckSplitter.jl:5-81)
  primary_conn      BisectCost(AffinePrimaryConnectivityModel, 0.01), Pi = Equi(A', 64)       (PrimaryConnectivityCosts.jl, PartwiseCounts.jl)
  secondary_conn    FlipBisectCost(AffineSecondaryConnectivityModel, 0.01), Pi as above       (SecondaryConnectivityCosts.jl, BisectCost...:70-127)
  primary_edge      LazyBisectCost(AffinePrimaryEdgeCutModel, 0.01)                           (PrimaryEdgeCutCosts.jl)
  secondary_edge    FlipBisectIndex(AffineSecondaryEdgeCutModel)                              (SecondaryEdgeCutCosts.jl, BisectIndex...:87-166)
  convex_constrained  ConvexTotalSplitter(ConstrainedCost(nets, VertexCount, ceil(1.5 n / K))), K = 8, banded n = 2^14
                                                                                              (bin/test_table_constrained_splits.jl:26-40)
  map_objective     bottleneck_value of a random MapPartition (K = 64), connectivity          (Costs.jl:52-66)
  prefix            dominancecount + dominancesum build and 2^20 queries                      (SparsePrefixMatrices.jl; CPU = b-ary DominanceCount
                                                                                               restated / Fenwick sweep for the sums)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import chainb200 as cp  # noqa: E402
import pyoracle as ref  # noqa: E402
from chainb200 import synth, synth_torch  # noqa: E402


def timed(fn, reps):
    best, out = None, None
    for _ in range(reps):
        cp.synchronize()
        t0 = time.perf_counter()
        out = fn()
        cp.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best * 1e3, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cases", default="bisect_index,primary_conn,secondary_conn,primary_edge,secondary_edge,convex_constrained,map_objective,prefix")
    args = ap.parse_args()
    cp.init(0)
    cases = args.cases.split(",")
    K = 64
    A = synth_torch.erdos_renyi(args.n, 10)
    dA = cp.device_matrix(A)
    Pi = ref.partition_stripe(ref.adjointpattern(A), K, cp.EquiSplitter())
    lines = []

    def emit(line):
        lines.append(line)
        print(json.dumps(line), flush=True)

    def split_case(name, mtd, Pi_arg, matrix=A, dmatrix=dA, k=K):
        call = (lambda M: cp.partition_stripe(M, k, mtd, Pi_arg)) if Pi_arg is not None else (lambda M: cp.partition_stripe(M, k, mtd))
        call(dmatrix)
        t_res, g = timed(lambda: call(dmatrix), args.reps)
        t_e2e, g2 = timed(lambda: call(matrix), args.reps)
        t0 = time.perf_counter()
        r = ref.partition_stripe(matrix, k, mtd, Pi_arg) if Pi_arg is not None else ref.partition_stripe(matrix, k, mtd)
        t_cpu = (time.perf_counter() - t0) * 1e3
        emit({"case": name, "n": matrix.n, "nnz": matrix.nnz, "K": k, "gpu_resident_ms": t_res, "gpu_e2e_ms": t_e2e, "cpu_oracle_ms": t_cpu,
              "identical": bool(np.array_equal(g.spl, r.spl) and np.array_equal(g2.spl, r.spl)), "speedup_e2e": t_cpu / t_e2e})

    net = cp.AffineConnectivityModel(0, 10, 1, 100)
    if "bisect_index" in cases:
        split_case("bisect_index: BisectIndexBottleneckSplitter(connectivity)", cp.BisectIndexBottleneckSplitter(net), None)
    if "primary_conn" in cases:
        split_case("primary_conn: BisectCost(AffinePrimaryConnectivityModel(0,10,1,0,100), 0.01), Pi = Equi(A', 64)",
                   cp.BisectCostBottleneckSplitter(cp.AffinePrimaryConnectivityModel(0, 10, 1, 0, 100), 0.01), Pi)
    if "secondary_conn" in cases:
        split_case("secondary_conn: FlipBisectCost(AffineSecondaryConnectivityModel(0,10,1,0,100), 0.01)",
                   cp.FlipBisectCostBottleneckSplitter(cp.AffineSecondaryConnectivityModel(0, 10, 1, 0, 100), 0.01), Pi)
    if "primary_edge" in cases:
        split_case("primary_edge: LazyBisectCost(AffinePrimaryEdgeCutModel(0,10,1,100), 0.01)",
                   cp.LazyBisectCostBottleneckSplitter(cp.AffinePrimaryEdgeCutModel(0, 10, 1, 100), 0.01), Pi)
    if "secondary_edge" in cases:
        split_case("secondary_edge: FlipBisectIndex(AffineSecondaryEdgeCutModel(0,10,1,100))",
                   cp.FlipBisectIndexBottleneckSplitter(cp.AffineSecondaryEdgeCutModel(0, 10, 1, 100)), Pi)
    if "convex_constrained" in cases:
        B = synth.banded(1 << 14, 16, 4)
        dB = cp.device_matrix(B)
        for k in (4, 8, 16):
            mdl = cp.ConstrainedCost(cp.AffineConnectivityModel(0, 0, 0, 1), cp.VertexCount(), int(np.ceil(B.n / k * 1.5)))
            split_case("convex_constrained: ConvexTotalSplitter(ConstrainedCost(nets, VertexCount, ceil(1.5 n / K))), banded n = 2^14",
                       cp.ConvexTotalSplitter(mdl), None, matrix=B, dmatrix=dB, k=k)
            if k == 8:  # the "dynamic" row of the same table
                split_case("dynamic_constrained: DynamicTotalSplitter(ConstrainedCost(nets, VertexCount, ceil(1.5 n / K))), banded n = 2^14",
                           cp.DynamicTotalSplitter(mdl), None, matrix=B, dmatrix=dB, k=k)
        dB.close()
    if "map_objective" in cases:
        rng = np.random.default_rng(1)
        Phi = cp.MapPartition(K, rng.integers(1, K + 1, A.n))
        cp.bottleneck_value(dA, Phi, net)
        t_res, g = timed(lambda: cp.bottleneck_value(dA, Phi, net), args.reps)
        t0 = time.perf_counter()
        r = ref.bottleneck_value(A, Phi, net)
        t_cpu = (time.perf_counter() - t0) * 1e3
        emit({"case": "map_objective: bottleneck_value(A, MapPartition(K = 64), connectivity)", "n": A.n, "nnz": A.nnz, "gpu_resident_ms": t_res,
              "cpu_oracle_ms": t_cpu, "identical": bool(g == r), "note": "the CPU side permutes the columns in numpy, then the step oracle"})
    if "prefix" in cases:
        rng = np.random.default_rng(2)
        val = rng.integers(0, 2**64, A.nnz, dtype=np.uint64)
        Q = 1 << 20
        i, j = rng.integers(1, A.m + 2, Q), rng.integers(1, A.n + 2, Q)
        C0 = cp.dominancecount(A); C0.close()
        t_bc, C = timed(lambda: cp.dominancecount(A), 1)
        t_bs, S = timed(lambda: cp.dominancesum(A, val), 1)
        C.query(i, j)
        t_qc, gc = timed(lambda: C.query(i, j), args.reps)
        t_qs, gs = timed(lambda: S.query(i, j), args.reps)
        t0 = time.perf_counter()
        rc = ref.dominancecount(A, i, j, hint=cp.SparseHint())
        t_cpu_c = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        rs = ref.prefix_query(A.m, A.n, A.nnz, A.colptr, A.rowval, val, i, j)
        t_cpu_s = (time.perf_counter() - t0) * 1e3
        emit({"case": "prefix: dominancecount / dominancesum, 2^20 uniform queries", "n": A.n, "nnz": A.nnz, "Q": Q,
              "gpu_build_count_ms": t_bc, "gpu_build_sum_ms": t_bs, "gpu_query_count_ms": t_qc, "gpu_query_sum_ms": t_qs,
              "cpu_count_build_plus_queries_ms": t_cpu_c, "cpu_sum_sweep_ms": t_cpu_s,
              "identical": bool(np.array_equal(gc, rc) and np.array_equal(gs, rs)),
              "note": "gpu times include the upload of the host arrays and the download of the answers; CPU count = the restated b-ary DominanceCount "
                      "(SparseHint), CPU sum = the oracle's offline Fenwick sweep (not the reference's layout)"})
        C.close(); S.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"next_{args.tag}.jsonl"), "w") as fh:
        for line in lines:
            fh.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
