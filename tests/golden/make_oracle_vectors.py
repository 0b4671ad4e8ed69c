"""Regenerates tests/golden/oracle_vectors.json: split vectors of the CPU oracle on the reference's six fixture matrices.

The reference has no golden outputs (its tests are property tests) and Julia is not available, so these vectors
are NOT reference outputs: they pin the oracle against itself -- a regression guard for refactors of oracle/ (every vector
here was produced after the oracle passed the reference's property tests, and the device results equal them bit for bit).

Run:  python tests/golden/make_oracle_vectors.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import chainb200 as cp  # noqa: E402
import pyoracle as ref  # noqa: E402
from helpers import load_fixtures  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.json")


def cases(A):
    """(label, callable) pairs; labels are stable identifiers"""
    net = cp.AffineConnectivityModel(0, 10, 1, 100)
    out = []
    for K in (2, 5, 9):
        Pi = ref.partition_stripe(ref.adjointpattern(A), K, cp.EquiSplitter())
        mt = [("DynamicBottleneck", cp.DynamicBottleneckSplitter(net), None), ("DynamicTotal", cp.DynamicTotalSplitter(net), None),
              ("BisectCost_0.01", cp.BisectCostBottleneckSplitter(net, 0.01), None), ("LazyBisectCost_0.1", cp.LazyBisectCostBottleneckSplitter(net, 0.1), None),
              ("BisectIndex", cp.BisectIndexBottleneckSplitter(net), None), ("ConvexTotal", cp.ConvexTotalSplitter(net), None),
              ("ConvexTotal_w6", cp.ConvexTotalSplitter(cp.ConstrainedCost(cp.AffineConnectivityModel(0, 0, 0, 1), cp.VertexCount(), max(6, -(-A.n // K) + 2))), None),
              ("DynamicTotal_w6", cp.DynamicTotalSplitter(cp.ConstrainedCost(cp.AffineConnectivityModel(0, 0, 0, 1), cp.VertexCount(), max(6, -(-A.n // K) + 2))), None),
              ("PrimaryConn_BisectCost", cp.BisectCostBottleneckSplitter(cp.AffinePrimaryConnectivityModel(0, 2, 1, 3, 6), 0.01), Pi),
              ("SecondaryConn_FlipBisectIndex", cp.FlipBisectIndexBottleneckSplitter(cp.AffineSecondaryConnectivityModel(0, 2, 1, 3, 6)), Pi),
              ("PrimaryEdge_LazyBisect", cp.LazyBisectCostBottleneckSplitter(cp.AffinePrimaryEdgeCutModel(0, 2, 1, 5), 0.01), Pi),
              ("SecondaryEdge_FlipBisectCost", cp.FlipBisectCostBottleneckSplitter(cp.AffineSecondaryEdgeCutModel(0, 2, 1, 5), 0.01), Pi)]
        if A.m == A.n:
            sym = cp.AffineMonotonizedSymmetricConnectivityModel(0, 0, 1, 100, 2)
            mt.append(("MonoSym_LazyBisect_0.1", cp.LazyBisectCostBottleneckSplitter(sym, 0.1), None))
        for label, mtd, pi in mt:
            out.append((f"partition_stripe/{label}/K={K}", (lambda mtd=mtd, pi=pi, K=K: ref.partition_stripe(A, K, mtd, pi) if pi is not None else ref.partition_stripe(A, K, mtd))))
    Pi4 = ref.pack_stripe(ref.adjointpattern(A), cp.EquiChunker(4))
    blk = cp.BlockComponentCostModel(int, 1, 3, (1, cp.identity), (1, cp.identity))
    packs = [("DynamicTotalChunker_block_w8", cp.DynamicTotalChunker(cp.ConstrainedCost(blk, cp.VertexCount(), 8)), Pi4),
             ("ConvexTotalChunker_nets_w8", cp.ConvexTotalChunker(cp.ConstrainedCost(cp.AffineConnectivityModel(0, 0, 0, 1), cp.VertexCount(), 8)), None),
             ("OverlapChunker_0.9_8", cp.OverlapChunker(0.9, 8), None), ("StrictChunker_8", cp.StrictChunker(8), None)]
    for label, mtd, pi in packs:
        out.append((f"pack_stripe/{label}", (lambda mtd=mtd, pi=pi: ref.pack_stripe(A, mtd, pi) if pi is not None else ref.pack_stripe(A, mtd))))
    return out


def generate():
    vectors = {}
    for name, A in load_fixtures().items():
        for label, fn in cases(A):
            vectors[f"{name}|{label}"] = [int(x) for x in fn().spl]
    return vectors


if __name__ == "__main__":
    v = generate()
    with open(OUT, "w") as fh:
        json.dump(v, fh, indent=0, sort_keys=True)
    print(len(v), "vectors ->", OUT)
