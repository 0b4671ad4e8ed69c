#!/usr/bin/env python
"""bench_configs.py -- all five BASELINE.json configs on one B200, each beside the CPU oracle.

For every config: generate the synthetic matrix (GPU generators, same hash as synth.py), run the
public call with HOST arrays (e2e: H2D + oracle construction + search + split vector D2H), run it
again with the pattern resident in HBM, time the single-threaded CPU restatement of the reference on
the same input (full size unless --cpu-scale shrinks it), and compare the split vectors bit for bit.
Writes one JSON line per config to stdout and gpurun_out/configs_<tag>.jsonl.

This is NOT the driver's bench (that is bench.py, config 2); it is the evidence for "bit-exact on all
five configs + end-to-end time next to the CPU path".
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

AFF = cp.AffineConnectivityModel(0, 10, 1, 100)


def timed(fn, reps=1):
    best = None
    out = None
    for _ in range(reps):
        cp.synchronize()
        t0 = time.perf_counter()
        out = fn()
        cp.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out


def run_config(name, A, gpu_call, cpu_call, cmp, reps, extra=None, skip_cpu=False):
    dA = cp.device_matrix(A)
    gpu_call(A)  # warm-up (pool growth, module load)
    t_e2e, out_e2e = timed(lambda: gpu_call(A), reps)
    cp.profile_enable(True)
    cp.profile_reset()
    t_res, out_res = timed(lambda: gpu_call(dA), reps)
    prof = cp.profile_get()
    cp.profile_enable(False)
    line = {"config": name, "m": A.m, "n": A.n, "nnz": A.nnz, "gpu_e2e_ms": t_e2e * 1e3, "gpu_resident_ms": t_res * 1e3,
            "h2d_bytes": (A.nnz + A.n + 1) * 8,
            "phases_ms": {k: round(v["ms"] / reps, 3) for k, v in prof.items()},
            "phase_gbs": {k: round(v["bytes"] / 1e9 / (v["ms"] / 1e3), 1) for k, v in prof.items() if v["ms"] > 0 and v["bytes"] > 0}}
    if not skip_cpu:
        t0 = time.perf_counter()
        out_cpu = cpu_call(A)
        line["cpu_oracle_ms"] = (time.perf_counter() - t0) * 1e3
        line["cpu_phases_s"] = list(ref.last_seconds)
        line["identical"] = bool(cmp(out_e2e, out_cpu) and cmp(out_res, out_cpu))
        line["speedup_e2e"] = line["cpu_oracle_ms"] / line["gpu_e2e_ms"]
    try:
        line["bisection"] = cp.bisect_stats()
    except Exception:
        pass
    if extra:
        line.update(extra(out_e2e))
    dA.close()
    return line


def same_split(a, b):
    return np.array_equal(a.spl, b.spl)


def same_pair(a, b):
    return np.array_equal(a[0].spl, b[0].spl) and np.array_equal(a[1].spl, b[1].spl)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every config (1.0 = BASELINE sizes)")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--skip-cpu", default="")
    args = ap.parse_args()
    cp.init(0)
    want = set(args.configs.split(","))
    skip_cpu = set(args.skip_cpu.split(",")) if args.skip_cpu else set()
    lines = []
    sc = args.scale

    if "1" in want:
        g = max(8, int(round(256 * np.sqrt(sc))))
        A = synth.laplacian5(g)
        mtd = cp.DynamicBottleneckSplitter(AFF)
        lines.append(run_config(f"C1 laplacian {g}x{g}, K=8, DynamicBottleneckSplitter", A, lambda M: cp.partition_stripe(M, 8, mtd),
                                lambda M: ref.partition_stripe(M, 8, mtd), same_split, args.reps, skip_cpu="1" in skip_cpu))
        print(json.dumps(lines[-1]), flush=True)
    if "2" in want:
        n = max(1000, int(1_000_000 * sc))
        A = synth_torch.erdos_renyi(n, 10)
        mtd = cp.BisectCostBottleneckSplitter(AFF, 0.01)
        lines.append(run_config(f"C2 Erdos-Renyi n={n}, K=64, BisectCost eps=0.01", A, lambda M: cp.partition_stripe(M, 64, mtd),
                                lambda M: ref.partition_stripe(M, 64, mtd), same_split, args.reps, skip_cpu="2" in skip_cpu))
        print(json.dumps(lines[-1]), flush=True)
    if "3" in want:
        scale = max(10, int(round(24 + np.log2(sc)))) if sc < 1 else 24
        A = synth_torch.rmat(scale, 16 << scale)
        K = min(1024, max(2, A.n // 64))
        mtd = cp.LazyBisectCostBottleneckSplitter(AFF, 0.01)
        lines.append(run_config(f"C3 R-MAT scale {scale}, K={K}, LazyBisectCost eps=0.01", A, lambda M: cp.partition_stripe(M, K, mtd),
                                lambda M: ref.partition_stripe(M, K, mtd), same_split, args.reps, skip_cpu="3" in skip_cpu))
        print(json.dumps(lines[-1]), flush=True)
    if "4" in want:
        n = max(1024, int((1 << 22) * sc))
        A = synth_torch.banded(n, 64)
        X = cp.adjointpattern(A)
        Pi = cp.pack_stripe(A, cp.EquiChunker(4))
        blk = cp.BlockComponentCostModel(int, 1, 3, (1, cp.identity), (1, cp.identity))
        m1 = cp.DynamicTotalChunker(blk, 8)
        m2 = cp.ConvexTotalChunker(cp.ConstrainedCost(cp.AffineConnectivityModel(0, 0, 0, 1), cp.VertexCount(), 8))
        lines.append(run_config(f"C4a banded n={n} bw=64, pack_stripe DynamicTotalChunker(block_model, 8)", X, lambda M: cp.pack_stripe(M, m1, Pi),
                                lambda M: ref.pack_stripe(M, m1, Pi), same_split, args.reps, extra=lambda o: {"chunks": int(o.K)}, skip_cpu="4" in skip_cpu))
        print(json.dumps(lines[-1]), flush=True)
        lines.append(run_config(f"C4b banded n={n} bw=64, pack_stripe ConvexTotalChunker(connectivity, 8)", X, lambda M: cp.pack_stripe(M, m2),
                                lambda M: ref.pack_stripe(M, m2), same_split, args.reps, extra=lambda o: {"chunks": int(o.K)}, skip_cpu="4" in skip_cpu))
        print(json.dumps(lines[-1]), flush=True)
    if "5" in want:
        n = max(1024, int((1 << 23) * sc))
        A = synth_torch.random_geometric(n)
        s = cp.AffineMonotonizedSymmetricConnectivityModel(0, 0, 1, 100, 90)
        mtd = cp.LazyBisectCostBottleneckSplitter(s, 0.1)
        K = min(256, max(2, n // 64))
        meth = cp.AlternatingPartitioner(mtd, mtd)
        lines.append(run_config(f"C5 random-geometric n={n}, K={K}, partition_plaid AlternatingPartitioner(LazyBisect(sym, 0.1) x2)", A,
                                lambda M: cp.partition_plaid(M, K, meth), lambda M: ref.partition_plaid(M, K, meth), same_pair, args.reps,
                                skip_cpu="5" in skip_cpu))
        print(json.dumps(lines[-1]), flush=True)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"configs_{args.tag}.jsonl"), "w") as fh:
        for l in lines:
            fh.write(json.dumps(l) + "\n")
    bad = [l["config"] for l in lines if l.get("identical") is False]
    if bad:
        print("MISMATCH:", bad)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
