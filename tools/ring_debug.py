#!/usr/bin/env python
"""One large case per process with the ring probes on: prints the split-vector checksum or the library's error message
(which carries the record a timed-out mbarrier wait leaves).  Usage: python tools/ring_debug.py <case>"""
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np

import chainb200 as cp
from chainb200 import synth, synth_torch

AFF = cp.AffineConnectivityModel(0, 10, 1, 100)
sym = cp.AffineMonotonizedSymmetricConnectivityModel(0, 0, 1, 100, 4)
case = sys.argv[1]
cp.init(0)
if case == "rmat20":
    A, f, K, eps = synth_torch.rmat(20, 16 << 20), AFF, 256, 0.01
elif case == "rmat18":
    A, f, K, eps = synth_torch.rmat(18, 16 << 18), AFF, 256, 0.01
elif case == "er1m":
    A, f, K, eps = synth_torch.erdos_renyi(1_000_000, 10), AFF, 64, 0.01
elif case == "rgg":
    A, f, K, eps = synth_torch.random_geometric(1 << 20), sym, 128, 0.01
elif case == "lap512":
    A, f, K, eps = synth.laplacian5(512), AFF, 100, 0.001
elif case == "c3":
    A, f, K, eps = synth_torch.rmat(24, 16 << 24), AFF, 1024, 0.01
else:
    raise SystemExit("unknown case")
dA = cp.device_matrix(A)
mtd = cp.LazyBisectCostBottleneckSplitter(f, eps)
out = {}
for ring in (("1",) if os.environ.get("CPB_RINGDBG_ONLY") else ("0", "1")):
    os.environ["CPB_PROBE_RING"] = ring
    try:
        cp.partition_stripe(dA, K, mtd)
        cp.synchronize()
        t0 = time.perf_counter()
        r = cp.partition_stripe(dA, K, mtd)
        cp.synchronize()
        out[ring] = r.spl
        print(case, "ring", ring, "ok", round((time.perf_counter() - t0) * 1e3, 2), "ms", cp.bisect_stats(), flush=True)
    except Exception as e:
        print(case, "ring", ring, "FAILED:", e, flush=True)
        sys.exit(3)
if "0" in out:
    print(case, "identical", bool(np.array_equal(out["0"], out["1"])))
