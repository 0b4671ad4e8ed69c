"""a23 on the device beside the CPU oracle: the queue routine is sequential on both (profiles/r02_session2.md)."""
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np
import chainb200 as cp
import pyoracle as ref
from chainb200 import synth
cp.init(0)
f = cp.AffineConnectivityModel(0, 10, 1, 100)
tab = cp.ColumnBlockComponentCostModel(int, 7, lambda w: int(20 * np.sqrt(w)))
for n in (2000, 20000):
    A = synth.erdos_renyi(n, 8)
    for name, call in (("ConcaveTotalChunker(connectivity)", lambda impl: impl.pack_stripe(A, cp.ConcaveTotalChunker(f))),
                       ("ConcaveTotalChunker(column-block, concave beta)", lambda impl: impl.pack_stripe(A, cp.ConcaveTotalChunker(tab))),
                       ("ConcaveTotalSplitter(connectivity), K = 4", lambda impl: impl.partition_stripe(A, 4, cp.ConcaveTotalSplitter(f)))):
        call(cp)
        t0 = time.perf_counter(); g = call(cp); tg = time.perf_counter() - t0
        t0 = time.perf_counter(); r = call(ref); tr = time.perf_counter() - t0
        print(f"n = {n}: {name}: device {tg * 1e3:.1f} ms, CPU oracle {tr * 1e3:.1f} ms, identical {bool(np.array_equal(g.spl, r.spl))}", flush=True)
