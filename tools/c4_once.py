import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np
import chainb200 as cp
from chainb200 import synth_torch
cp.init(0)
n = 1 << 22
A = synth_torch.banded(n, 64)
X = cp.adjointpattern(A)
Pi = cp.pack_stripe(A, cp.EquiChunker(4))
blk = cp.BlockComponentCostModel(int, 1, 3, (1, cp.identity), (1, cp.identity))
m1 = cp.DynamicTotalChunker(blk, 8)
m2 = cp.ConvexTotalChunker(cp.ConstrainedCost(cp.AffineConnectivityModel(0, 0, 0, 1), cp.VertexCount(), 8))
dX = cp.device_matrix(X)
for name, call in (("C4a", lambda: cp.pack_stripe(dX, m1, Pi)), ("C4b", lambda: cp.pack_stripe(dX, m2))):
    for rep in range(2):
        cp.synchronize(); t0 = time.perf_counter(); r = call(); cp.synchronize(); print(name, round((time.perf_counter() - t0) * 1e3, 2), "ms", flush=True)
