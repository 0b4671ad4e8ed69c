"""Where does the host time of pack_stripe (C4a) go?  cProfile of the Python mirror around the C call."""
import os, sys, time, cProfile, pstats
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
dX = cp.device_matrix(X)
for rep in range(2):
    cp.pack_stripe(dX, m1, Pi)
pr = cProfile.Profile()
cp.synchronize(); t0 = time.perf_counter()
pr.enable(); r = cp.pack_stripe(dX, m1, Pi); pr.disable()
cp.synchronize(); print("C4a call", round((time.perf_counter() - t0) * 1e3, 2), "ms (under cProfile)")
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
