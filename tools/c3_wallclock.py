import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import chainb200 as cp
from chainb200 import synth_torch
cp.init(0)
AFF = cp.AffineConnectivityModel(0, 10, 1, 100)
A = synth_torch.rmat(24, 16 << 24)
dA = cp.device_matrix(A)
mtd = cp.LazyBisectCostBottleneckSplitter(AFF, 0.01)
cp.partition_stripe(dA, 1024, mtd)
ts = []
for rep in range(40):
    cp.synchronize(); t0 = time.perf_counter(); Phi = cp.partition_stripe(dA, 1024, mtd); cp.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
ts = np.array(ts)
print("fence", os.environ.get("CPB_STAGE_FENCE", "1"), "40 reps: min %.1f median %.1f max %.1f ms, >40 ms: %d" % (ts.min(), np.median(ts), ts.max(), (ts > 40).sum()), np.round(ts, 1).tolist(), flush=True)
