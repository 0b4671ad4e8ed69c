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
out = {}
for wmin in ("0", "33554432", "0", "33554432"):
    os.environ["CPB_WINDOWED_SCATTER_MIN"] = wmin
    cp.partition_stripe(dA, 1024, mtd)
    ts = []
    for rep in range(3):
        cp.synchronize(); t0 = time.perf_counter(); Phi = cp.partition_stripe(dA, 1024, mtd); cp.synchronize(); ts.append(time.perf_counter() - t0)
    cp.profile_enable(True); cp.profile_reset()
    Phi = cp.partition_stripe(dA, 1024, mtd); cp.synchronize()
    prof = cp.profile_get(); cp.profile_enable(False)
    out[wmin] = Phi.spl.copy()
    print("windowed_min", wmin, "nnz", A.nnz, "resident ms", ["%.1f" % (t * 1e3) for t in ts], {k: round(v["ms"], 2) for k, v in prof.items() if k in ("k_link_prev", "k_rs_scatter", "k_rs_hist", "build_links", "k_probe_stream", "k_lt_count")}, flush=True)
assert (out["0"] == out["33554432"]).all()
print("identical split vectors")
