import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import chainb200 as cp
from chainb200 import synth_torch
cp.init(0)
f = cp.AffineConnectivityModel(0, 10, 1, 100)
A = synth_torch.rmat(24, 16 << 24)
dA = cp.device_matrix(A)
for rep in range(4):
    cp.profile_enable(True); cp.profile_reset()
    t0 = time.perf_counter()
    o = cp.oracle_stripe(f, dA); o.query(np.array([1]), np.array([A.n + 1])); o.close()
    cp.synchronize(); dt = time.perf_counter() - t0
    p = cp.profile_get(); cp.profile_enable(False)
    k = p["k_wm_level"]
    print("rep", rep, "wall ms %.1f" % (dt * 1e3), "k_wm_level: %.1f us/launch, %.0f GB/s" % (1e3 * k["ms"] / k["launches"], k["bytes"] / 1e9 / (k["ms"] / 1e3)), {n: round(v["ms"], 2) for n, v in p.items()}, flush=True)
