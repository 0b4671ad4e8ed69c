import os, sys, json, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import chainb200 as cp
from chainb200 import synth_torch
cp.init(0)
AFF = cp.AffineConnectivityModel(0, 10, 1, 100)
A = synth_torch.rmat(24, 16 << 24)
dA = cp.device_matrix(A)
mtd = cp.LazyBisectCostBottleneckSplitter(AFF, 0.01)
ref = None
for plan, depth in [(1, 4), (0, 4), (0, 3), (0, 2), (0, 1), (1, 3), (1, 2)]:
    os.environ["CPB_BISECT_PLAN"] = str(plan); os.environ["CPB_BISECT_DEPTH"] = str(depth)
    cp.partition_stripe(dA, 1024, mtd)
    cp.profile_enable(True); cp.profile_reset()
    t0 = time.perf_counter(); Phi = cp.partition_stripe(dA, 1024, mtd); cp.synchronize(); dt = time.perf_counter() - t0
    prof = cp.profile_get(); cp.profile_enable(False)
    st = cp.bisect_stats()
    if ref is None: ref = Phi.spl.copy()
    assert (ref == Phi.spl).all()
    k = prof.get("k_probe_stream", {"ms": 0, "launches": 0})
    print(f"plan={plan} P={2**depth-1} total_ms={dt*1e3:.1f} probe_ms={k['ms']:.1f} launches={k['launches']} ms/launch={k['ms']/max(k['launches'],1):.2f} speculated={st['speculated']} rounds={st['rounds']}", flush=True)
