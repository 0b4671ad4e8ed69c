"""Is the wall-clock spread of back-to-back C3-size solves CFS throttling of the (spinning) host thread?  Prints the cgroup's CPU
quota and the throttling counters around 30 solves.  Run from the repo root on a GPU box."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import chainb200 as cp
from chainb200 import synth_torch


def read(path):
    try:
        return open(path).read().strip().replace("\n", " | ")
    except Exception as e:
        return "n/a (%s)" % type(e).__name__


def stat():
    for p in ("/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/cpu/cpu.stat"):
        if os.path.exists(p):
            return read(p)
    return "n/a"


print("cpu.max:", read("/sys/fs/cgroup/cpu.max"), "| cfs_quota_us:", read("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"), "| nproc:", os.cpu_count(),
      "| affinity:", len(os.sched_getaffinity(0)), flush=True)
cp.init(0)
A = synth_torch.rmat(24, 16 << 24)
dA = cp.device_matrix(A)
mtd = cp.LazyBisectCostBottleneckSplitter(cp.AffineConnectivityModel(0, 10, 1, 100), 0.01)
cp.partition_stripe(dA, 1024, mtd)
print("before:", stat(), flush=True)
ts = []
for rep in range(30):
    cp.synchronize(); t0 = time.perf_counter(); cp.partition_stripe(dA, 1024, mtd); cp.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("after: ", stat(), flush=True)
ts = np.array(ts)
print("30 solves: min %.1f median %.1f max %.1f ms" % (ts.min(), np.median(ts), ts.max()), np.round(ts, 1).tolist())
