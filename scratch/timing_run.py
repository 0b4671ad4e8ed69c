import os, sys, shutil
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import chainb200 as cp
from chainb200 import synth_torch
cp.init(0)
A = synth_torch.erdos_renyi(1_000_000, 10)
dA = cp.device_matrix(A)
mtd = cp.BisectCostBottleneckSplitter(cp.AffineConnectivityModel(0, 10, 1, 100), 0.01)
for i in range(3):
    cp.partition_stripe(dA, 64, mtd)
os.environ["CPB_PROBE_TIMING_DUMP"] = "1"
cp.partition_stripe(dA, 64, mtd)
