"""C3-sized link construction only (for ncu): python tools/onesweep_prof.py [scale]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import chainb200 as cp
from chainb200 import synth_torch
cp.init(0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 24
A = synth_torch.rmat(scale, 16 << scale)
dA = cp.device_matrix(A)
ocl = cp.oracle_stripe(cp.AffineConnectivityModel(0, 10, 1, 100), dA)
buf = torch.zeros(A.nnz + A.n + 8, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for rep in range(2):
    ocl.links_partial(1, A.m + 1, buf.data_ptr())
    cp.synchronize()
print("done")
