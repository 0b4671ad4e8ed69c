"""Upload time of C3-sized Int64 index arrays from pageable memory vs the number of narrowing threads (CPB_H2D_THREADS)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import chainb200 as cp
cp.init(0)
n = 1 << 24
nnz = 263_000_000
rng = np.random.default_rng(0)
deg = np.full(n, nnz // n, dtype=np.int64); deg[: nnz - deg.sum()] += 1
colptr = np.concatenate([[1], 1 + np.cumsum(deg)]).astype(np.int64)
rowval = np.empty(nnz, dtype=np.int64)
base = rng.integers(1, n + 1, 1 << 20, dtype=np.int64)
for o in range(0, nnz, 1 << 20):
    c = min(1 << 20, nnz - o)
    rowval[o:o + c] = base[:c]
# rows inside a column need not be sorted for the upload path (only range-checked)
A = cp.SparseMatrixCSC(n, n, colptr, rowval)
print("cores", os.cpu_count(), "bytes", colptr.nbytes + rowval.nbytes, flush=True)
for T in [int(x) for x in os.environ.get("THREADS", "12,16,12,16,12,16").split(",")]:
    os.environ["CPB_H2D_THREADS"] = str(T)
    ts = []
    for rep in range(8):
        cp.synchronize(); t0 = time.perf_counter()
        dA = cp.device_matrix(A); cp.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
        dA.close()
    print("threads", T, "upload ms", ["%.1f" % t for t in ts], flush=True)
