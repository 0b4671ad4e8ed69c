"""GPU check of the one-sweep link construction (csrc/onesweep.cu) against the round-1 tile-histogram sort and a numpy
restatement of SparseColorArrays.jl:103-118 (last-seen sweep), plus timings at C3.  Usage: python tools/onesweep_check.py [--c3]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import torch
import chainb200 as cp
from chainb200 import synth, synth_torch

cp.init(0)
AFF = cp.AffineConnectivityModel(0, 10, 1, 100)


def links_numpy(A):
    """prev[q] = 1 + position of the previous nonzero of the same row (0: none) -- stable argsort by row"""
    row = np.asarray(A.rowval) - 1
    order = np.argsort(row, kind="stable")
    prev = np.zeros(len(row), dtype=np.int64)
    same = row[order[1:]] == row[order[:-1]]
    prev[order[1:][same]] = order[:-1][same] + 1
    return prev


def device_links(dA, A):
    ocl = cp.oracle_stripe(AFF, dA)
    buf = torch.zeros(A.nnz + A.n + 8, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ne = ocl.links_partial(1, A.m + 1, buf.data_ptr())
    cp.synchronize()
    out = buf[:ne].cpu().numpy().astype(np.int64) & 0xffffffff
    ocl.close()
    return out


def check(name, A, with_numpy=True):
    dA = cp.device_matrix(A)
    res = {}
    for os_on in ("0", "1"):
        for wmin in ("0", "1"):
            os.environ["CPB_ONESWEEP"] = os_on
            os.environ["CPB_NO_ROW_SEGMENTS"] = "1"
            os.environ["CPB_WINDOWED_SCATTER_MIN"] = wmin
            res[(os_on, wmin)] = device_links(dA, A)
    base = res[("0", "0")]
    ok = all(np.array_equal(base, v) for v in res.values())
    if with_numpy:
        ok = ok and np.array_equal(base, links_numpy(A))
    print(name, "nnz", A.nnz, "identical" if ok else "DIFFERENT", flush=True)
    for k in ("CPB_ONESWEEP", "CPB_NO_ROW_SEGMENTS", "CPB_WINDOWED_SCATTER_MIN"):
        os.environ.pop(k, None)
    assert ok, name
    return dA


rng = np.random.default_rng(7)
check("empty", cp.SparseMatrixCSC(5, 4, [1, 1, 1, 1, 1], np.zeros(0, dtype=np.int64)))
check("one", cp.SparseMatrixCSC(3, 2, [1, 2, 2], np.array([2])))
check("laplacian 64", synth.laplacian5(64))
check("er 20000", synth.erdos_renyi(20000, 10))
# one very heavy row + tile-boundary sizes
for n in (4095, 4096, 4097, 12289):
    rows = [np.unique(np.concatenate([rng.integers(1, 50, 3), [7]])) for _ in range(n)]
    check("heavy row n=%d" % n, cp.SparseMatrixCSC(50, n, np.cumsum([1] + [len(r) for r in rows]), np.concatenate(rows)))
check("rmat 16", synth_torch.rmat(16, 16 << 16))
check("rmat 20", synth_torch.rmat(20, 16 << 20))

if "--c3" in sys.argv:
    A = synth_torch.rmat(24, 16 << 24)
    dA = check("rmat 24 (C3)", A, with_numpy=False)
    mtd = cp.LazyBisectCostBottleneckSplitter(AFF, 0.01)
    spl = {}
    for os_on in ("0", "1", "0", "1"):
        os.environ["CPB_ONESWEEP"] = os_on
        cp.partition_stripe(dA, 1024, mtd)
        ts = []
        for rep in range(3):
            cp.synchronize(); t0 = time.perf_counter(); Phi = cp.partition_stripe(dA, 1024, mtd); cp.synchronize(); ts.append(time.perf_counter() - t0)
        cp.profile_enable(True); cp.profile_reset()
        Phi = cp.partition_stripe(dA, 1024, mtd); cp.synchronize()
        prof = cp.profile_get(); cp.profile_enable(False)
        spl[os_on] = Phi.spl.copy()
        keys = ("k_expand_columns", "k_os_hist", "k_os_pass", "k_link_prev", "k_rs_scatter", "k_rs_hist", "build_links", "k_probe_stream")
        print("onesweep", os_on, "resident ms", ["%.1f" % (t * 1e3) for t in ts], {k: round(v["ms"], 2) for k, v in prof.items() if k in keys}, flush=True)
    assert (spl["0"] == spl["1"]).all()
    print("identical split vectors")
if "--er" in sys.argv:
    # uniform digits, forced through the sort path
    A = synth_torch.erdos_renyi(8_000_000, 16)
    dA = check("er 8M x 16 (sort path forced)", A, with_numpy=False)
    os.environ["CPB_NO_ROW_SEGMENTS"] = "1"
    ocl = cp.oracle_stripe(AFF, dA)
    buf = torch.zeros(A.nnz + A.n + 8, dtype=torch.int32, device="cuda")
    for os_on in ("0", "1", "0", "1"):
        os.environ["CPB_ONESWEEP"] = os_on
        ocl.links_partial(1, A.m + 1, buf.data_ptr()); cp.synchronize()
        cp.profile_enable(True); cp.profile_reset()
        ocl.links_partial(1, A.m + 1, buf.data_ptr()); cp.synchronize()
        prof = cp.profile_get(); cp.profile_enable(False)
        keys = ("k_expand_columns", "k_os_hist", "k_os_pass", "k_link_prev", "k_rs_scatter", "k_rs_hist", "build_links")
        print("er onesweep", os_on, "nnz", A.nnz, {k: round(v["ms"], 2) for k, v in prof.items() if k in keys}, flush=True)
    os.environ.pop("CPB_NO_ROW_SEGMENTS"); os.environ.pop("CPB_ONESWEEP")
print("onesweep check ok")
