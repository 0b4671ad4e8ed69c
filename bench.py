#!/usr/bin/env python
"""bench.py -- the hot path of ChainPartitioners.jl on B200, measured per the driver contract.

Workload (N = 1): BASELINE.json configs[1] -- Erdos-Renyi 1M x 1M, 10 nnz/column, K = 64,
``partition_stripe(A, 64, BisectCostBottleneckSplitter(AffineConnectivityModel(0,10,1,100), 0.01))``.
A "step" is one full partition_stripe: oracle construction (link building + dominance index) and
the bisection probes.  ``value`` times it with the CSC pattern already resident in HBM; ``e2e``
times the public call with HOST (pinned) colptr/rowval, host->device copies inside the timed region
and the split vector read back.  N > 1: one process per GPU, each partitioning its own matrix of
the same shape (independent problems, no data-path collective) -> weak scaling.

``--impl reference`` times the reference's CPU algorithm for the same call on the host cores: the
C++ restatement in oracle/ (Julia is not installed in this image, so the reference itself cannot run;
kind = "port", single-threaded because the reference is).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_COLS = int(os.environ.get("CPB_BENCH_N", 1_000_000))
NNZ_PER_COL = 10
K_PARTS = 64
EPS = 0.01
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` at this workload
# (profiles/r01_ncu_summary.md: k_probe_stream 45.0 MB read + 0.2 MB write; k_lt_fill 79.6 + 37.9 MB; k_lt_link
# 76.3 + 36.2 MB; k_lt_count 44.0 MB read; k_rs_scatter 42-82 MB read + 52-56 MB write; k_wm_level 40.05 + 1-3 MB), bytes
# dram read + write per launch from the ncu --set full captures (profiles/r01_ncu_summary.md sections 2 and 6)
TRAFFIC = {"k_probe_stream": 46.8e6, "k_lt_fill": 118.7e6, "k_lt_link": 113.9e6, "k_lt_count": 44.0e6, "k_rs_scatter": 122e6, "k_wm_level": 42e6,
           "k_expand_columns": 4.1e6}
METRIC = "partition_stripe throughput (BisectCostBottleneckSplitter, Erdos-Renyi 1Mx1M, K=64)"
UNIT = "partitions/s"


def workload():
    import chainb200 as cp
    from chainb200 import synth

    cache = os.path.join("/tmp", f"cpb_er_{N_COLS}_{NNZ_PER_COL}.npz")
    A = None
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            A = cp.SparseMatrixCSC(N_COLS, N_COLS, z["colptr"], z["rowval"])
        except Exception:
            A = None
    if A is None:
        A = synth.erdos_renyi(N_COLS, NNZ_PER_COL)
        try:  # atomic publish: concurrent ranks may generate the same matrix
            tmp = f"{cache}.{os.getpid()}.tmp.npz"
            np.savez(tmp, colptr=A.colptr, rowval=A.rowval)
            os.replace(tmp, cache)
        except OSError:
            pass
    f = cp.AffineConnectivityModel(0, 10, 1, 100)
    return A, f, cp.BisectCostBottleneckSplitter(f, EPS)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_run(A, mtd, steps, warmup):
    import pyoracle as ref

    for _ in range(warmup):
        ref.partition_stripe(A, K_PARTS, mtd)
    t0 = time.perf_counter()
    secs = [0.0, 0.0]
    for _ in range(steps):
        Phi = ref.partition_stripe(A, K_PARTS, mtd)
        secs[0] += ref.last_seconds[0]
        secs[1] += ref.last_seconds[1]
    dt = time.perf_counter() - t0
    return dt / steps, [s / steps for s in secs], Phi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    config = {"workload": f"configs[1]: Erdos-Renyi {N_COLS}x{N_COLS}, {NNZ_PER_COL} nnz/col, K={K_PARTS}, BisectCostBottleneckSplitter eps={EPS}, "
                          "AffineConnectivityModel(0,10,1,100)",
              "l2": "flushed between timed steps (256 MiB write)", "per_gpu": "one independent matrix per GPU"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        A, f, mtd = workload()
        steps = max(1, min(args.steps, 20))
        per, secs, _ = cpu_reference_run(A, mtd, steps, min(args.warmup, 1))
        v = 1.0 / per
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
                "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                 "sample": f"full workload x{steps}; oracle build {secs[0]*1e3:.1f} ms + bisection {secs[1]*1e3:.1f} ms per step; "
                                           "C++ restatement of the Julia reference (b-ary DominanceCount + windowed binary searches), 1 thread"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import chainb200 as cp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    cp.init(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        cp.synchronize()

    A, f, mtd = workload()
    # pinned host copies of the Julia-layout arrays (the e2e inputs)
    colptr_pin = torch.from_numpy(A.colptr).pin_memory()
    rowval_pin = torch.from_numpy(A.rowval).pin_memory()
    A_pin = cp.SparseMatrixCSC(A.m, A.n, colptr_pin.numpy(), rowval_pin.numpy())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush.zero_()
        torch.cuda.synchronize()

    dA = cp.device_matrix(A_pin)

    def step_resident():
        return cp.partition_stripe(dA, K_PARTS, mtd)

    def step_e2e():
        return cp.partition_stripe(A_pin, K_PARTS, mtd)

    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- value: inputs resident in HBM ----
    cp.profile_enable(True)
    cp.profile_reset()
    barrier()
    dev_ms = 0.0
    for _ in range(args.steps):
        flush_l2()
        cp.timer_start()
        Phi = step_resident()
        dev_ms += cp.timer_stop()
    barrier()
    launches = cp.launch_count()
    prof = cp.profile_get()
    cp.profile_enable(False)

    # ---- e2e: host buffers through the public call ----
    barrier()
    e2e_ms = 0.0
    e2e_samples = []
    for _ in range(args.steps):
        flush_l2()
        t0 = time.perf_counter()
        Phi2 = step_e2e()
        cp.synchronize()
        e2e_samples.append((time.perf_counter() - t0) * 1e3)
        e2e_ms += e2e_samples[-1]
    barrier()
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    assert np.array_equal(Phi.spl, Phi2.spl)

    # ---- the same end-to-end call for a SparseMatrixCSC{Tv, Int32} (half the upload; informative, not the headline) ----
    colptr32_pin = torch.from_numpy(A.colptr.astype(np.int32)).pin_memory()
    rowval32_pin = torch.from_numpy(A.rowval.astype(np.int32)).pin_memory()

    def step_e2e_i32():
        d32 = cp.device_matrix_i32(A.m, A.n, colptr32_pin.numpy(), rowval32_pin.numpy())
        try:
            return cp.partition_stripe(d32, K_PARTS, mtd)
        finally:
            d32.close()

    for _ in range(3):
        step_e2e_i32()
    n32 = max(3, min(args.steps, 50))
    e2e32_samples = []
    for _ in range(n32):
        flush_l2()
        t0 = time.perf_counter()
        Phi3 = step_e2e_i32()
        cp.synchronize()
        e2e32_samples.append((time.perf_counter() - t0) * 1e3)
    e2e32_ms = float(np.mean(e2e32_samples))
    assert np.array_equal(Phi.spl, Phi3.spl)

    # ---- cost-oracle query throughput (the other half of BASELINE.json's metric) ----
    Q = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    qa = torch.randint(1, A.n + 2, (Q,), device="cuda", generator=g, dtype=torch.int64)
    qb = torch.randint(1, A.n + 2, (Q,), device="cuda", generator=g, dtype=torch.int64)
    qj, qjp = torch.minimum(qa, qb).contiguous(), torch.maximum(qa, qb).contiguous()
    qout = torch.empty(Q, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ocl = cp.oracle_stripe(f, dA)
    ocl.query_device(qj.data_ptr(), qjp.data_ptr(), qout.data_ptr(), Q)  # builds the dominance index, warms up
    cp.synchronize()
    q_ms = 0.0
    for _ in range(3):
        flush_l2()
        cp.timer_start()
        ocl.query_device(qj.data_ptr(), qjp.data_ptr(), qout.data_ptr(), Q)
        q_ms += cp.timer_stop()
    queries_per_s = 3 * Q / (q_ms / 1e3)
    q_check = qout[:4096].cpu().numpy()
    q_ref_args = (qj[:4096].cpu().numpy(), qjp[:4096].cpu().numpy())
    ocl.close()

    from chainb200 import parallel

    dev_ms_max, e2e_ms_max = parallel.max_over_ranks([dev_ms, e2e_ms], device="cuda")
    value = world * args.steps / (dev_ms_max / 1e3)
    e2e_value = world * args.steps / (e2e_ms_max / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # the dominant kernel of the step = the kernel-level profile entry with the largest total time
        kernels = {nm: v for nm, v in prof.items() if nm.startswith("k_") and v["ms"] > 0}
        top = max(kernels, key=lambda nm: kernels[nm]["ms"]) if kernels else None

        def roof(nm, note):
            k = kernels[nm]
            ach = (k["bytes"] / 1e9) / (k["ms"] / 1e3)
            return {"bound": "hbm", "kernel": nm, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": TRAFFIC.get(nm),
                    "peak_source": peak_src, "launches_per_step": k["launches"] / max(args.steps, 1), "avg_launch_us": 1e3 * k["ms"] / max(k["launches"], 1),
                    "algorithmic_bytes_per_launch": k["bytes"] / max(k["launches"], 1), "share_of_step": k["ms"] / max(dev_ms, 1e-9), "note": note}

        NOTES = {"k_probe_stream": "greedy feasibility probes of the 15 most likely thresholds of the bisection tree, one 8-CTA cluster each; latency-bound "
                                   "(K sequential parts x 2 cluster barriers), bytes = ONE pass over the link array per launch (SURVEY 8d G4) although every "
                                   "threshold streams it (mostly from L2)",
                 "k_lt_fill": "build_links: every nonzero takes a slot of its row's segment through an atomic cursor and stores its position; "
                              "bytes = read row, write 4 B per nonzero (random 4-byte stores)",
                 "k_lt_link": "build_links: every slot scans its row segment for the largest column below its own and writes the link (scattered 4-byte stores)",
                 "k_lt_count": "build_links: row histogram (one atomic per nonzero)",
                 "k_rs_scatter": "stable radix scatter of (row, position) pairs inside build_links (rows heavier than 128 nonzeros)",
                 "k_wm_level": "one bit level of the wavelet-matrix dominance index",
                 "k_expand_columns": "column of every nonzero from the offsets (marks + running maximum per tile); "
                                     "the 40 MB it writes are consumed from L2 by the next kernels (dram traffic under ncu: 4 MB read, ~0 written)"}
        roofline = roof(top, NOTES.get(top, "")) if top else None
        roofline_all = [roof(nm, NOTES.get(nm, "")) for nm in sorted(kernels, key=lambda nm: -kernels[nm]["ms"])]
        phases = {nm: round(v["ms"] / args.steps, 4) for nm, v in prof.items()}
        # CPU baseline on rank 0 only, N = 1 only, bounded sample
        cpu = None
        if world == 1:
            per, secs, Phi_ref = cpu_reference_run(A, mtd, 3, 0)
            assert np.array_equal(Phi_ref.spl, Phi.spl), "GPU result differs from the CPU oracle"
            import pyoracle as ref

            assert np.array_equal(ref.oracle_query(f, A, *q_ref_args, hint=cp.SparseHint()), q_check), "oracle queries differ from the CPU oracle"
            cpu = {"value": 1.0 / per, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"full workload x3 (oracle build {secs[0]*1e3:.1f} ms + bisection {secs[1]*1e3:.1f} ms per step); C++ restatement of the "
                             "Julia reference, single thread (the reference has no threads)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
                "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms_max / args.steps, "median_ms_per_step": float(np.median(e2e_samples)),
                        "max_ms": float(np.max(e2e_samples)), "h2d_bytes_per_step": int((A.nnz + A.n + 1) * 8 + 512), "d2h_bytes_per_step": int((K_PARTS + 1) * 8 + 64)},
                "e2e_int32_index": {"ms_per_step": e2e32_ms, "median_ms_per_step": float(np.median(e2e32_samples)), "max_ms": float(np.max(e2e32_samples)), "h2d_bytes_per_step": int(colptr32_pin.numel() * 4 + rowval32_pin.numel() * 4),
                                    "note": "same public call for a SparseMatrixCSC{Tv,Int32} (cpb_matrix_create_i32), rank 0; the headline e2e is the Int64 form"},
                "oracle_queries_per_s": queries_per_s * world, "oracle_query_batch": {"Q": Q, "ms": q_ms / 3, "pattern": "uniform random (j <= j') pairs, device-resident"},
                "gpu_launches": int(launches), "roofline": roofline, "roofline_all_kernels": roofline_all, "cpu_baseline": cpu, "clocks": sampler.summary(),
                "phases_ms_per_step": phases, "bisection": cp.bisect_stats(), "parity": "split vector identical to the CPU oracle" if cpu else "checked at N=1"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
