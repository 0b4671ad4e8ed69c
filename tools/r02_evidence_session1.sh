#!/bin/bash
# round 2, call 2: ring probes + sharded (emulated) parity, bench line, A/B of the probe kernels, launch list + set full
set -x
mkdir -p gpurun_out
nproc > gpurun_out/host_r02b.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host_r02b.txt; nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader >> gpurun_out/host_r02b.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "probe_ring or sharded or config or speculation or large_inputs" > gpurun_out/pytest_r02b_ring.log 2>&1
RING_RC=$?
echo "ring pytest rc=$RING_RC" >> gpurun_out/pytest_r02b_ring.log
tail -15 gpurun_out/pytest_r02b_ring.log
if [ $RING_RC -ne 0 ]; then export CPB_PROBE_RING=0; echo "RING DISABLED for the rest of the call"; fi
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_r02b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02b.log
tail -14 gpurun_out/pytest_r02b.log
timeout 900 python bench.py > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r02b.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_r02b_reference.json 2> gpurun_out/bench_r02b_reference.err; echo "ref rc=$?"
# A/B of the two probe kernels at C3 and C2 (resident + e2e), 0 = register tiles, 1 = ring
for ring in 0 1; do
  CPB_PROBE_RING=$ring timeout 300 python bench.py --configs none --steps 5 > gpurun_out/ab_c3_ring$ring.json 2> gpurun_out/ab_c3_ring$ring.err
  CPB_PROBE_RING=$ring CPB_BENCH_HEADLINE=C2 timeout 300 python bench.py --configs none --steps 50 > gpurun_out/ab_c2_ring$ring.json 2> gpurun_out/ab_c2_ring$ring.err
done
python - <<'PY'
import json
for f in ["bench_r02b", "ab_c3_ring0", "ab_c3_ring1", "ab_c2_ring0", "ab_c2_ring1"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "res ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "pinned", round(d["e2e_pinned"]["ms_per_step"], 2), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), round(d["roofline"]["avg_launch_us"], 1), "us", {k: v for k, v in d["phases_ms_per_step"].items() if v > 0.05})
        for c in d.get("configs", []):
            print("   ", c["config"], "res", round(c["ms_per_step"], 2), "e2e", round(c["e2e_ms"], 2), "cpu", round(c["cpu_baseline"]["ms"]), "identical", c["identical"], c["roofline"]["kernel"], round(c["roofline"]["frac"], 4))
    except Exception as e:
        print(f, "unreadable", e)
PY
# launch list at C3 (headline only) and at C2
timeout 600 python bench.py --steps 2 --warmup 3 --configs none > gpurun_out/plain_c3.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02b_c3.csv python bench.py --steps 2 --warmup 3 --configs none > gpurun_out/ncu_c3_list.log 2>&1
echo "launch list rc=$?"
# --set full: one super-capture for the C3 kernels (no source import: small report), one with source for the probe kernel at C2
timeout 600 python bench.py --steps 1 --warmup 3 --configs none > gpurun_out/plain_c3b.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none -k 'regex:k_probe_ring|k_probe_stream|k_rs_scatter|k_rs_hist|k_link_values|k_scatter_pairs|k_lt_count|k_ub_count|k_expand_columns' -s 60 -c 18 -o gpurun_out/prof_r02b_c3 python bench.py --steps 1 --warmup 3 --configs none > gpurun_out/ncu_c3_full.log 2>&1
echo "set full c3 rc=$?"
CPB_BENCH_HEADLINE=C2 timeout 300 python bench.py --steps 2 --warmup 3 --configs none > gpurun_out/plain_c2.log 2>&1 && \
CPB_BENCH_HEADLINE=C2 timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_probe_ring|k_lt_fill|k_lt_link|k_lt_count' -s 20 -c 6 -o gpurun_out/prof_r02b_c2 python bench.py --steps 2 --warmup 3 --configs none > gpurun_out/ncu_c2_full.log 2>&1
echo "set full c2 rc=$?"
timeout 300 python tools/c4_once.py > gpurun_out/plain_c4.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k 'regex:k_window_hist|k_cost_table|k_chunk_warp|k_lt_link|k_lt_fill|k_chunk_combine|k_chain|k_convex' -s 10 -c 24 -o gpurun_out/prof_r02b_c4 python tools/c4_once.py > gpurun_out/ncu_c4_full.log 2>&1
echo "set full c4 rc=$?"
for r in prof_r02b_c3 prof_r02b_c2 prof_r02b_c4; do
  if [ -f gpurun_out/$r.ncu-rep ]; then
    ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
    sz=$(stat -c %s gpurun_out/$r.ncu-rep)
    if [ $sz -gt 18000000 ]; then rm -f gpurun_out/$r.ncu-rep; echo "$r.ncu-rep ($sz bytes) dropped, raw csv kept"; fi
  fi
done
if [ -f gpurun_out/prof_r02b_c2.ncu-rep ]; then ncu -i gpurun_out/prof_r02b_c2.ncu-rep --page source --csv -k regex:k_probe_ring > gpurun_out/prof_r02b_c2.probe_source.csv 2>/dev/null; fi
du -sh gpurun_out; ls -la gpurun_out | head -50
