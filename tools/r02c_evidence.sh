#!/bin/bash
# final evidence of round 2 (second session): tests, bench (+ reference arm), launch list, ncu --set full, fuzz
set -x
mkdir -p gpurun_out
nproc > gpurun_out/host_r02c.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host_r02c.txt; nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader >> gpurun_out/host_r02c.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_r02c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02c.log
tail -12 gpurun_out/pytest_r02c.log
timeout 900 python bench.py > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r02c.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_r02c_reference.json 2> gpurun_out/bench_r02c_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ["bench_r02c"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "res ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), "roof", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), round(d["roofline"]["avg_launch_us"], 1), "us", {k: v for k, v in d["phases_ms_per_step"].items() if v > 0.05})
        for c in d.get("configs", []):
            print("   ", c["config"], "res", round(c["ms_per_step"], 2), "e2e", round(c["e2e_ms"], 2), "cpu", round(c["cpu_baseline"]["ms"]), "identical", c["identical"], c["roofline"]["kernel"], round(c["roofline"]["frac"], 4))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 300 python bench.py --steps 2 --warmup 3 --configs none > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r02c_c3.csv python bench.py --steps 2 --warmup 3 --configs none > gpurun_out/ncu_launches_r02c.log 2>&1
timeout 900 ncu --set full --clock-control none -k 'regex:k_probe_stream|k_os_pass|k_os_hist|k_os_scatter|k_ub_count|k_ub_greedy|k_ub_grid|k_expand_columns' -s 40 -c 14 -o gpurun_out/prof_r02c_c3 python bench.py --steps 1 --warmup 3 --configs none > gpurun_out/ncu_full_r02c_c3.log 2>&1
ncu -i gpurun_out/prof_r02c_c3.ncu-rep --page raw --csv > gpurun_out/prof_r02c_c3.raw.csv 2>/dev/null
CPB_BENCH_HEADLINE=C2 timeout 600 ncu --set full --clock-control none -k 'regex:k_probe_stream|k_lt_fill|k_lt_link|k_lt_count|k_ub_count|k_expand_columns' -s 30 -c 8 -o gpurun_out/prof_r02c_c2 python bench.py --steps 2 --warmup 3 --configs none > gpurun_out/ncu_full_r02c_c2.log 2>&1
ncu -i gpurun_out/prof_r02c_c2.ncu-rep --page raw --csv > gpurun_out/prof_r02c_c2.raw.csv 2>/dev/null
rm -f gpurun_out/prof_r02c_c3.ncu-rep
ls -la gpurun_out/prof_r02c_* gpurun_out/launches_r02c_c3.csv
timeout 120 python tests/fuzz_parity.py --seconds 60 --seed 41 --large 0.2 > gpurun_out/fuzz_r02c.log 2>&1; echo "fuzz rc=$?"; head -1 gpurun_out/fuzz_r02c.log | cut -c1-300
timeout 300 python tools/concave_timing.py > gpurun_out/concave_timing.log 2>&1; cat gpurun_out/concave_timing.log
