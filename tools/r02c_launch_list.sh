#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --configs none > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:k_' -c 600 --csv --log-file gpurun_out/launches_r02c_c3.csv python bench.py --steps 2 --warmup 3 --configs none > gpurun_out/ncu_launches_r02c.log 2>&1
wc -l gpurun_out/launches_r02c_c3.csv
