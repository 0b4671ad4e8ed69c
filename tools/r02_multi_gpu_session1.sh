#!/bin/bash
# N GPUs (N = number visible): the C-ABI sharded solve -- correctness on all ranks, then the bench line
set -x
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/run_sharded_check.py > gpurun_out/sharded_check_$N.log 2>&1; echo "check rc=$?" >> gpurun_out/sharded_check_$N.log
grep -E "^\{|rc=|Error|error" gpurun_out/sharded_check_$N.log | cut -c1-400 | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_${N}gpu.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N", d["n_gpus"], "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "sharded", d["sharded"], {k: v for k, v in d["phases_ms_per_step"].items() if v > 0.2})
except Exception as e:
    print("unreadable", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_${N}gpu_reference.json 2>&1; echo "ref rc=$?"
