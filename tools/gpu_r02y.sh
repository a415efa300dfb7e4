#!/bin/bash
# the driver's scaling commands with default options (end-to-end leg included) on one 8-GPU box
mkdir -p gpurun_out
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_r02y_${n}gpu.json 2> gpurun_out/bench_r02y_${n}gpu.err
  echo "N=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02y_${n}gpu.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"], d["clocks"])
except Exception as ex:
    print("no line:", ex); print(open("gpurun_out/bench_r02y_${n}gpu.err").read()[-1500:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 2>&1 | tail -c 600
