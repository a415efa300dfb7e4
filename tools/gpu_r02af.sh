#!/bin/bash
mkdir -p gpurun_out
run() { # n config tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500+$1)) bench.py --gpus $1 --config $2 --steps 20 --warmup 5 --e2e-steps 0 > gpurun_out/bench_r02af_$3.json 2> gpurun_out/bench_r02af_$3.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02af_$3.json").read().strip().splitlines()[-1])
    print("$3", d["n_gpus"], d["ms_per_step"], d["value"], d.get("per_rank",{}).get("assembly_ms"))
except Exception as ex:
    print("no line:", ex); print(open("gpurun_out/bench_r02af_$3.err").read()[-1500:])
PY
}
run 8 5 cfg5_8gpu
run 8 3 cfg3_8gpu
run 4 3 cfg3_4gpu
