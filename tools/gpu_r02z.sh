#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_r02z_${n}gpu.json 2> gpurun_out/bench_r02z_${n}gpu.err
  echo "N=$n rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02z_${n}gpu.json").read().strip().splitlines()[-1])
    print(d["n_gpus"], d["ms_per_step"], d["e2e"])
except Exception as ex:
    print("no line:", ex); print(open("gpurun_out/bench_r02z_${n}gpu.err").read()[-1500:])
PY
done
