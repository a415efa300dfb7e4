#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi_rank.py -x -q 2>&1 | tail -2
for N in 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 0 > gpurun_out/bench_r02j_${N}gpu.json 2> gpurun_out/bench_r02j_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r02j_${N}gpu.json").read().strip().splitlines()[-1])
print($N, "ms/step", d["ms_per_step"], "value", d["value"], d.get("per_rank"))
PY
done
