#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi_rank.py -x -q 2>&1 | tail -1
run() { # N config tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2962$1 bench.py --gpus $1 --config $2 --steps 20 --warmup 3 --e2e-steps 0 > gpurun_out/bench_r02t_cfg$2_$1gpu.json 2> gpurun_out/bench_r02t_cfg$2_$1gpu.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02t_cfg$2_$1gpu.json").read().strip().splitlines()[-1])
    pr=d.get("per_rank") or {}
    print("cfg $2 N=$1 ms/step %.4f value %.4g"%(d["ms_per_step"], d["value"]), "assembly", pr.get("assembly_ms"), "allreduce", pr.get("exchange_ms"))
except Exception as ex:
    print("cfg $2 N=$1 FAILED", ex); print(open("gpurun_out/bench_r02t_cfg$2_$1gpu.err").read()[-1200:])
PY
}
run 8 3; run 4 3; run 2 3; run 8 5; run 8 2
timeout 600 python bench.py --config 3 --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_r02t_cfg3_1gpu.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02t_cfg3_1gpu.json').read().strip().splitlines()[-1]); print('cfg 3 N=1 ms/step', d['ms_per_step'])"
