#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02ag_pytest.log 2>&1; tail -4 gpurun_out/r02ag_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02ag.json 2> gpurun_out/bench_r02ag.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02ag.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"])
PY
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02ag_ref.json 2> gpurun_out/bench_r02ag_ref.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02ag_ref.json").read().strip().splitlines()[-1])
print(d["impl"], d["value"], d["steps"], d["warmup"], d["cpu_baseline"]["cores"])
PY
