#!/bin/bash
mkdir -p gpurun_out
# DRAM traffic of one assembly at cfg 3 with the final kernels (records + energy + two column launches)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cl2_ -s 12 -c 4 --csv --log-file gpurun_out/traffic_r02ai.csv python tools/clvar.py --reps 1 > gpurun_out/traffic_r02ai.log 2>&1
tail -13 gpurun_out/traffic_r02ai.csv | cut -c1-60,200-
# end-of-round verification
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02ai_pytest.log 2>&1; tail -4 gpurun_out/r02ai_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02ai.json 2> gpurun_out/bench_r02ai.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r02ai.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"])
PY
