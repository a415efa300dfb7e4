#!/bin/bash
# last GPU call of round 2 (12 GPU-minutes left): targeted parity tests on the rebuilt library (pfa_create validation of
# owned_nodes, partition object), smoke(), the bench line at HEAD and the ncu launch list of the bench command
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_zzzz_gpu_column_lane.py tests/test_gpu_owner_partition.py tests/test_abi_cpp_host.py -x -q -m gpu > gpurun_out/r02aj_pytest.log 2>&1; tail -3 gpurun_out/r02aj_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 240 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02aj.json 2> gpurun_out/bench_r02aj.err; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_r02aj.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"], d["clocks"])
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02aj.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/launches_r02aj.log 2>&1
tail -n 12 gpurun_out/launches_r02aj.csv | cut -c1-40,150-
