#!/bin/bash
# verification of the elected-issue kernel as the default: full GPU suite, smoke(), bench line
mkdir -p gpurun_out
S=$(date +%s); timeout -k 5 300 python -m pytest tests -x -q -m gpu > gpurun_out/r02an_pytest.log 2>&1; tail -3 gpurun_out/r02an_pytest.log; echo "pytest -m gpu took $(( $(date +%s) - S )) s"
timeout -k 5 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 5 200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02an.json 2> gpurun_out/bench_r02an.err; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_r02an.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"], d["clocks"])
except Exception as e:
    print("bench line unreadable:", e)
PY
