#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zzz_gpu_config_size.py -x -q 2>&1 | tail -3
for c in 1 2 4 4le 5; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/bench_r02k_cfg$c.json 2> gpurun_out/bench_r02k_cfg$c.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02k_cfg$c.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("cfg $c", d["metric"], "ms/step %.4f"%d["ms_per_step"], "value %.4g"%d["value"], "hbm frac %.3f"%r["frac"], "fp64 frac %.3f"%r["fp64"]["frac"], "kernel", r["kernel"], "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]))
except Exception as ex:
    print("cfg $c FAILED", ex); print(open("gpurun_out/bench_r02k_cfg$c.err").read()[-1500:])
PY
done
