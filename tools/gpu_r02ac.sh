#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/clvar_r02ac.jsonl; : > $OUT
for r in 1 2; do
for d in 0 1 2; do PFA_CL_DEEP=$d timeout 300 python tools/clvar.py --tag deep${d}_run$r >> $OUT; done
done
cat $OUT | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'], d['n'], 'p',d['p'], 'ms %.3f'%d['kernel_ms'], 'min %.3f'%d['kernel_ms_min'])
"
PFA_CL_DEEP=1 timeout 600 python -m pytest tests/test_zzzz_gpu_column_lane.py tests/test_gpu_owner_partition.py -x -q -m gpu 2>&1 | tail -3
