#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/clvar_r02ae.jsonl; : > $OUT
for r in 1 2; do
timeout 300 python tools/clvar.py --tag default_run$r >> $OUT
PFA_LIB=polyfem_b200/libpfa_shfl.so timeout 300 python tools/clvar.py --tag shfl_rows_run$r >> $OUT
done
PFA_LIB=polyfem_b200/libpfa_shfl.so timeout 300 python tools/clvar.py --n 44 --p 1 --reps 20 --tag shfl_rows_p1_cfg2 >> $OUT
timeout 300 python tools/clvar.py --n 44 --p 1 --reps 20 --tag default_p1_cfg2 >> $OUT
cat $OUT | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'], d['n'], 'p',d['p'], 'ms %.3f'%d['kernel_ms'], 'min %.3f'%d['kernel_ms_min'])
"
PFA_LIB=polyfem_b200/libpfa_shfl.so timeout 600 python -m pytest tests/test_zzzz_gpu_column_lane.py tests/test_gpu_owner_partition.py -x -q -m gpu 2>&1 | tail -3
