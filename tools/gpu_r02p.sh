#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zzzz_gpu_column_lane.py tests/test_gpu_owner_partition.py tests/test_zzz_gpu_config_size.py -x -q 2>&1 | tail -2
OUT=gpurun_out/clvar_r02p.jsonl; : > $OUT
timeout 300 python tools/clvar.py --tag p2z >> $OUT
PFA_CL_NO_P2Z=1 timeout 300 python tools/clvar.py --tag p2s >> $OUT
cat $OUT | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'], d['n'], 'p',d['p'], 'ms %.3f'%d['kernel_ms'], d['energy'], d['vsum'])
"
