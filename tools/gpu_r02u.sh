#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/clvar_r02u.jsonl; : > $OUT
for c in 48 24 12 8; do PFA_CL_CHUNK_STEPS=$c timeout 300 python tools/clvar.py --tag chunk$c >> $OUT; done
for b in 4096 65536; do PFA_CL_BUCKET=$b PFA_CL_CHUNK_STEPS=12 timeout 300 python tools/clvar.py --tag chunk12_bucket$b >> $OUT; done
PFA_CL_SMALL_ROWS=84 PFA_CL_CHUNK_STEPS=12 timeout 300 python tools/clvar.py --tag chunk12_small84 >> $OUT
cat $OUT | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'], d['n'], 'p',d['p'], 'ms %.3f'%d['kernel_ms'], 'min %.3f'%d['kernel_ms_min'])
"
