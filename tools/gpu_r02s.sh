#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zzzz_gpu_column_lane.py -x -q 2>&1 | tail -1
OUT=gpurun_out/clvar_r02s.jsonl; : > $OUT
timeout 300 python tools/clvar.py --n 44 --p 1 --tag p1_pad >> $OUT
timeout 300 python tools/clvar.py --n 119 --p 1 --tag p1_n119_pad >> $OUT
PFA_LIB=polyfem_b200/libpfa_p1s8.so timeout 300 python tools/clvar.py --n 44 --p 1 --tag p1_pad_s8 >> $OUT
PFA_LIB=polyfem_b200/libpfa_p1s8.so timeout 300 python tools/clvar.py --n 119 --p 1 --tag p1_n119_pad_s8 >> $OUT
timeout 300 python tools/clvar.py --tag p2 >> $OUT
cat $OUT | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'], d['n'], 'p',d['p'], 'ms %.3f'%d['kernel_ms'], 'fill %.3f'%d['zero_fill_ms'], d['energy'], d['vsum'])
"
