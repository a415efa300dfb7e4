#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zzzz_gpu_column_lane.py tests/test_gpu_owner_partition.py -x -q 2>&1 | tail -2
OUT=gpurun_out/clvar_r02q.jsonl; : > $OUT
timeout 300 python tools/clvar.py --tag depth >> $OUT
timeout 300 python tools/clvar.py --n 44 --p 1 --tag p1_depth4 >> $OUT
timeout 300 python tools/clvar.py --n 119 --p 1 --tag p1_n119 >> $OUT
timeout 300 python tools/clvar.py --n 44 --p 1 --flags 8 --tag p1_rowlane >> $OUT
timeout 300 python tools/clvar.py --n 119 --p 1 --flags 8 --tag p1_n119_rowlane >> $OUT
cat $OUT | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'], d['n'], 'p',d['p'], 'ms %.3f'%d['kernel_ms'], 'fill %.3f'%d['zero_fill_ms'], d['energy'], d['vsum'])
"
