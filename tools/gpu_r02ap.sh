#!/bin/bash
# streamed vertex-weighted entry step (MODE 3, PFA_CL_P2Y=1) against the default (MODE 2) with the elected TMA issue
mkdir -p gpurun_out
OUT=gpurun_out/clvar_r02ap.jsonl; : > $OUT
timeout -k 5 60 python tools/clvar.py --tag mode2 >> $OUT 2>/dev/null
PFA_CL_P2Y=1 timeout -k 5 60 python tools/clvar.py --tag mode3_streamed >> $OUT 2>/dev/null
PFA_CL_P2Y=1 timeout -k 5 60 python -m pytest tests/test_zzzz_gpu_column_lane.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
python - <<'PY'
import json
for l in open("gpurun_out/clvar_r02ap.jsonl"):
    d=json.loads(l); print(d['tag'], 'ms %.4f'%d['kernel_ms'], 'min %.4f'%d['kernel_ms_min'], 'E %.12e'%d['energy'])
PY
