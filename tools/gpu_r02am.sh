#!/bin/bash
# kernel variants of the P2 / P1 column kernels (tools/clvar.py, cfg 3 and cfg 2): one elected lane issues the TMA copies of a step
# (e), B-lane update in one batch (b), flush with the next block's strip rows loaded early (f), TMA + elected issue for P1 (te)
mkdir -p gpurun_out
OUT=gpurun_out/clvar_r02am.jsonl; : > $OUT
run() { PFA_LIB=polyfem_b200/libpfa$1.so timeout -k 5 90 python tools/clvar.py --n $2 --p $3 --tag "lib$1" >> $OUT 2>> gpurun_out/clvar_r02am.err; }
run "" 69 2; run _e 69 2; run _ebf 69 2
run "" 44 1; run _te 44 1
run _ebf 69 2; run "" 69 2
python - <<'PY'
import json
for l in open("gpurun_out/clvar_r02am.jsonl"):
    d=json.loads(l); print(d['tag'], d['n'], 'p',d['p'], 'ms %.4f'%d['kernel_ms'], 'min %.4f'%d['kernel_ms_min'], 'E %.12e'%d['energy'], 'vsum %.6e'%d['vsum'])
PY
tail -3 gpurun_out/clvar_r02am.err
for L in _ebf; do PFA_LIB=polyfem_b200/libpfa$L.so timeout -k 5 150 python -m pytest tests/test_zzzz_gpu_column_lane.py tests/test_gpu_owner_partition.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2; done
