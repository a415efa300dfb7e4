#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/clvar_r02g.jsonl; : > $OUT
for v in "" _odd _sel; do
  export PFA_LIB=polyfem_b200/libpfa$v.so
  echo "== variant '$v'"
  timeout 600 python -m pytest tests/test_zzzz_gpu_column_lane.py -x -q 2>&1 | tail -1
  timeout 300 python tools/clvar.py --tag "v$v" >> $OUT
done
export PFA_LIB=polyfem_b200/libpfa.so
for w in 2 4 6; do PFA_CL_WARPS_PER_SM=$w timeout 300 python tools/clvar.py --tag "warps$w" >> $OUT; done
PFA_CL_WARPS_PER_SM=4 timeout 300 python tools/clvar.py --n 44 --p 1 --tag "p1warps4" >> $OUT
PFA_CL_WARPS_PER_SM=8 timeout 300 python tools/clvar.py --n 44 --p 1 --tag "p1warps8" >> $OUT
timeout 300 python tools/clvar.py --n 44 --p 1 --tag "p1" >> $OUT
cat $OUT
