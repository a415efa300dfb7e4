#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/clvar_r02f.jsonl; : > $OUT
for v in "" _tma _mb12 _s1rg0mb12 _s1rg0 _s2rg0mb12 _s1mb12; do
  export PFA_LIB=polyfem_b200/libpfa$v.so
  echo "== variant '$v'"
  timeout 600 python -m pytest tests/test_zzzz_gpu_column_lane.py -x -q 2>&1 | tail -1
  timeout 300 python tools/clvar.py --tag "v$v" >> $OUT
  timeout 300 python tools/clvar.py --n 44 --p 1 --tag "p1$v" >> $OUT
done
cat $OUT
