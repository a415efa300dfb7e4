#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zzzz_gpu_column_lane.py -x -q 2>&1 | tail -1
OUT=gpurun_out/clvar_r02n.jsonl; : > $OUT
timeout 300 python tools/clvar.py --tag stcs >> $OUT
PFA_CL_BUCKET=4096 timeout 300 python tools/clvar.py --tag stcs_b4096 >> $OUT
cat $OUT | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['tag'], d['n'], 'p',d['p'], 'ms %.3f'%d['kernel_ms'], d['energy'], d['vsum'])
"
for b in 16384 4096; do
PFA_CL_BUCKET=$b timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cl2_ -s 12 -c 4 --csv --log-file gpurun_out/traffic_r02n_$b.csv python tools/clvar.py --reps 1 > gpurun_out/traffic_r02n.log 2>&1
done
python - <<'PY'
import csv
for f in ("gpurun_out/traffic_r02n_16384.csv","gpurun_out/traffic_r02n_4096.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>14 and r[0].isdigit()]
    R=sum(float(r[14]) for r in rows if r[12]=="dram__bytes_read.sum"); W=sum(float(r[14]) for r in rows if r[12]=="dram__bytes_write.sum")
    print(f, "read GB %.2f write GB %.2f total %.2f ratio %.2f"%(R/1e9,W/1e9,(R+W)/1e9,(R+W)/5.884e9))
PY
