#!/bin/bash
# ncu launch list of the bench command with the final kernels (elected TMA issue)
mkdir -p gpurun_out
timeout -k 5 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02ao.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/launches_r02ao.log 2>&1
tail -n 5 gpurun_out/launches_r02ao.csv | cut -c1-40,150-
