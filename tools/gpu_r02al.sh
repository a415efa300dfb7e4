#!/bin/bash
# ncu --set full capture of the final default kernels at the HEADLINE size (n = 69, P2): records, energy, the two column launches
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cl2_ -s 12 -c 4 -o gpurun_out/cl2_r02al python tools/clvar.py --reps 1 > gpurun_out/ncu_r02al.log 2>&1
tail -3 gpurun_out/ncu_r02al.log; ls -la gpurun_out/cl2_r02al.ncu-rep
