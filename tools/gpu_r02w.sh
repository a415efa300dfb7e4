#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_psd_generic.py tests/test_abi_cpp_host.py -x -q -m gpu > gpurun_out/r02w_pytest.log 2>&1; tail -6 gpurun_out/r02w_pytest.log
# launch list of the bench command (no CPU baseline, no end-to-end leg: kernels of the timed device loop only)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02w.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/launches_r02w.log 2>&1
tail -3 gpurun_out/launches_r02w.log | cut -c1-400; wc -l gpurun_out/launches_r02w.csv
