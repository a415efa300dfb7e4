#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02h_pytest.log 2>&1; tail -4 gpurun_out/r02h_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err; tail -c 3000 gpurun_out/bench_r02h.json; tail -3 gpurun_out/bench_r02h.err
