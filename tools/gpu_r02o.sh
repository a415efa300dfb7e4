#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_large_index.py -x -q 2>&1 | tail -3
timeout 1500 python bench.py --config 4le --n 32 --flags 32 --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/bench_r02o_cfg4le_n32.json 2> gpurun_out/bench_r02o_cfg4le_n32.err
tail -c 1500 gpurun_out/bench_r02o_cfg4le_n32.json; tail -5 gpurun_out/bench_r02o_cfg4le_n32.err
