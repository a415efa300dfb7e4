#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi_rank.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02i_${N}gpu.json 2> gpurun_out/bench_r02i_${N}gpu.err
tail -c 2500 gpurun_out/bench_r02i_${N}gpu.json; tail -3 gpurun_out/bench_r02i_${N}gpu.err
