#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fixed_corotational.py tests/test_gpu_viscous_damping.py tests/test_gpu_mooney_rivlin.py tests/test_gpu_saint_venant_and_curved.py -q -m gpu > gpurun_out/r02x_pytest.log 2>&1; tail -12 gpurun_out/r02x_pytest.log
