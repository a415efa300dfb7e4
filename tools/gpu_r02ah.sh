#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_goldens_materials.py -q -m gpu > gpurun_out/r02ah_pytest.log 2>&1; tail -8 gpurun_out/r02ah_pytest.log
