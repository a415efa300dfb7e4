#!/bin/bash
# last seconds of the round's GPU budget: the tests most sensitive to rounding changes, under the new default (MODE 3)
timeout -k 3 45 python -m pytest tests/test_newton_gpu.py tests/test_gpu_owner_partition.py tests/test_zzz_gpu_config_size.py -x -q -m gpu 2>&1 | tail -2
