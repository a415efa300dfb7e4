#!/bin/bash
# full GPU suite at HEAD (the driver's round-end command), timed
mkdir -p gpurun_out
/usr/bin/time -v timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/r02ak_pytest.log 2> gpurun_out/r02ak_time.log; tail -4 gpurun_out/r02ak_pytest.log; grep -E "Elapsed|Maximum resident" gpurun_out/r02ak_time.log
