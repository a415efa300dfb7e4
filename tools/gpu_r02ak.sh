#!/bin/bash
# full GPU suite at HEAD (the driver's round-end command), timed
mkdir -p gpurun_out
S=$(date +%s); timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/r02ak_pytest.log 2>&1; tail -4 gpurun_out/r02ak_pytest.log
echo "pytest -m gpu took $(( $(date +%s) - S )) s"
