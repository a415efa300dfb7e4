#!/bin/bash
# round 2, second GPU session: first run of the v2 owner-computes kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zzzz_gpu_column_lane.py -x -q > gpurun_out/r02b_pytest_cl.log 2>&1
tail -5 gpurun_out/r02b_pytest_cl.log
cat > /tmp/cl_one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from polyfem_b200 import capi, mesh as M, tables
n=int(sys.argv[1]); p=int(sys.argv[2]); mesh = M.kuhn_cube(n, p); t = tables.reference_tables(p)
lam, mu = M.lame_from_E_nu(1e5, 0.3); x = M.random_displacement(mesh)
h = capi.Handle("NeoHookean", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu)
xd = torch.from_numpy(np.ascontiguousarray(x[: h.ndof])).cuda()
e = torch.zeros(1, dtype=torch.float64, device="cuda"); g = torch.zeros(h.ndof, dtype=torch.float64, device="cuda"); v = torch.zeros(h.nnz, dtype=torch.float64, device="cuda")
for _ in range(int(sys.argv[3])): h.grad_hess_raw(xd, e, g, v)
h.synchronize()
print("ok", float(e.item()))
PY
timeout 600 compute-sanitizer --tool memcheck python /tmp/cl_one.py 4 2 1 > gpurun_out/r02b_memcheck.log 2>&1
tail -3 gpurun_out/r02b_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python /tmp/cl_one.py 4 2 1 > gpurun_out/r02b_racecheck.log 2>&1
tail -3 gpurun_out/r02b_racecheck.log
timeout 600 python tools/clbench.py --n 69 --p 2 --reps 10 > gpurun_out/clbench_r02b.jsonl 2> gpurun_out/clbench_r02b.err
cat gpurun_out/clbench_r02b.jsonl
timeout 300 python tools/clbench.py --n 44 --p 1 --reps 10 > gpurun_out/clbench_r02b_p1.jsonl 2>> gpurun_out/clbench_r02b.err
cat gpurun_out/clbench_r02b_p1.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cl2_ -s 8 -c 4 -o gpurun_out/cl2_r02b python /tmp/cl_one.py 40 2 3 > gpurun_out/ncu_r02b.log 2>&1
tail -3 gpurun_out/ncu_r02b.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02b_pytest_all.log 2>&1
tail -5 gpurun_out/r02b_pytest_all.log
