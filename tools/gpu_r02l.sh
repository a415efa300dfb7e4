#!/bin/bash
mkdir -p gpurun_out
# DRAM traffic of one assembly at cfg 3 (records + energy + two column kernels): 2 passes per kernel
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:cl2_ -s 12 -c 4 --csv --log-file gpurun_out/traffic_r02l.csv python tools/clvar.py --reps 1 > gpurun_out/traffic_r02l.log 2>&1
# launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_r02l.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/launches_r02l.log 2>&1
# full capture of the final kernels at n = 40
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cl2_ -s 8 -c 4 -o gpurun_out/cl2_r02l python /tmp/cl_one.py 40 2 3 > gpurun_out/ncu_r02l.log 2>&1
# sanitizer logs: racecheck + memcheck of the default path (P2 and P1) and of the owner-partition handle
timeout 600 compute-sanitizer --tool racecheck python /tmp/cl_one.py 5 2 1 > gpurun_out/racecheck_r02l_p2.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python /tmp/cl_one.py 6 1 1 > gpurun_out/racecheck_r02l_p1.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python /tmp/cl_one.py 5 2 1 > gpurun_out/memcheck_r02l_p2.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python /tmp/cl_one.py 6 1 1 > gpurun_out/memcheck_r02l_p1.log 2>&1
PFA_ROW_LANE=1 timeout 600 compute-sanitizer --tool racecheck python /tmp/cl_one.py 4 2 1 > gpurun_out/racecheck_r02l_rowlane.log 2>&1
tail -2 gpurun_out/racecheck_r02l_*.log gpurun_out/memcheck_r02l_*.log
cat gpurun_out/traffic_r02l.csv | tail -14
