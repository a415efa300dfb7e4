#!/bin/bash
# round 2, first GPU session: time the round-1 column-lane kernels, microbench4, ncu capture at n=40
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
timeout 600 python tools/clbench.py --n 69 --p 2 --reps 10 > gpurun_out/clbench_r02a.jsonl 2> gpurun_out/clbench_r02a.err
timeout 300 python tools/clbench.py --n 44 --p 1 --reps 10 > gpurun_out/clbench_r02a_p1.jsonl 2>> gpurun_out/clbench_r02a.err
timeout 300 ./tools/microbench4 > gpurun_out/microbench4_r02a.jsonl 2>&1
cat > /tmp/cl_one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from polyfem_b200 import capi, mesh as M, tables
n=int(sys.argv[1]); mesh = M.kuhn_cube(n, 2); t = tables.reference_tables(2)
lam, mu = M.lame_from_E_nu(1e5, 0.3); x = M.random_displacement(mesh)
h = capi.Handle("NeoHookean", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu, flags=capi.FLAG_COLUMN_LANE)
xd = torch.from_numpy(np.ascontiguousarray(x[: h.ndof])).cuda()
e = torch.zeros(1, dtype=torch.float64, device="cuda"); g = torch.zeros(h.ndof, dtype=torch.float64, device="cuda"); v = torch.zeros(h.nnz, dtype=torch.float64, device="cuda")
for _ in range(3): h.grad_hess_raw(xd, e, g, v)
h.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:collane -s 6 -c 3 -o gpurun_out/collane_r02a python /tmp/cl_one.py 40 > gpurun_out/ncu_r02a.log 2>&1
ls -la gpurun_out
