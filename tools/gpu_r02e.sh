#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zzzz_gpu_column_lane.py -x -q 2>&1 | tail -3
OUT=gpurun_out/clvar_r02e.jsonl; : > $OUT
timeout 300 python tools/clvar.py --tag base >> $OUT
timeout 300 python tools/clvar.py --n 44 --p 1 --tag p1 >> $OUT
cat $OUT
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cl2_columns -s 4 -c 2 -o gpurun_out/cl2_r02e python /tmp/cl_one.py 40 2 3 > gpurun_out/ncu_r02e.log 2>&1
tail -2 gpurun_out/ncu_r02e.log
