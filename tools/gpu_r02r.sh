#!/bin/bash
mkdir -p gpurun_out
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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cl2_ -s 6 -c 3 -o gpurun_out/cl2_r02r_p1 python /tmp/cl_one.py 80 1 3 > gpurun_out/ncu_r02r.log 2>&1
tail -2 gpurun_out/ncu_r02r.log
