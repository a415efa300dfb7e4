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
# full capture of the final default kernels (records + the two column launches, vertex-weighted entry step) at n = 40
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cl2_ -s 8 -c 4 -o gpurun_out/cl2_r02aa python /tmp/cl_one.py 40 2 3 > gpurun_out/ncu_r02aa.log 2>&1
tail -2 gpurun_out/ncu_r02aa.log
# sanitizer runs of the generic-kernel additions of this round (tiny cases; the oracle side runs on the host)
T="tests/test_gpu_fixed_corotational.py::test_fixed_corotational_equals_oracle[1-4-0.2] tests/test_gpu_fixed_corotational.py::test_fixed_corotational_equals_oracle[2-3-0.1] tests/test_gpu_viscous_damping.py::test_viscous_damping_equals_oracle[2-3] tests/test_gpu_mooney_rivlin.py::test_mooney_rivlin_equals_oracle[2-3] tests/test_gpu_mooney_rivlin.py::test_projection[2-2] tests/test_gpu_psd_generic.py::test_generic_projection_equals_oracle[NeoHookean-4-1-0.02]"
timeout 1200 compute-sanitizer --tool memcheck python -m pytest $T -q -m gpu -p no:cacheprovider > gpurun_out/memcheck_r02aa_materials.log 2>&1; tail -3 gpurun_out/memcheck_r02aa_materials.log
timeout 1200 compute-sanitizer --tool racecheck python -m pytest $T -q -m gpu -p no:cacheprovider > gpurun_out/racecheck_r02aa_materials.log 2>&1; tail -3 gpurun_out/racecheck_r02aa_materials.log
