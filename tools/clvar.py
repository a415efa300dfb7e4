#!/usr/bin/env python
"""One timing line for the default NeoHookean path of the library selected by PFA_LIB (kernel-variant experiments):
  PFA_LIB=polyfem_b200/libpfa_x.so PFA_CL_SMALL_ROWS=96 python tools/clvar.py --n 69 --p 2 --tag x
Prints {"tag", "n", "p", "kernel_ms", "kernel_ms_min", "zero_fill_ms", "energy"}; device-resident buffers, library CUDA events."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from polyfem_b200 import capi, mesh as M, tables  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=69)
ap.add_argument("--p", type=int, default=2)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--tag", default="")
a = ap.parse_args()
mesh = M.kuhn_cube(a.n, a.p)
t = tables.reference_tables(a.p)
lam, mu = M.lame_from_E_nu(1e5, 0.3)
x = M.random_displacement(mesh)
h = capi.Handle("NeoHookean", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu, flags=a.flags)
xd = torch.from_numpy(np.ascontiguousarray(x[: h.ndof])).cuda()
e = torch.zeros(1, dtype=torch.float64, device="cuda")
g = torch.zeros(h.ndof, dtype=torch.float64, device="cuda")
v = torch.zeros(h.nnz, dtype=torch.float64, device="cuda")
for _ in range(3):
    h.grad_hess_raw(xd, e, g, v)
h.synchronize()
h.profile_enable(True)
for _ in range(a.reps):
    h.grad_hess_raw(xd, e, g, v)
recs = h.profile_read()
kern = [ms for (k, ms) in recs if "assemble" in k]
fill = [ms for (k, ms) in recs if "zero_fill" in k]
env = {k: v_ for k, v_ in os.environ.items() if k.startswith("PFA_")}
print(json.dumps({"tag": a.tag, "n": a.n, "p": a.p, "elements": mesh.n_elements, "kernel_ms": float(np.mean(kern)), "kernel_ms_min": float(np.min(kern)),
                  "zero_fill_ms": float(np.mean(fill)) if fill else 0.0, "energy": float(e.item()), "vsum": float(v.sum().item()), "env": env}), flush=True)
