#!/usr/bin/env python
"""Kernel-variant timing: one mesh, device-resident buffers, prints per-kernel ms.
Usage: PFA_LIB=polyfem_b200/libpfa_x.so python tools/kbench.py [--n 40] [--p 2] [--material NeoHookean]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from polyfem_b200 import capi, mesh as M, tables

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=40)
ap.add_argument("--p", type=int, default=2)
ap.add_argument("--material", default="NeoHookean")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--linear", action="store_true")
a = ap.parse_args()
mesh = M.kuhn_cube(a.n, a.p)
x = M.random_displacement(mesh)
t = tables.reference_tables(a.p)
lam, mu = M.lame_from_E_nu(1e5, 0.3)
h = capi.Handle(a.material, mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu)
xd = torch.from_numpy(x[: h.ndof] if h.size == 3 else x[: h.ndof]).cuda()
e = torch.zeros(1, dtype=torch.float64, device="cuda")
g = torch.zeros(h.ndof, dtype=torch.float64, device="cuda")
v = torch.zeros(h.nnz, dtype=torch.float64, device="cuda")
run = (lambda: h.linear_stiffness_raw(v)) if a.linear else (lambda: h.grad_hess_raw(xd, e, g, v))
for _ in range(3):
    run()
h.synchronize()
h.profile_enable(True)
for _ in range(a.reps):
    run()
recs = h.profile_read()
names = sorted({r[0] for r in recs})
out = {nm: float(np.mean([ms for (k, ms) in recs if k == nm])) for nm in names}
print(os.environ.get("PFA_LIB", "default"), f"n_el={mesh.n_elements}", {k: round(vv, 4) for k, vv in out.items()},
      "Mel/s(kernel)=%.1f" % (mesh.n_elements / max(vv for k, vv in out.items() if "assemble" in k) / 1e3))
