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
ap.add_argument("--project", action="store_true", help="also time the Dirichlet projection (x = 0 face clamped) and is_step_valid")
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

if a.project:
    import time
    on_face = np.flatnonzero(mesh.node_xyz[:, 0] < 1e-12)
    dofs = (on_face[:, None] * 3 + np.arange(3)[None, :]).reshape(-1).astype(np.int32)
    t0 = time.perf_counter()
    h.set_constrained_dofs(dofs)
    setup = time.perf_counter() - t0
    gr = torch.zeros(h.ndof_reduced, dtype=torch.float64, device="cuda")
    vr = torch.zeros(h.nnz_reduced, dtype=torch.float64, device="cuda")
    for _ in range(2):
        h.project_hessian(v, 1.0, out=vr)
        h.project_gradient(g, 1.0, out=gr)
        h.is_step_valid(xd)
    h.synchronize()
    h.profile_enable(True)
    for _ in range(a.reps):
        h.project_hessian(v, 0.5, out=vr)
        h.project_gradient(g, 0.5, out=gr)
        h.is_step_valid(xd)
    recs = h.profile_read()
    for _ in range(2):
        h.grad_hess_reduced_raw(xd, 0.5, e, gr, vr)
    h.synchronize()
    h.profile_enable(True)
    for _ in range(a.reps):
        h.grad_hess_reduced_raw(xd, 0.5, e, gr, vr)
    fr = h.profile_read()
    print("fused reduced assembly:", {nm: round(float(np.mean([ms for (k, ms) in fr if k == nm])), 4) for nm in sorted({r[0] for r in fr})})
    names = sorted({r[0] for r in recs})
    out = {nm: round(float(np.mean([ms for (k, ms) in recs if k == nm])), 4) for nm in names}
    gb = (8 * h.nnz_reduced * 2 + 4 * h.nnz_reduced) / 1e9
    print("projection:", f"constrained={dofs.size} ndof_red={h.ndof_reduced} nnz_red={h.nnz_reduced} setup_s={setup:.3f}", out,
          "project_hessian GB/s=%.0f" % (gb / (out["project_hessian(gather)"] * 1e-3)))
