#!/usr/bin/env python
"""Kernel timings of every BASELINE.json config that fits one GPU (device-resident buffers).

  python tools/cfgbench.py [--cfg 1 2 4L 4E] [--reps 10] > profiles/cfgbench_rNN.jsonl

One JSON line per config: elements, dofs, nnz, kernel ms (library CUDA events on the launching
stream), zero-fill ms, elements/s, nnz/s and the roofline fraction SURVEY.md §8(d) asks for:
HBM (B_alg bytes/element over the measured copy bandwidth) for P1/P2, FP64 (F_alg flops per
element over the measured DFMA peak, tools/microbench.cu: 33.8 TFLOP/s) for P4.
cfg 3 (the headline) is bench.py's job."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from polyfem_b200 import capi, mesh as M, tables  # noqa: E402

DFMA_PEAK_TFLOPS = 33.8  # profiles/microbench_r01.jsonl

CFG = {
    # name: (material, p, n, linear, label)
    "1": ("LinearElasticity", 1, 20, True, "cfg 1 LinearElasticity P1 n=20 stiffness"),
    "2": ("NeoHookean", 1, 44, False, "cfg 2 NeoHookean P1 n=44 E+g+H"),
    "2L": ("LinearElasticity", 1, 44, True, "LinearElasticity P1 n=44 stiffness"),
    "3s": ("NeoHookean", 2, 40, False, "NeoHookean P2 n=40 E+g+H (small cfg 3)"),
    "4L": ("Laplacian", 4, 32, True, "cfg 4 Laplacian P4 n=32 stiffness"),
    "4E": ("LinearElasticity", 4, 16, True, "cfg 4 LinearElasticity P4 n=16 stiffness (n=32 has nnz > 2^31)"),
    "5s": ("NeoHookean", 1, 60, False, "cfg 5 per-GPU share: NeoHookean P1 n=60 (1.30 M tets ~ 10.1 M / 8) E+g+H"),
    "5m": ("Mass", 1, 60, True, "cfg 5 per-GPU share: Mass P1 n=60 (1.30 M tets), mass quadrature order 2"),
    "3m": ("Mass", 2, 40, True, "Mass P2 n=40 (384 k tets), mass quadrature order 4"),
}


def hbm_peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", nargs="*", default=["1", "2", "4L", "4E", "5s"])
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    for key in a.cfg:
        material, p, n, linear, label = CFG[key]
        mesh = M.kuhn_cube(n, p)
        if material == "Mass":
            t = tables.reference_tables(p, tables.quadrature_order(p, is_mass=True))
            h = capi.Handle(material, mesh.conn, mesh.n_bases, t["weights"], None, vertices=mesh.vertices, ref_vals=t["val"], density=1000.0)
        else:
            t = tables.reference_tables(p)
            h = capi.Handle(material, mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu)
        x = M.random_displacement(mesh)
        xd = torch.from_numpy(np.ascontiguousarray(x[: h.ndof])).cuda()
        e = torch.zeros(1, dtype=torch.float64, device="cuda")
        g = torch.zeros(h.ndof, dtype=torch.float64, device="cuda")
        v = torch.zeros(h.nnz, dtype=torch.float64, device="cuda")
        run = (lambda: h.linear_stiffness_raw(v)) if linear else (lambda: h.grad_hess_raw(xd, e, g, v))
        for _ in range(3):
            run()
        h.synchronize()
        h.profile_enable(True)
        for _ in range(a.reps):
            run()
        recs = h.profile_read()
        kern = [ms for (k, ms) in recs if "assemble" in k]
        fill = [ms for (k, ms) in recs if "zero_fill" in k]
        name = sorted({k for (k, ms) in recs if "assemble" in k})[0]
        k_ms, f_ms = float(np.mean(kern)), float(np.mean(fill)) if fill else 0.0
        n_el, n_loc, n_qp = mesh.n_elements, h.n_loc, t["weights"].shape[0]
        size = h.size
        b_alg = 4 * n_loc + 80 + (16 if material != "Laplacian" else 0) + (0 if linear else 16 * h.ndof / n_el) + 8 * h.nnz / n_el
        N = n_loc * size
        if material == "Mass":
            f_alg = n_qp * n_loc * n_loc
        elif material == "Laplacian":
            f_alg = n_qp * 3 * n_loc * n_loc
        elif linear:
            f_alg = n_qp * (72 * N + 6 * N * N)
        else:
            f_alg = n_qp * (2 * (81 * n_loc + 27 * n_loc * (n_loc + 1) / 2) + 300)
        hbm = b_alg * n_el / (k_ms * 1e-3) / 1e9
        fp64 = f_alg * n_el / (k_ms * 1e-3) / 1e12
        line = {"cfg": key, "workload": label, "kernel": name, "elements": n_el, "dofs": int(h.ndof), "nnz": int(h.nnz),
                "kernel_ms": k_ms, "zero_fill_ms": f_ms, "elements_per_s": n_el / ((k_ms + f_ms) * 1e-3),
                "nnz_per_s": h.nnz / ((k_ms + f_ms) * 1e-3),
                "B_alg": b_alg, "hbm_GBs": hbm, "hbm_frac": hbm / hbm_peak(),
                "F_alg": f_alg, "fp64_TFLOPs": fp64, "fp64_frac": fp64 / DFMA_PEAK_TFLOPS,
                "setup_seconds": h.setup_seconds()}
        print(json.dumps(line), flush=True)
        del h, v, g, xd
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
