#!/usr/bin/env python
"""Row-lane (PFA_FLAG_ROW_LANE) against column-lane (default) NeoHookean assembly on one GPU, device-resident buffers.

  python tools/clbench.py [--n 69] [--p 2] [--reps 10] > profiles/clbench_rNN.jsonl

One JSON line per path: kernel ms (library CUDA events on the launching stream; for the column-lane path the records kernel
and the column kernels are one record), zero-fill ms, elements/s; then one line comparing the two outputs on the device
(largest |difference| over the largest |entry|, for values and gradient) and whether two column-lane calls agree bit for bit."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from polyfem_b200 import capi, mesh as M, tables  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=69)
    ap.add_argument("--p", type=int, default=2)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    mesh = M.kuhn_cube(a.n, a.p)
    t = tables.reference_tables(a.p)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    x = M.random_displacement(mesh)
    out = {}
    for name, flags in (("row_lane", capi.FLAG_ROW_LANE), ("column_lane", 0)):
        h = capi.Handle("NeoHookean", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu, flags=flags)
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
        kname = sorted({k for (k, ms) in recs if "assemble" in k})[0]
        k_ms, f_ms = float(np.mean(kern)), float(np.mean(fill)) if fill else 0.0
        line = {"path": name, "kernel": kname, "elements": mesh.n_elements, "nnz": int(h.nnz), "kernel_ms": k_ms, "kernel_ms_min": float(np.min(kern)),
                "zero_fill_ms": f_ms, "elements_per_s": mesh.n_elements / ((k_ms + f_ms) * 1e-3), "setup_seconds": h.setup_seconds()}
        print(json.dumps(line), flush=True)
        out[name] = (float(e.item()), g.clone(), v.clone())
        if name == "column_lane":
            h.grad_hess_raw(xd, e, g, v)
            h.synchronize()
            out["repeat"] = bool(torch.equal(v, out[name][2]) and torch.equal(g, out[name][1]) and float(e.item()) == out[name][0])
        del h, v, g, xd
        torch.cuda.empty_cache()
    (e0, g0, v0), (e1, g1, v1) = out["row_lane"], out["column_lane"]
    print(json.dumps({"compare": "column_lane vs row_lane", "energy_rel": abs(e1 - e0) / abs(e0),
                      "gradient_rel_max": float((g1 - g0).abs().max() / g0.abs().max()),
                      "values_rel_max": float((v1 - v0).abs().max() / v0.abs().max()),
                      "column_lane_bitwise_repeatable": out["repeat"]}), flush=True)


if __name__ == "__main__":
    main()
