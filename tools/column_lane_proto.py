#!/usr/bin/env python
"""CPU prototype of the "column lanes" data flow proposed for round 2: checks the
indexing (incidence lists, row positions, symmetric placement) against the oracle's global Hessian
and prints the per-node statistics the host scheduler has to balance. Test infrastructure only
(uses the oracle); nothing here is on a product path.

  python tools/column_lane_proto.py [cells per side, default 3] [order, default 2]

Data structures a kernel would get (all built once per mesh):
  inc_off[n_bases+1], inc[(e << 4) | i]      elements incident to each node, with the node's local index
  rowpos[t][j] = 3 * k_j                     position of row node g_j in adj(b) for incidence t = (b, e, i)
Flow per incidence (b, e, i) and component m (one lane): the element's row (i, m) of H_e, 30 values
H_e[(i,m),(j,n)], is by symmetry the element's share of CSC column (b, m) at rows (g_j, n):
  column_run(b, m)[rowpos[t][j] + n] += H_e[(i,m),(j,n)]
and column_run(b, m) is the contiguous slice values[9*off_b + m*3*deg_b : ... + 3*deg_b]."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle  # noqa: E402
from polyfem_b200 import dist as D, mesh as M  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    mesh = M.kuhn_cube(n, p, jitter=0.1)
    x = M.random_displacement(mesh)
    ref = pyoracle.problem_from_mesh(mesh, "NeoHookean")
    H = ref.assemble_hessian(x)
    conn = mesh.conn.astype(np.int64)
    ne, nl = conn.shape
    nb = mesh.n_bases
    adj_off, adj = D.block_pattern_numpy(mesh.conn, nb)
    deg = np.diff(adj_off)

    # incidence lists (counting sort by node)
    node = conn.reshape(-1)
    order = np.argsort(node, kind="stable")
    inc = ((order // nl) << 4) | (order % nl)
    inc_off = np.zeros(nb + 1, dtype=np.int64)
    np.add.at(inc_off, node + 1, 1)
    inc_off = np.cumsum(inc_off)

    values = np.zeros(9 * adj.size)
    rowpos_all = np.zeros((inc.size, nl), dtype=np.int32)
    for b in range(nb):
        rows = adj[adj_off[b]:adj_off[b + 1]]
        run = [values[9 * adj_off[b] + m * 3 * deg[b]: 9 * adj_off[b] + (m + 1) * 3 * deg[b]] for m in range(3)]
        for t in range(inc_off[b], inc_off[b + 1]):
            e, i = int(inc[t]) >> 4, int(inc[t]) & 15
            assert conn[e, i] == b
            k = np.searchsorted(rows, conn[e])
            assert np.array_equal(rows[k], conn[e])
            rowpos_all[t] = 3 * k
            He = ref.local_hessian(e, x).reshape(3 * nl, 3 * nl)  # [(i,m),(j,n)]
            for m in range(3):
                row = He[3 * i + m].reshape(nl, 3)  # [j][n]
                for j in range(nl):
                    run[m][3 * k[j]: 3 * k[j] + 3] += row[j]
    # compare with the oracle's CSC values (same layout: column (b,m) = contiguous run)
    scale = np.abs(H.values).max()
    err = np.abs(values - H.values).max() / scale
    print(f"P{p} n={n}: {ne} elements, {nb} nodes, nnz {H.values.size}; column-lane assembly vs oracle: max err {err:.2e} of the largest entry")
    assert err <= 1e-13
    R = np.diff(inc_off)
    print(f"incident elements per node: min {R.min()}, mean {R.mean():.2f}, max {R.max()}; row nodes per column: mean {deg.mean():.1f}, max {deg.max()}")
    print(f"column buffer per node (9*deg doubles): mean {72 * deg.mean() / 1024:.2f} KB, max {72 * deg.max() / 1024:.2f} KB; "
          f"tables: inc {4 * inc.size / 1e6:.2f} MB, rowpos (uint8) {rowpos_all.size / 1e6:.2f} MB")
    hist = np.bincount(R)
    print("histogram of incident-element counts:", {int(c): int(v) for c, v in enumerate(hist) if v})


if __name__ == "__main__":
    main()
