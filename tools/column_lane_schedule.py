#!/usr/bin/env python
"""Feasibility numbers for the host scheduler of the "column lanes" design, CPU only.

  python tools/column_lane_schedule.py [cells per side, default 14] [order, default 2]

A warp has 10 slots (3 lanes each). A slot works through a queue of column nodes, one incident element per
step; nodes are taken in spatial (Morton) order so that the records a warp needs stay close. Greedy packing:
every new node goes to the slot with the least work so far, within a window of nodes per warp ("unit").
Reports, per unit size: lane utilisation (useful slot-steps / (10 x steps of the longest slot)), the shared
memory a warp needs for its lane-private strips (rows = 3*deg of the largest node a slot group holds, 128 B
per row and 5-slot group) and the number of distinct elements a warp touches per step (record reads)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polyfem_b200 import dist as D, mesh as M  # noqa: E402


def morton(xyz):
    q = np.clip(((xyz - xyz.min(0)) / (xyz.max(0) - xyz.min(0)) * 1023).astype(np.int64), 0, 1023)

    def spread(v):
        r = np.zeros_like(v)
        for b in range(10):
            r |= ((v >> b) & 1) << (3 * b)
        return r
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    mesh = M.kuhn_cube(n, p)
    conn = mesh.conn.astype(np.int64)
    nb = mesh.n_bases
    adj_off, adj = D.block_pattern_numpy(mesh.conn, nb)
    deg = np.diff(adj_off)
    R = np.bincount(conn.reshape(-1), minlength=nb)
    order = np.argsort(morton(mesh.node_xyz), kind="stable")
    print(f"P{p} n={n}: {mesh.n_elements} elements, {nb} nodes; incident elements per node mean {R.mean():.2f} max {R.max()}, row nodes mean {deg.mean():.1f} max {deg.max()}")
    for unit_nodes in (16, 32, 64, 128):
        steps_tot, useful_tot, smem = 0, 0, []
        for u0 in range(0, nb - unit_nodes + 1, unit_nodes * 5):  # sample units
            nodes = order[u0:u0 + unit_nodes]
            nodes = nodes[np.argsort(-R[nodes], kind="stable")]  # longest first
            load = np.zeros(10, dtype=np.int64)
            rows = np.zeros(10, dtype=np.int64)
            for b in nodes:
                s = int(np.argmin(load))
                load[s] += R[b]
                rows[s] = max(rows[s], 3 * deg[b])
            steps_tot += int(load.max())
            useful_tot += int(load.sum())
            smem.append(128 * (rows[:5].max() + rows[5:].max()))
        util = useful_tot / (10 * steps_tot)
        print(f"  unit of {unit_nodes:3d} nodes: lane utilisation {util * 30 / 32:.2f} (slots {util:.2f}), strips {np.mean(smem) / 1024:.1f} KB per warp (max {np.max(smem) / 1024:.1f})")


if __name__ == "__main__":
    main()
