"""Host logic of the multi-GPU path on CPU: world_size-2/3 gloo processes partition the mesh,
assemble their own elements with the oracle, run the interface exchange, and every rank's owned
columns / dofs must equal the single-process result. Second half: the owner-computes
form (no exchange; pfa_partition_create + the column-lane data flow per rank + the energy all-reduce),
which is what `bench.py --gpus N` runs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, p, out_dir, combined=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle
        from polyfem_b200 import dist as pdist, mesh as M, tables
        mesh = M.kuhn_cube(n, p, jitter=0.1)
        x = M.random_displacement(mesh)
        t = tables.reference_tables(p)
        lam, mu = M.lame_from_E_nu(1e5, 0.3)
        part = pdist.partition_elements(mesh, rank, world, align=6 * n * n if combined else 1)
        adj_off, adj = pdist.block_pattern_numpy(part.conn, part.n_bases)
        nb = part.n_bases
        # local assembly of own elements with the oracle, in local numbering
        own_conn = part.conn[:part.n_own_elements]
        prob = pyoracle.OracleProblem("NeoHookean", own_conn, part.vertices, nb, t["points"], t["weights"], t["grad"], lam=lam, mu=mu)
        x_loc = x.reshape(-1, 3)[part.l2g].reshape(-1)
        e_loc = prob.assemble_energy(x_loc)
        g_loc = prob.assemble_gradient(x_loc)
        H = prob.assemble_hessian(x_loc)
        # scatter the own-element CSC into the ghost-widened local pattern
        nnz = 9 * adj.size
        values = np.zeros(nnz)
        outer_w = np.zeros(3 * nb + 1, dtype=np.int64)
        deg = np.diff(adj_off)
        for b in range(nb):
            for c in range(3):
                outer_w[3 * b + c] = 9 * adj_off[b] + c * 3 * deg[b]
        outer_w[-1] = nnz
        for col in range(3 * nb):
            rows = H.inner[H.outer[col]:H.outer[col + 1]]
            b = col // 3
            lst = adj[adj_off[b]:adj_off[b + 1]]
            k = np.searchsorted(lst, rows // 3)
            values[outer_w[col] + 3 * k + rows % 3] = H.values[H.outer[col]:H.outer[col + 1]]
        e_t = torch.tensor([e_loc], dtype=torch.float64)
        g_t = torch.from_numpy(g_loc.copy())
        v_t = torch.from_numpy(values)
        if combined:  # values and gradient in one tensor, cuts on whole cell layers (bench.py's multi-GPU path)
            vg = torch.cat([v_t, g_t])
            ex = pdist.InterfaceExchange(None, part, rank, world, torch.device("cpu"), block_pattern=(adj_off, adj), grad_offset=nnz)
            ex.reduce_combined(e_t, vg)
            v_t, g_t = vg[:nnz].clone(), vg[nnz:].clone()
        else:
            ex = pdist.InterfaceExchange(None, part, rank, world, torch.device("cpu"), block_pattern=(adj_off, adj))
            ex.reduce(e_t, g_t, v_t)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), e=e_t.numpy(), g=g_t.numpy(), v=v_t.numpy(), adj_off=adj_off, adj=adj,
                 l2g=part.l2g, owner=part.owner, n_own=part.n_own_elements, n_ghost=part.n_ghost_elements)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,combined", [(2, False), (3, False), (2, True)])
def test_partition_exchange_matches_single_process(tmp_path, oracle, world, combined):
    from polyfem_b200 import mesh as M
    n, p = 3, 2
    port = 29650 + world + (10 if combined else 0)
    mp.spawn(_worker, args=(world, port, n, p, str(tmp_path), combined), nprocs=world, join=True)
    mesh = M.kuhn_cube(n, p, jitter=0.1)
    x = M.random_displacement(mesh)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    e_ref, g_ref, H_ref = ref.assemble_energy(x), ref.assemble_gradient(x), ref.assemble_hessian(x)
    Hs = H_ref.to_scipy().tocsc()
    owned_nodes_seen = np.zeros(mesh.n_bases, dtype=int)
    total_own = 0
    for r in range(world):
        d = np.load(os.path.join(tmp_path, f"rank{r}.npz"))
        total_own += int(d["n_own"])
        assert abs(d["e"][0] - e_ref) <= 1e-12 * abs(e_ref)
        l2g, owner, adj_off, adj = d["l2g"], d["owner"], d["adj_off"], d["adj"]
        mine = np.nonzero(owner == r)[0]
        owned_nodes_seen[l2g[mine]] += 1
        g = d["g"].reshape(-1, 3)
        assert np.abs(g[mine] - g_ref.reshape(-1, 3)[l2g[mine]]).max() <= 1e-12 * np.abs(g_ref).max()
        scale = np.abs(H_ref.values).max()
        for b in mine[:: max(1, mine.size // 60)]:
            gb = l2g[b]
            rows_l = adj[adj_off[b]:adj_off[b + 1]]
            deg = rows_l.size
            for c in range(3):
                col = Hs.getcol(3 * gb + c)
                # owned column must have exactly the global row set
                exp_rows = np.sort(col.indices)
                got_rows = np.sort((l2g[rows_l][:, None] * 3 + np.arange(3)[None, :]).reshape(-1))
                assert np.array_equal(exp_rows, got_rows)
                vals = d["v"][9 * adj_off[b] + c * 3 * deg: 9 * adj_off[b] + (c + 1) * 3 * deg].reshape(deg, 3)
                dense = np.zeros(3 * mesh.n_bases)
                dense[(l2g[rows_l][:, None] * 3 + np.arange(3)[None, :]).reshape(-1)] = vals.reshape(-1)
                assert np.abs(dense - col.toarray().ravel()).max() <= 1e-12 * scale
    assert np.all(owned_nodes_seen == 1)  # every node has exactly one owner
    assert total_own == mesh.n_elements


def _owner_worker(rank, world, port, n, p, out_dir, emul_path):
    """One rank of the owner-computes form (the default of `bench.py --gpus N`): pfa_partition_create, the column-lane
    data flow on the rank's own + ghost elements for the nodes it owns (CPU emulation of the kernels), and the one collective of
    a step - the all-reduce of the energy of the own elements."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle
        from polyfem_b200 import dist as pdist, mesh as M
        from test_collane2_emulation import load_emul, node_adjacency, run_emulation
        mesh = M.kuhn_cube(n, p, jitter=0.1)
        x = M.random_displacement(mesh)[: mesh.n_bases * 3]
        part = pdist.partition_owner_computes(mesh, rank, world)

        class Local:  # the rank's mesh: own elements first, then the ghost elements around its owned nodes
            conn, vertices, n_bases, n_elements = part.conn, part.vertices, part.n_bases, part.conn.shape[0]
        Local.p = mesh.p
        x_loc = np.ascontiguousarray(x.reshape(-1, 3)[part.l2g].reshape(-1))
        prob, _, g, v, _ = run_emulation(load_emul(emul_path), pyoracle, Local, x_loc, 96, 3 if p == 2 else 0, owned=part.owned)
        e_own = float(prob.assemble_energy_per_element(x_loc)[: part.n_own_elements].sum())  # ghost elements belong to other ranks
        e_t = torch.tensor([e_own], dtype=torch.float64)
        dist.all_reduce(e_t, op=dist.ReduceOp.SUM)
        adj_off, adj = node_adjacency(Local)
        np.savez(os.path.join(out_dir, f"owner_rank{rank}.npz"), e=e_t.numpy(), g=g, v=v, adj_off=adj_off, adj=adj, l2g=part.l2g, owned=part.owned,
                 n_own=part.n_own_elements)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,p,n", [(2, 2, 3), (3, 1, 4)])
def test_owner_computes_step_matches_single_process(tmp_path, oracle, world, p, n):
    from polyfem_b200 import mesh as M
    from test_collane2_emulation import build_emul
    emul_path = build_emul(str(tmp_path / "libcollane2_emul.so"))
    mp.spawn(_owner_worker, args=(world, 29700 + world + p, n, p, str(tmp_path), emul_path), nprocs=world, join=True)
    mesh = M.kuhn_cube(n, p, jitter=0.1)
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    e_ref, g_ref = ref.assemble_energy(x), ref.assemble_gradient(x).reshape(-1, 3)
    Hs = ref.assemble_hessian(x).to_scipy().tocsc()
    scale = np.abs(Hs.data).max()
    seen = np.zeros(mesh.n_bases, dtype=int)
    total_own = 0
    for r in range(world):
        d = np.load(os.path.join(tmp_path, f"owner_rank{r}.npz"))
        total_own += int(d["n_own"])
        assert abs(d["e"][0] - e_ref) <= 1e-12 * abs(e_ref)  # every rank holds the all-reduced energy
        l2g, owned, adj_off, adj = d["l2g"], d["owned"], d["adj_off"], d["adj"]
        g, v = d["g"].reshape(-1, 3), d["v"]
        for b in range(l2g.size):
            sl = slice(9 * adj_off[b], 9 * adj_off[b + 1])
            if not owned[b]:
                assert np.isnan(v[sl]).all() and np.isnan(g[b]).all()  # columns / dofs of other ranks are never written
                continue
            gb = int(l2g[b])
            seen[gb] += 1
            assert np.abs(g[b] - g_ref[gb]).max() <= 1e-12 * np.abs(g_ref).max()
            rows_l = adj[adj_off[b]:adj_off[b + 1]]
            deg = rows_l.size
            rows_g = (l2g[rows_l][:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
            for c in range(3):
                col = Hs.getcol(3 * gb + c)
                assert np.array_equal(np.sort(col.indices), np.sort(rows_g))  # the full global row set is present on the owner
                vals = v[9 * adj_off[b] + c * 3 * deg: 9 * adj_off[b] + (c + 1) * 3 * deg]
                assert np.abs(vals - np.asarray(Hs[rows_g, 3 * gb + c].todense()).ravel()).max() <= 1e-12 * scale
    assert np.all(seen == 1) and total_own == mesh.n_elements  # every node finished by exactly one rank, every element owned once

