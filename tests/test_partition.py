"""pfa_partition_create (include/pfa.h): the element partition of the multi-GPU owner-computes path, on the CPU.

Invariants: contiguous own blocks that tile the element range, every node owned by exactly one rank (the rank of the
first element touching it), every element incident to an owned node present on that rank (own or ghost), balanced
incidence counts; and the consequence the multi-GPU path rests on - the column-lane data flow run on a rank's local mesh
yields the FINISHED columns and gradient entries of the nodes it owns (checked against the single-mesh oracle through
global node ids), so no interface exchange is needed."""
import numpy as np
import pytest

from helpers import REL_TOL
from polyfem_b200 import capi, dist as pdist, mesh as M
from test_collane2_emulation import emul, node_adjacency, run_emulation  # noqa: F401  (fixture + helpers)


@pytest.mark.parametrize("p,n,world", [(1, 5, 2), (2, 4, 3), (2, 6, 8), (1, 3, 5)])
def test_partition_invariants(p, n, world):
    mesh = M.kuhn_cube(n, p, jitter=0.1)
    inc = np.bincount(mesh.conn.reshape(-1), minlength=mesh.n_bases)
    first = np.full(mesh.n_bases, mesh.n_elements)
    np.minimum.at(first, mesh.conn.reshape(-1), np.repeat(np.arange(mesh.n_elements), mesh.conn.shape[1]))
    owner = np.full(mesh.n_bases, -1)
    own_blocks, work = [], []
    for r in range(world):
        q = capi.partition(mesh.conn, mesh.n_bases, world, r)
        own = q["elements"][:q["n_own"]]
        assert q["n_own"] == 0 or np.array_equal(own, np.arange(own[0], own[0] + own.size))
        own_blocks.append(own)
        assert np.array_equal(q["l2g"][q["conn"]], mesh.conn[q["elements"]])  # local numbering is consistent
        mine = q["l2g"][q["owned"] == 1]
        assert (owner[mine] == -1).all()
        owner[mine] = r
        if own.size:
            assert ((first[mine] >= own[0]) & (first[mine] <= own[-1])).all()  # owner = rank of the first touching element
        # every element incident to an owned node is on this rank
        touches = np.isin(mesh.conn, mine).any(axis=1)
        assert np.isin(np.nonzero(touches)[0], q["elements"]).all()
        # ... and no ghost element is superfluous
        assert touches[q["elements"][q["n_own"]:]].all()
        work.append(int(inc[mine].sum()))
    assert (owner >= 0).all()
    assert np.array_equal(np.concatenate(own_blocks), np.arange(mesh.n_elements))
    assert sum(work) == mesh.conn.size
    assert max(work) <= 1.0 * sum(work) / world + inc.max() * mesh.conn.shape[1]  # balanced up to one element's worth


def test_partition_rejects_bad_arguments():
    mesh = M.kuhn_cube(2, 1)
    with pytest.raises(capi.PfaError):
        capi.partition(mesh.conn, mesh.n_bases, 2, 2)
    bad = mesh.conn.copy()
    bad[0, 0] = mesh.n_bases
    with pytest.raises(capi.PfaError):
        capi.partition(bad, mesh.n_bases, 2, 0)


@pytest.mark.parametrize("world", [2, 3])
def test_owned_columns_are_finished_without_exchange(emul, oracle, world):  # noqa: F811
    mesh = M.kuhn_cube(3, 2, jitter=0.2)
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    H = ref.assemble_hessian(x).to_scipy().tocsc()
    g_ref = ref.assemble_gradient(x)
    e_ref = ref.assemble_energy_per_element(x)
    seen = 0
    for rank in range(world):
        part = pdist.partition_owner_computes(mesh, rank, world)

        class Local:  # what run_emulation reads of a mesh
            p, conn, vertices, n_bases, n_elements = mesh.p, part.conn, part.vertices, part.n_bases, part.conn.shape[0]
        x_loc = np.ascontiguousarray(x.reshape(-1, 3)[part.l2g].reshape(-1))
        _, e, g, v, _ = run_emulation(emul, oracle, Local, x_loc, 96, 1, owned=part.owned)
        adj_off, adj = node_adjacency(Local)
        for b in range(part.n_bases):
            sl = slice(9 * adj_off[b], 9 * adj_off[b + 1])
            if not part.owned[b]:
                assert np.isnan(v[sl]).all() and np.isnan(g[3 * b:3 * b + 3]).all()  # untouched
                continue
            gb = int(part.l2g[b])
            deg = adj_off[b + 1] - adj_off[b]
            rows_g = part.l2g[adj[adj_off[b]:adj_off[b + 1]]]
            for m in range(3):
                col = H[:, 3 * gb + m]
                assert col.nnz == 3 * deg, "an owned node misses neighbours on its rank"
                mine = v[9 * adj_off[b] + m * 3 * deg: 9 * adj_off[b] + (m + 1) * 3 * deg].reshape(deg, 3)
                want = np.asarray(H[(3 * rows_g[:, None] + np.arange(3)[None, :]).reshape(-1), 3 * gb + m].todense()).reshape(deg, 3)
                assert np.abs(mine - want).max() <= REL_TOL * np.abs(want).max()
            assert np.abs(g[3 * b:3 * b + 3] - g_ref[3 * gb:3 * gb + 3]).max() <= REL_TOL * np.abs(g_ref).max()
            seen += 1
    assert seen == mesh.n_bases
    assert e_ref.size == mesh.n_elements
