"""The host side of pfa_create that needs no device, through its C entry points (include/pfa.h: pfa_host_pattern_*,
pfa_host_element_order): the sparsity pattern + slot map that replace the first-call path of SparseMatrixCache
(reference utils/MatrixCache.cpp:88-213) and the internal element order. Checked against an independent numpy restatement, against
the oracle's CSC pattern (bit for bit, after the size x size expansion the device performs) and by invariants."""
import numpy as np
import pytest

from polyfem_b200 import capi, mesh as M


def _numpy_adjacency(conn, n_bases):
    nl = conn.shape[1]
    rows = np.repeat(conn, nl, axis=1).reshape(-1)
    cols = np.tile(conn, (1, nl)).reshape(-1)
    pairs = np.unique(np.stack([cols, rows], axis=1), axis=0)  # sorted by column node, then row node
    adj_off = np.zeros(n_bases + 1, dtype=np.int64)
    np.add.at(adj_off, pairs[:, 0] + 1, 1)
    return np.cumsum(adj_off).astype(np.int32), pairs[:, 1].astype(np.int32)


def _expand_csc(adj_off, adj, size):
    """what expand_inner_kernel does on the device: column (b, n) lists rows size*a + m for the row nodes a of b, ascending"""
    deg = np.diff(adj_off)
    outer = np.concatenate([[0], np.cumsum(np.repeat(deg * size, size))]).astype(np.int32)
    inner = []
    for b in range(adj_off.size - 1):
        rows = (size * adj[adj_off[b]:adj_off[b + 1]][:, None] + np.arange(size)[None, :]).reshape(-1)
        inner.extend([rows] * size)
    return outer, np.concatenate(inner).astype(np.int32)


@pytest.mark.parametrize("p,n,jitter", [(1, 3, 0.0), (2, 2, 0.2), (3, 2, 0.0), (4, 1, 0.0), (2, 4, 0.1)])
def test_pattern_and_slot_map(oracle, p, n, jitter):
    mesh = M.kuhn_cube(n, p, jitter=jitter)
    adj_off, adj, slot = capi.host_pattern(mesh.conn, mesh.n_bases)
    ref_off, ref_adj = _numpy_adjacency(mesh.conn, mesh.n_bases)
    assert adj_off.dtype == np.int32 and np.array_equal(adj_off, ref_off) and np.array_equal(adj, ref_adj)
    # symmetric, every node lists itself, rows ascending inside a column
    for b in range(mesh.n_bases):
        col = adj[adj_off[b]:adj_off[b + 1]]
        assert (np.diff(col) > 0).all() and b in col
    # slot[e][i][j]: position of row node conn[e][i] in the column list of node conn[e][j]
    ci = mesh.conn[:, :, None].repeat(mesh.conn.shape[1], axis=2)
    cj = mesh.conn[:, None, :].repeat(mesh.conn.shape[1], axis=1)
    assert np.array_equal(adj[slot], ci)
    assert (slot >= adj_off[cj]).all() and (slot < adj_off[cj + 1]).all()
    # the CSC pattern of the assembled matrix (oracle = restatement of SparseMatrixCache, bit-pinned against the reference's
    # MatrixCache.cpp in tests/test_oracle_cache_vs_reference.py) is this pattern expanded by size x size
    if mesh.n_elements <= 400:
        prob = oracle.problem_from_mesh(mesh, "NeoHookean")
        H = prob.assemble_hessian(np.zeros(mesh.n_bases * 3))
        outer, inner = _expand_csc(adj_off, adj, 3)
        assert outer.tobytes() == H.outer.astype(np.int32).tobytes() and inner.tobytes() == H.inner.astype(np.int32).tobytes()


def test_pattern_with_unused_nodes_and_bad_input():
    mesh = M.kuhn_cube(2, 1)
    adj_off, adj, _ = capi.host_pattern(mesh.conn, mesh.n_bases + 3)  # three nodes without elements: empty columns
    assert (np.diff(adj_off)[-3:] == 0).all() and adj.size == adj_off[-1]
    bad = mesh.conn.copy()
    bad[0, 0] = mesh.n_bases  # out of range
    with pytest.raises(capi.PfaError) as ei:
        capi.host_pattern(bad, mesh.n_bases)
    assert ei.value.code == capi.PFA_ERR_INVALID


@pytest.mark.parametrize("n", [1, 3, 8, 12])
def test_element_order_is_a_locality_preserving_permutation(n):
    mesh = M.kuhn_cube(n, 1, jitter=0.1)
    perm = capi.host_element_order(mesh.vertices)
    assert np.array_equal(np.sort(perm), np.arange(mesh.n_elements))
    assert np.array_equal(perm, capi.host_element_order(mesh.vertices))  # deterministic
    if n >= 8:
        # a window of 48 consecutive elements (8 cells) of the curve is a compact block; in the caller's x-slowest order it is a
        # row of cells: compare the mean extent of the windows
        cen = mesh.vertices.reshape(-1, 4, 3).mean(axis=1)

        def extent(c):
            w = c[: (c.shape[0] // 48) * 48].reshape(-1, 48, 3)
            return (w.max(axis=1) - w.min(axis=1)).sum(axis=1).mean()
        assert extent(cen[perm]) < 0.8 * extent(cen)
