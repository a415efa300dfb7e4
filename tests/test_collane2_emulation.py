"""CPU emulation of the owner-computes (column-lane) kernels against the oracle.

`polyfem_b200/csrc/pfa_collane2.h` holds the element record, the per-lane column math and the host schedule of the
owner-computes kernels (`pfa_collane2.cu`, the default NeoHookean P1/P2 path). `tests/collane2_emul.cpp` walks
chunks / groups / steps / triples / lanes on the CPU with exactly those functions and tables (strips start as NaN: first
contributions are stored, not added); here its energy, gradient and `values[]` are compared with the oracle
(pattern layout = the library's: column 3b+m at 9*adj_off[b] + m*3*deg(b)), 1e-12 like the GPU parity tests.
Each entry is summed in a fixed order, so two runs must agree bit for bit."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close
from polyfem_b200 import mesh as M, tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_emul(out):
    """compiles tests/collane2_emul.cpp (the kernels' own header on the CPU) into the shared library `out`"""
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", os.path.join(ROOT, "tests", "collane2_emul.cpp"), "-o", out],
                   check=True)
    return out


def load_emul(path):
    lib = ctypes.CDLL(path)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    lib.collane2_emulate.argtypes = [ctypes.c_int] * 4 + [ip, ip, ip, dp, dp, dp, dp, dp, dp, ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.POINTER(ctypes.c_uint8), ctypes.c_double, dp, dp, dp, ctypes.POINTER(ctypes.c_int64)]
    return lib


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    return load_emul(build_emul(str(tmp_path_factory.mktemp("collane") / "libcollane2_emul.so")))


def node_adjacency(mesh):
    """adj_off, adj (ascending, with the node itself): the node-block pattern the library builds on the device."""
    nl = mesh.conn.shape[1]
    a = np.repeat(mesh.conn, nl, axis=1).reshape(-1)
    b = np.tile(mesh.conn, (1, nl)).reshape(-1)
    pairs = np.unique(np.stack([a, b], axis=1), axis=0)  # sorted by (a, b)
    adj_off = np.zeros(mesh.n_bases + 1, dtype=np.int32)
    np.add.at(adj_off, pairs[:, 0] + 1, 1)
    return np.cumsum(adj_off).astype(np.int32), np.ascontiguousarray(pairs[:, 1], dtype=np.int32)


def run_emulation(lib, oracle, mesh, x, small_rows, structured=0, chunk_steps=24, owned=None, scale=1.0, lam_mu=None):
    t = tables.reference_tables(mesh.p)
    prob = oracle.problem_from_mesh(mesh, "NeoHookean")
    ne, nl, nq = mesh.n_elements, mesh.conn.shape[1], t["weights"].size
    jit, det = np.zeros((ne, 9)), np.zeros(ne)
    for e in range(ne):
        d, j, _ = prob.assembly_values(e)
        jit[e], det[e] = j[0].reshape(9), d[0]
    adj_off, adj = node_adjacency(mesh)
    conn = np.ascontiguousarray(mesh.conn, dtype=np.int32)
    if lam_mu is None:
        l0, m0 = M.lame_from_E_nu(1e5, 0.3)
        lam_mu = (np.full(ne, l0), np.full(ne, m0))
    lam, mu = (np.ascontiguousarray(a, dtype=np.float64) for a in lam_mu)
    mstride = lam.size // ne
    nnz = 9 * adj.size
    energy, grad, values, stats = np.zeros(1), np.full(mesh.n_bases * 3, np.nan), np.full(nnz, np.nan), np.zeros(8, dtype=np.int64)
    P = lambda a, ty=ctypes.c_double: a.ctypes.data_as(ctypes.POINTER(ty))  # noqa: E731
    own = None if owned is None else P(np.ascontiguousarray(owned, dtype=np.uint8), ctypes.c_uint8)
    rc = lib.collane2_emulate(nl, nq, ne, mesh.n_bases, P(conn, ctypes.c_int32), P(adj_off, ctypes.c_int32), P(adj, ctypes.c_int32), P(jit), P(det),
                              P(np.ascontiguousarray(t["weights"])), P(np.ascontiguousarray(t["grad"])), P(lam), P(mu), mstride, P(np.ascontiguousarray(x)),
                              small_rows, chunk_steps, structured, own, scale, P(energy), P(grad), P(values), P(stats, ctypes.c_int64))
    assert rc == 0, f"emulation failed with code {rc}"
    return prob, float(energy[0]), grad, values, stats


@pytest.mark.parametrize("p,n,small_rows,structured", [(1, 3, 96, 0), (2, 2, 96, 0), (2, 3, 96, 1), (2, 3, 1000, 0), (2, 2, 0, 1), (2, 5, 96, 1), (1, 7, 96, 0),
                                                       (2, 3, 96, 2), (2, 4, 0, 2), (2, 3, 96, 3), (2, 4, 0, 3)])
def test_column_lane_data_flow_equals_oracle(emul, oracle, p, n, small_rows, structured):
    """structured = 1: the column step that uses the structural zeros of the P2 reference gradients (P2S); 2: the vertex-weighted
    entry step for the symmetric 4-point rule (P2Z); 3: the same, streamed."""
    mesh = M.kuhn_cube(n, p, jitter=0.2)
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    prob, e, g, v, stats = run_emulation(emul, oracle, mesh, x, small_rows, structured)
    H = prob.assemble_hessian(x)
    assert v.size == H.values.size and not np.isnan(v).any()  # every column was flushed exactly once
    e_ref = prob.assemble_energy(x)
    assert abs(e - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(g, prob.assemble_gradient(x))
    assert_values_close(H.outer, H.inner, v, H.values)
    # schedule statistics: every (element, node) incidence is worked on exactly once
    assert stats[5] == mesh.n_elements * mesh.conn.shape[1]
    if small_rows == 0:
        assert stats[0] == 0 and stats[1] > 0
    if small_rows == 1000:
        assert stats[1] == 0
    # fixed summation order: bitwise reproducible
    _, e2, g2, v2, _ = run_emulation(emul, oracle, mesh, x, small_rows, structured)
    assert e2 == e and np.array_equal(g2, g) and np.array_equal(v2, v)


def test_column_lane_nan_propagation(emul, oracle):
    mesh = M.kuhn_cube(2, 2)
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    nodes = mesh.conn[5]
    x.reshape(-1, 3)[nodes[1]] += 3.0 * (mesh.node_xyz[nodes[0]] - mesh.node_xyz[nodes[1]])
    prob, e, g, v, _ = run_emulation(emul, oracle, mesh, x, 96)
    H = prob.assemble_hessian(x)
    assert np.isnan(e) and np.isnan(prob.assemble_energy(x))
    assert_vector_close(g, prob.assemble_gradient(x))
    assert_values_close(H.outer, H.inner, v, H.values)


@pytest.mark.parametrize("world", [2, 3])
def test_owned_columns_need_no_exchange(emul, oracle, world):
    """The multi-GPU consequence of owner-computes columns: the element partition of polyfem_b200/dist.py gives
    every rank all elements incident to its owned nodes (own + ghost elements), so the column-lane data flow run on a rank's
    local mesh already yields the FINISHED columns and gradient entries of the nodes it owns - no interface exchange.
    Checked per rank against the single-mesh oracle, matching rows through global node ids (local numbering differs)."""
    from polyfem_b200 import dist as pdist
    mesh = M.kuhn_cube(3, 2, jitter=0.2)
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    H = ref.assemble_hessian(x).to_scipy().tocsc()
    g_ref = ref.assemble_gradient(x)
    owned_total = 0
    for rank in range(world):
        part = pdist.partition_elements(mesh, rank, world)
        # the local mesh with the geometry of own AND ghost elements (global element ids in partition order)
        bounds = pdist.element_ranges(mesh.n_elements, world)
        elem_rank = np.searchsorted(np.asarray(bounds[1:]), np.arange(mesh.n_elements), side="right")
        node_owner = np.full(mesh.n_bases, world)
        np.minimum.at(node_owner, mesh.conn.reshape(-1), np.repeat(elem_rank, mesh.conn.shape[1]))
        ghost = np.nonzero((node_owner[mesh.conn] == rank).any(axis=1) & (elem_rank != rank))[0]
        elems = np.concatenate([part.own_elements, ghost])
        assert elems.size == part.conn.shape[0] and np.array_equal(part.l2g[part.conn], mesh.conn[elems])

        class Local:  # what run_emulation reads of a mesh
            p, conn, vertices, n_bases, n_elements = mesh.p, part.conn, mesh.vertices[elems], part.n_bases, elems.size
        x_loc = np.ascontiguousarray(x.reshape(-1, 3)[part.l2g].reshape(-1))
        # only the columns of owned nodes are scheduled (what a rank of the multi-GPU path does); the rest stays untouched
        _, _, g, v, _ = run_emulation(emul, oracle, Local, x_loc, 96, 1, owned=(part.owner == rank))
        adj_off, adj = node_adjacency(Local)
        for b in np.nonzero(part.owner != rank)[0]:
            assert np.isnan(v[9 * adj_off[b]:9 * adj_off[b + 1]]).all() and np.isnan(g[3 * b:3 * b + 3]).all()
        for b in np.nonzero(part.owner == rank)[0]:
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
            owned_total += 1
    assert owned_total == mesh.n_bases  # every node is owned by exactly one rank


def test_scale_chunks_and_per_quadrature_point_materials(emul, oracle):
    """Form weight (scale), the chunk size of the group hand-out (results must not depend on it, bit for bit) and
    (lambda, mu) given per (element, quadrature point)."""
    mesh = M.kuhn_cube(2, 2, jitter=0.2)
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    rng = np.random.default_rng(3)
    ne, nq = mesh.n_elements, 4
    l0, m0 = M.lame_from_E_nu(1e5, 0.3)
    # the oracle takes one (lambda, mu) per element: replicate them over the quadrature points for the [element][qp] layout
    lam_e, mu_e = l0 * (1.0 + 0.3 * rng.random(ne)), m0 * (1.0 + 0.3 * rng.random(ne))
    lam, mu = np.repeat(lam_e[:, None], nq, axis=1), np.repeat(mu_e[:, None], nq, axis=1)
    # (chunk_steps = 10 also switches the emulation's schedule to spatial buckets of 5 elements: same sums per entry in another
    # node order - the values agree to rounding, not bit for bit, with the unbucketed order)
    _, e_b, g_b, v_b, _ = run_emulation(emul, oracle, mesh, x, 96, 1, chunk_steps=10, scale=0.25, lam_mu=(lam, mu))
    prob, e, g, v, _ = run_emulation(emul, oracle, mesh, x, 96, 1, chunk_steps=1, scale=0.25, lam_mu=(lam, mu))
    assert np.allclose(v_b, v, rtol=1e-13, atol=1e-13 * np.abs(v).max()) and np.allclose(g_b, g, rtol=1e-13, atol=1e-13 * np.abs(g).max())
    _, e2, g2, v2, st = run_emulation(emul, oracle, mesh, x, 96, 1, chunk_steps=10 ** 6, scale=0.25, lam_mu=(lam, mu))
    assert e2 == e and np.array_equal(g2, g) and np.array_equal(v2, v)
    from polyfem_b200 import tables as T
    t = T.reference_tables(2)
    prob = type(prob)("NeoHookean", mesh.conn, mesh.vertices, mesh.n_bases, t["points"], t["weights"], t["grad"], lam=lam_e, mu=mu_e, basis_order=2)
    H = prob.assemble_hessian(x)
    assert abs(e - 0.25 * prob.assemble_energy(x)) <= REL_TOL * abs(e)
    assert_vector_close(g, 0.25 * prob.assemble_gradient(x))
    assert_values_close(H.outer, H.inner, v, 0.25 * H.values)


@pytest.mark.parametrize("p,structured", [(1, 0), (2, 2)])
def test_unstructured_mesh(emul, oracle, p, structured):
    """A Delaunay mesh with elements in random order: vertex valences from 1 to 30+ incident tets, odd incidence counts (idle
    half-steps), node degrees the Kuhn cube does not have."""
    from unstructured import delaunay_mesh
    mesh = delaunay_mesh(150, p)
    x = 0.02 * np.random.default_rng(1).uniform(-1, 1, mesh.n_bases * 3) * mesh.h
    prob, e, g, v, stats = run_emulation(emul, oracle, mesh, x, 96, structured, chunk_steps=10)
    H = prob.assemble_hessian(x)
    assert not np.isnan(v).any() and stats[5] == mesh.n_elements * mesh.conn.shape[1]
    assert abs(e - prob.assemble_energy(x)) <= REL_TOL * abs(e)
    assert_vector_close(g, prob.assemble_gradient(x))
    assert_values_close(H.outer, H.inner, v, H.values)
