"""Property tests of the reference restated on the synthetic cube:
  tests/test_assembler.cpp:148-315  closed-form NeoHookean == autodiff NeoHookean (1e-12)
  tests/test_assembler.cpp:24-84    NL Hessian of LinearElasticity == linear stiffness (1e-8)
  tests/test_form_derivatives.cpp:169-230  gradient / Hessian vs finite differences
and the structural facts of SURVEY.md §8 (node counts, nnz closed forms, first/steady call).
"""
import numpy as np
import pytest

from helpers import make_case


@pytest.mark.parametrize("p", [1, 2])
def test_neohookean_closed_form_equals_autodiff(oracle, p):
    mesh, x, _ = make_case(2, p, jitter=0.2)
    nh = oracle.problem_from_mesh(mesh, "NeoHookean")
    rng = np.random.default_rng(1)
    for e in rng.choice(mesh.n_elements, 6, replace=False):
        ea, eb = nh.local_energy(e, x), nh.local_energy(e, x, True)
        ga, gb = nh.local_gradient(e, x), nh.local_gradient(e, x, True)
        ha, hb = nh.local_hessian(e, x), nh.local_hessian(e, x, True)
        assert abs(ea - eb) <= 1e-12 * abs(ea)
        assert np.abs(ga - gb).max() <= 1e-12 * np.abs(ga).max()
        assert np.abs(ha - hb).max() <= 1e-12 * np.abs(ha).max()
        assert np.abs(ha - ha.T).max() <= 1e-12 * np.abs(ha).max()


def test_neohookean_nan_iff_inverted(oracle):
    mesh, x, _ = make_case(2, 1)
    nh = oracle.problem_from_mesh(mesh, "NeoHookean")
    assert np.isfinite(nh.assemble_energy(x))
    xi = x.copy()
    nodes = mesh.conn[7]
    # collapse element 7 through its opposite face: det F < 0 -> log gives NaN
    xi.reshape(-1, 3)[nodes[1]] += 3.0 * (mesh.node_xyz[nodes[0]] - mesh.node_xyz[nodes[1]])
    epe = nh.assemble_energy_per_element(xi)
    assert np.isnan(epe).any() and np.isnan(nh.assemble_energy(xi))
    assert np.isnan(nh.assemble_gradient(xi)).any()


@pytest.mark.parametrize("p", [1, 2])
def test_linear_elasticity_nl_hessian_equals_stiffness(oracle, p):
    mesh, _, _ = make_case(2, p, jitter=0.2)
    le = oracle.problem_from_mesh(mesh, "LinearElasticity", n_threads=2)
    K = le.assemble()
    rng = np.random.default_rng(0)
    for _ in range(3):  # the reference uses 10 random displacements; cache reused across calls
        x = rng.uniform(-1, 1, le.ndof)
        H = le.assemble_hessian(x)
        assert np.array_equal(K.outer, H.outer) and np.array_equal(K.inner, H.inner)
        assert np.abs(K.values - H.values).max() < 1e-8
    x = rng.uniform(-1, 1, le.ndof)
    Ks = K.to_scipy()
    assert abs(le.assemble_energy(x) - 0.5 * x @ (Ks @ x)) < 1e-10 * abs(x @ (Ks @ x))
    assert np.abs(le.assemble_gradient(x) - Ks @ x).max() < 1e-10 * np.abs(Ks @ x).max()


@pytest.mark.parametrize("p", [1, 2])
def test_derivatives_vs_finite_differences(oracle, p):
    mesh, x, _ = make_case(2, p, jitter=0.1)
    nh = oracle.problem_from_mesh(mesh, "NeoHookean")
    rng = np.random.default_rng(3)
    d = rng.standard_normal(x.size)
    eps = 1e-7 * mesh.h  # displacement scale is 0.05 h
    g = nh.assemble_gradient(x)
    H = nh.assemble_hessian(x).to_scipy()
    fd_e = (nh.assemble_energy(x + eps * d) - nh.assemble_energy(x - eps * d)) / (2 * eps)
    assert abs(fd_e - g @ d) <= 1e-4 * abs(g @ d)
    fd_g = (nh.assemble_gradient(x + eps * d) - nh.assemble_gradient(x - eps * d)) / (2 * eps)
    assert np.abs(fd_g - H @ d).max() <= 1e-4 * np.abs(H @ d).max()


@pytest.mark.parametrize("p,n", [(1, 4), (2, 3), (3, 2), (4, 1)])
def test_pattern_closed_forms(oracle, p, n):
    mesh, x, _ = make_case(n, p)
    assert mesh.n_bases == (p * n + 1) ** 3
    assert mesh.n_elements == 6 * n ** 3
    lap = oracle.problem_from_mesh(mesh, "Laplacian")
    L = lap.assemble()
    if p == 1:
        assert L.nnz == 15 * n ** 3 + 21 * n ** 2 + 9 * n + 1
    if p == 2:
        assert L.nnz == 230 * n ** 3 + 138 * n ** 2 + 24 * n + 1
    le = oracle.problem_from_mesh(mesh, "LinearElasticity")
    assert le.assemble().nnz == 9 * L.nnz
    # Laplacian annihilates constants; stiffness is symmetric
    Ls = L.to_scipy()
    assert np.abs(Ls @ np.ones(Ls.shape[0])).max() < 1e-11
    assert abs(Ls - Ls.T).max() < 1e-12


def test_first_call_and_slot_map_call_agree(oracle):
    """First assemble_hessian builds the pattern from triplets (MatrixCache.cpp:88-100); later
    calls go through the slot map (:101-112). Same pattern, same values up to summation order."""
    mesh, x, _ = make_case(2, 2)
    nh = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=3)
    H1 = nh.assemble_hessian(x)
    H2 = nh.assemble_hessian(x)
    H3 = nh.assemble_hessian(0.5 * x)
    assert np.array_equal(H1.outer, H2.outer) and np.array_equal(H1.inner, H2.inner)
    assert np.array_equal(H1.inner, H3.inner)
    assert np.abs(H1.values - H2.values).max() <= 1e-13 * np.abs(H1.values).max()
    assert np.all(np.diff(H1.outer) >= 0)
    for c in range(0, H1.n, 37):
        col = H1.inner[H1.outer[c]:H1.outer[c + 1]]
        assert np.all(np.diff(col) > 0)


def test_threads_and_cache_modes_agree(oracle):
    mesh, x, _ = make_case(2, 2, jitter=0.15)
    a = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=1, use_cache=True)
    b = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=4, use_cache=False)
    assert abs(a.assemble_energy(x) - b.assemble_energy(x)) <= 1e-13 * abs(a.assemble_energy(x))
    ga, gb = a.assemble_gradient(x), b.assemble_gradient(x)
    assert np.abs(ga - gb).max() <= 1e-12 * np.abs(ga).max()
    Ha, Hb = a.assemble_hessian(x), b.assemble_hessian(x)
    assert np.array_equal(Ha.inner, Hb.inner)
    assert np.abs(Ha.values - Hb.values).max() <= 1e-12 * np.abs(Ha.values).max()


def test_project_to_psd_properties(oracle):
    rng = np.random.default_rng(5)
    a = rng.standard_normal((12, 12))
    a = a + a.T
    p = oracle.project_to_psd(a)
    w = np.linalg.eigvalsh(p)
    assert w.min() > -1e-10 * abs(w).max()
    assert np.abs(oracle.project_to_psd(p) - p).max() < 1e-10 * np.abs(p).max()  # idempotent
    spd = a @ a.T + np.eye(12)
    assert np.array_equal(oracle.project_to_psd(spd), spd)                        # identity on PSD input
    wa, va = np.linalg.eigh(a)
    ref = (va * np.maximum(wa, 0)) @ va.T
    assert np.abs(p - ref).max() < 1e-10 * np.abs(a).max()


def test_nodes_without_elements_give_empty_columns(oracle):
    """n_bases larger than the nodes the elements touch (a sub-mesh of a bigger space, the local meshes of a partition): the
    reference's matrix has size n_bases * dim with empty columns there (SparseMatrixCache::init(size), MatrixCache.cpp:28-47),
    the gradient entries are zero and the values are those of the mesh without the extra nodes."""
    from polyfem_b200 import mesh as M, tables
    mesh = M.kuhn_cube(2, 2, jitter=0.2)
    t = tables.reference_tables(2)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    nb = mesh.n_bases + 3
    x0 = M.random_displacement(mesh)[: mesh.n_bases * 3]
    x = np.concatenate([x0, np.full(9, 9.0)])
    big = oracle.OracleProblem("NeoHookean", mesh.conn, mesh.vertices, nb, t["points"], t["weights"], t["grad"], lam=lam, mu=mu)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    H, H0 = big.assemble_hessian(x), ref.assemble_hessian(x0)
    assert H.outer.size == 3 * nb + 1 and np.all(H.outer[-10:] == H0.values.size)
    assert np.array_equal(H.inner, H0.inner) and np.array_equal(H.values, H0.values)
    g = big.assemble_gradient(x)
    assert np.array_equal(g[: x0.size], ref.assemble_gradient(x0)) and not g[x0.size:].any()
    assert big.assemble_energy(x) == ref.assemble_energy(x0)


def test_project_to_psd_known_answers(oracle):
    """Hand-computed answers for the restatement of ipc::project_to_psd (the library's source is absent: documented behaviour =
    symmetric eigendecomposition, negative eigenvalues clamped to zero, a PSD input returned unchanged)."""
    # 2 x 2, eigenvalues 3 and -1 with eigenvectors (1, 1)/sqrt(2), (1, -1)/sqrt(2): projection = 3 v v^T
    out = oracle.project_to_psd(np.array([[1.0, 2.0], [2.0, 1.0]]))
    assert np.abs(out - np.array([[1.5, 1.5], [1.5, 1.5]])).max() <= 1e-15
    # diagonal: clamps entry by entry
    out = oracle.project_to_psd(np.diag([2.0, -1.0, 0.5]))
    assert np.abs(out - np.diag([2.0, 0.0, 0.5])).max() <= 1e-15
    # 3 x 3 with a known spectrum: Q diag(4, -2, 1) Q^T for the rotation Q about (1, 1, 1)/sqrt(3) by 120 degrees composed
    # with a Householder reflection - built here, answer = Q diag(4, 0, 1) Q^T
    w = np.array([1.0, 2.0, 2.0]) / 3.0
    Q = np.eye(3) - 2.0 * np.outer(w, w)
    A = Q @ np.diag([4.0, -2.0, 1.0]) @ Q.T
    out = oracle.project_to_psd(A)
    assert np.abs(out - Q @ np.diag([4.0, 0.0, 1.0]) @ Q.T).max() <= 1e-14
    # an indefinite 2 x 2 block embedded in a 6 x 6 matrix that is otherwise PSD and decoupled: only the block changes
    B = np.zeros((6, 6))
    B[0, 0], B[1, 1], B[4, 4], B[5, 5] = 3.0, 0.25, 7.0, 0.0
    B[2:4, 2:4] = [[1.0, 2.0], [2.0, 1.0]]
    out = oracle.project_to_psd(B)
    want = B.copy()
    want[2:4, 2:4] = 1.5
    assert np.abs(out - want).max() <= 1e-15
    # already PSD (including a zero eigenvalue): returned bit for bit
    P = np.array([[2.0, 1.0, 0.0], [1.0, 2.0, 0.0], [0.0, 0.0, 0.0]])
    assert np.array_equal(oracle.project_to_psd(P), P)
    rng = np.random.default_rng(11)
    G = rng.standard_normal((12, 12))
    S = G @ G.T
    assert np.array_equal(oracle.project_to_psd(S), S)
