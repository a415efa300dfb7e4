"""The oracle's NeoHookean local gradient / Hessian against the REFERENCE'S OWN function bodies.

`oracle/refmath/` compiles `compute_energy_aux`, `compute_energy_aux_gradient_fast`, `compute_energy_hessian_aux_fast`
(with `hat`, `cross`; assembler/NeoHookeanElasticity.cpp:338-658) verbatim from /root/reference against a small dense-matrix
stand-in (Eigen is not installed) into oracle/_ref/libnhref.so. `tools/make_golden.py` ran them on 12
single-element cases (P1..P4, jittered tets, one inverted element) and committed inputs and outputs as
tests/golden/nh_local.npz, which is what travels to the GPU box. Tolerance: 1e-13 of the largest entry
(the two sides order their sums differently)."""
import ctypes
import os

import numpy as np
import pytest

from polyfem_b200 import tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "nh_local.npz"))
TOL = 1e-13


def cases():
    return range(int(GOLD["n_cases"]))


def one_element_problem(oracle, k):
    p = int(GOLD[f"p_{k}"])
    t = tables.reference_tables(p)
    verts = GOLD[f"vertices_{k}"]
    u = GOLD[f"u_{k}"]
    nl = u.shape[0]
    conn = np.arange(nl, dtype=np.int32)[None, :]
    prob = oracle.OracleProblem("NeoHookean", conn, verts[None], nl, t["points"], t["weights"], t["grad"],
                                lam=float(GOLD["lambda"]), mu=float(GOLD["mu"]))
    return prob, u.reshape(-1), nl


def close(a, b):
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    ok = ~np.isnan(b)
    if ok.any():
        assert np.abs(a[ok] - b[ok]).max() <= TOL * np.abs(b[ok]).max()


@pytest.mark.parametrize("k", cases())
def test_oracle_local_math_equals_reference_functions(oracle, k):
    prob, x, nl = one_element_problem(oracle, k)
    e_ref = float(GOLD[f"energy_{k}"])
    for e in (prob.local_energy(0, x), prob.assemble_energy(x)):
        assert np.isnan(e) == np.isnan(e_ref) and (np.isnan(e_ref) or abs(e - e_ref) <= TOL * abs(e_ref))
    close(prob.local_gradient(0, x), GOLD[f"gradient_{k}"])
    close(prob.local_hessian(0, x).reshape(3 * nl, 3 * nl), GOLD[f"hessian_{k}"])
    # and through the global loops + SparseMatrixCache scatter: a one-element mesh assembles to H_e itself
    H = prob.assemble_hessian(x)
    close(np.asarray(H.to_scipy().todense()), GOLD[f"hessian_{k}"])
    close(prob.assemble_gradient(x), GOLD[f"gradient_{k}"])


def test_golden_has_the_inverted_case():
    assert any(np.isnan(GOLD[f"gradient_{k}"]).any() for k in cases())


def test_live_against_libnhref_when_present(oracle):
    """More random elements, directly against oracle/_ref/libnhref.so (build container only)."""
    path = os.path.join(ROOT, "oracle", "_ref", "libnhref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libnhref.so not built (no reference tree)")
    from polyfem_b200 import mesh as M
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    for f in (lib.ref_nh_energy, lib.ref_nh_gradient, lib.ref_nh_hessian):
        f.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp]

    def P(a):
        return a.ctypes.data_as(dp)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    for p, n in [(1, 2), (2, 2), (3, 1)]:
        mesh = M.kuhn_cube(n, p, jitter=0.2)
        x = M.random_displacement(mesh, scale=0.05 if p < 3 else 0.01)
        t = tables.reference_tables(p)
        ref = oracle.problem_from_mesh(mesh, "NeoHookean")
        nl, nq = mesh.conn.shape[1], t["weights"].size
        for e in range(min(mesh.n_elements, 10)):
            edges = mesh.vertices[e][1:] - mesh.vertices[e][0]
            jac_it = np.ascontiguousarray(np.repeat(np.linalg.inv(edges).T[None], nq, 0).reshape(nq, 9))
            da = np.ascontiguousarray(np.linalg.det(edges) * t["weights"])
            u = np.ascontiguousarray(x.reshape(-1, 3)[mesh.conn[e]].reshape(-1))
            g, H = np.zeros(nl * 3), np.zeros((nl * 3, nl * 3))
            assert lib.ref_nh_gradient(nl, nq, P(u), P(np.ascontiguousarray(t["grad"])), P(jac_it), P(da), lam, mu, P(g)) == 0
            assert lib.ref_nh_hessian(nl, nq, P(u), P(np.ascontiguousarray(t["grad"])), P(jac_it), P(da), lam, mu, P(H)) == 0
            close(ref.local_gradient(e, x), g)
            close(ref.local_hessian(e, x).reshape(3 * nl, 3 * nl), H)
            en = np.zeros(1)
            assert lib.ref_nh_energy(nl, nq, P(u), P(np.ascontiguousarray(t["grad"])), P(jac_it), P(da), lam, mu, P(en)) == 0
            assert abs(ref.local_energy(e, x) - en[0]) <= TOL * abs(en[0])


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_linear_local_blocks_equal_reference_functions(oracle, p):
    """LinearElasticity::assemble / Laplacian::assemble / Mass::assemble(LinearAssemblerData) of the reference
    (LinearElasticity.cpp:29-63, Laplacian.cpp:13-26, Mass.cpp:5-23, compiled verbatim like the NeoHookean functions)
    against the oracle's local blocks, and against the assembled matrix of the one-element mesh."""
    verts = GOLD[f"lin_vertices_p{p}"]
    t = tables.reference_tables(p)
    tm = tables.reference_tables(p, tables.quadrature_order(p, is_mass=True))
    nl = t["grad"].shape[1]
    conn = np.arange(nl, dtype=np.int32)[None, :]
    lam, mu, rho = float(GOLD["lambda"]), float(GOLD["mu"]), float(GOLD["rho"])
    le = oracle.OracleProblem("LinearElasticity", conn, verts[None], nl, t["points"], t["weights"], t["grad"], lam=lam, mu=mu)
    lap = oracle.OracleProblem("Laplacian", conn, verts[None], nl, t["points"], t["weights"], t["grad"])
    mass = oracle.OracleProblem("Mass", conn, verts[None], nl, tm["points"], tm["weights"], tm["grad"], ref_vals=tm["val"], density=rho)
    g_le, g_lap, g_mass = GOLD[f"le_blocks_p{p}"], GOLD[f"lap_blocks_p{p}"], GOLD[f"mass_blocks_p{p}"]
    s_le, s_lap, s_mass = np.abs(g_le).max(), np.abs(g_lap).max(), np.abs(g_mass).max()
    for i in range(nl):
        for j in range(nl):
            assert np.abs(le.local_stiffness(0, i, j) - g_le[i, j]).max() <= TOL * s_le
            assert abs(lap.local_stiffness(0, i, j)[0] - g_lap[i, j]) <= TOL * s_lap
            assert np.abs(mass.local_stiffness(0, i, j) - g_mass[i, j]).max() <= TOL * s_mass
    # LinearAssembler::assemble (Assembler.cpp:228-250): entry (g_i*size + m, g_j*size + n) = block(n*size + m), j <= i
    # computed and mirrored; on a one-element mesh the assembled matrix is exactly that table
    for prob, blocks, size in ((le, g_le, 3), (mass, g_mass, 3)):
        K = np.asarray(prob.assemble().to_scipy().todense())
        ref = np.zeros_like(K)
        for i in range(nl):
            for j in range(i + 1):
                blk = blocks[i, j].reshape(size, size)  # [n][m]
                ref[i * size:(i + 1) * size, j * size:(j + 1) * size] = blk.T
                ref[j * size:(j + 1) * size, i * size:(i + 1) * size] = blk
        assert np.abs(K - ref).max() <= TOL * np.abs(ref).max()
    K = np.asarray(lap.assemble().to_scipy().todense())
    ref = np.tril(g_lap) + np.tril(g_lap, -1).T
    assert np.abs(K - ref).max() <= TOL * np.abs(ref).max()


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_geometry_equals_reference_finalize3d(oracle, p):
    """det, J^-T and grad_t_m of the oracle against the reference's own ElementAssemblyValues::finalize3d
    (ElementAssemblyValues.cpp:65-104, compiled verbatim; its 3x3 inverse is the stand-in's cofactor formula)."""
    verts = GOLD[f"lin_vertices_p{p}"]
    t = tables.reference_tables(p)
    nl = t["grad"].shape[1]
    conn = np.arange(nl, dtype=np.int32)[None, :]
    prob = oracle.OracleProblem("Laplacian", conn, verts[None], nl, t["points"], t["weights"], t["grad"])
    det, jit, gt = prob.assembly_values(0)
    for got, ref in ((det, GOLD[f"geo_det_p{p}"]), (jit, GOLD[f"geo_jac_it_p{p}"]), (gt, GOLD[f"geo_grad_t_m_p{p}"])):
        assert np.abs(got - ref).max() <= TOL * np.abs(ref).max()
    # and the numpy form the parity tests feed to the library as "general geometry" input
    edges = verts[1:] - verts[0]
    assert np.abs(np.linalg.inv(edges).T - GOLD[f"geo_jac_it_p{p}"][0]).max() <= 1e-12 * np.abs(GOLD[f"geo_jac_it_p{p}"][0]).max()
    assert abs(np.linalg.det(edges) - GOLD[f"geo_det_p{p}"][0]) <= 1e-13 * abs(GOLD[f"geo_det_p{p}"][0])


@pytest.mark.parametrize("k", cases())
def test_linear_elasticity_energy_in_a_nonlinear_solve_equals_reference_function(oracle, k):
    """LinearElasticity::compute_energy -> compute_energy_aux<double> (LinearElasticity.cpp:65-68, 103-132, with get_local_disp /
    compute_disp_grad_at_quad of utils/ElasticityUtils.hpp:70-136, all compiled verbatim) against the oracle's energy of the
    linear material on the same one-element meshes; and E = 1/2 u^T K u with the pinned stiffness, the closed form the
    CUDA path evaluates. The gradient of this path is autodiff of the same function in the reference (exact derivative):
    g = K u is checked against it by finite differences in tests/test_oracle_properties.py."""
    LE = np.load(os.path.join(ROOT, "tests", "golden", "le_energy.npz"))
    p = int(GOLD[f"p_{k}"])
    t = tables.reference_tables(p)
    verts, u = GOLD[f"vertices_{k}"], GOLD[f"u_{k}"]
    nl = u.shape[0]
    conn = np.arange(nl, dtype=np.int32)[None, :]
    prob = oracle.OracleProblem("LinearElasticity", conn, verts[None], nl, t["points"], t["weights"], t["grad"],
                                lam=float(GOLD["lambda"]), mu=float(GOLD["mu"]))
    e_ref = float(LE[f"le_energy_{k}"])
    x = u.reshape(-1)
    assert np.isfinite(e_ref) and e_ref > 0
    assert abs(prob.assemble_energy(x) - e_ref) <= TOL * e_ref
    K = prob.assemble().to_scipy()
    assert abs(0.5 * float(x @ (K @ x)) - e_ref) <= 1e-12 * e_ref
