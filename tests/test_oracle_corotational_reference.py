"""The oracle's FixedCorotational local energy / gradient / Hessian against the REFERENCE'S OWN function bodies AND its own SVD.

`oracle/refmath/fc_glue.cpp` compiles `FixedCorotational::compute_energy_aux`, `compute_energy_aux_gradient_fast`,
`compute_energy_hessian_aux_fast` and the six `*_from_singular_values` / `*_from_def_grad` functions
(assembler/FixedCorotational.cpp:293-436, 592-827) together with `fastSVD3d` and its helpers (utils/svd.hpp:134-317: analytic
eigenvalues of A^T A, eigenvectors by cofactors) verbatim from /root/reference against the dense-matrix stand-in `mini_eigen.hpp`
into oracle/_ref/libfcref.so. `tools/make_golden.py corotational` ran them on 12 single-element cases (P1..P3; tiny, moderate and
large displacements, an inverted element, a rigid rotation) and committed inputs and outputs as tests/golden/fc_local.npz.
Tolerance: 1e-12 of the largest entry - the oracle diagonalises F^T F by Jacobi rotations, the reference by the trigonometric
formula, and the two agree to about 3e-15 on these cases."""
import ctypes
import os

import numpy as np
import pytest

from polyfem_b200 import tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "fc_local.npz"))
TOL = 1e-12


def problem(oracle, k):
    p = int(GOLD[f"p_{k}"])
    t = tables.reference_tables(p)
    u = GOLD[f"u_{k}"]
    nl = u.shape[0]
    prob = oracle.OracleProblem("FixedCorotational", np.arange(nl, dtype=np.int32)[None, :], GOLD[f"vertices_{k}"][None], nl, t["points"], t["weights"],
                                t["grad"], lam=float(GOLD["lambda"]), mu=float(GOLD["mu"]))
    return prob, u.reshape(-1), nl


@pytest.mark.parametrize("k", range(int(GOLD["n_cases"])))
def test_oracle_equals_reference_functions(oracle, k):
    prob, x, nl = problem(oracle, k)
    H_ref, g_ref, e_ref = GOLD[f"hessian_{k}"], GOLD[f"gradient_{k}"], float(GOLD[f"energy_{k}"])
    hs = np.abs(H_ref).max()
    # energy and stress are measured against the scale of the tangent (they vanish for the rotation case)
    h = float(np.linalg.norm(GOLD[f"vertices_{k}"][1] - GOLD[f"vertices_{k}"][0]))
    assert abs(prob.local_energy(0, x) - e_ref) <= TOL * max(abs(e_ref), hs * h * h * 1e-3)
    assert np.abs(prob.local_gradient(0, x) - g_ref).max() <= TOL * max(np.abs(g_ref).max(), hs * h * 1e-3)
    assert np.abs(prob.local_hessian(0, x).reshape(3 * nl, 3 * nl) - H_ref).max() <= TOL * hs
    assert np.abs(np.asarray(prob.assemble_hessian(x).to_scipy().todense()) - H_ref).max() <= TOL * hs


def test_golden_covers_inversion_and_rotation():
    dets, energies = [], []
    for k in range(int(GOLD["n_cases"])):
        energies.append(float(GOLD[f"energy_{k}"]))
    assert min(abs(e) for e in energies) < 1e-12 * max(energies)  # the rigid rotation
    assert (GOLD["svd_S"][:, 2] < 0).any()  # a reflected matrix: the last singular value carries the sign


def test_signed_svd_matches_the_reference_svd(oracle):
    """the reference's own SVD outputs (utils/svd.hpp) of six matrices, among them near-identity ones where the trigonometric
    eigenvalue formula is at its least accurate: singular values, the rotation U V^T and the reconstruction"""
    for A, U, S, V in zip(GOLD["svd_A"], GOLD["svd_U"], GOLD["svd_S"], GOLD["svd_V"]):
        assert np.abs(U @ np.diag(S) @ V.T - A).max() <= 1e-9  # what the reference itself achieves
        s_np = np.linalg.svd(A, compute_uv=False)
        assert np.abs(np.abs(S) - s_np).max() <= 1e-9
        assert abs(np.linalg.det(U) - 1) < 1e-9 and abs(np.linalg.det(V) - 1) < 1e-9 and np.sign(S[2]) == np.sign(np.linalg.det(A))


def test_live_against_libfcref_when_present(oracle):
    path = os.path.join(ROOT, "oracle", "_ref", "libfcref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libfcref.so not built (no reference tree)")
    from polyfem_b200 import mesh as M
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ref_fc_local.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp]

    def P(a):
        return a.ctypes.data_as(dp)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    for p, n, scale in [(1, 2, 0.3), (2, 2, 0.1), (3, 1, 0.03)]:
        mesh = M.kuhn_cube(n, p, jitter=0.2)
        t = tables.reference_tables(p)
        x = M.random_displacement(mesh, scale=scale, seed=5)[: mesh.n_bases * 3]
        prob = oracle.problem_from_mesh(mesh, "FixedCorotational")
        nl, nq = mesh.conn.shape[1], t["weights"].size
        grads = np.ascontiguousarray(t["grad"])
        for e in range(min(mesh.n_elements, 12)):
            det, jit, _ = prob.assembly_values(e)
            jac_it, da = np.ascontiguousarray(jit.reshape(nq, 9)), np.ascontiguousarray(det * t["weights"])
            u = np.ascontiguousarray(x.reshape(-1, 3)[mesh.conn[e]].reshape(-1))
            en, g, H = np.zeros(1), np.zeros(nl * 3), np.zeros((nl * 3, nl * 3))
            assert lib.ref_fc_local(nl, nq, P(u), P(grads), P(jac_it), P(da), lam, mu, P(en), P(g), P(H)) == 0
            assert abs(prob.local_energy(e, x) - en[0]) <= TOL * abs(en[0])
            assert np.abs(prob.local_gradient(e, x) - g).max() <= TOL * np.abs(g).max()
            assert np.abs(prob.local_hessian(e, x).reshape(3 * nl, 3 * nl) - H).max() <= TOL * np.abs(H).max()
