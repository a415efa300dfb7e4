"""The oracle's SaintVenant local energy / gradient / Hessian against the REFERENCE'S OWN code path, autodiff included.

`oracle/refmath/sv_glue.cpp` compiles `SaintVenantElasticity::compute_energy_aux<T>`, `stress<T, N>`, `strain_from_disp_grad`,
`assemble_gradient`, `assemble_hessian` (assembler/SaintVenantElasticity.cpp:9-20, 61-70, 88-129, 206-266), the dispatch functions
`gradient_from_energy` / `hessian_from_energy` (utils/ElasticityUtils.cpp:81-270) and `ElasticityTensor::set_from_lambda_mu` /
`operator()` (assembler/MatParams.cpp:91-123, 211-253) - all extracted verbatim at build time - over the reference's OWN forward-mode
scalars `utils/autodiff.h`, included unmodified, into oracle/_ref/libsvref.so. `tools/make_golden.py saint_venant` ran it on 12
single-element cases (P1..P4; tiny and moderate displacements and a state compressed to 40 %, whose tangent is indefinite) and
committed inputs and outputs as tests/golden/sv_local.npz. Tolerance 1e-13 of the largest entry."""
import ctypes
import os

import numpy as np
import pytest

from polyfem_b200 import tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "sv_local.npz"))
TOL = 1e-13


def problem(oracle, k):
    p = int(GOLD[f"p_{k}"])
    t = tables.reference_tables(p)
    u = GOLD[f"u_{k}"]
    nl = u.shape[0]
    prob = oracle.OracleProblem("SaintVenant", np.arange(nl, dtype=np.int32)[None, :], GOLD[f"vertices_{k}"][None], nl, t["points"], t["weights"],
                                t["grad"], lam=float(GOLD["lambda"]), mu=float(GOLD["mu"]))
    return prob, u.reshape(-1), nl


def close(a, b):
    assert np.abs(a - b).max() <= TOL * np.abs(b).max()


@pytest.mark.parametrize("k", range(int(GOLD["n_cases"])))
def test_oracle_equals_reference_code_path(oracle, k):
    prob, x, nl = problem(oracle, k)
    e_ref = float(GOLD[f"energy_{k}"])
    assert abs(prob.local_energy(0, x) - e_ref) <= TOL * abs(e_ref)
    close(prob.local_gradient(0, x), GOLD[f"gradient_{k}"])
    close(prob.local_hessian(0, x).reshape(3 * nl, 3 * nl), GOLD[f"hessian_{k}"])
    close(np.asarray(prob.assemble_hessian(x).to_scipy().todense()), GOLD[f"hessian_{k}"])
    close(prob.assemble_gradient(x), GOLD[f"gradient_{k}"])


def test_golden_has_an_indefinite_tangent():
    assert any(np.linalg.eigvalsh(0.5 * (GOLD[f"hessian_{k}"] + GOLD[f"hessian_{k}"].T)).min() < -1e-6 * np.abs(GOLD[f"hessian_{k}"]).max()
               for k in range(int(GOLD["n_cases"])))


def test_live_against_libsvref_when_present(oracle):
    path = os.path.join(ROOT, "oracle", "_ref", "libsvref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libsvref.so not built (no reference tree)")
    from polyfem_b200 import mesh as M
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ref_sv_local.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp]

    def P(a):
        return a.ctypes.data_as(dp)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    for p, n, scale in [(1, 2, 0.3), (2, 2, 0.2), (3, 1, 0.1)]:
        mesh = M.kuhn_cube(n, p, jitter=0.2)
        t = tables.reference_tables(p)
        x = M.random_displacement(mesh, scale=scale, seed=5)[: mesh.n_bases * 3]
        prob = oracle.problem_from_mesh(mesh, "SaintVenant")
        nl, nq = mesh.conn.shape[1], t["weights"].size
        grads = np.ascontiguousarray(t["grad"])
        for e in range(min(mesh.n_elements, 12)):
            det, jit, _ = prob.assembly_values(e)
            jac_it, da = np.ascontiguousarray(jit.reshape(nq, 9)), np.ascontiguousarray(det * t["weights"])
            u = np.ascontiguousarray(x.reshape(-1, 3)[mesh.conn[e]].reshape(-1))
            en, g, H = np.zeros(1), np.zeros(nl * 3), np.zeros((nl * 3, nl * 3))
            assert lib.ref_sv_local(nl, nq, P(u), P(grads), P(jac_it), P(da), lam, mu, P(en), P(g), P(H)) == 0
            assert abs(prob.local_energy(e, x) - en[0]) <= TOL * abs(en[0])
            close(prob.local_gradient(e, x), g)
            close(prob.local_hessian(e, x).reshape(3 * nl, 3 * nl), H)


LE = np.load(os.path.join(ROOT, "tests", "golden", "le_nl_local.npz"))


@pytest.mark.parametrize("k", range(int(LE["n_cases"])))
def test_linear_elasticity_nl_path_equals_reference_autodiff(oracle, k):
    """LinearElasticity as an NLAssembler: the reference differentiates compute_energy_aux<T> with its own autodiff scalars
    (LinearElasticity.cpp:65-132); tests/golden/le_nl_local.npz holds its outputs (ref_le_nl_local of oracle/_ref/libsvref.so,
    `tools/make_golden.py le_nl`). The oracle's gradient / Hessian of that path were property-pinned until now."""
    p = int(LE[f"p_{k}"])
    t = tables.reference_tables(p)
    u = LE[f"u_{k}"]
    nl = u.shape[0]
    prob = oracle.OracleProblem("LinearElasticity", np.arange(nl, dtype=np.int32)[None, :], LE[f"vertices_{k}"][None], nl, t["points"], t["weights"],
                                t["grad"], lam=float(LE["lambda"]), mu=float(LE["mu"]))
    x = u.reshape(-1)
    e_ref = float(LE[f"energy_{k}"])
    assert abs(prob.local_energy(0, x) - e_ref) <= TOL * abs(e_ref)
    close(prob.local_gradient(0, x), LE[f"gradient_{k}"])
    close(prob.local_hessian(0, x).reshape(3 * nl, 3 * nl), LE[f"hessian_{k}"])
    close(prob.assemble_gradient(x), LE[f"gradient_{k}"])
    close(np.asarray(prob.assemble_hessian(x).to_scipy().todense()), LE[f"hessian_{k}"])
