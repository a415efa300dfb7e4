"""The oracle's MooneyRivlin local energy / gradient / Hessian against the REFERENCE'S OWN code path, autodiff included.

`oracle/refmath/mr_glue.cpp` compiles `MooneyRivlinElasticity::elastic_energy<T>` (MooneyRivlinElasticity.hpp:25-47), GenericElastic's
`compute_energy_aux`, `compute_gradient_from_stress`, `compute_hessian_from_stress` (GenericElastic.hpp:92-212, 268-351; the STRESS
autodiff mode is the reference's default), `compute_B_block`, `first_invariant` / `second_invariant`, `determinant` - all extracted
verbatim at build time - over the reference's OWN forward-mode scalars `utils/autodiff.h`, included unmodified, against the
dense-matrix stand-in `mini_eigen.hpp` (Eigen is not installed) into oracle/_ref/libmrref.so. `tools/make_golden.py mooney` ran it
on 12 single-element cases (P1..P4) and committed inputs and outputs as tests/golden/mr_local.npz. Tolerance 1e-13 of the largest
entry: the reference differentiates with respect to F (9 variables) and contracts, the oracle with respect to the local dofs."""
import ctypes
import os

import numpy as np
import pytest

from polyfem_b200 import tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "mr_local.npz"))
TOL = 1e-13


def problem(oracle, k):
    p = int(GOLD[f"p_{k}"])
    t = tables.reference_tables(p)
    u = GOLD[f"u_{k}"]
    nl = u.shape[0]
    prob = oracle.OracleProblem("MooneyRivlin", np.arange(nl, dtype=np.int32)[None, :], GOLD[f"vertices_{k}"][None], nl, t["points"], t["weights"],
                                t["grad"], lam=float(GOLD["c1"]), mu=float(GOLD["c2"]), param3=float(GOLD["k"]))
    return prob, u.reshape(-1), nl


def close(a, b):
    assert np.abs(a - b).max() <= TOL * np.abs(b).max()


@pytest.mark.parametrize("k", range(int(GOLD["n_cases"])))
def test_oracle_equals_reference_code_path(oracle, k):
    prob, x, nl = problem(oracle, k)
    e_ref = float(GOLD[f"energy_{k}"])
    # psi = c1 (I1~ - 3) + c2 (I2~ - 3) + ...: for tiny strains the energy is a difference of numbers near 3 (c1 + c2) per unit volume,
    # and both sides carry that rounding
    verts = GOLD[f"vertices_{k}"]
    vol = abs(np.linalg.det(verts[1:] - verts[0])) / 6.0
    assert abs(prob.local_energy(0, x) - e_ref) <= TOL * max(abs(e_ref), 3.0 * (float(GOLD["c1"]) + float(GOLD["c2"])) * vol)
    close(prob.local_gradient(0, x), GOLD[f"gradient_{k}"])
    close(prob.local_hessian(0, x).reshape(3 * nl, 3 * nl), GOLD[f"hessian_{k}"])
    close(np.asarray(prob.assemble_hessian(x).to_scipy().todense()), GOLD[f"hessian_{k}"])
    close(prob.assemble_gradient(x), GOLD[f"gradient_{k}"])


def test_live_against_libmrref_when_present(oracle):
    path = os.path.join(ROOT, "oracle", "_ref", "libmrref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libmrref.so not built (no reference tree)")
    from polyfem_b200 import mesh as M
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ref_mr_local.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, dp, dp]

    def P(a):
        return a.ctypes.data_as(dp)
    c1, c2, kk = 9000.0, 4000.0, 60000.0
    for p, n, scale in [(1, 2, 0.2), (2, 2, 0.08), (3, 1, 0.03)]:
        mesh = M.kuhn_cube(n, p, jitter=0.2)
        t = tables.reference_tables(p)
        x = M.random_displacement(mesh, scale=scale, seed=5)[: mesh.n_bases * 3]
        prob = oracle.problem_from_mesh(mesh, "MooneyRivlin", c1=c1, c2=c2, k=kk)
        nl, nq = mesh.conn.shape[1], t["weights"].size
        grads = np.ascontiguousarray(t["grad"])
        for e in range(min(mesh.n_elements, 12)):
            det, jit, _ = prob.assembly_values(e)
            jac_it, da = np.ascontiguousarray(jit.reshape(nq, 9)), np.ascontiguousarray(det * t["weights"])
            u = np.ascontiguousarray(x.reshape(-1, 3)[mesh.conn[e]].reshape(-1))
            en, g, H = np.zeros(1), np.zeros(nl * 3), np.zeros((nl * 3, nl * 3))
            assert lib.ref_mr_local(nl, nq, P(u), P(grads), P(jac_it), P(da), c1, c2, kk, P(en), P(g), P(H)) == 0
            assert abs(prob.local_energy(e, x) - en[0]) <= TOL * abs(en[0])
            close(prob.local_gradient(e, x), g)
            close(prob.local_hessian(e, x).reshape(3 * nl, 3 * nl), H)
