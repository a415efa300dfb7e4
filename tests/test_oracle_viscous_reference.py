"""The oracle's ViscousDamping local energy / gradient / Hessian against the REFERENCE'S OWN function bodies.

`oracle/refmath/vd_glue.cpp` compiles `ViscousDamping::compute_energy`, `assemble_gradient`, `assemble_hessian`,
`compute_stress_aux` and `compute_stress_grad_aux` (assembler/ViscousDamping.cpp:5-62, 122-229, 297-342) verbatim from
/root/reference against the dense-matrix stand-in `mini_eigen.hpp` (Eigen is not installed) into oracle/_ref/libvdref.so.
`tools/make_golden.py viscous` ran them on 9 single-element cases (P1..P3, jittered tets, three time-step sizes) and committed
inputs and outputs as tests/golden/vd_local.npz, which is what travels to the GPU box. Tolerance: 1e-13 of the largest entry."""
import ctypes
import os

import numpy as np
import pytest

from polyfem_b200 import tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "vd_local.npz"))
TOL = 1e-13


def problem(oracle, k):
    p = int(GOLD[f"p_{k}"])
    t = tables.reference_tables(p)
    u, u_prev = GOLD[f"u_{k}"], GOLD[f"u_prev_{k}"]
    nl = u.shape[0]
    prob = oracle.OracleProblem("ViscousDamping", np.arange(nl, dtype=np.int32)[None, :], GOLD[f"vertices_{k}"][None], nl, t["points"], t["weights"],
                                t["grad"], lam=float(GOLD["psi"]), mu=float(GOLD["phi"]))
    prob.set_previous(u_prev.reshape(-1), float(GOLD[f"dt_{k}"]))
    return prob, u.reshape(-1), nl


def close(a, b):
    assert np.abs(a - b).max() <= TOL * np.abs(b).max()


@pytest.mark.parametrize("k", range(int(GOLD["n_cases"])))
def test_oracle_equals_reference_functions(oracle, k):
    prob, x, nl = problem(oracle, k)
    e_ref = float(GOLD[f"energy_{k}"])
    for e in (prob.local_energy(0, x), prob.assemble_energy(x)):
        assert abs(e - e_ref) <= TOL * abs(e_ref)
    close(prob.local_gradient(0, x), GOLD[f"gradient_{k}"])
    close(prob.local_hessian(0, x).reshape(3 * nl, 3 * nl), GOLD[f"hessian_{k}"])
    # through the global loops + SparseMatrixCache scatter: a one-element mesh assembles to the local matrix itself
    close(np.asarray(prob.assemble_hessian(x).to_scipy().todense()), GOLD[f"hessian_{k}"])
    close(prob.assemble_gradient(x), GOLD[f"gradient_{k}"])


def test_live_against_libvdref_when_present(oracle):
    """Elements of a jittered mesh, directly against oracle/_ref/libvdref.so (build container only), including the
    'previous displacement of another size' branch."""
    path = os.path.join(ROOT, "oracle", "_ref", "libvdref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libvdref.so not built (no reference tree)")
    from polyfem_b200 import mesh as M
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ref_vd_local.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, dp, dp]

    def P(a):
        return None if a is None else a.ctypes.data_as(dp)
    psi, phi, dt = 12.0, 35.0, 0.02
    for p, n in [(1, 2), (2, 2), (3, 1)]:
        mesh = M.kuhn_cube(n, p, jitter=0.2)
        t = tables.reference_tables(p)
        x0 = M.random_displacement(mesh, scale=0.1, seed=1)[: mesh.n_bases * 3]
        x = x0 + M.random_displacement(mesh, scale=0.03, seed=2)[: mesh.n_bases * 3]
        prob = oracle.problem_from_mesh(mesh, "ViscousDamping", psi=psi, phi=phi)
        prob.set_previous(x0, dt)
        nl, nq = mesh.conn.shape[1], t["weights"].size
        grads = np.ascontiguousarray(t["grad"])
        for e in range(min(mesh.n_elements, 12)):
            det, jit, _ = prob.assembly_values(e)
            jac_it, da = np.ascontiguousarray(jit.reshape(nq, 9)), np.ascontiguousarray(det * t["weights"])
            u = np.ascontiguousarray(x.reshape(-1, 3)[mesh.conn[e]].reshape(-1))
            up = np.ascontiguousarray(x0.reshape(-1, 3)[mesh.conn[e]].reshape(-1))
            en, g, H = np.zeros(1), np.zeros(nl * 3), np.zeros((nl * 3, nl * 3))
            assert lib.ref_vd_local(nl, nq, P(u), P(up), P(grads), P(jac_it), P(da), dt, psi, phi, P(en), P(g), P(H)) == 0
            assert abs(prob.local_energy(e, x) - en[0]) <= TOL * abs(en[0])
            close(prob.local_gradient(e, x), g)
            close(prob.local_hessian(e, x).reshape(3 * nl, 3 * nl), H)
            assert lib.ref_vd_local(nl, nq, P(u), None, P(grads), P(jac_it), P(da), dt, psi, phi, P(en), P(g), P(H)) == 0
            assert en[0] == 0.0 and not g.any() and not H.any()
