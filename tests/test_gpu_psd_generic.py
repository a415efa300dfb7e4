"""project_to_psd (Assembler.cpp:693-694: ipc::project_to_psd on every local Hessian of an NLAssembler) beyond the NeoHookean
P1 / P2 kernels of round 1: the generic kernel's projection for NeoHookean P3 / P4 and SaintVenant P1 .. P4 (P4: N = 105 is odd, the matrix is padded
with a zero row and column for the round-robin Jacobi ordering, one warp per CTA), and LinearElasticity,
where the flag has no effect (its element stiffness is PSD). The oracle's restatement (own Jacobi eigen-solver, pinned by
hand-computed answers in tests/test_oracle_properties.py since ipc-toolkit's source is absent) is the checker; both are
non-expansive maps of the same local matrices computed by different eigen-solvers, so the bar is 1e-10 of the row scale."""
import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close, gpu_handle, make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("material,p,n,scale", [("NeoHookean", 3, 2, 0.02), ("SaintVenant", 1, 4, 0.2), ("SaintVenant", 2, 3, 0.2), ("SaintVenant", 3, 2, 0.2),
                                                  ("NeoHookean", 4, 1, 0.02), ("SaintVenant", 4, 1, 0.2)])
def test_generic_projection_equals_oracle(oracle, material, p, n, scale):
    mesh, x, t = make_case(n, p, jitter=0.1, scale=scale)
    x = x[: mesh.n_bases * 3]
    if material == "SaintVenant" and p == 4:
        x = x - 0.7 * mesh.node_xyz.reshape(-1)  # one cell: random noise alone leaves it convex; compress it (S < 0)
    ref = oracle.problem_from_mesh(mesh, material, n_threads=2)
    h = gpu_handle(mesh, material, t)
    H0 = ref.assemble_hessian(x)
    H1 = ref.assemble_hessian(x, project_to_psd=True)
    assert np.abs(H0.values - H1.values).max() > 1e-4 * np.abs(H0.values).max(), "projection inactive: test is vacuous"
    h.profile_enable(True)
    v = h.hessian(x, project_to_psd=True)
    assert any("psd" in k for (k, ms) in h.profile_read())
    assert_values_close(H1.outer, H1.inner, v, H1.values, tol=1e-10, what="projected hessian")
    e, g, v2 = h.grad_hess(x, project_to_psd=True)
    assert abs(e - ref.assemble_energy(x)) <= REL_TOL * abs(e)
    assert_vector_close(g, ref.assemble_gradient(x))
    assert_values_close(H1.outer, H1.inner, v2, H1.values, tol=1e-10, what="projected hessian (fused)")
    # elements whose local matrix is already PSD are scattered unchanged: at x = 0 the projected call equals the plain one
    z = np.zeros_like(x)
    assert_values_close(H0.outer, H0.inner, h.hessian(z, project_to_psd=True), h.hessian(z), tol=1e-13, what="PSD state")


def test_linear_elasticity_ignores_the_flag_and_p5_fails_loudly(oracle):
    from polyfem_b200 import capi
    mesh, x, t = make_case(3, 2, jitter=0.1)
    x = x[: mesh.n_bases * 3]
    h = gpu_handle(mesh, "LinearElasticity", t)
    ref = oracle.problem_from_mesh(mesh, "LinearElasticity")
    H1 = ref.assemble_hessian(x, project_to_psd=True)
    assert_values_close(H1.outer, H1.inner, h.hessian(x, project_to_psd=True), H1.values, tol=1e-12, what="LinearElasticity, projected")
    # a 44-node element (N = 132 > 128; P5 would be 168): the two local matrices do not fit the shared memory of an SM - refused,
    # never silently unprojected. (Synthetic element: one cell, made-up gradients; the refusal comes before any launch.)
    rng = np.random.default_rng(0)
    nl, nq = 44, 4
    conn = np.arange(nl, dtype=np.int32)[None, :]
    verts = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]], dtype=np.float64)
    lam, mu = 57692.3, 38461.5
    h5 = capi.Handle("NeoHookean", conn, nl, np.full(nq, 1.0 / 24), rng.standard_normal((nq, nl, 3)), vertices=verts, lam=lam, mu=mu)
    with pytest.raises(capi.PfaError) as ei:
        h5.hessian(np.zeros(h5.ndof), project_to_psd=True)
    assert ei.value.code == capi.PFA_ERR_UNSUPPORTED
