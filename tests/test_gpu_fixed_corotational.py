"""FixedCorotational (assembler/FixedCorotational.cpp) on the GPU against the oracle's restatement (pinned
by tests/test_oracle_fixed_corotational.py): the per-point signed SVD, stress and 9 x 9 stiffness in the kernel, contracted by the
warp. P1 .. P3, an inverted state, curved P2 elements, project_to_psd, and the host-side assembler class. The bar is 1e-11 of the
row scale for the Hessian: both sides take singular vectors from the eigenvectors of F^T F, whose conditioning enters the tangent
through 1 / (sigma_k + sigma_l)."""
import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close, gpu_handle, make_case
from polyfem_b200 import mesh as M, tables
from test_gpu_saint_venant_and_curved import geometry_arrays
from test_oracle_saint_venant_and_curved import curved_geometry, iso_problem

pytestmark = pytest.mark.gpu


def check(h, ref, x, tol=1e-11):
    H = ref.assemble_hessian(x)
    outer, inner = h.pattern()
    assert outer.tobytes() == H.outer.tobytes() and inner.tobytes() == H.inner.tobytes()
    e, g, v = h.grad_hess(x)
    e_ref = ref.assemble_energy(x)
    assert np.isfinite(e_ref) and abs(e - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(g, ref.assemble_gradient(x), tol=tol)
    assert_values_close(H.outer, H.inner, v, H.values, tol=tol)
    epe = h.energy_per_element(x)
    assert np.abs(epe - ref.assemble_energy_per_element(x)).max() <= REL_TOL * np.abs(epe).max()
    assert_vector_close(h.gradient(x), ref.assemble_gradient(x), tol=tol)


@pytest.mark.parametrize("p,n,scale", [(1, 4, 0.2), (2, 3, 0.1), (3, 2, 0.05), (1, 3, 0.9)])
def test_fixed_corotational_equals_oracle(oracle, p, n, scale):
    mesh, x, t = make_case(n, p, jitter=0.2, scale=scale, seed=3)
    x = x[: mesh.n_bases * 3]
    check(gpu_handle(mesh, "FixedCorotational", t), oracle.problem_from_mesh(mesh, "FixedCorotational", n_threads=2), x)


def test_rest_state_and_rejections(oracle):
    from polyfem_b200 import capi
    mesh, _, t = make_case(3, 2, jitter=0.1)
    h = gpu_handle(mesh, "FixedCorotational", t)
    z = np.zeros(mesh.n_bases * 3)
    K = oracle.problem_from_mesh(mesh, "LinearElasticity").assemble()
    e, g, v = h.grad_hess(z)
    assert abs(e) <= 1e-20 * np.abs(K.values).max() and np.abs(g).max() <= 1e-12 * np.abs(K.values).max()
    assert_values_close(K.outer, K.inner, v, K.values, tol=1e-11, what="tangent at rest vs LinearElasticity stiffness")
    with pytest.raises(capi.PfaError) as ei:
        h.linear_stiffness()
    assert ei.value.code == capi.PFA_ERR_UNSUPPORTED


def test_curved_p2_elements_and_projection(oracle):
    from polyfem_b200 import capi
    mesh = M.kuhn_cube(3, 2, jitter=0.1)
    t = tables.reference_tables(2)
    x = M.random_displacement(mesh, scale=0.2)[: mesh.n_bases * 3]
    ref = iso_problem(oracle, mesh, "FixedCorotational", curved_geometry(mesh), n_threads=2)
    jit, da = geometry_arrays(ref, mesh, t)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    h = capi.Handle("FixedCorotational", mesh.conn, mesh.n_bases, t["weights"], t["grad"], jac_it=jit, da=da, lam=lam, mu=mu)
    check(h, ref, x)
    H0 = ref.assemble_hessian(x)
    v0 = H0.values.copy()
    H1 = ref.assemble_hessian(x, project_to_psd=True)
    assert np.abs(v0 - H1.values).max() > 1e-4 * np.abs(v0).max(), "projection inactive: test is vacuous"
    assert_values_close(H1.outer, H1.inner, h.hessian(x, project_to_psd=True), H1.values, tol=1e-10, what="projected hessian")


def test_assembler_class(oracle):
    from polyfem_b200 import assembler as A
    mesh, x, t = make_case(3, 2, jitter=0.1, scale=0.1)
    x = x[: mesh.n_bases * 3]
    a = A.make_assembler("FixedCorotational")
    a.set_materials([], {"E": 1e5, "nu": 0.3})
    bases = A.FESpace.from_mesh(mesh)
    cache = A.AssemblyValsCache(mesh.p)
    ref = oracle.problem_from_mesh(mesh, "FixedCorotational")
    d = x.reshape(-1, 1)
    e = a.assemble_energy(True, bases, bases, cache, 0.0, 1.0, d, d)
    assert abs(e - ref.assemble_energy(x)) <= REL_TOL * abs(e)
    g = a.assemble_gradient(True, mesh.n_bases, bases, bases, cache, 0.0, 1.0, d, d)
    assert_vector_close(np.asarray(g).reshape(-1), ref.assemble_gradient(x), tol=1e-11)
