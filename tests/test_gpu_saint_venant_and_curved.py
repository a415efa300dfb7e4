"""SURVEY.md §8f rank 4 on the GPU, against the oracle (1e-12, helpers.py):

* another ElasticityNLAssembler material through the same gather / scatter machinery: SaintVenant (closed forms P = F S and
  its tangent in the generic kernel; the oracle differentiates the reference's energy expression by forward-mode autodiff),
  P1 .. P3, affine and per-quadrature-point geometry;
* genuinely NON-AFFINE elements: isoparametric P2 geometry with curved edges (the Jacobian varies over the quadrature points,
  finalize3d over P2 geometric bases, ElementAssemblyValues.cpp:65-104), J^-T and det*w per quadrature point handed to
  pfa_create as the reference's AssemblyValsCache would hold them - NeoHookean, SaintVenant and LinearElasticity (energy,
  gradient, Hessian / stiffness), and the Laplacian stiffness."""
import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close, gpu_handle, make_case
from polyfem_b200 import mesh as M, tables
from test_oracle_saint_venant_and_curved import curved_geometry, iso_problem

pytestmark = pytest.mark.gpu


def check_nl(h, ref, x):
    H = ref.assemble_hessian(x)
    outer, inner = h.pattern()
    assert outer.tobytes() == H.outer.tobytes() and inner.tobytes() == H.inner.tobytes()
    e, g, v = h.grad_hess(x)
    e_ref = ref.assemble_energy(x)
    assert abs(e - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(g, ref.assemble_gradient(x))
    assert_values_close(H.outer, H.inner, v, H.values)
    assert_vector_close(h.gradient(x), ref.assemble_gradient(x))
    assert abs(h.energy(x) - e_ref) <= REL_TOL * abs(e_ref)
    epe = h.energy_per_element(x)
    assert np.abs(epe - ref.assemble_energy_per_element(x)).max() <= REL_TOL * np.abs(epe).max()


@pytest.mark.parametrize("p,n", [(1, 4), (2, 3), (3, 2)])
def test_saint_venant_equals_oracle(oracle, p, n):
    mesh, x, t = make_case(n, p, jitter=0.2, scale=0.3)
    x = x[: mesh.n_bases * 3]
    check_nl(gpu_handle(mesh, "SaintVenant", t), oracle.problem_from_mesh(mesh, "SaintVenant", n_threads=2), x)


def test_saint_venant_rejects_what_it_does_not_have(oracle):
    from polyfem_b200 import capi
    mesh, x, t = make_case(2, 1)
    h = gpu_handle(mesh, "SaintVenant", t)
    with pytest.raises(capi.PfaError) as ei:
        h.linear_stiffness()
    assert ei.value.code == capi.PFA_ERR_UNSUPPORTED


def geometry_arrays(ref, mesh, t):
    """jac_it [n_el, n_qp, 9] and da [n_el, n_qp] as the reference's ElementAssemblyValues hold them."""
    nq = t["weights"].size
    jit, da = np.zeros((mesh.n_elements, nq, 9)), np.zeros((mesh.n_elements, nq))
    for e in range(mesh.n_elements):
        det, j, _ = ref.assembly_values(e)
        jit[e], da[e] = j.reshape(nq, 9), det * t["weights"]
    return jit, da


@pytest.mark.parametrize("material", ["NeoHookean", "SaintVenant", "LinearElasticity"])
def test_curved_p2_elements_equal_oracle(oracle, material):
    from polyfem_b200 import capi
    mesh = M.kuhn_cube(3, 2, jitter=0.1)
    t = tables.reference_tables(2)
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    ref = iso_problem(oracle, mesh, material, curved_geometry(mesh), n_threads=2)
    jit, da = geometry_arrays(ref, mesh, t)
    assert np.abs(jit - jit[:, :1]).max() > 1e-3 * np.abs(jit).max()  # really non-affine
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    h = capi.Handle(material, mesh.conn, mesh.n_bases, t["weights"], t["grad"], jac_it=jit, da=da, lam=lam, mu=mu)
    check_nl(h, ref, x)
    if material == "LinearElasticity":
        K = ref.assemble()
        assert_values_close(K.outer, K.inner, h.linear_stiffness(), K.values, what="stiffness")


def test_curved_p2_laplacian_stiffness(oracle):
    from polyfem_b200 import capi
    mesh = M.kuhn_cube(3, 2, jitter=0.1)
    t = tables.reference_tables(2)
    ref = iso_problem(oracle, mesh, "Laplacian", curved_geometry(mesh))
    jit, da = geometry_arrays(ref, mesh, t)
    h = capi.Handle("Laplacian", mesh.conn, mesh.n_bases, t["weights"], t["grad"], jac_it=jit, da=da)
    K = ref.assemble()
    outer, inner = h.pattern()
    assert outer.tobytes() == K.outer.tobytes() and inner.tobytes() == K.inner.tobytes()
    assert_values_close(K.outer, K.inner, h.linear_stiffness(), K.values, what="stiffness")
