"""MooneyRivlin (MooneyRivlinElasticity.hpp:26-47 through GenericElastic) on the GPU against the oracle
(1e-12, helpers.py). The reference and the oracle differentiate the energy expression by forward-mode autodiff; the kernel
uses the chain rule over (I1, I2, J) in closed form (pfa_kernels.cu). The oracle's restatement is pinned by the known answers
of tests/test_oracle_mooney_rivlin.py. P1 .. P4, per-element (c1, c2, k) that change between calls, curved P2 geometry,
project_to_psd, and the host-side assembler class."""
import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close, make_case
from polyfem_b200 import mesh as M, tables
from test_gpu_saint_venant_and_curved import check_nl, geometry_arrays
from test_oracle_saint_venant_and_curved import curved_geometry

pytestmark = pytest.mark.gpu

C1, C2, K = 11000.0, 7000.0, 90000.0


def handle(mesh, t, c1=C1, c2=C2, k=K, **kw):
    from polyfem_b200 import capi
    return capi.Handle("MooneyRivlin", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=c1, mu=c2, param3=k, **kw)


@pytest.mark.parametrize("p,n", [(1, 4), (2, 3), (3, 2), (4, 1)])
def test_mooney_rivlin_equals_oracle(oracle, p, n):
    mesh, x, t = make_case(n, p, jitter=0.2, scale={1: 0.2, 2: 0.08, 3: 0.03, 4: 0.02}[p])
    x = x[: mesh.n_bases * 3]
    assert np.isfinite(oracle.problem_from_mesh(mesh, "MooneyRivlin", c1=C1, c2=C2, k=K).assemble_energy(x))
    check_nl(handle(mesh, t), oracle.problem_from_mesh(mesh, "MooneyRivlin", c1=C1, c2=C2, k=K, n_threads=2), x)


def test_per_element_parameters_and_update(oracle):
    from polyfem_b200 import capi
    mesh, x, t = make_case(3, 2, jitter=0.1, scale=0.08)
    x = x[: mesh.n_bases * 3]
    rng = np.random.default_rng(3)
    ne = mesh.n_elements
    c1, c2, k = C1 * rng.uniform(0.5, 1.5, ne), C2 * rng.uniform(0.5, 1.5, ne), K * rng.uniform(0.5, 1.5, ne)
    h = handle(mesh, t, c1, c2, k)
    check_nl(h, oracle.problem_from_mesh(mesh, "MooneyRivlin", c1=c1, c2=c2, k=k), x)
    # new values through pfa_set_material_params (the elements are re-ordered internally: rows are gathered on the device)
    c1b, c2b, kb = c1[::-1].copy(), 2.0 * c2, 0.5 * k
    h.set_materials(c1b, c2b, 1, param3=kb)
    check_nl(h, oracle.problem_from_mesh(mesh, "MooneyRivlin", c1=c1b, c2=c2b, k=kb), x)
    with pytest.raises(capi.PfaError):  # two-parameter call on a three-parameter material
        h.set_materials(c1, c2, 1)
    with pytest.raises(capi.PfaError) as ei:  # pfa_create without k
        capi.Handle("MooneyRivlin", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=C1, mu=C2)
    assert ei.value.code == capi.PFA_ERR_INVALID
    with pytest.raises(capi.PfaError) as ei:
        h.linear_stiffness()
    assert ei.value.code == capi.PFA_ERR_UNSUPPORTED


def test_parameters_per_quadrature_point(oracle):
    """material_stride == n_qp (what the host shim sends: one value per element and quadrature point)"""
    mesh, x, t = make_case(3, 2, jitter=0.1, scale=0.08)
    x = x[: mesh.n_bases * 3]
    rng = np.random.default_rng(4)
    ne, nq = mesh.n_elements, t["weights"].size
    c1, c2, k = C1 * rng.uniform(0.5, 1.5, ne), C2 * rng.uniform(0.5, 1.5, ne), K * rng.uniform(0.5, 1.5, ne)
    h = handle(mesh, t, np.repeat(c1, nq), np.repeat(c2, nq), np.repeat(k, nq))
    check_nl(h, oracle.problem_from_mesh(mesh, "MooneyRivlin", c1=c1, c2=c2, k=k), x)
    h.set_materials(np.repeat(c2, nq), np.repeat(c1, nq), nq, param3=np.repeat(2.0 * k, nq))
    check_nl(h, oracle.problem_from_mesh(mesh, "MooneyRivlin", c1=c2, c2=c1, k=2.0 * k), x)


def test_curved_p2_elements(oracle):
    from polyfem_b200 import capi
    mesh = M.kuhn_cube(3, 2, jitter=0.1)
    t = tables.reference_tables(2)
    x = M.random_displacement(mesh, scale=0.08)[: mesh.n_bases * 3]
    lat = np.array(tables.P_NODES_LATTICE[2], dtype=np.int32)
    ref = oracle.OracleProblem("MooneyRivlin", mesh.conn, mesh.vertices, mesh.n_bases, t["points"], t["weights"], t["grad"], lam=C1, mu=C2, param3=K,
                               basis_order=2, node_lattice=lat, n_threads=2, geom_order=2, geom_lattice=lat, geom_nodes=curved_geometry(mesh))
    jit, da = geometry_arrays(ref, mesh, t)
    h = capi.Handle("MooneyRivlin", mesh.conn, mesh.n_bases, t["weights"], t["grad"], jac_it=jit, da=da, lam=C1, mu=C2, param3=K)
    check_nl(h, ref, x)


@pytest.mark.parametrize("p,n", [(1, 3), (2, 2)])
def test_projection(oracle, p, n):
    mesh, x, t = make_case(n, p, jitter=0.1, scale=0.2 if p == 1 else 0.12)
    x = x[: mesh.n_bases * 3]
    ref = oracle.problem_from_mesh(mesh, "MooneyRivlin", c1=C1, c2=C2, k=K, n_threads=2)
    h = handle(mesh, t)
    H0 = ref.assemble_hessian(x)
    H1 = ref.assemble_hessian(x, project_to_psd=True)
    assert np.abs(H0.values - H1.values).max() > 1e-4 * np.abs(H0.values).max(), "projection inactive: test is vacuous"
    assert_values_close(H1.outer, H1.inner, h.hessian(x, project_to_psd=True), H1.values, tol=1e-10, what="projected hessian")


def test_assembler_class(oracle):
    from polyfem_b200 import assembler as A
    mesh, x, t = make_case(3, 2, jitter=0.1, scale=0.08)
    x = x[: mesh.n_bases * 3]
    a = A.make_assembler("MooneyRivlin")
    assert a.name() == "MooneyRivlin" and not a.is_linear()
    a.set_materials([], {"c1": C1, "c2": C2, "k": K})
    bases = A.FESpace.from_mesh(mesh)
    cache = A.AssemblyValsCache(mesh.p)
    ref = oracle.problem_from_mesh(mesh, "MooneyRivlin", c1=C1, c2=C2, k=K)
    disp = x.reshape(-1, 1)
    e = a.assemble_energy(True, bases, bases, cache, 0.0, 1.0, disp, disp)
    assert abs(e - ref.assemble_energy(x)) <= REL_TOL * abs(e)
    g = a.assemble_gradient(True, mesh.n_bases, bases, bases, cache, 0.0, 1.0, disp, disp)
    assert_vector_close(np.asarray(g).reshape(-1), ref.assemble_gradient(x))
    with pytest.raises(RuntimeError):
        A.make_assembler("MooneyRivlin").set_materials([], {"E": 1e5, "nu": 0.3})
