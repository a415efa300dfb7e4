"""ViscousDamping (assembler/ViscousDamping.cpp) on the GPU against the oracle's restatement of the
reference's explicit formulas (9 x 9 second-derivative tensors; pinned by tests/test_oracle_viscous_damping.py). The kernel uses
the closed form documented in include/pfa.h (SaintVenant's tangent with substituted operands). P1 .. P3, affine and curved P2
elements, project_to_psd, the "no previous displacement" branch, and the host-side assembler class."""
import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close, make_case
from polyfem_b200 import mesh as M, tables
from test_gpu_saint_venant_and_curved import check_nl, geometry_arrays
from test_oracle_saint_venant_and_curved import curved_geometry

pytestmark = pytest.mark.gpu

PSI, PHI, DT = 30.0, 20.0, 0.05


def handle(mesh, t, **kw):
    from polyfem_b200 import capi
    return capi.Handle("ViscousDamping", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=PSI, mu=PHI, **kw)


def states(mesh):
    x0 = M.random_displacement(mesh, scale=0.1, seed=1)[: mesh.n_bases * 3]
    return x0, x0 + M.random_displacement(mesh, scale=0.03, seed=2)[: mesh.n_bases * 3]


@pytest.mark.parametrize("p,n", [(1, 4), (2, 3), (3, 2)])
def test_viscous_damping_equals_oracle(oracle, p, n):
    mesh, _, t = make_case(n, p, jitter=0.2)
    x0, x = states(mesh)
    ref = oracle.problem_from_mesh(mesh, "ViscousDamping", psi=PSI, phi=PHI, n_threads=2)
    h = handle(mesh, t)
    # no previous displacement yet: zeros, as the reference returns
    e, g, v = h.grad_hess(x)
    assert e == 0.0 and not g.any() and not v.any() and not h.energy_per_element(x).any()
    ref.set_previous(x0, DT)
    h.set_previous(x0, DT)
    check_nl(h, ref, x)
    # another time step size and previous state
    ref.set_previous(x, 0.5 * DT)
    h.set_previous(x, 0.5 * DT)
    check_nl(h, ref, x0)
    h.set_previous(None, DT)
    assert h.energy(x) == 0.0


def test_curved_p2_elements_and_projection(oracle):
    from polyfem_b200 import capi
    mesh = M.kuhn_cube(3, 2, jitter=0.1)
    t = tables.reference_tables(2)
    x0, x = states(mesh)
    lat = np.array(tables.P_NODES_LATTICE[2], dtype=np.int32)
    ref = oracle.OracleProblem("ViscousDamping", mesh.conn, mesh.vertices, mesh.n_bases, t["points"], t["weights"], t["grad"], lam=PSI, mu=PHI,
                               basis_order=2, node_lattice=lat, n_threads=2, geom_order=2, geom_lattice=lat, geom_nodes=curved_geometry(mesh))
    jit, da = geometry_arrays(ref, mesh, t)
    h = capi.Handle("ViscousDamping", mesh.conn, mesh.n_bases, t["weights"], t["grad"], jac_it=jit, da=da, lam=PSI, mu=PHI)
    ref.set_previous(x0, DT)
    h.set_previous(x0, DT)
    check_nl(h, ref, x)
    H0 = ref.assemble_hessian(x)
    v0 = H0.values.copy()
    H1 = ref.assemble_hessian(x, project_to_psd=True)
    assert_values_close(H1.outer, H1.inner, h.hessian(x, project_to_psd=True), H1.values, tol=1e-10, what="projected hessian")
    with pytest.raises(capi.PfaError) as ei:
        h.linear_stiffness()
    assert ei.value.code == capi.PFA_ERR_UNSUPPORTED
    with pytest.raises(capi.PfaError):
        h.set_previous(x0, 0.0)


def test_assembler_class(oracle):
    from polyfem_b200 import assembler as A
    mesh, _, t = make_case(3, 2, jitter=0.1)
    x0, x = states(mesh)
    a = A.make_assembler("ViscousDamping")
    a.set_materials([], {"psi": PSI, "phi": PHI})
    bases = A.FESpace.from_mesh(mesh)
    cache = A.AssemblyValsCache(mesh.p)
    ref = oracle.problem_from_mesh(mesh, "ViscousDamping", psi=PSI, phi=PHI)
    ref.set_previous(x0, DT)
    d, dp = x.reshape(-1, 1), x0.reshape(-1, 1)
    e = a.assemble_energy(True, bases, bases, cache, 0.0, DT, d, dp)
    assert abs(e - ref.assemble_energy(x)) <= REL_TOL * abs(e)
    g = a.assemble_gradient(True, mesh.n_bases, bases, bases, cache, 0.0, DT, d, dp)
    assert_vector_close(np.asarray(g).reshape(-1), ref.assemble_gradient(x))
    Hm = a.assemble_hessian(True, mesh.n_bases, False, bases, bases, cache, 0.0, DT, d, dp)
    H = ref.assemble_hessian(x)
    assert_values_close(H.outer, H.inner, np.asarray(Hm.data), H.values)
    # displacement_prev of another size: the first step of a simulation (ViscousDamping.cpp:125-126)
    assert a.assemble_energy(True, bases, bases, cache, 0.0, DT, d, np.zeros((0, 1))) == 0.0
