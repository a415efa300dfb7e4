"""Oracle additions of round 2, on the CPU:

* SaintVenant (assembler/SaintVenantElasticity.cpp:219-266 with the isotropic tensor of MatParams.cpp:211-253): the energy
  expression is restated, gradient / Hessian are its forward-mode derivatives as in the reference; pinned by properties
  (the reference's own tests for it need polyfem-data): finite differences, symmetry, Hessian at x = 0 == linear stiffness,
  the closed form psi = mu E:E + lambda/2 tr(E)^2 on one element.
* isoparametric geometry (finalize3d summing over P2 geometric bases, ElementAssemblyValues.cpp:65-104): with straight
  edges it reproduces the P1-geometry values; with curved edges the Jacobian varies over the quadrature points and the
  derivatives stay consistent."""
import numpy as np
import pytest

from polyfem_b200 import mesh as M, tables


def curved_geometry(mesh, amplitude=0.08, seed=5):
    """Isoparametric P2 geometry nodes [n_el, 10, 3]: the mesh's own P2 nodes with the edge midpoints pushed off their edges
    (the same displacement for a shared node, so the curved mesh stays conforming)."""
    rng = np.random.default_rng(seed)
    shift = amplitude * mesh.h * rng.uniform(-1, 1, size=mesh.node_xyz.shape)
    is_mid = np.ones(mesh.n_bases, dtype=bool)
    is_mid[np.unique(mesh.conn[:, :4])] = False
    xyz = mesh.node_xyz + shift * is_mid[:, None]
    return np.ascontiguousarray(xyz[mesh.conn])


def iso_problem(oracle, mesh, material, gnodes, **kw):
    t = tables.reference_tables(mesh.p)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    lat = np.array(tables.P_NODES_LATTICE[mesh.p], dtype=np.int32)
    return oracle.OracleProblem(material, mesh.conn, mesh.vertices, mesh.n_bases, t["points"], t["weights"], t["grad"], lam=lam, mu=mu,
                                basis_order=mesh.p, node_lattice=lat, geom_order=mesh.p, geom_lattice=lat, geom_nodes=gnodes, **kw)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_saint_venant_derivatives(oracle, p):
    mesh = M.kuhn_cube(2, p, jitter=0.2)
    x = M.random_displacement(mesh, scale=0.3)[: mesh.n_bases * 3]
    prob = oracle.problem_from_mesh(mesh, "SaintVenant")
    g, H = prob.assemble_gradient(x), prob.assemble_hessian(x).to_scipy()
    d = np.random.default_rng(0).standard_normal(x.size)
    eps = 1e-6
    fd = (prob.assemble_energy(x + eps * d) - prob.assemble_energy(x - eps * d)) / (2 * eps)
    assert abs(g @ d - fd) <= 1e-7 * abs(fd)
    gd = (prob.assemble_gradient(x + eps * d) - prob.assemble_gradient(x - eps * d)) / (2 * eps)
    assert np.abs(H @ d - gd).max() <= 1e-7 * np.abs(gd).max()
    assert abs(H - H.T).max() <= 1e-12 * abs(H).max()
    K = oracle.problem_from_mesh(mesh, "LinearElasticity").assemble().to_scipy()
    H0 = prob.assemble_hessian(0 * x).to_scipy()
    assert abs(H0 - K).max() <= 1e-13 * abs(K).max()


def test_saint_venant_energy_closed_form(oracle):
    mesh = M.kuhn_cube(1, 1, jitter=0.2)
    x = M.random_displacement(mesh, scale=0.4)[: mesh.n_bases * 3]
    prob = oracle.problem_from_mesh(mesh, "SaintVenant")
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    total = 0.0
    for e in range(mesh.n_elements):
        det, jit, gt = prob.assembly_values(e)
        u = x.reshape(-1, 3)[mesh.conn[e]]
        F = np.eye(3) + u.T @ gt[0]
        E = 0.5 * (F.T @ F - np.eye(3))
        total += (mu * np.sum(E * E) + 0.5 * lam * np.trace(E) ** 2) * det[0] / 6.0
    assert abs(prob.assemble_energy(x) - total) <= 1e-13 * abs(total)


def test_isoparametric_geometry_with_straight_edges_equals_p1_geometry(oracle):
    mesh = M.kuhn_cube(2, 2, jitter=0.2)
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    straight = np.ascontiguousarray(mesh.node_xyz[mesh.conn])
    a = oracle.problem_from_mesh(mesh, "NeoHookean")
    b = iso_problem(oracle, mesh, "NeoHookean", straight)
    for e in (0, 7):
        da, ja, ga = a.assembly_values(e)
        db, jb, gb = b.assembly_values(e)
        assert np.allclose(da, db, rtol=1e-12) and np.allclose(ja, jb, rtol=1e-11, atol=1e-11 * np.abs(ja).max())
    Ha, Hb = a.assemble_hessian(x), b.assemble_hessian(x)
    assert np.array_equal(Ha.inner, Hb.inner) and np.abs(Ha.values - Hb.values).max() <= 1e-11 * np.abs(Ha.values).max()


def test_curved_elements_are_not_affine_and_stay_consistent(oracle):
    mesh = M.kuhn_cube(2, 2, jitter=0.1)
    prob = iso_problem(oracle, mesh, "NeoHookean", curved_geometry(mesh))
    det, jit, _ = prob.assembly_values(3)
    assert det.min() > 0 and (det.max() - det.min()) > 1e-3 * det.mean()  # the Jacobian varies over the quadrature points
    assert np.abs(jit - jit[0]).max() > 1e-3 * np.abs(jit).max()
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    g, H = prob.assemble_gradient(x), prob.assemble_hessian(x).to_scipy()
    d = np.random.default_rng(1).standard_normal(x.size)
    eps = 1e-6 * mesh.h
    gd = (prob.assemble_gradient(x + eps * d) - prob.assemble_gradient(x - eps * d)) / (2 * eps)
    assert np.abs(H @ d - gd).max() <= 1e-6 * np.abs(gd).max()
    fd = (prob.assemble_energy(x + eps * d) - prob.assemble_energy(x - eps * d)) / (2 * eps)
    assert abs(g @ d - fd) <= 1e-6 * max(abs(fd), np.abs(g).max() * 1e-3)


def test_curved_geometry_equals_the_reference_finalize3d(oracle):
    """tests/golden/geom_iso.npz: det, jac_it and grad_t_m of CURVED elements (isoparametric P2 geometry under P2 and P3 bases) as
    returned by the reference's own ElementAssemblyValues::finalize3d (ElementAssemblyValues.cpp:65-104, compiled verbatim into
    oracle/_ref/libgeomref.so; `tools/make_golden.py geom_iso`). 1e-13 of the largest entry."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geom_iso.npz"))
    lat2 = np.array(tables.P_NODES_LATTICE[2], dtype=np.int32)
    for k in range(int(gold["n_cases"])):
        p = int(gold[f"p_{k}"])
        t = tables.reference_tables(p)
        nl = t["grad"].shape[1]
        prob = oracle.OracleProblem("NeoHookean", np.arange(nl, dtype=np.int32)[None, :], gold[f"vertices_{k}"][None], nl, t["points"], t["weights"],
                                    t["grad"], lam=1.0, mu=1.0, basis_order=p, node_lattice=np.array(tables.P_NODES_LATTICE[p], dtype=np.int32),
                                    geom_order=2, geom_lattice=lat2, geom_nodes=gold[f"geom_nodes_{k}"][None])
        det, jit, gt = prob.assembly_values(0)
        assert np.abs(det - gold[f"det_{k}"]).max() <= 1e-13 * np.abs(gold[f"det_{k}"]).max()
        assert np.abs(jit - gold[f"jac_it_{k}"]).max() <= 1e-13 * np.abs(gold[f"jac_it_{k}"]).max()
        assert np.abs(gt - gold[f"grad_t_m_{k}"]).max() <= 1e-13 * np.abs(gold[f"grad_t_m_{k}"]).max()
        assert np.abs(jit - jit[:1]).max() > 1e-3 * np.abs(jit).max()  # the Jacobian really varies over the element
