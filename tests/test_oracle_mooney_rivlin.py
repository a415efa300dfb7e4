"""MooneyRivlin (MooneyRivlinElasticity.hpp:26-47 through GenericElastic, autodiff): known answers that pin the oracle's
restatement without the reference binary.
  * rest state: psi = 0 and the stress vanishes (I1~ = I2~ = 3, ln J = 0), so energy and gradient are zero at x = 0;
  * small strains: the tangent at rest is the isotropic linear-elastic one with shear modulus 2 (c1 + c2) and bulk modulus k,
    i.e. the oracle's LinearElasticity stiffness for mu = 2 (c1 + c2), lambda = k - 2 mu / 3 (textbook linearisation);
  * a pure dilation x = s X: F = (1 + s) I, the isochoric part is the identity: energy = vol * k/2 ln^2((1 + s)^3) exactly;
  * rigid rotations leave the energy unchanged (objectivity); gradient / Hessian are consistent with finite differences of
    energy / gradient (the autodiff `pow` rule added for this material)."""
import numpy as np

from polyfem_b200 import mesh as M

C1, C2, K = 11000.0, 7000.0, 90000.0


def _problem(oracle, mesh, **kw):
    return oracle.problem_from_mesh(mesh, "MooneyRivlin", c1=C1, c2=C2, k=K, **kw)


def test_rest_state_and_small_strain_limit(oracle):
    mesh = M.kuhn_cube(2, 2, jitter=0.15)
    pb = _problem(oracle, mesh)
    z = np.zeros(mesh.n_bases * 3)
    assert abs(pb.assemble_energy(z)) < 1e-18
    g = pb.assemble_gradient(z)
    H = pb.assemble_hessian(z)
    assert np.abs(g).max() <= 1e-12 * np.abs(H.values).max()
    mu = 2.0 * (C1 + C2)
    lam = K - 2.0 * mu / 3.0
    E = mu * (3 * lam + 2 * mu) / (lam + mu)
    nu = lam / (2 * (lam + mu))
    le = oracle.problem_from_mesh(mesh, "LinearElasticity", E=E, nu=nu)
    S = le.assemble()
    assert np.array_equal(S.outer, H.outer) and np.array_equal(S.inner, H.inner)
    assert np.abs(S.values - H.values).max() <= 1e-11 * np.abs(S.values).max()


def test_pure_dilation_energy(oracle):
    mesh = M.kuhn_cube(2, 1)
    pb = _problem(oracle, mesh)
    s = 0.07
    X = np.zeros((mesh.n_bases, 3))
    for e in range(mesh.conn.shape[0]):
        X[mesh.conn[e]] = mesh.vertices[e]
    x = (s * X).reshape(-1)
    want = 1.0 * K / 2.0 * np.log((1.0 + s) ** 3) ** 2
    assert abs(pb.assemble_energy(x) - want) <= 1e-12 * want


def test_objectivity_and_finite_differences(oracle):
    mesh = M.kuhn_cube(2, 2, jitter=0.1)
    pb = _problem(oracle, mesh)
    x = M.random_displacement(mesh, scale=0.05)[: mesh.n_bases * 3]
    e0 = pb.assemble_energy(x)
    # rotate the deformed configuration about the origin: x' = R (X + u) - X
    X = np.zeros((mesh.n_bases, 3))
    t = __import__("polyfem_b200.tables", fromlist=["x"])
    lat = np.array(t.P_NODES_LATTICE[2], dtype=np.float64) / 2.0
    for e in range(mesh.conn.shape[0]):
        v = mesh.vertices[e]
        X[mesh.conn[e]] = v[0] + lat @ (v[1:] - v[0])
    th = 0.7
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]]) @ np.array([[1, 0, 0], [0, np.cos(0.3), -np.sin(0.3)], [0, np.sin(0.3), np.cos(0.3)]])
    xr = ((X + x.reshape(-1, 3)) @ R.T - X).reshape(-1)
    assert abs(pb.assemble_energy(xr) - e0) <= 1e-11 * abs(e0)
    g = pb.assemble_gradient(x)
    H = pb.assemble_hessian(x)
    rng = np.random.default_rng(1)
    d = rng.standard_normal(x.size)
    h = 1e-6
    fd_e = (pb.assemble_energy(x + h * d) - pb.assemble_energy(x - h * d)) / (2 * h)
    assert abs(fd_e - g @ d) <= 1e-6 * max(abs(g @ d), np.abs(g).max())
    fd_g = (pb.assemble_gradient(x + h * d) - pb.assemble_gradient(x - h * d)) / (2 * h)
    Hd = np.zeros_like(d)
    for c in range(x.size):
        rows = H.inner[H.outer[c]:H.outer[c + 1]]
        Hd[rows] += H.values[H.outer[c]:H.outer[c + 1]] * d[c]
    assert np.abs(fd_g - Hd).max() <= 1e-6 * np.abs(Hd).max()
