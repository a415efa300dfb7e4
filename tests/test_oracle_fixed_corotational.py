"""FixedCorotational (assembler/FixedCorotational.cpp: psi = mu sum (sigma_i - 1)^2 + lambda/2 (prod sigma - 1)^2 on the signed
singular values of F, stress and stiffness through the SVD): known answers for the oracle's restatement (own 3 x 3 SVD in the
manner of utils/svd.hpp: eigenvectors of F^T F, U from F V, sigma_2 negated for det F < 0).
  * rest: zero energy and stress, tangent = the LinearElasticity stiffness of the same (lambda, mu);
  * a rigid rotation of the whole body: zero energy and stress (sigma = 1);
  * uniform dilation x = s X: energy = vol (3 mu s^2 + lambda/2 ((1 + s)^3 - 1)^2);
  * gradient / Hessian are the finite differences of energy / gradient, also for an inverted state (det F < 0 in some elements:
    allow_inversion() is true for this material)."""
import numpy as np

from polyfem_b200 import mesh as M


def test_rest_state_is_linear_elasticity(oracle):
    mesh = M.kuhn_cube(2, 2, jitter=0.15)
    pb = oracle.problem_from_mesh(mesh, "FixedCorotational")
    z = np.zeros(mesh.n_bases * 3)
    assert abs(pb.assemble_energy(z)) < 1e-20
    H = pb.assemble_hessian(z)
    assert np.abs(pb.assemble_gradient(z)).max() <= 1e-12 * np.abs(H.values).max()
    S = oracle.problem_from_mesh(mesh, "LinearElasticity").assemble()
    assert np.array_equal(S.outer, H.outer) and np.array_equal(S.inner, H.inner)
    assert np.abs(S.values - H.values).max() <= 1e-11 * np.abs(S.values).max()


def test_rotation_and_dilation(oracle):
    mesh = M.kuhn_cube(2, 1, jitter=0.1)
    pb = oracle.problem_from_mesh(mesh, "FixedCorotational")
    X = mesh.node_xyz
    th = 0.9
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]]) @ np.array([[1, 0, 0], [0, np.cos(0.4), -np.sin(0.4)], [0, np.sin(0.4), np.cos(0.4)]])
    xr = (X @ R.T - X).reshape(-1)
    H = pb.assemble_hessian(xr)
    assert abs(pb.assemble_energy(xr)) <= 1e-20 * np.abs(H.values).max()
    assert np.abs(pb.assemble_gradient(xr)).max() <= 1e-12 * np.abs(H.values).max()
    s = 0.06
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    want = 1.0 * (3 * mu * s * s + lam / 2 * ((1 + s) ** 3 - 1) ** 2)
    assert abs(pb.assemble_energy((s * X).reshape(-1)) - want) <= 1e-12 * want


def _fd_check(pb, x):
    g = pb.assemble_gradient(x)
    H = pb.assemble_hessian(x).to_scipy()
    d = np.random.default_rng(0).standard_normal(x.size)
    h = 1e-6
    fd_e = (pb.assemble_energy(x + h * d) - pb.assemble_energy(x - h * d)) / (2 * h)
    assert abs(fd_e - g @ d) <= 1e-6 * max(abs(g @ d), np.abs(g).max())
    fd_g = (pb.assemble_gradient(x + h * d) - pb.assemble_gradient(x - h * d)) / (2 * h)
    assert np.abs(fd_g - H @ d).max() <= 1e-6 * np.abs(H @ d).max()
    assert abs(H - H.T).max() <= 1e-11 * abs(H).max()


def test_finite_differences(oracle):
    mesh = M.kuhn_cube(2, 2, jitter=0.15)
    pb = oracle.problem_from_mesh(mesh, "FixedCorotational")
    _fd_check(pb, M.random_displacement(mesh, scale=0.1)[: mesh.n_bases * 3])


def test_finite_differences_with_inverted_elements(oracle):
    mesh = M.kuhn_cube(2, 1, jitter=0.1)
    pb = oracle.problem_from_mesh(mesh, "FixedCorotational")
    x = M.random_displacement(mesh, scale=0.9, seed=3)[: mesh.n_bases * 3]
    epe = pb.assemble_energy_per_element(x)
    assert np.isfinite(epe).all()
    _fd_check(pb, x)
