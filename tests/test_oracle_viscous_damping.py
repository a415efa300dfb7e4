"""ViscousDamping (assembler/ViscousDamping.cpp: R = psi |dE/dt|^2 + phi/2 tr(dE/dt)^2, dE/dt = sym(dF/dt^T F),
dF/dt = (F - F_prev) / dt): known answers for the oracle's restatement.
  * no previous displacement: everything is zero (the reference's size check);
  * x == x_prev: zero energy and gradient;
  * uniform stretching x_prev = 0, x = s X: dE/dt = s (1 + s) / dt I, energy = vol (3 psi + 9/2 phi) (s (1 + s) / dt)^2;
  * the reference's gradient and Hessian formulas (explicit 9 x 9 tensors) are the derivatives of its energy in x: finite differences."""
import numpy as np

from polyfem_b200 import mesh as M

PSI, PHI, DT = 30.0, 20.0, 0.05


def _problem(oracle, mesh, **kw):
    return oracle.problem_from_mesh(mesh, "ViscousDamping", psi=PSI, phi=PHI, **kw)


def test_without_previous_and_at_rest(oracle):
    mesh = M.kuhn_cube(2, 2, jitter=0.1)
    pb = _problem(oracle, mesh)
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    assert pb.assemble_energy(x) == 0.0 and not pb.assemble_gradient(x).any() and not pb.assemble_hessian(x).values.any()
    pb.set_previous(x, DT)
    assert pb.assemble_energy(x) == 0.0
    H = pb.assemble_hessian(x)
    assert np.abs(pb.assemble_gradient(x)).max() <= 1e-14 * np.abs(H.values).max()
    S = H.to_scipy()
    assert abs(S - S.T).max() <= 1e-12 * np.abs(H.values).max()
    pb.set_previous(None, DT)
    assert pb.assemble_energy(x) == 0.0


def test_uniform_stretching(oracle):
    mesh = M.kuhn_cube(2, 1)
    pb = _problem(oracle, mesh)
    s = 0.04
    x = (s * mesh.node_xyz).reshape(-1)
    pb.set_previous(np.zeros_like(x), DT)
    rate = s * (1 + s) / DT
    want = 1.0 * (3 * PSI + 4.5 * PHI) * rate * rate
    assert abs(pb.assemble_energy(x) - want) <= 1e-12 * want


def test_derivatives_by_finite_differences(oracle):
    mesh = M.kuhn_cube(2, 2, jitter=0.15)
    pb = _problem(oracle, mesh)
    x0 = M.random_displacement(mesh, scale=0.1, seed=1)[: mesh.n_bases * 3]
    x = x0 + M.random_displacement(mesh, scale=0.03, seed=2)[: mesh.n_bases * 3]
    pb.set_previous(x0, DT)
    g = pb.assemble_gradient(x)
    H = pb.assemble_hessian(x).to_scipy()
    d = np.random.default_rng(0).standard_normal(x.size)
    h = 1e-6
    fd_e = (pb.assemble_energy(x + h * d) - pb.assemble_energy(x - h * d)) / (2 * h)
    assert abs(fd_e - g @ d) <= 1e-7 * max(abs(g @ d), np.abs(g).max())
    fd_g = (pb.assemble_gradient(x + h * d) - pb.assemble_gradient(x - h * d)) / (2 * h)
    assert np.abs(fd_g - H @ d).max() <= 1e-7 * np.abs(H @ d).max()
    assert abs(H - H.T).max() <= 1e-12 * abs(H).max()
