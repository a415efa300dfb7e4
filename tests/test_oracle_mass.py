"""Oracle Mass assembler (assembler/Mass.cpp:5-23 through LinearAssembler::assemble,
Assembler.cpp:157-384) and InertiaForm (solver/forms/InertiaForm.cpp:17-34). The reference has no
stand-alone known-answer test for Mass (tests/test_form_derivatives.cpp:482,534 only feed it to
InertiaForm's finite-difference check), so the pins are the closed forms of the definition:
P1 local mass matrix rho*V/20*(1+delta_ij), total mass = rho * volume for every order, stored
zeros off the block diagonal, and the pattern of the elastic stiffness."""
import numpy as np
import pytest

from polyfem_b200 import mesh as M


def _component(K, c, n_bases):
    """Scalar matrix of component c (rows/cols c, c+3, ...) as dense."""
    A = K.to_scipy().toarray()
    return A[c::3, c::3]


def test_p1_local_mass_closed_form(oracle):
    mesh = M.kuhn_cube(1, 1, jitter=0.0)  # 6 tets of volume 1/6
    rho = 2.5
    K = oracle.problem_from_mesh(mesh, "Mass", rho=rho).assemble()
    ref = np.zeros((mesh.n_bases, mesh.n_bases))
    for e in range(mesh.n_elements):
        v = mesh.vertices[e]
        vol = abs(np.linalg.det(v[1:] - v[0])) / 6.0
        for i in range(4):
            for j in range(4):
                ref[mesh.conn[e, i], mesh.conn[e, j]] += rho * vol / 20.0 * (2.0 if i == j else 1.0)
    for c in range(3):
        assert np.abs(_component(K, c, mesh.n_bases) - ref).max() <= 1e-15
    A = K.to_scipy().toarray()
    for m in range(3):
        for n in range(3):
            if m != n:
                assert not A[m::3, n::3].any()  # stored zeros


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_total_mass_and_pattern(oracle, p):
    mesh = M.kuhn_cube(2 if p < 4 else 1, p, jitter=0.15)
    rho = 1.7
    K = oracle.problem_from_mesh(mesh, "Mass", rho=rho).assemble()
    vol = sum(abs(np.linalg.det(v[1:] - v[0])) / 6.0 for v in mesh.vertices)
    for c in range(3):
        assert abs(_component(K, c, mesh.n_bases).sum() - rho * vol) <= 1e-13 * rho * vol  # partition of unity
    S = oracle.problem_from_mesh(mesh, "LinearElasticity").assemble()
    assert K.outer.tobytes() == S.outer.tobytes() and K.inner.tobytes() == S.inner.tobytes()
    A = K.to_scipy()
    assert abs(A - A.T).max() <= 1e-16
    w = np.linalg.eigvalsh(_component(K, 0, mesh.n_bases))
    assert w.min() > 0  # consistent mass matrix is SPD


def test_inertia_form(oracle):
    mesh = M.kuhn_cube(2, 2, jitter=0.1)
    K = oracle.problem_from_mesh(mesh, "Mass", rho=3.0).assemble()
    rng = np.random.default_rng(0)
    x, xt = rng.standard_normal(3 * mesh.n_bases), rng.standard_normal(3 * mesh.n_bases)
    e, g = oracle.inertia(K, x, xt)
    # gradient is the derivative of the value (finite differences, like test_form_derivatives.cpp)
    d = rng.standard_normal(x.size)
    h = 1e-6
    fd = (oracle.inertia(K, x + h * d, xt)[0] - oracle.inertia(K, x - h * d, xt)[0]) / (2 * h)
    assert abs(fd - g @ d) <= 1e-6 * abs(g @ d)
    assert oracle.inertia(K, xt, xt)[0] == 0.0
