"""Newton iteration counts: the same fixed Newton loop (tests/newton_loop.py) driven by the CPU
oracle and by the CUDA path must take the same number of iterations with the same step sizes
(north_star: "Newton iteration counts must be identical on the configs"; polysolve itself is
not available offline)."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import gpu_handle, make_case
from newton_loop import newton_solve

pytestmark = pytest.mark.gpu


def _problem(mesh, stretch, linear_init=False):
    z = mesh.node_xyz[:, 2]
    bottom, top = np.where(z < 1e-12)[0], np.where(z > 1 - 1e-12)[0]
    x0 = np.zeros(mesh.n_bases * 3)
    # clamp the bottom face, pull and twist the top face; interior starts undeformed
    ang = 0.35 * stretch
    xy = mesh.node_xyz[top, :2] - 0.5
    rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
    x0.reshape(-1, 3)[top, :2] = xy @ rot.T - xy
    x0.reshape(-1, 3)[top, 2] = stretch
    if linear_init:  # start from the top-face motion scaled by height (no inverted element at the start)
        a = 0.35 * stretch * z
        c, s_ = np.cos(a), np.sin(a)
        xy_all = mesh.node_xyz[:, :2] - 0.5
        x0.reshape(-1, 3)[:, 0] = c * xy_all[:, 0] - s_ * xy_all[:, 1] - xy_all[:, 0]
        x0.reshape(-1, 3)[:, 1] = s_ * xy_all[:, 0] + c * xy_all[:, 1] - xy_all[:, 1]
        x0.reshape(-1, 3)[:, 2] = stretch * z
    fixed = np.concatenate([bottom, top])
    mask = np.ones(mesh.n_bases * 3, dtype=bool)
    mask.reshape(-1, 3)[fixed] = False
    return x0, np.where(mask)[0]


# (basis order, cells per side, top-face stretch, start from the interpolated motion)
CASES = [(1, 5, 0.25, False), (2, 3, 0.25, False), (2, 3, -0.3, True), (1, 4, 2.5, False)]


@pytest.mark.parametrize("p,n,stretch,linear_init", CASES)
def test_newton_iteration_counts_match(oracle, p, n, stretch, linear_init):
    mesh, _, t = make_case(n, p)
    x0, free = _problem(mesh, stretch, linear_init)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=2)
    h = gpu_handle(mesh, "NeoHookean", t)
    outer, inner = h.pattern()
    ndof = h.ndof

    def asm_ref(x, hessian=True):
        e, g = ref.assemble_energy(x), ref.assemble_gradient(x)
        if not hessian:
            return e, g, None
        H = ref.assemble_hessian(x)
        return e, g, sp.csc_matrix((H.values, H.inner, H.outer), shape=(ndof, ndof))

    def asm_gpu(x, hessian=True):
        if not hessian:
            return h.energy(x), h.gradient(x), None
        e, g, v = h.grad_hess(x)
        return e, g, sp.csc_matrix((v, inner, outer), shape=(ndof, ndof))

    x_ref, hist_ref = newton_solve(asm_ref, x0, free)
    x_gpu, hist_gpu = newton_solve(asm_gpu, x0, free)
    assert len(hist_ref) >= 3, "problem too easy to say anything about iteration counts"
    assert hist_ref[-1][1] == 0.0 and hist_ref[-1][0] <= 1e-8 * hist_ref[0][0], f"reference loop did not converge: {hist_ref}"
    assert len(hist_gpu) == len(hist_ref), (hist_ref, hist_gpu)
    assert [s[1:] for s in hist_gpu] == [s[1:] for s in hist_ref], (hist_ref, hist_gpu)  # step sizes, halvings
    if stretch == 2.5:
        assert any(s[2] > 0 for s in hist_ref), "this case is meant to exercise the line search"
    assert np.abs(x_gpu - x_ref).max() <= 1e-9 * max(1.0, np.abs(x_ref).max())


@pytest.mark.parametrize("p,n,stretch,linear_init", [CASES[1], CASES[3]])
def test_newton_on_the_reduced_system(oracle, p, n, stretch, linear_init):
    """The loop NLProblem runs: Dirichlet-reduced gradient / Hessian and the is_step_valid probe.
    CPU side: oracle assembly + the oracle's BCLagrangianForm::project_* restatement;
    GPU side: pfa_grad_hess_reduced (fused projection) + pfa_is_step_valid. Same iteration
    counts, step sizes and solution."""
    from newton_loop import newton_solve_reduced
    mesh, _, t = make_case(n, p)
    x0, free = _problem(mesh, stretch, linear_init)
    ndof = mesh.n_bases * 3
    constrained = np.setdiff1d(np.arange(ndof), free)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=2)
    h = gpu_handle(mesh, "NeoHookean", t)
    h.set_constrained_dofs(constrained)
    o_r, i_r = h.reduced_pattern()
    nred = h.ndof_reduced
    assert nred == free.size

    def asm_ref(x):
        e, g, H = ref.assemble_energy(x), ref.assemble_gradient(x), ref.assemble_hessian(x)
        R = oracle.project_hessian(H, constrained)
        return e, oracle.project_gradient(g, constrained), sp.csc_matrix((R.values, R.inner, R.outer), shape=(nred, nred))

    def probe_ref(x):
        g = ref.assemble_gradient(x)  # ElasticForm::is_step_valid: gradient, then NaN check
        return (not np.isnan(g).any()), ref.assemble_energy(x)

    def asm_gpu(x):
        e, g, v = h.grad_hess_reduced(x)
        return e, g, sp.csc_matrix((v, i_r, o_r), shape=(nred, nred))

    x_ref, hist_ref = newton_solve_reduced(asm_ref, probe_ref, x0, free)
    x_gpu, hist_gpu = newton_solve_reduced(asm_gpu, h.is_step_valid, x0, free)
    assert len(hist_ref) >= 3 and hist_ref[-1][1] == 0.0
    assert len(hist_gpu) == len(hist_ref), (hist_ref, hist_gpu)
    assert [s_[1:] for s_ in hist_gpu] == [s_[1:] for s_ in hist_ref], (hist_ref, hist_gpu)
    assert np.abs(x_gpu - x_ref).max() <= 1e-9 * max(1.0, np.abs(x_ref).max())
