"""A fixed, documented Newton + backtracking loop (test infrastructure).

polysolve (the reference's nonlinear solver, cmake/recipes/polysolve.cmake:11) is not available
offline, so "identical Newton iteration counts" (BASELINE.json north_star) is checked with this
loop run twice on the same problem: once on the CPU oracle's assembly, once on the CUDA path.
It mirrors what the reference's callbacks do per iteration: value, gradient,
Hessian, a step-validity test that rejects NaN energies/gradients (ElasticForm::is_step_valid,
solver/forms/ElasticForm.cpp:388-396) and an Armijo backtracking line search.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def newton_solve(assemble, x0, free, max_it=30, rel_grad_tol=1e-8, c1=1e-4, max_ls=30):
    """assemble(x) -> (energy, gradient, csc_matrix). DOFs outside `free` stay at x0.
    Stops when |g_free|_inf <= rel_grad_tol * (its value at x0): far above rounding noise, so that
    every accept/reject decision of the line search is the same for two assemblies that agree to
    1e-12. Returns (x, history), history = [(|g_free|_inf, step size, line-search halvings), ...]."""
    x = x0.copy()
    hist = []
    gn0 = None
    for _ in range(max_it):
        e, g, H = assemble(x)
        gf = g[free]
        gn = float(np.abs(gf).max())
        gn0 = gn if gn0 is None else gn0
        if gn <= rel_grad_tol * gn0:
            hist.append((gn, 0.0, 0))
            break
        Hff = H[free][:, free].tocsc()
        dx = spla.spsolve(Hff, -gf)
        if not np.all(np.isfinite(dx)) or float(dx @ gf) >= 0.0:
            dx = -gf  # not a descent direction: gradient step (polysolve falls back similarly)
        slope = float(dx @ gf)
        alpha, halvings = 1.0, 0
        while halvings < max_ls:
            xt = x.copy()
            xt[free] += alpha * dx
            et, gt, _ = assemble(xt, hessian=False)
            if np.isfinite(et) and np.all(np.isfinite(gt)) and et <= e + c1 * alpha * slope:
                break
            alpha *= 0.5
            halvings += 1
        x = xt
        hist.append((gn, alpha, halvings))
    return x, hist


def newton_solve_reduced(assemble_reduced, probe, x0, free, max_it=30, rel_grad_tol=1e-8, c1=1e-4, max_ls=30):
    """The same loop on the Dirichlet-reduced system, the way NLProblem drives its forms
    (solver/NLProblem.cpp:565-640): assemble_reduced(x_full) -> (energy, reduced gradient, reduced
    csc Hessian) is NLProblem::value/gradient/hessian after BCLagrangianForm::project_*;
    probe(x_full) -> (is_step_valid, energy) is the line-search trial (is_step_valid + value).
    `free` = not_constraints_ (sorted). Returns (x, history) like newton_solve."""
    x = x0.copy()
    hist = []
    gn0 = None
    for _ in range(max_it):
        e, gf, Hff = assemble_reduced(x)
        gn = float(np.abs(gf).max())
        gn0 = gn if gn0 is None else gn0
        if gn <= rel_grad_tol * gn0:
            hist.append((gn, 0.0, 0))
            break
        dx = spla.spsolve(Hff.tocsc(), -gf)
        if not np.all(np.isfinite(dx)) or float(dx @ gf) >= 0.0:
            dx = -gf
        slope = float(dx @ gf)
        alpha, halvings = 1.0, 0
        while halvings < max_ls:
            xt = x.copy()
            xt[free] += alpha * dx
            ok, et = probe(xt)
            if ok and np.isfinite(et) and et <= e + c1 * alpha * slope:
                break
            alpha *= 0.5
            halvings += 1
        x = xt
        hist.append((gn, alpha, halvings))
    return x, hist
