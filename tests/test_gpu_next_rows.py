"""GPU parity of the rows right after the assembly: Dirichlet projection of
gradient / Hessian and the line-search validity probe, against the oracle restatements.
Gathers are exact: projected values must be BIT-identical to scale * full values."""
import numpy as np
import pytest

from helpers import REL_TOL, gpu_handle, make_case

pytestmark = pytest.mark.gpu


def _constraints(mesh, rng):
    """Dirichlet set like a clamped face: all dofs of nodes on x = 0, plus single components elsewhere."""
    on_face = np.flatnonzero(mesh.node_xyz[:, 0] < 1e-12)
    dofs = (on_face[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
    extra = rng.choice(mesh.n_bases * 3, size=7, replace=False)
    return rng.permutation(np.union1d(dofs, extra)).astype(np.int32)


@pytest.mark.parametrize("p,n", [(1, 4), (2, 3)])
def test_projection_pattern_and_values(oracle, p, n):
    mesh, x, t = make_case(n, p)
    rng = np.random.default_rng(5)
    h = gpu_handle(mesh, "NeoHookean", t)
    e, g, v = h.grad_hess(x)
    outer, inner = h.pattern()
    full = oracle.CSC(h.ndof, outer, inner, v)
    constrained = _constraints(mesh, rng)
    h.set_constrained_dofs(constrained)
    ref = oracle.project_hessian(full, constrained)
    assert h.ndof_reduced == ref.n and h.nnz_reduced == ref.inner.size
    o_r, i_r = h.reduced_pattern()
    assert o_r.dtype == np.int32 and o_r.tobytes() == ref.outer.tobytes() and i_r.tobytes() == ref.inner.tobytes()
    assert np.array_equal(h.project_hessian(v), ref.values)
    assert np.array_equal(h.project_hessian(v, scale=0.37), 0.37 * ref.values)
    g_ref = oracle.project_gradient(g, constrained)
    assert np.array_equal(h.project_gradient(g), g_ref)
    assert np.array_equal(h.project_gradient(g, scale=-2.5), -2.5 * g_ref)
    # a second constraint set replaces the first; the empty set is the identity
    h.set_constrained_dofs([])
    assert h.ndof_reduced == h.ndof and h.nnz_reduced == h.nnz
    o2, i2 = h.reduced_pattern()
    assert o2.tobytes() == outer.tobytes() and i2.tobytes() == inner.tobytes()
    assert np.array_equal(h.project_hessian(v), v)


def test_projection_stays_on_the_device(oracle):
    import torch
    mesh, x, t = make_case(3, 2)
    h = gpu_handle(mesh, "NeoHookean", t)
    constrained = _constraints(mesh, np.random.default_rng(6))
    h.set_constrained_dofs(constrained)
    xd = torch.from_numpy(x).cuda()
    e = torch.zeros(1, dtype=torch.float64, device="cuda")
    g = torch.zeros(h.ndof, dtype=torch.float64, device="cuda")
    v = torch.zeros(h.nnz, dtype=torch.float64, device="cuda")
    h.grad_hess_raw(xd, e, g, v)
    gr = torch.zeros(h.ndof_reduced, dtype=torch.float64, device="cuda")
    vr = torch.zeros(h.nnz_reduced, dtype=torch.float64, device="cuda")
    h.project_gradient(g, 1.0, out=gr)
    h.project_hessian(v, 1.0, out=vr)
    h.synchronize()
    outer, inner = h.pattern()
    ref = oracle.project_hessian(oracle.CSC(h.ndof, outer, inner, v.cpu().numpy()), constrained)
    assert np.array_equal(vr.cpu().numpy(), ref.values)
    assert np.array_equal(gr.cpu().numpy(), oracle.project_gradient(g.cpu().numpy(), constrained))
    po, pi = h.reduced_pattern_device_ptrs()
    assert po and pi


def test_constraint_errors(oracle):
    from polyfem_b200 import capi
    mesh, x, t = make_case(2, 1)
    h = gpu_handle(mesh, "NeoHookean", t)
    with pytest.raises(capi.PfaError):
        h.project_hessian(np.zeros(h.nnz), out=np.zeros(h.nnz))  # no constraint set yet
    with pytest.raises(capi.PfaError) as ei:
        h.set_constrained_dofs([0, h.ndof])  # out of range
    assert ei.value.code == capi.PFA_ERR_INVALID


@pytest.mark.parametrize("p,n", [(1, 4), (2, 3)])
def test_is_step_valid_matches_gradient_nan_check(oracle, p, n):
    """ElasticForm::is_step_valid: valid <=> the assembled gradient has no NaN; the energy of the
    same pass equals assemble_energy."""
    mesh, x, t = make_case(n, p)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    h = gpu_handle(mesh, "NeoHookean", t)
    valid, e = h.is_step_valid(x)
    assert valid and not np.isnan(ref.assemble_gradient(x)).any()
    e_ref = ref.assemble_energy(x)
    assert abs(e - e_ref) <= REL_TOL * abs(e_ref)
    xi = x.copy()
    nodes = mesh.conn[5]
    xi.reshape(-1, 3)[nodes[1]] += 3.0 * (mesh.node_xyz[nodes[0]] - mesh.node_xyz[nodes[1]])  # inverts element 5
    valid, e = h.is_step_valid(xi)
    assert not valid and np.isnan(ref.assemble_gradient(xi)).any()
    assert np.isnan(e) == np.isnan(ref.assemble_energy(xi))
    valid, e = h.is_step_valid(x, want_energy=False)
    assert valid and e is None


@pytest.mark.parametrize("p,n", [(1, 4), (2, 3)])
@pytest.mark.parametrize("psd", [False, True])
def test_fused_reduced_assembly_equals_assemble_then_project(oracle, p, n, psd):
    """pfa_grad_hess_reduced == project(pfa_grad_hess): same reduced pattern, values within the
    summation-order tolerance, gradient / energy scaled by the Form weight."""
    from helpers import assert_values_close, assert_vector_close
    mesh, x, t = make_case(n, p, scale=0.12 if psd else 0.05)
    h = gpu_handle(mesh, "NeoHookean", t)
    constrained = _constraints(mesh, np.random.default_rng(8))
    h.set_constrained_dofs(constrained)
    w = 0.25
    e_full, g_full, v_full = h.grad_hess(x, project_to_psd=psd)
    outer, inner = h.pattern()
    ref = oracle.project_hessian(oracle.CSC(h.ndof, outer, inner, w * v_full), constrained)
    g_ref = oracle.project_gradient(w * g_full, constrained)
    e, g, v = h.grad_hess_reduced(x, scale=w, project_to_psd=psd)
    assert abs(e - w * e_full) <= REL_TOL * abs(e_full)
    assert_vector_close(g, g_ref)
    assert_values_close(ref.outer, ref.inner, v, ref.values, what="fused reduced hessian")
    # against the oracle's own assembly as well
    H = oracle.problem_from_mesh(mesh, "NeoHookean").assemble_hessian(x, project_to_psd=psd)
    ref2 = oracle.project_hessian(oracle.CSC(h.ndof, H.outer, H.inner, w * H.values), constrained)
    assert_values_close(ref2.outer, ref2.inner, v, ref2.values, what="fused reduced hessian vs oracle")
    # full-size call after a reduced one still works (shared staging buffers)
    e2, g2, v2 = h.grad_hess(x, project_to_psd=psd)
    assert_values_close(outer, inner, v2, v_full)


def test_fused_reduced_assembly_unsupported_paths(oracle):
    from polyfem_b200 import capi
    mesh, x, t = make_case(2, 3)
    h = gpu_handle(mesh, "NeoHookean", t)
    h.set_constrained_dofs([0, 1, 2])
    with pytest.raises(capi.PfaError) as ei:
        h.grad_hess_reduced(x)
    assert ei.value.code == capi.PFA_ERR_UNSUPPORTED  # P3 goes through the generic kernel: use pfa_project_*


@pytest.mark.parametrize("p,n,n_first", [(1, 4, 37), (2, 3, 50), (2, 3, 0), (3, 2, 11)])
def test_two_part_assembly_equals_one_launch(oracle, p, n, n_first):
    """pfa_grad_hess_part(FIRST) + (REST) == pfa_grad_hess (row-lane and generic kernels); the
    internal re-ordering keeps the two element groups apart."""
    import torch
    from helpers import assert_values_close, assert_vector_close
    from polyfem_b200 import capi, mesh as M
    mesh, x, t = make_case(n, p, jitter=0.1, scale=0.05 if p < 3 else 0.01)  # P3 overshoots: keep det F > 0
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    h = capi.Handle("NeoHookean", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu,
                    n_first_elements=n_first)
    e_ref, g_ref, v_ref = h.grad_hess(x)
    assert np.isfinite(e_ref)
    H = oracle.problem_from_mesh(mesh, "NeoHookean").assemble_hessian(x)
    assert_values_close(H.outer, H.inner, v_ref, H.values)
    xd = torch.from_numpy(x).cuda()
    e = torch.full((1,), 7.0, dtype=torch.float64, device="cuda")
    g = torch.full((h.ndof,), 7.0, dtype=torch.float64, device="cuda")
    v = torch.full((h.nnz,), 7.0, dtype=torch.float64, device="cuda")
    h.grad_hess_part_raw(xd, e, g, v, 1)
    h.synchronize()
    if n_first == 0:
        assert float(e.item()) == 0.0 and not v.any() and not g.any()
    else:
        # the first part alone is the assembly of the first n_first elements
        sub = oracle.OracleProblem("NeoHookean", mesh.conn[:n_first], mesh.vertices[:n_first], mesh.n_bases,
                                   t["points"], t["weights"], t["grad"], lam=lam, mu=mu)
        assert_vector_close(g.cpu().numpy(), sub.assemble_gradient(x))
    h.grad_hess_part_raw(xd, e, g, v, 2)
    h.synchronize()
    assert abs(float(e.item()) - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(g.cpu().numpy(), g_ref)
    outer, inner = h.pattern()
    assert_values_close(outer, inner, v.cpu().numpy(), v_ref)
    with pytest.raises(capi.PfaError):
        h.grad_hess_part_raw(x, np.zeros(1), np.zeros(h.ndof), np.zeros(h.nnz), 1)  # host pointers


def _mass_handle(mesh, rho, **kw):
    from polyfem_b200 import capi, tables
    t = tables.reference_tables(mesh.p, tables.quadrature_order(mesh.p, is_mass=True))
    return capi.Handle("Mass", mesh.conn, mesh.n_bases, t["weights"], None, vertices=mesh.vertices,
                       ref_vals=t["val"], density=rho, **kw), t


@pytest.mark.parametrize("p,n", [(1, 4), (2, 3), (3, 2), (4, 1)])
def test_mass_matrix(oracle, p, n):
    from helpers import assert_values_close
    mesh, x, t = make_case(n, p, jitter=0.2)
    rng = np.random.default_rng(11)
    rho = 1.0 + rng.random(mesh.n_elements)  # per-element density
    ref = oracle.problem_from_mesh(mesh, "Mass", rho=rho).assemble()
    h, tm = _mass_handle(mesh, rho)
    assert h.size == 3 and h.nnz == ref.inner.size
    outer, inner = h.pattern()
    assert outer.tobytes() == ref.outer.tobytes() and inner.tobytes() == ref.inner.tobytes()
    v = h.linear_stiffness()
    assert_values_close(ref.outer, ref.inner, v, ref.values, what="mass")
    col = np.repeat(np.arange(h.ndof), np.diff(outer))
    off_diag = (inner % 3) != (col % 3)
    assert off_diag.sum() == 6 * (h.nnz // 9) and not v[off_diag].any()  # stored zeros off the block diagonal


def test_inertia_form_on_device(oracle):
    import torch
    from helpers import assert_vector_close
    mesh, x, t = make_case(3, 2, jitter=0.1)
    h, tm = _mass_handle(mesh, 1000.0)
    ref = oracle.problem_from_mesh(mesh, "Mass", rho=1000.0).assemble()
    Mv = h.linear_stiffness()
    rng = np.random.default_rng(12)
    xt = x + 1e-3 * rng.standard_normal(x.size)
    e_ref, g_ref = oracle.inertia(ref, x, xt)
    e, g = h.inertia(Mv, x, xt)
    assert abs(e - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(g, g_ref)
    assert_vector_close(h.symv(Mv, x), ref.to_scipy() @ x, what="symv")
    # device-resident: Newton matrix of an implicit-Euler step = dt^2 * elastic + mass, same pattern
    he = gpu_handle(mesh, "NeoHookean", t)
    assert he.nnz == h.nnz
    dt = 1e-3
    _, _, Hv = he.grad_hess(x)
    Hd = torch.from_numpy(dt * dt * Hv).cuda()
    Md = torch.from_numpy(Mv).cuda()
    h.axpy(1.0, Md, Hd)
    h.synchronize()
    assert np.array_equal(Hd.cpu().numpy(), dt * dt * Hv + Mv)
    ed = torch.zeros(1, dtype=torch.float64, device="cuda")
    gd = torch.zeros(h.ndof, dtype=torch.float64, device="cuda")
    h.inertia_raw(Md, torch.from_numpy(x).cuda(), torch.from_numpy(xt).cuda(), ed, gd)
    h.synchronize()
    assert abs(float(ed.item()) - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(gd.cpu().numpy(), g_ref)
