"""GPU parity tests proper: CUDA path (through the C ABI) vs the CPU oracle on the same
seeded inputs. Bars: pattern bit-exact; energy / gradient / Hessian values
within 1e-12 relative (row-scale guarded); NaN <=> NaN."""
import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close, gpu_handle, make_case

pytestmark = pytest.mark.gpu

CASES = [(1, 4), (2, 3), (3, 2), (4, 1)]


@pytest.mark.parametrize("p,n", CASES)
def test_pattern_bit_exact(oracle, p, n):
    mesh, x, t = make_case(n, p)
    ref = oracle.problem_from_mesh(mesh, "LinearElasticity").assemble()
    h = gpu_handle(mesh, "LinearElasticity", t)
    outer, inner = h.pattern()
    assert outer.dtype == np.int32 and inner.dtype == np.int32
    assert outer.tobytes() == ref.outer.tobytes()
    assert inner.tobytes() == ref.inner.tobytes()
    lap = oracle.problem_from_mesh(mesh, "Laplacian").assemble()
    hl = gpu_handle(mesh, "Laplacian", t)
    o2, i2 = hl.pattern()
    assert o2.tobytes() == lap.outer.tobytes() and i2.tobytes() == lap.inner.tobytes()


@pytest.mark.parametrize("p,n", CASES[:3])
@pytest.mark.parametrize("jitter", [0.0, 0.2])
def test_neohookean_energy_gradient_hessian(oracle, p, n, jitter):
    # P3 bases overshoot between nodes: keep the random field small enough that det F > 0
    mesh, x, t = make_case(n, p, jitter=jitter, scale=0.05 if p < 3 else 0.01)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=2)
    h = gpu_handle(mesh, "NeoHookean", t)
    e_ref, g_ref, H_ref = ref.assemble_energy(x), ref.assemble_gradient(x), ref.assemble_hessian(x)
    assert np.isfinite(e_ref)
    assert abs(h.energy(x) - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(h.gradient(x), g_ref)
    assert_values_close(H_ref.outer, H_ref.inner, h.hessian(x), H_ref.values)
    epe = h.energy_per_element(x)
    epe_ref = ref.assemble_energy_per_element(x)
    assert np.abs(epe - epe_ref).max() <= REL_TOL * np.abs(epe_ref).max()
    # fused entry
    e, g, v = h.grad_hess(x)
    assert abs(e - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(g, g_ref)
    assert_values_close(H_ref.outer, H_ref.inner, v, H_ref.values)
    # steady-state (slot-map) call of the oracle with another displacement, same handle
    x2 = 0.3 * x
    H2 = ref.assemble_hessian(x2)
    assert_values_close(H2.outer, H2.inner, h.hessian(x2), H2.values)


def test_neohookean_nan_propagation(oracle):
    mesh, x, t = make_case(3, 2)
    xi = x.copy()
    nodes = mesh.conn[11]
    xi.reshape(-1, 3)[nodes[1]] += 3.0 * (mesh.node_xyz[nodes[0]] - mesh.node_xyz[nodes[1]])
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    h = gpu_handle(mesh, "NeoHookean", t)
    e_ref = ref.assemble_energy(xi)
    assert np.isnan(e_ref) and np.isnan(h.energy(xi))
    g_ref, g = ref.assemble_gradient(xi), h.gradient(xi)
    assert np.array_equal(np.isnan(g_ref), np.isnan(g)) and np.isnan(g).any()
    assert_vector_close(g, g_ref)
    H_ref = ref.assemble_hessian(xi)
    assert_values_close(H_ref.outer, H_ref.inner, h.hessian(xi), H_ref.values)
    assert np.array_equal(np.isnan(ref.assemble_energy_per_element(xi)), np.isnan(h.energy_per_element(xi)))


@pytest.mark.parametrize("p,n", CASES)
def test_linear_elasticity_stiffness_and_nl_path(oracle, p, n):
    mesh, x, t = make_case(n, p, jitter=0.2)
    ref = oracle.problem_from_mesh(mesh, "LinearElasticity", n_threads=2)
    h = gpu_handle(mesh, "LinearElasticity", t)
    K = ref.assemble()
    assert_values_close(K.outer, K.inner, h.linear_stiffness(), K.values, what="stiffness")
    if p <= 2:  # the oracle's autodiff NL path is O(N^2) per operation
        e_ref, g_ref = ref.assemble_energy(x), ref.assemble_gradient(x)
        assert abs(h.energy(x) - e_ref) <= REL_TOL * abs(e_ref)
        assert_vector_close(h.gradient(x), g_ref)
        H = ref.assemble_hessian(x)
        assert_values_close(H.outer, H.inner, h.hessian(x), H.values)


@pytest.mark.parametrize("p,n", [(1, 3), (2, 2), (3, 1)])
def test_linear_elasticity_quadrature_kernels(oracle, p, n):
    """The stiffness kernels that keep the quadrature loop (generic kernel): reached with per-qp
    geometry input or per-qp materials; the affine / per-element case goes through the
    reference-moment kernel and is covered by the test above. All must agree with the oracle."""
    from polyfem_b200 import capi, mesh as M
    mesh, x, t = make_case(n, p, jitter=0.2)
    rng = np.random.default_rng(21)
    lam = rng.uniform(4e4, 8e4, mesh.n_elements)
    mu = rng.uniform(2e4, 5e4, mesh.n_elements)
    K = oracle.OracleProblem("LinearElasticity", mesh.conn, mesh.vertices, mesh.n_bases, t["points"], t["weights"], t["grad"],
                             lam=lam, mu=mu).assemble()
    nq = t["weights"].size
    e = mesh.vertices[:, 1:, :] - mesh.vertices[:, :1, :]
    jit = np.linalg.inv(e).transpose(0, 2, 1)
    jac_it = np.repeat(jit[:, None, :, :], nq, axis=1).reshape(mesh.n_elements, nq, 9)
    da = np.linalg.det(e)[:, None] * t["weights"][None, :]
    h_geo = capi.Handle("LinearElasticity", mesh.conn, mesh.n_bases, t["weights"], t["grad"], jac_it=jac_it, da=da, lam=lam, mu=mu)
    assert_values_close(K.outer, K.inner, h_geo.linear_stiffness(), K.values, what="stiffness, per-qp geometry")
    h_mat = capi.Handle("LinearElasticity", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices,
                        lam=np.repeat(lam, nq), mu=np.repeat(mu, nq))
    assert_values_close(K.outer, K.inner, h_mat.linear_stiffness(), K.values, what="stiffness, per-qp materials")
    h_aff = capi.Handle("LinearElasticity", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu)
    assert_values_close(K.outer, K.inner, h_aff.linear_stiffness(), K.values, what="stiffness, reference moments")


@pytest.mark.parametrize("p,n", CASES)
def test_laplacian_stiffness(oracle, p, n):
    mesh, x, t = make_case(n, p, jitter=0.2)
    ref = oracle.problem_from_mesh(mesh, "Laplacian")
    h = gpu_handle(mesh, "Laplacian", t)
    assert h.size == 1 and h.ndof == mesh.n_bases
    K = ref.assemble()
    assert_values_close(K.outer, K.inner, h.linear_stiffness(), K.values, what="laplacian")


def test_general_geometry_input_equals_affine(oracle):
    """jac_it / da given per quadrature point (non-affine input form) reproduce the affine path."""
    from polyfem_b200 import capi, mesh as M
    mesh, x, t = make_case(3, 2, jitter=0.2)
    e = mesh.vertices[:, 1:, :] - mesh.vertices[:, :1, :]
    jit = np.linalg.inv(e).transpose(0, 2, 1)  # (J^-1)^T, J rows = edges
    det = np.linalg.det(e)
    nq = t["weights"].size
    jac_it = np.repeat(jit[:, None, :, :], nq, axis=1).reshape(mesh.n_elements, nq, 9)
    da = det[:, None] * t["weights"][None, :]
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    h = capi.Handle("NeoHookean", mesh.conn, mesh.n_bases, t["weights"], t["grad"], jac_it=jac_it, da=da, lam=lam, mu=mu)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    H = ref.assemble_hessian(x)
    e1, g1, v1 = h.grad_hess(x)
    assert abs(e1 - ref.assemble_energy(x)) <= REL_TOL * abs(e1)
    assert_vector_close(g1, ref.assemble_gradient(x))
    assert_values_close(H.outer, H.inner, v1, H.values)


@pytest.mark.parametrize("p,n", [(3, 2), (4, 1)])
def test_laplacian_tile_kernel_per_qp_geometry(oracle, p, n):
    """The register-tiled Laplacian kernel with jac_it / da given per quadrature point."""
    from polyfem_b200 import capi
    mesh, x, t = make_case(n, p, jitter=0.2)
    e = mesh.vertices[:, 1:, :] - mesh.vertices[:, :1, :]
    jit = np.linalg.inv(e).transpose(0, 2, 1)
    det = np.linalg.det(e)
    nq = t["weights"].size
    jac_it = np.repeat(jit[:, None, :, :], nq, axis=1).reshape(mesh.n_elements, nq, 9)
    da = det[:, None] * t["weights"][None, :]
    h = capi.Handle("Laplacian", mesh.conn, mesh.n_bases, t["weights"], t["grad"], jac_it=jac_it, da=da)
    K = oracle.problem_from_mesh(mesh, "Laplacian").assemble()
    assert_values_close(K.outer, K.inner, h.linear_stiffness(), K.values, what="laplacian per-qp geometry")


def test_per_element_and_per_qp_materials(oracle):
    from polyfem_b200 import capi
    mesh, x, t = make_case(2, 2)
    rng = np.random.default_rng(9)
    lam = rng.uniform(4e4, 8e4, mesh.n_elements)
    mu = rng.uniform(2e4, 5e4, mesh.n_elements)
    ref = oracle.OracleProblem("NeoHookean", mesh.conn, mesh.vertices, mesh.n_bases, t["points"], t["weights"], t["grad"], lam=lam, mu=mu)
    H = ref.assemble_hessian(x)
    h = capi.Handle("NeoHookean", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu)
    assert_values_close(H.outer, H.inner, h.hessian(x), H.values)
    nq = t["weights"].size
    h2 = capi.Handle("NeoHookean", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices,
                     lam=np.repeat(lam, nq), mu=np.repeat(mu, nq))
    assert_values_close(H.outer, H.inner, h2.hessian(x), H.values)
    # set_materials (t changed): back to constants
    h.set_materials(np.full(mesh.n_elements, lam[0]), np.full(mesh.n_elements, mu[0]))
    ref2 = oracle.OracleProblem("NeoHookean", mesh.conn, mesh.vertices, mesh.n_bases, t["points"], t["weights"], t["grad"], lam=lam[0], mu=mu[0])
    H2 = ref2.assemble_hessian(x)
    assert_values_close(H2.outer, H2.inner, h.hessian(x), H2.values)


def test_device_pointers_and_profile(oracle):
    import torch
    mesh, x, t = make_case(3, 2)
    h = gpu_handle(mesh, "NeoHookean", t)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    xd = torch.from_numpy(x).cuda()
    e = torch.zeros(1, dtype=torch.float64, device="cuda")
    g = torch.zeros(h.ndof, dtype=torch.float64, device="cuda")
    v = torch.zeros(h.nnz, dtype=torch.float64, device="cuda")
    h.profile_enable(True)
    h.grad_hess_raw(xd, e, g, v)
    h.synchronize()
    recs = h.profile_read()
    assert any("assemble" in r[0] for r in recs) and all(r[1] >= 0 for r in recs)
    H = ref.assemble_hessian(x)
    assert_values_close(H.outer, H.inner, v.cpu().numpy(), H.values)
    assert_vector_close(g.cpu().numpy(), ref.assemble_gradient(x))
    assert h.launch_count() >= 3


def test_interface_mirror_reads_like_reference_test(oracle):
    """tests/test_assembler.cpp:24-84 ("hessian_lin") through the mirrored operator interface."""
    from polyfem_b200 import assembler as A, mesh as M
    mesh = M.kuhn_cube(3, 2, jitter=0.1)
    bases = gbases = A.FESpace.from_mesh(mesh)
    cache = A.AssemblyValsCache(mesh.p)
    asm = A.make_assembler("LinearElasticity")
    asm.set_size(3)
    asm.set_materials([], {"E": 1e5, "nu": 0.3})
    n_basis = mesh.n_bases
    stiffness = asm.assemble(True, n_basis, bases, gbases, cache, 0)
    mat_cache = A.MatrixCache()
    rng = np.random.default_rng(2)
    for _ in range(10):
        disp = rng.uniform(-1, 1, (n_basis * 3, 1))
        hessian = asm.assemble_hessian(True, n_basis, False, bases, gbases, cache, 0, 0, disp, np.zeros(0), mat_cache)
        assert np.array_equal(hessian.indices, stiffness.indices)
        assert abs(hessian - stiffness).max() < 1e-8
    with pytest.raises(RuntimeError):
        A.make_assembler("Ogden")  # an assembler outside the path: refused loudly, as AssemblerUtils::make_assembler does for unknown names


@pytest.mark.parametrize("p,n", [(1, 5), (2, 4)])
@pytest.mark.parametrize("flags", [1, 2, 3])
def test_create_flags_do_not_change_results(oracle, p, n, flags):
    """pfa_mesh_desc.flags (caller element order kept; values[] cleared inside the kernel) change
    the schedule only: same pattern, same values, per-element energies in the caller's order."""
    mesh, x, t = make_case(n, p, jitter=0.15)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=2)
    h = gpu_handle(mesh, "NeoHookean", t, flags=flags)
    H = ref.assemble_hessian(x)
    for rep in range(3):  # repeated calls: the in-kernel zero fill must not leave stale sums
        e, g, v = h.grad_hess(x)
        assert abs(e - ref.assemble_energy(x)) <= REL_TOL * abs(e)
        assert_vector_close(g, ref.assemble_gradient(x))
        assert_values_close(H.outer, H.inner, v, H.values)
    epe = h.energy_per_element(x)
    epe_ref = ref.assemble_energy_per_element(x)
    assert np.abs(epe - epe_ref).max() <= REL_TOL * np.abs(epe_ref).max()


def test_two_quadrature_tables_in_one_process(oracle):
    """The reference gradients of the column side live in one __constant__ slot per basis order;
    handles with different tables for the same order must still give their own results."""
    mesh, x, t = make_case(3, 2)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    h1 = gpu_handle(mesh, "NeoHookean", t)
    # same rule with the quadrature points listed in reverse order: a different table, same integrals
    t2 = {k: (v[::-1].copy() if k in ("points", "weights", "val", "grad") else v) for k, v in t.items()}
    h2 = gpu_handle(mesh, "NeoHookean", t2)
    H = ref.assemble_hessian(x)
    for h in (h1, h2, h1, h2):
        assert_values_close(H.outer, H.inner, h.hessian(x), H.values)


@pytest.mark.parametrize("p,n,scale", [(1, 4, 0.25), (2, 3, 0.12)])
def test_project_to_psd(oracle, p, n, scale):
    """assemble_hessian(project_to_psd=true) (Assembler.cpp:693-694): per-element eigenvalue clamp.
    ipc-toolkit's source is absent (parity unpinned): the GPU Jacobi solver is compared
    with the oracle's restatement; both are non-expansive maps of the same local matrices, so the
    bar is 1e-10 of the row scale instead of 1e-12."""
    mesh, x, t = make_case(n, p, jitter=0.1, scale=scale)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=2)
    h = gpu_handle(mesh, "NeoHookean", t)
    H0 = ref.assemble_hessian(x)
    H1 = ref.assemble_hessian(x, project_to_psd=True)
    assert np.abs(H0.values - H1.values).max() > 1e-3 * np.abs(H0.values).max(), "projection inactive: test is vacuous"
    v = h.hessian(x, project_to_psd=True)
    assert_values_close(H1.outer, H1.inner, v, H1.values, tol=1e-10, what="projected hessian")
    # fused call: energy and gradient are not affected by the projection
    e, g, v2 = h.grad_hess(x, project_to_psd=True)
    assert abs(e - ref.assemble_energy(x)) <= REL_TOL * abs(e)
    assert_vector_close(g, ref.assemble_gradient(x))
    assert_values_close(H1.outer, H1.inner, v2, H1.values, tol=1e-10, what="projected hessian (fused)")
    # a PSD-everywhere state (x = 0) is returned unchanged, to the last bit of the unprojected path
    z = np.zeros_like(x)
    assert np.array_equal(h.hessian(z, project_to_psd=True), h.hessian(z)) or np.abs(h.hessian(z, project_to_psd=True) - ref.assemble_hessian(z, project_to_psd=True).values).max() <= 1e-10 * np.abs(H0.values).max()


@pytest.mark.parametrize("k", range(12))
def test_against_the_reference_own_function_outputs(oracle, k):
    """CUDA path vs tests/golden/nh_local.npz: gradient and Hessian of single-element meshes as returned by
    the reference's own compute_energy_aux_gradient_fast / compute_energy_hessian_aux_fast (compiled from
    /root/reference, tests/test_oracle_reference_math.py). 1e-12 of the largest entry; NaN <=> NaN."""
    import os
    from polyfem_b200 import capi, tables
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nh_local.npz"))
    p = int(gold[f"p_{k}"])
    t = tables.reference_tables(p)
    u = gold[f"u_{k}"]
    nl = u.shape[0]
    conn = np.arange(nl, dtype=np.int32)[None, :]
    h = capi.Handle("NeoHookean", conn, nl, t["weights"], t["grad"], vertices=gold[f"vertices_{k}"][None],
                    lam=float(gold["lambda"]), mu=float(gold["mu"]))
    e, g, v = h.grad_hess(u.reshape(-1))
    outer, inner = h.pattern()
    e_ref = float(gold[f"energy_{k}"])
    assert np.isnan(e) == np.isnan(e_ref) and (np.isnan(e_ref) or abs(e - e_ref) <= REL_TOL * abs(e_ref))
    assert h.nnz == (3 * nl) ** 2  # one element: the matrix is the dense local Hessian
    H = np.zeros((3 * nl, 3 * nl))
    col = np.repeat(np.arange(3 * nl), np.diff(outer))
    H[inner, col] = v
    for got, ref in ((g, gold[f"gradient_{k}"]), (H, gold[f"hessian_{k}"])):
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        ok = ~np.isnan(ref)
        if ok.any():
            assert np.abs(got[ok] - ref[ok]).max() <= REL_TOL * np.abs(ref[ok]).max()
