"""The owner-computes (column-lane) kernels on the GPU (polyfem_b200/csrc/pfa_collane2.cu, the default NeoHookean P1/P2 path
since round 2) against the oracle: same tolerances as everywhere (tests/helpers.py), plus what only this path promises - two
calls agree bit for bit - and agreement with the round-1 row-lane reduction kernels (PFA_FLAG_ROW_LANE). The data flow is
validated on the CPU as well (tests/test_collane2_emulation.py)."""
import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close, gpu_handle, make_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p,n,jitter", [(1, 3, 0.2), (2, 2, 0.2), (2, 4, 0.0), (1, 6, 0.0), (2, 8, 0.1), (2, 1, 0.0), (1, 1, 0.0), (2, 13, 0.1), (1, 21, 0.1)])
def test_column_lane_equals_oracle_and_is_reproducible(oracle, p, n, jitter):
    from polyfem_b200 import capi
    mesh, x, t = make_case(n, p, jitter=jitter)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=4)
    h = gpu_handle(mesh, "NeoHookean", t)
    h.profile_enable(True)
    e, g, v = h.grad_hess(x)
    names = [k for (k, ms) in h.profile_read()]
    assert any("column_lane" in k for k in names), f"the column-lane kernels did not run: {names}"
    e_ref, g_ref, H = ref.assemble_energy(x), ref.assemble_gradient(x), ref.assemble_hessian(x)
    assert abs(e - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(g, g_ref)
    assert_values_close(H.outer, H.inner, v, H.values)
    # Hessian-only entry and a second fused call: energy, gradient and values bit for bit the same (fixed summation order)
    v2 = h.hessian(x)
    e3, g3, v3 = h.grad_hess(x)
    assert np.array_equal(v2, v) and np.array_equal(v3, v) and np.array_equal(g3, g) and e3 == e
    # and the row-lane reduction kernels agree to rounding
    h0 = gpu_handle(mesh, "NeoHookean", t, flags=capi.FLAG_ROW_LANE)
    h0.profile_enable(True)
    e0, g0, v0 = h0.grad_hess(x)
    assert not any("column_lane" in k for (k, ms) in h0.profile_read())
    assert_vector_close(g, g0)
    assert_values_close(H.outer, H.inner, v, v0)
    # paths the column lanes do not cover fall back to the row-lane kernels on the same handle
    assert_vector_close(h.gradient(x), g_ref)
    assert abs(h.energy(x) - e_ref) <= REL_TOL * abs(e_ref)


def test_column_lane_nan_propagation(oracle):
    from polyfem_b200 import capi
    mesh, x, t = make_case(3, 2)
    xi = x.copy()
    nodes = mesh.conn[11]
    xi.reshape(-1, 3)[nodes[1]] += 3.0 * (mesh.node_xyz[nodes[0]] - mesh.node_xyz[nodes[1]])
    ref = oracle.problem_from_mesh(mesh, "NeoHookean")
    h = gpu_handle(mesh, "NeoHookean", t)
    e, g, v = h.grad_hess(xi)
    assert np.isnan(e) and np.isnan(ref.assemble_energy(xi))
    assert_vector_close(g, ref.assemble_gradient(xi))
    H = ref.assemble_hessian(xi)
    assert_values_close(H.outer, H.inner, v, H.values)


@pytest.mark.parametrize("p,n_points", [(1, 400), (2, 250)])
def test_unstructured_mesh_on_the_gpu(oracle, p, n_points):
    """Delaunay mesh, elements in random order (tests/unstructured.py): irregular valences, both strip classes, idle half-steps."""
    from polyfem_b200 import tables
    from unstructured import delaunay_mesh
    mesh = delaunay_mesh(n_points, p)
    t = tables.reference_tables(p)
    x = 0.02 * np.random.default_rng(1).uniform(-1, 1, mesh.n_bases * 3) * mesh.h
    ref = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=4)
    h = gpu_handle(mesh, "NeoHookean", t)
    h.profile_enable(True)
    e, g, v = h.grad_hess(x)
    assert any("column_lane" in k for (k, ms) in h.profile_read())
    H = ref.assemble_hessian(x)
    outer, inner = h.pattern()
    assert outer.tobytes() == H.outer.tobytes() and inner.tobytes() == H.inner.tobytes()
    assert abs(e - ref.assemble_energy(x)) <= REL_TOL * abs(e)
    assert_vector_close(g, ref.assemble_gradient(x))
    assert_values_close(H.outer, H.inner, v, H.values)
    e2, g2, v2 = h.grad_hess(x)
    assert e2 == e and np.array_equal(g2, g) and np.array_equal(v2, v)
