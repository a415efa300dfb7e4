"""The opt-in column-lane (owner-computes) kernels on the GPU (PFA_FLAG_COLUMN_LANE, polyfem_b200/csrc/pfa_collane.cu) against
the oracle: same tolerances as the row-lane path (tests/helpers.py), plus what only this path promises - two calls agree
bit for bit. The data flow was validated on the CPU first (tests/test_collane_emulation.py); this file was written after
the round-1 GPU budget was spent, so its first run is the driver's round-end suite. It sorts last on purpose: the default
path does not depend on it."""
import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close, gpu_handle, make_case

# The kernels under test have never run on a GPU (see the module docstring); until they have, a failure here is reported
# as "xfailed" and a success as "xpassed", so that the suite of the measured default path stays readable. Remove the mark
# once the first GPU run is in.
pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(reason="opt-in column-lane kernels: first GPU run (written after the round-1 GPU budget was spent)", strict=False)]


@pytest.mark.parametrize("p,n,jitter", [(1, 3, 0.2), (2, 2, 0.2), (2, 4, 0.0), (1, 6, 0.0), (2, 8, 0.1)])
def test_column_lane_equals_oracle_and_is_reproducible(oracle, p, n, jitter):
    from polyfem_b200 import capi
    mesh, x, t = make_case(n, p, jitter=jitter)
    ref = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=4)
    h = gpu_handle(mesh, "NeoHookean", t, flags=capi.FLAG_COLUMN_LANE)
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
    # and the default (row-lane) handle agrees to rounding
    h0 = gpu_handle(mesh, "NeoHookean", t)
    e0, g0, v0 = h0.grad_hess(x)
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
    h = gpu_handle(mesh, "NeoHookean", t, flags=capi.FLAG_COLUMN_LANE)
    e, g, v = h.grad_hess(xi)
    assert np.isnan(e) and np.isnan(ref.assemble_energy(xi))
    assert_vector_close(g, ref.assemble_gradient(xi))
    H = ref.assemble_hessian(xi)
    assert_values_close(H.outer, H.inner, v, H.values)
