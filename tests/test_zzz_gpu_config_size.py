"""GPU against the oracle at the size of the BASELINE configurations (driver-run): cfg 1 (LinearElasticity P1, n = 20,
48 000 tets) and cfg 2 (NeoHookean P1, n = 44, 511 104 tets) whole, cfg 3 (NeoHookean P2) at n = 30 (162 000 tets, 56 M nnz -
the whole n = 69 mesh would need the oracle's 7 GB slot map plus one 5.5 GB value buffer per thread; its properties are
checked in tests/test_zzz_gpu_fullsize.py). Pattern byte-identical, energy / gradient / values within 1e-12 (helpers.py).
The oracle runs on all host threads; a few seconds to about a minute per case."""
import os

import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close, gpu_handle, make_case

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 4


def test_cfg1_linear_elasticity_p1_n20(oracle):
    mesh, x, t = make_case(20, 1)
    ref = oracle.problem_from_mesh(mesh, "LinearElasticity", n_threads=THREADS)
    K = ref.assemble()
    h = gpu_handle(mesh, "LinearElasticity", t)
    outer, inner = h.pattern()
    assert outer.tobytes() == K.outer.tobytes() and inner.tobytes() == K.inner.tobytes()
    assert_values_close(K.outer, K.inner, h.linear_stiffness(), K.values, what="stiffness")


@pytest.mark.parametrize("p,n", [(1, 44), (2, 30)], ids=["cfg2_neohookean_p1_n44", "cfg3_neohookean_p2_n30"])
def test_neohookean_at_configuration_size(oracle, p, n):
    mesh, x, t = make_case(n, p)
    x = x[: mesh.n_bases * 3]
    ref = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=THREADS)
    H = ref.assemble_hessian(x)
    g_ref, e_ref = ref.assemble_gradient(x), ref.assemble_energy(x)
    h = gpu_handle(mesh, "NeoHookean", t)
    outer, inner = h.pattern()
    assert outer.tobytes() == H.outer.tobytes() and inner.tobytes() == H.inner.tobytes()
    e, g, v = h.grad_hess(x)
    assert abs(e - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(g, g_ref)
    assert_values_close(H.outer, H.inner, v, H.values)
    # the round-1 reduction kernels at the same size
    from polyfem_b200 import capi
    h0 = gpu_handle(mesh, "NeoHookean", t, flags=capi.FLAG_ROW_LANE)
    e0, g0, v0 = h0.grad_hess(x)
    assert abs(e0 - e_ref) <= REL_TOL * abs(e_ref)
    assert_vector_close(g0, g_ref)
    assert_values_close(H.outer, H.inner, v0, H.values)
