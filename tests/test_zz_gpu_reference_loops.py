"""CUDA path vs tests/golden/nl_loops.npz: energy, gradient and CSC Hessian of small multi-element meshes as returned by
the reference's OWN global loops NLAssembler::assemble_energy / assemble_gradient / assemble_hessian (Assembler.cpp:495-771,
compiled from /root/reference over its own NeoHookean local functions and its unmodified MatrixCache.cpp; see
tests/test_oracle_loops_vs_reference.py for how the golden is made). Pattern bit-exact; values within 1e-12 (north_star's
tolerance, tests/helpers.py); NaN <=> NaN on the mesh with inverted elements."""
import os

import numpy as np
import pytest

from helpers import REL_TOL, assert_values_close, assert_vector_close, gpu_handle
from test_oracle_loops_vs_reference import GOLD_PATH, loop_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", range(4))
def test_cuda_path_equals_reference_global_loops(k):
    G = np.load(GOLD_PATH)
    name, mesh, x = loop_cases()[k]
    assert np.array_equal(G[f"x_{name}"], x), "golden inputs are stale: rerun tools/make_golden.py"
    h = gpu_handle(mesh, "NeoHookean")
    outer, inner = h.pattern()
    assert outer.tobytes() == G[f"outer_{name}"].astype(outer.dtype).tobytes()
    assert inner.tobytes() == G[f"inner_{name}"].astype(inner.dtype).tobytes()
    e_ref, g_ref, v_ref = float(G[f"energy_{name}_t1"]), G[f"gradient_{name}_t1"], G[f"values_{name}_t1"]
    for e, g, v in (h.grad_hess(x), (h.energy(x), h.gradient(x), h.hessian(x))):
        assert np.isnan(e) == np.isnan(e_ref) and (np.isnan(e_ref) or abs(e - e_ref) <= REL_TOL * abs(e_ref))
        assert_vector_close(g, g_ref)
        assert_values_close(outer, inner, v, v_ref)
    if np.isfinite(e_ref):  # assemble_energy_per_element (Assembler.cpp:533-572)
        epe_ref = G[f"energy_per_element_{name}"]
        assert np.abs(h.energy_per_element(x) - epe_ref).max() <= REL_TOL * np.abs(epe_ref).max()


@pytest.mark.parametrize("k", range(6))
def test_cuda_path_equals_reference_linear_loop(k):
    """pfa_linear_stiffness vs tests/golden/linear_loops.npz: the matrix the reference's own LinearAssembler::assemble
    (Assembler.cpp:157-384) builds from its own LinearElasticity / Laplacian / Mass local blocks."""
    from polyfem_b200 import capi, tables
    from test_oracle_loops_vs_reference import LINEAR_GOLD_PATH, RHO, linear_cases
    G = np.load(LINEAR_GOLD_PATH)
    name, material, mesh = linear_cases()[k]
    assert np.array_equal(G[f"vertices_{name}"], mesh.vertices), "golden inputs are stale: rerun tools/make_golden.py"
    if material == "Mass":
        t = tables.reference_tables(mesh.p, tables.quadrature_order(mesh.p, is_mass=True))
        h = capi.Handle("Mass", mesh.conn, mesh.n_bases, t["weights"], None, vertices=mesh.vertices, ref_vals=t["val"], density=RHO)
    else:
        h = gpu_handle(mesh, material)
    outer, inner = h.pattern()
    assert outer.tobytes() == G[f"outer_{name}"].astype(outer.dtype).tobytes()
    assert inner.tobytes() == G[f"inner_{name}"].astype(inner.dtype).tobytes()
    assert_values_close(outer, inner, h.linear_stiffness(), G[f"values_{name}"], what=name)


@pytest.mark.parametrize("k", range(12))
def test_cuda_linear_elasticity_energy_equals_reference_function(k):
    """pfa_energy of a LinearElasticity handle (a linear material inside a nonlinear solve) vs tests/golden/le_energy.npz:
    LinearElasticity::compute_energy of the reference (LinearElasticity.cpp:65-68, 103-132 compiled verbatim) on the
    one-element meshes of nh_local.npz, P1-P4."""
    from polyfem_b200 import capi, tables
    here = os.path.dirname(os.path.abspath(__file__))
    G = np.load(os.path.join(here, "golden", "nh_local.npz"))
    LE = np.load(os.path.join(here, "golden", "le_energy.npz"))
    t = tables.reference_tables(int(G[f"p_{k}"]))
    u = G[f"u_{k}"]
    nl = u.shape[0]
    conn = np.arange(nl, dtype=np.int32)[None, :]
    h = capi.Handle("LinearElasticity", conn, nl, t["weights"], t["grad"], vertices=G[f"vertices_{k}"][None],
                    lam=float(G["lambda"]), mu=float(G["mu"]))
    e_ref = float(LE[f"le_energy_{k}"])
    assert abs(h.energy(u.reshape(-1)) - e_ref) <= REL_TOL * e_ref

