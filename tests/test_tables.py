"""Reference-element data vs the reference's own generated code.

Golden file tests/golden/ref_tables.npz was produced by tools/make_golden.py from
/root/reference's auto_p_bases.cpp / TetQuadrature.cpp compiled unmodified (oracle/_ref).
Also restates the reference's tests/test_bases.cpp:891-950 (Kronecker delta, partition of
unity) and tests/test_quadrature.cpp:237-246 (weights sum to 1/6, points in [0,1]).
"""
import os

import numpy as np
import pytest

from polyfem_b200 import tables

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_tables.npz"))


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_node_order_matches_reference(p):
    assert np.array_equal(tables.p_nodes(p), GOLD[f"nodes_p{p}"])


@pytest.mark.parametrize("p", [1, 2, 3, 4])
@pytest.mark.parametrize("order", [1, 2, 4, 6, 8])
def test_basis_matches_reference_generated_code(p, order):
    pts, _ = tables.tet_quadrature(order)
    val, grad = tables.p_basis(p, pts)
    # 1e-10 is the reference's own margin for P1/P2 formulas (tests/test_bases.cpp:799-859);
    # the two evaluations differ only by rounding of differently factored polynomials.
    assert np.abs(val - GOLD[f"val_p{p}_q{order}"]).max() < 1e-12
    assert np.abs(grad - GOLD[f"grad_p{p}_q{order}"]).max() < 2e-13 * p ** 3


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_basis_at_extra_points(p):
    val, grad = tables.p_basis(p, GOLD["extra_points"])
    assert np.abs(val - GOLD[f"val_p{p}_extra"]).max() < 1e-12
    assert np.abs(grad - GOLD[f"grad_p{p}_extra"]).max() < 2e-13 * p ** 3


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_kronecker_and_partition_of_unity(p):
    nodes = tables.p_nodes(p)
    val, grad = tables.p_basis(p, nodes)
    assert np.abs(val - np.eye(nodes.shape[0])).max() < 1e-13
    pts, _ = tables.tet_quadrature(6)
    val, grad = tables.p_basis(p, pts)
    assert np.abs(val.sum(axis=1) - 1).max() < 1e-13
    assert np.abs(grad.sum(axis=1)).max() < 1e-12


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_gradient_vs_finite_differences(p):
    pts = GOLD["extra_points"]
    _, grad = tables.p_basis(p, pts)
    h = 1e-6
    for d in range(3):
        e = np.zeros(3)
        e[d] = h
        vp, _ = tables.p_basis(p, pts + e)
        vm, _ = tables.p_basis(p, pts - e)
        assert np.abs((vp - vm) / (2 * h) - grad[:, :, d]).max() < 1e-7


@pytest.mark.parametrize("order", range(1, 9))
def test_quadrature_weights(order):
    pts, w = tables.tet_quadrature(order)
    assert abs(w.sum() - 1.0 / 6.0) < 1e-14
    assert pts.min() >= 0 and pts.max() <= 1
    assert pts.shape == (w.size, 3)


def test_quadrature_point_counts_and_order_rule():
    assert [tables.tet_quadrature(o)[1].size for o in (1, 2, 4, 6)] == [1, 4, 11, 23]
    assert [tables.quadrature_order(p) for p in (1, 2, 3, 4)] == [1, 2, 4, 6]
    assert [tables.quadrature_order(p, True) for p in (1, 2, 3, 4)] == [2, 4, 6, 8]


def test_quadrature_integrates_polynomials_exactly():
    # int_T x^a y^b z^c = a! b! c! / (a+b+c+3)!
    from math import factorial as f
    for order in (1, 2, 4, 6):
        pts, w = tables.tet_quadrature(order)
        for a in range(order + 1):
            for b in range(order + 1 - a):
                for c in range(order + 1 - a - b):
                    exact = f(a) * f(b) * f(c) / f(a + b + c + 3)
                    got = (w * pts[:, 0] ** a * pts[:, 1] ** b * pts[:, 2] ** c).sum()
                    assert abs(got - exact) < 1e-14


def test_live_reference_lib_if_present():
    """When oracle/_ref/libpfref.so exists (it travels to the GPU box), the committed data
    must still equal what the reference's sources produce."""
    import ctypes
    path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref", "libpfref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built")
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.pfref_tet_quadrature.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int]
    for order in range(1, 9):
        pts = np.zeros((512, 3))
        w = np.zeros(512)
        n = lib.pfref_tet_quadrature(order, pts.ctypes.data_as(dp), w.ctypes.data_as(dp), 512)
        mp, mw = tables.tet_quadrature(order)
        assert n == mw.size
        assert np.array_equal(pts[:n], mp) and np.array_equal(w[:n], mw)


def test_quadrature_order_and_lame_conversion_against_reference_functions():
    """AssemblerUtils::quadrature_order (AssemblerUtils.cpp:201-245) and convert_to_lambda / convert_to_mu
    (utils/ElasticityUtils.cpp:12-23), compiled verbatim from /root/reference (oracle/_ref/libmiscref.so): the values
    below were produced by those functions (tools/make_golden.py prints them) and are checked live where the library
    exists."""
    import ctypes
    import os
    from polyfem_b200 import mesh as M, tables
    golden_orders = {("Mass", 1): 2, ("Mass", 2): 4, ("Mass", 3): 6, ("Mass", 4): 8,
                     ("NeoHookean", 1): 1, ("NeoHookean", 2): 2, ("NeoHookean", 3): 4, ("NeoHookean", 4): 6,
                     ("Laplacian", 1): 1, ("Laplacian", 4): 6, ("LinearElasticity", 2): 2}
    for (name, p), order in golden_orders.items():
        assert tables.quadrature_order(p, is_mass=(name == "Mass")) == order
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    assert (lam, mu) == (57692.30769230769, 38461.53846153846)  # bit for bit what the reference functions return
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libmiscref.so")
    if os.path.exists(path):
        L = ctypes.CDLL(path)
        L.ref_quadrature_order.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.ref_convert_to_lambda.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double]
        L.ref_convert_to_lambda.restype = ctypes.c_double
        L.ref_convert_to_mu.argtypes = [ctypes.c_double, ctypes.c_double]
        L.ref_convert_to_mu.restype = ctypes.c_double
        for (name, p), order in golden_orders.items():
            assert L.ref_quadrature_order(name.encode(), p, 0, 3) == order
        rng = np.random.default_rng(5)
        for _ in range(50):
            E, nu = float(rng.uniform(1, 1e7)), float(rng.uniform(-0.9, 0.49))
            lam, mu = M.lame_from_E_nu(E, nu)
            assert lam == L.ref_convert_to_lambda(1, E, nu) and mu == L.ref_convert_to_mu(E, nu)
