"""Oracle restatement of BCLagrangianForm::project_gradient / project_hessian
(solver/forms/lagrangian/BCLagrangianForm.cpp:149-155, 167-213): (1) against an independent dense
formulation (the defining property: constrained rows / columns deleted, stored entries - explicit zeros
included - kept and ascending per column; the reference has no unit test for these functions), and
(2) against the REFERENCE'S OWN function bodies and map-building loops (:121-137), extracted from
/root/reference at build time and compiled verbatim against oracle/refmath's stand-ins
(oracle/_ref/libbcref.so); their results on fixed cases are committed as tests/golden/bc_projection.npz."""
import ctypes
import os

import numpy as np
import pytest

from oracle import pyoracle


def _random_csc(n, density, rng, explicit_zeros=True):
    mask = rng.random((n, n)) < density
    mask |= np.eye(n, dtype=bool)
    vals = rng.standard_normal((n, n))
    if explicit_zeros:
        vals[rng.random((n, n)) < 0.2] = 0.0  # stored zeros must survive
    outer, inner, values = [0], [], []
    for c in range(n):
        rows = np.flatnonzero(mask[:, c])
        inner.extend(rows.tolist())
        values.extend(vals[rows, c].tolist())
        outer.append(len(inner))
    return pyoracle.CSC(n, np.array(outer, np.int32), np.array(inner, np.int32), np.array(values)), mask, vals


def test_project_hessian_matches_dense_row_column_deletion():
    rng = np.random.default_rng(3)
    for n, k in [(1, 0), (7, 3), (23, 9), (40, 0), (40, 39)]:
        csc, mask, vals = _random_csc(n, 0.3, rng)
        constrained = rng.permutation(n)[:k]
        red = pyoracle.project_hessian(csc, constrained)
        keep = np.setdiff1d(np.arange(n), constrained)
        assert red.n == keep.size and red.outer[0] == 0 and red.outer[-1] == red.inner.size
        sub_mask, sub_vals = mask[np.ix_(keep, keep)], vals[np.ix_(keep, keep)]
        assert red.inner.size == sub_mask.sum()  # stored zeros are kept
        for c in range(keep.size):
            rows = red.inner[red.outer[c]:red.outer[c + 1]]
            assert np.all(np.diff(rows) > 0)
            assert np.array_equal(rows, np.flatnonzero(sub_mask[:, c]))
            assert np.array_equal(red.values[red.outer[c]:red.outer[c + 1]], sub_vals[rows, c])


def test_project_gradient_keeps_unconstrained_entries_in_order():
    rng = np.random.default_rng(4)
    g = rng.standard_normal(31)
    constrained = np.array([30, 0, 7, 8, 15])
    out = pyoracle.project_gradient(g, constrained)
    keep = np.setdiff1d(np.arange(31), constrained)
    assert np.array_equal(out, g[keep])
    assert np.array_equal(pyoracle.project_gradient(g, []), g)


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BCLIB = os.path.join(ROOT, "oracle", "_ref", "libbcref.so")


def golden_cases():
    """(csc, constrained, gradient) cases shared by the golden writer (tools/make_golden.py) and the tests."""
    rng = np.random.default_rng(77)
    out = []
    for n, k in [(1, 0), (9, 4), (30, 11), (30, 0), (30, 29), (64, 20)]:
        csc, _, _ = _random_csc(n, 0.25, rng)
        constrained = rng.permutation(n)[:k].astype(np.int32)
        out.append((csc, constrained, rng.standard_normal(n)))
    return out


def reference_projection(csc, constrained, grad):
    """Through oracle/_ref/libbcref.so: (not_constraints, old_to_new, projected gradient, (outer, inner, values))."""
    L = ctypes.CDLL(BCLIB)
    ip, dp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)
    L.ref_bc_maps.argtypes = [ctypes.c_int, ctypes.c_int, ip, ip, ip]
    L.ref_bc_project_gradient.argtypes = [ctypes.c_int, ctypes.c_int, ip, dp, dp]
    L.ref_bc_project_hessian.argtypes = [ctypes.c_int, ctypes.c_int, ip, ctypes.c_long, ip, ip, dp, ip, ip, dp]
    L.ref_bc_project_hessian.restype = ctypes.c_long
    n = csc.n
    c = np.ascontiguousarray(constrained, dtype=np.int32)
    cp = c.ctypes.data_as(ip)
    nc, o2n = np.zeros(n, np.int32), np.zeros(n, np.int32)
    n_red = L.ref_bc_maps(n, c.size, cp, nc.ctypes.data_as(ip), o2n.ctypes.data_as(ip))
    g = np.ascontiguousarray(grad, dtype=np.float64)
    gout = np.zeros(n)
    assert L.ref_bc_project_gradient(n, c.size, cp, g.ctypes.data_as(dp), gout.ctypes.data_as(dp)) == n_red
    outer, inner = np.ascontiguousarray(csc.outer, np.int32), np.ascontiguousarray(csc.inner, np.int32)
    vals = np.ascontiguousarray(csc.values)
    o_r, i_r, v_r = np.zeros(n + 1, np.int32), np.zeros(max(inner.size, 1), np.int32), np.zeros(max(inner.size, 1))
    nnz = L.ref_bc_project_hessian(n, c.size, cp, inner.size, outer.ctypes.data_as(ip), inner.ctypes.data_as(ip), vals.ctypes.data_as(dp),
                                   o_r.ctypes.data_as(ip), i_r.ctypes.data_as(ip), v_r.ctypes.data_as(dp))
    return nc[:n_red].copy(), o2n, gout[:n_red].copy(), (o_r[:n_red + 1].copy(), i_r[:nnz].copy(), v_r[:nnz].copy())


def _check_against(ref_result, csc, constrained, grad):
    nc, o2n, g_ref, (o_r, i_r, v_r) = ref_result
    red = pyoracle.project_hessian(csc, constrained)
    assert red.outer.tobytes() == np.asarray(o_r, np.int32).tobytes()
    assert red.inner.tobytes() == np.asarray(i_r, np.int32).tobytes()
    assert np.array_equal(red.values, v_r)  # a gather: bit identical
    assert np.array_equal(pyoracle.project_gradient(grad, constrained), g_ref)
    keep = np.setdiff1d(np.arange(csc.n), constrained)
    assert np.array_equal(nc, keep)  # not_constraints_ is the sorted complement
    exp = -np.ones(csc.n, np.int64)
    exp[keep] = np.arange(keep.size)
    assert np.array_equal(o2n, exp)


def test_projection_equals_golden_reference_results():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "bc_projection.npz"))
    for k, (csc, constrained, grad) in enumerate(golden_cases()):
        _check_against((gold[f"nc_{k}"], gold[f"o2n_{k}"], gold[f"g_{k}"], (gold[f"outer_{k}"], gold[f"inner_{k}"], gold[f"values_{k}"])),
                       csc, constrained, grad)


@pytest.mark.skipif(not os.path.exists(BCLIB), reason="oracle/_ref/libbcref.so not built (no reference tree)")
def test_projection_equals_reference_functions_live():
    for csc, constrained, grad in golden_cases():
        _check_against(reference_projection(csc, constrained, grad), csc, constrained, grad)
