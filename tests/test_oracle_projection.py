"""Oracle restatement of BCLagrangianForm::project_gradient / project_hessian
(solver/forms/lagrangian/BCLagrangianForm.cpp:149-155, 167-213) against an independent dense
formulation. The reference has no unit test for these two functions (grep of tests/*.cpp), so the
pin is the defining property: the result equals the dense matrix with constrained rows / columns
deleted, stored entries (explicit zeros included) are kept and stay ascending per column."""
import numpy as np

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
