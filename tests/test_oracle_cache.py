"""SparseMatrixCache semantics, pinned by the reference's self-contained known-answer test
tests/test_matrix.cpp:202-249 ("cache") restated verbatim against oracle/oracle.cpp."""
import numpy as np


def _feed(c, prune_mid):
    c.add_value(0, 0, 0, 1)
    c.add_value(0, 0, 1, 2)
    if prune_mid:
        c.prune()
    c.add_value(0, 9, 4, 3)
    c.add_value(0, 9, 4, 3)
    c.add_value(0, 9, 9, 4)


def _check(m):
    d = m.to_scipy()
    assert d[0, 0] == 1
    assert d[0, 1] == 2
    assert d[9, 4] == 6
    assert d[9, 9] == 4
    assert m.nnz == 4


def test_cache_known_answers(oracle):
    cache = oracle.Cache(10)
    _feed(cache, False)
    _check(cache.get_matrix())

    cache1 = oracle.Cache(10)
    _feed(cache1, True)
    m1 = cache1.get_matrix()
    _check(m1)

    cache2 = cache1.copy()  # SparseMatrixCache cache2(cache1): shares cache1's slot map
    _feed(cache2, True)
    m2 = cache2.get_matrix()
    _check(m2)
    assert np.array_equal(m1.outer, m2.outer) and np.array_equal(m1.inner, m2.inner)


def test_csc_layout_is_eigen_like(oracle):
    c = oracle.Cache(4)
    # unsorted insertion, duplicates, explicit zero
    for (i, j, v) in [(3, 1, 1.0), (0, 1, 2.0), (3, 1, 0.5), (2, 0, 0.0), (1, 3, -1.0)]:
        c.add_value(0, i, j, v)
    m = c.get_matrix()
    assert m.outer.tolist() == [0, 1, 3, 3, 4]
    assert m.inner.tolist() == [2, 0, 3, 1]           # ascending inside each column
    assert m.values.tolist() == [0.0, 2.0, 1.5, -1.0]  # explicit zero kept, duplicates summed
    assert m.outer.dtype == np.int32 and m.inner.dtype == np.int32
