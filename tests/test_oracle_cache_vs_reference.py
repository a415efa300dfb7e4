"""The oracle's SparseMatrixCache restatement against the REFERENCE'S OWN class: utils/MatrixCache.cpp is compiled
unmodified from /root/reference (oracle/refmath/cache_glue.cpp, shadow headers for Eigen / Types / Logger /
MaybeParallelFor -> oracle/_ref/libcacheref.so) and driven with the same call sequences as the oracle's cache:
first-call triplets with prunes in between, slot-map (second cache) reuse, thread copies and their merge.
Values are small integers, so every sum is exact and results must be identical bit for bit.

The live comparison runs where oracle/_ref/libcacheref.so exists (the build container; the file also travels to the
GPU box); `tools/make_golden.py` stored the reference's results of the same sequences in
tests/golden/cache_sequences.npz, which the oracle is checked against everywhere."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libcacheref.so")
needs_lib = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libcacheref.so not built (no reference tree)")


class RefCache:
    """ctypes view of the reference's SparseMatrixCache with the interface of pyoracle.Cache."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = ctypes.CDLL(LIB)
            vp = ctypes.c_void_p
            for name in ("refcache_new", "refcache_copy", "refcache_copy_ctor"):
                getattr(L, name).restype = vp
            L.refcache_new.argtypes = [ctypes.c_int]
            L.refcache_copy.argtypes = [vp]
            L.refcache_copy_ctor.argtypes = [vp]
            L.refcache_free.argtypes = [vp]
            L.refcache_add_value.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double]
            L.refcache_prune.argtypes = [vp]
            L.refcache_set_zero.argtypes = [vp]
            L.refcache_add.argtypes = [vp, vp]
            L.refcache_get_matrix.argtypes = [vp]
            L.refcache_get_matrix.restype = ctypes.c_long
            L.refcache_outer.restype = ctypes.POINTER(ctypes.c_int)
            L.refcache_inner.restype = ctypes.POINTER(ctypes.c_int)
            L.refcache_values.restype = ctypes.POINTER(ctypes.c_double)
            for name in ("refcache_outer", "refcache_inner", "refcache_values"):
                getattr(L, name).argtypes = [vp]
            cls._lib = L
        return cls._lib

    def __init__(self, size=None, _h=None):
        self.size = size
        self._h = _h if _h is not None else self.lib().refcache_new(int(size))

    def copy(self, ctor=False):
        c = RefCache(_h=(self.lib().refcache_copy_ctor if ctor else self.lib().refcache_copy)(self._h))
        c.size = self.size
        return c

    def add_value(self, e, i, j, v):
        self.lib().refcache_add_value(self._h, e, i, j, float(v))

    def prune(self):
        self.lib().refcache_prune(self._h)

    def set_zero(self):
        self.lib().refcache_set_zero(self._h)

    def add(self, other):
        self.lib().refcache_add(self._h, other._h)

    def get_matrix(self):
        L = self.lib()
        nnz = int(L.refcache_get_matrix(self._h))
        outer = np.ctypeslib.as_array(L.refcache_outer(self._h), shape=(self.size + 1,)).astype(np.int32)
        inner = np.ctypeslib.as_array(L.refcache_inner(self._h), shape=(max(nnz, 1),))[:nnz].astype(np.int32)
        vals = np.ctypeslib.as_array(L.refcache_values(self._h), shape=(max(nnz, 1),))[:nnz].copy()
        return outer, inner, vals


def oracle_matrix(c):
    m = c.get_matrix()
    return m.outer.astype(np.int32), m.inner.astype(np.int32), m.values.copy()


def same(a, b):
    assert np.array_equal(a[0], b[0]), "outer differs"
    assert np.array_equal(a[1], b[1]), "inner differs"
    assert np.array_equal(a[2], b[2]), "values differ"


def element_entries(rng, n_el, size, n_loc):
    """Per element: the (row, col) sequence a FEM element scatters (all pairs of its local dofs), fixed per element."""
    out = []
    for _ in range(n_el):
        dofs = rng.choice(size, n_loc, replace=False)
        out.append([(int(i), int(j)) for i in dofs for j in dofs])
    return out


def golden_sequences(make_cache, matrix_of):
    """The call sequences of the two tests below on ONE implementation; returns every matrix read on the way."""
    out = []
    rng = np.random.default_rng(1)
    size, n_el = 40, 25
    ents = element_entries(rng, n_el, size, 5)
    c = make_cache(size)
    for call in range(3):
        for e, pairs in enumerate(ents):
            for (i, j) in pairs:
                c.add_value(e, i, j, float(rng.integers(-3, 4)))
            if call == 0 and e % 7 == 3:
                c.prune()
        out.append(matrix_of(c))
    rng = np.random.default_rng(2)
    size, n_el, n_threads = 36, 18, 3
    ents = element_entries(rng, n_el, size, 4)
    c = make_cache(size)
    for call in range(3):
        c.set_zero()
        copies = [c.copy() for _ in range(n_threads)]
        for e, pairs in enumerate(ents):
            for (i, j) in pairs:
                copies[e * n_threads // n_el].add_value(e, i, j, float(rng.integers(-4, 5)))
        for cc in copies:
            cc.prune()
            c.add(cc)
        out.append(matrix_of(c))
    return out


def test_oracle_cache_equals_golden_reference_results(oracle):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "cache_sequences.npz"))
    got = golden_sequences(oracle.Cache, oracle_matrix)
    assert int(gold["n"]) == len(got)
    for k, m in enumerate(got):
        same((gold[f"outer_{k}"], gold[f"inner_{k}"], gold[f"values_{k}"]), m)


@needs_lib
def test_first_call_prune_and_slot_map_reuse(oracle):
    rng = np.random.default_rng(1)
    size, n_el = 40, 25
    ents = element_entries(rng, n_el, size, 5)
    ref, ora = RefCache(size), oracle.Cache(size)
    for e, pairs in enumerate(ents):  # first assembly: triplets, pruned now and then, explicit zeros included
        for (i, j) in pairs:
            v = float(rng.integers(-3, 4))
            ref.add_value(e, i, j, v)
            ora.add_value(e, i, j, v)
        if e % 7 == 3:
            ref.prune()
            ora.prune()
    same(ref.get_matrix(), oracle_matrix(ora))
    for _ in range(3):  # later assemblies go through the slot map, in the same per-element call order
        for e, pairs in enumerate(ents):
            for (i, j) in pairs:
                v = float(rng.integers(-5, 6))
                ref.add_value(e, i, j, v)
                ora.add_value(e, i, j, v)
        same(ref.get_matrix(), oracle_matrix(ora))


@needs_lib
@pytest.mark.parametrize("ctor", [False, True])
def test_thread_copies_and_merge(oracle, ctor):
    """NLAssembler::assemble_hessian's storage model (Assembler.cpp:669-767): per-thread copies of the caller's cache
    (made exactly as LocalThreadMatStorage does: copy(), init(main), copy(); or with the copy constructor), each
    assembling a block of elements, merged with operator+= and read with get_matrix; three calls (pattern-building
    call, then slot-map calls)."""
    rng = np.random.default_rng(2)
    size, n_el, n_threads = 36, 18, 3
    ents = element_entries(rng, n_el, size, 4)
    ref, ora = RefCache(size), oracle.Cache(size)
    for call in range(3):
        ref.set_zero()
        ora.set_zero()
        rcopies = [ref.copy(ctor) for _ in range(n_threads)]
        ocopies = [ora.copy() for _ in range(n_threads)]
        for e, pairs in enumerate(ents):
            t = e * n_threads // n_el
            for (i, j) in pairs:
                v = float(rng.integers(-4, 5))
                rcopies[t].add_value(e, i, j, v)
                ocopies[t].add_value(e, i, j, v)
        for rc, oc in zip(rcopies, ocopies):
            rc.prune()
            oc.prune()
            ref.add(rc)
            ora.add(oc)
        same(ref.get_matrix(), oracle_matrix(ora))


@needs_lib
def test_reference_known_answer_sequence(oracle):
    """tests/test_matrix.cpp:202-249 ("cache") on both implementations: 1, 2, 3+3, 4 at (0,0), (0,1), (9,4), (9,9)."""
    for make in (lambda: RefCache(10), lambda: oracle.Cache(10)):
        c = make()
        c.add_value(0, 0, 0, 1)
        c.add_value(0, 0, 1, 2)
        c.add_value(1, 9, 4, 3)
        c.prune()
        c.add_value(1, 9, 4, 3)
        c.add_value(2, 9, 9, 4)
        m = c.get_matrix() if isinstance(c, RefCache) else oracle_matrix(c)
        dense = np.zeros((10, 10))
        col = np.repeat(np.arange(10), np.diff(m[0]))
        dense[m[1], col] = m[2]
        exp = np.zeros((10, 10))
        exp[0, 0], exp[0, 1], exp[9, 4], exp[9, 9] = 1, 2, 6, 4
        assert np.array_equal(dense, exp) and m[2].size == 4
