"""The oracle's GLOBAL assembly (energy / gradient / Hessian over a multi-element mesh) against the REFERENCE'S OWN loops.

`oracle/refmath/loop_glue.cpp` compiles NLAssembler::assemble_energy / assemble_gradient / assemble_hessian and their
per-thread storage classes (assembler/Assembler.cpp:16-94, 495-531, 574-643, 645-771) verbatim from /root/reference, over
the reference's own NeoHookean local functions and its utils/MatrixCache.cpp compiled unmodified, into
oracle/_ref/libloopref.so. `tools/make_golden.py` ran them on the meshes of `loop_cases()` with 1 and 3 thread storages
and committed the results as tests/golden/nl_loops.npz, which is what travels to the GPU box. The linear loop
LinearAssembler::assemble (Assembler.cpp:157-384) is compiled the same way over the reference's own LinearElasticity /
Laplacian / Mass local functions: `linear_cases()` -> tests/golden/linear_loops.npz.

Checked: the CSC pattern (outer, inner) is identical; energy, gradient and values agree to 1e-13 of the largest entry
(the oracle and the reference cut the element range into different per-thread chunks, so sums are ordered differently);
a second assemble_hessian through the same matrix cache (the cached-pattern path every Newton iteration after the first
takes) returns the same matrix on both sides; the NaN pattern of a mesh with an inverted element is the same."""
import ctypes
import os

import numpy as np
import pytest

from polyfem_b200 import mesh as M, tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD_PATH = os.path.join(ROOT, "tests", "golden", "nl_loops.npz")
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libloopref.so")
TOL = 1e-13
THREADS = (1, 3)


def loop_cases():
    """(name, mesh, displacement): jittered Kuhn cubes, P1..P3, and one P1 case whose displacement inverts elements."""
    out = []
    for name, p, n, scale in (("p1", 1, 3, 0.05), ("p2", 2, 2, 0.05), ("p3", 3, 1, 0.01), ("p1_inverted", 1, 2, 0.6)):
        mesh = M.kuhn_cube(n, p, jitter=0.2)
        out.append((name, mesh, M.random_displacement(mesh, scale=scale)[: mesh.n_bases * 3]))
    return out


def _ptr(a, t=ctypes.c_double):
    return a.ctypes.data_as(ctypes.POINTER(t))


def reference_loops(oracle, mesh, x, threads):
    """Run the reference's own loops (build container only): (energy, gradient, outer, inner, values, energy per
    element [all NaN where the total is NaN: not run]). A second
    assemble_hessian through the same matrix cache must reproduce pattern and values bit for bit."""
    lib = ctypes.CDLL(LIB_PATH)
    dp, ip, vp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.c_void_p
    lib.refloop_new.restype = vp
    lib.refloop_new.argtypes = [ctypes.c_int] * 4 + [ip, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double]
    lib.refloop_free.argtypes = [vp]
    lib.refloop_energy.restype = ctypes.c_double
    lib.refloop_energy.argtypes = [vp, dp, ctypes.c_int]
    lib.refloop_gradient.argtypes = [vp, dp, ctypes.c_int, dp]
    lib.refloop_energy_per_element.argtypes = [vp, dp, ctypes.c_int, dp]
    lib.refloop_hessian.restype = ctypes.c_long
    lib.refloop_hessian.argtypes = [vp, dp, ctypes.c_int]
    for name, rt in (("refloop_outer", ip), ("refloop_inner", ip), ("refloop_values", dp)):
        getattr(lib, name).restype = rt
        getattr(lib, name).argtypes = [vp]

    prob = oracle.problem_from_mesh(mesh, "NeoHookean")
    t = tables.reference_tables(mesh.p)
    ne, nl, nq = mesh.n_elements, mesh.conn.shape[1], t["weights"].size
    det, jit = np.zeros((ne, nq)), np.zeros((ne, nq, 9))
    for e in range(ne):
        d, j, _ = prob.assembly_values(e)  # pinned against finalize3d by tests/test_oracle_reference_math.py
        det[e], jit[e] = d, j.reshape(nq, 9)
    conn = np.ascontiguousarray(mesh.conn, dtype=np.int32)
    grads, w = np.ascontiguousarray(t["grad"]), np.ascontiguousarray(t["weights"])
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    x = np.ascontiguousarray(x, dtype=np.float64)
    h = lib.refloop_new(ne, nl, nq, mesh.n_bases, _ptr(conn, ctypes.c_int), _ptr(grads), _ptr(jit), _ptr(det), _ptr(w), lam, mu)
    try:
        energy = lib.refloop_energy(h, _ptr(x), threads)
        grad = np.zeros(mesh.n_bases * 3)
        lib.refloop_gradient(h, _ptr(x), threads, _ptr(grad))
        epe = np.full(ne, np.nan)
        if np.isfinite(energy):  # with a NaN energy the reference's own debug assert (Assembler.cpp:565-569) aborts
            lib.refloop_energy_per_element(h, _ptr(x), threads, _ptr(epe))
        nnz = lib.refloop_hessian(h, _ptr(x), threads)
        outer = np.ctypeslib.as_array(lib.refloop_outer(h), shape=(mesh.n_bases * 3 + 1,)).copy()
        inner = np.ctypeslib.as_array(lib.refloop_inner(h), shape=(nnz,)).copy()
        values = np.ctypeslib.as_array(lib.refloop_values(h), shape=(nnz,)).copy()
        assert lib.refloop_hessian(h, _ptr(x), threads) == nnz  # cached-pattern path
        again = np.ctypeslib.as_array(lib.refloop_values(h), shape=(nnz,)).copy()
        assert np.array_equal(values, again, equal_nan=True), "the reference's cached-pattern pass changed the values"
        assert np.array_equal(outer, np.ctypeslib.as_array(lib.refloop_outer(h), shape=(mesh.n_bases * 3 + 1,)))
        assert np.array_equal(inner, np.ctypeslib.as_array(lib.refloop_inner(h), shape=(nnz,)))
    finally:
        lib.refloop_free(h)
    return energy, grad, outer, inner, values, epe


LINEAR_GOLD_PATH = os.path.join(ROOT, "tests", "golden", "linear_loops.npz")
LINEAR_IDS = {"LinearElasticity": 0, "Laplacian": 1, "Mass": 2}
RHO = 1000.0


def linear_cases():
    """(name, material, mesh): the global linear loop LinearAssembler::assemble (Assembler.cpp:157-384) on jittered cubes."""
    out = []
    for material, p, n in (("LinearElasticity", 1, 3), ("LinearElasticity", 2, 2), ("Laplacian", 2, 2), ("Laplacian", 3, 1),
                           ("Mass", 1, 3), ("Mass", 2, 1)):
        out.append((f"{material}_p{p}", material, M.kuhn_cube(n, p, jitter=0.2)))
    return out


def reference_linear_loop(oracle, material, mesh, threads):
    """Run the reference's own linear loop (build container only): (outer, inner, values) of the stiffness / mass matrix."""
    lib = ctypes.CDLL(LIB_PATH)
    dp, ip, vp = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int), ctypes.c_void_p
    lib.refloop_linear_new.restype = vp
    lib.refloop_linear_new.argtypes = [ctypes.c_int] * 5 + [ip, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double]
    lib.refloop_free.argtypes = [vp]
    lib.refloop_linear_assemble.restype = ctypes.c_long
    lib.refloop_linear_assemble.argtypes = [vp, ctypes.c_int]
    lib.refloop_cols.argtypes = [vp]
    for name, rt in (("refloop_outer", ip), ("refloop_inner", ip), ("refloop_values", dp)):
        getattr(lib, name).restype = rt
        getattr(lib, name).argtypes = [vp]

    prob = oracle.problem_from_mesh(mesh, material, rho=RHO)
    is_mass = material == "Mass"
    t = tables.reference_tables(mesh.p, tables.quadrature_order(mesh.p, is_mass=True) if is_mass else None)
    ne, nl, nq = mesh.n_elements, mesh.conn.shape[1], t["weights"].size
    det, gtm = np.zeros((ne, nq)), np.zeros((ne, nq, nl, 3))
    for e in range(ne):
        det[e], _, gtm[e] = prob.assembly_values(e)  # pinned against finalize3d by tests/test_oracle_reference_math.py
    conn = np.ascontiguousarray(mesh.conn, dtype=np.int32)
    w = np.ascontiguousarray(t["weights"])
    vals = np.ascontiguousarray(t["val"]) if is_mass else None
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    h = lib.refloop_linear_new(LINEAR_IDS[material], ne, nl, nq, mesh.n_bases, _ptr(conn, ctypes.c_int), None if is_mass else _ptr(gtm),
                               _ptr(vals) if is_mass else None, _ptr(det), _ptr(w), lam, mu, RHO)
    try:
        nnz = lib.refloop_linear_assemble(h, threads)
        cols = lib.refloop_cols(h)
        outer = np.ctypeslib.as_array(lib.refloop_outer(h), shape=(cols + 1,)).copy()
        inner = np.ctypeslib.as_array(lib.refloop_inner(h), shape=(nnz,)).copy()
        values = np.ctypeslib.as_array(lib.refloop_values(h), shape=(nnz,)).copy()
    finally:
        lib.refloop_free(h)
    return outer, inner, values


def golden():
    if not os.path.exists(GOLD_PATH):
        pytest.skip("tests/golden/nl_loops.npz missing (run tools/make_golden.py where the reference tree is mounted)")
    return np.load(GOLD_PATH)


def close(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    ok = ~np.isnan(b)
    if ok.any():
        assert np.abs(a[ok] - b[ok]).max() <= TOL * max(np.abs(b[ok]).max(), 1e-300)


@pytest.mark.parametrize("k", range(4))
@pytest.mark.parametrize("n_threads", (1, 4))
def test_oracle_global_assembly_equals_reference_loops(oracle, k, n_threads):
    G = golden()
    name, mesh, x = loop_cases()[k]
    assert np.array_equal(G[f"x_{name}"], x), "golden inputs are stale: rerun tools/make_golden.py"
    prob = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=n_threads)
    for pass_ in range(2):  # second pass: the oracle's cached-pattern path
        H = prob.assemble_hessian(x)
        assert np.array_equal(H.outer, G[f"outer_{name}"])  # the pattern does not depend on the number of storages
        assert np.array_equal(H.inner, G[f"inner_{name}"])
        for t in THREADS:
            close(H.values, G[f"values_{name}_t{t}"])
    for t in THREADS:
        close(prob.assemble_gradient(x), G[f"gradient_{name}_t{t}"])
        close([prob.assemble_energy(x)], [float(G[f"energy_{name}_t{t}"])])
    if np.isfinite(float(G[f"energy_{name}_t1"])):
        close(prob.assemble_energy_per_element(x), G[f"energy_per_element_{name}"])


def test_golden_covers_nan_and_thread_merge():
    G = golden()
    assert np.isnan(G["values_p1_inverted_t1"]).any() and np.isnan(G["gradient_p1_inverted_t1"]).any()
    assert not np.isnan(G["values_p2_t1"]).any()
    # 1 and 3 thread storages: same pattern, values equal up to the order of the merge
    for name in ("p1", "p2", "p3"):
        close(G[f"values_{name}_t3"], G[f"values_{name}_t1"])


def test_live_reference_loops_reproduce_the_golden(oracle):
    """Build container only: the committed golden is what libloopref.so returns today (bit for bit: chunks run serially)."""
    if not os.path.exists(LIB_PATH):
        pytest.skip("oracle/_ref/libloopref.so not built (no reference tree)")
    G = golden()
    for name, mesh, x in loop_cases():
        for t in THREADS:
            e, g, o, i, v, epe = reference_loops(oracle, mesh, x, t)
            assert np.array_equal(epe, G[f"energy_per_element_{name}"], equal_nan=True)
            assert np.array_equal(o, G[f"outer_{name}"]) and np.array_equal(i, G[f"inner_{name}"])
            for a, b in ((v, G[f"values_{name}_t{t}"]), (g, G[f"gradient_{name}_t{t}"]),
                         (np.array([e]), np.array([float(G[f"energy_{name}_t{t}"])]))):
                assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("k", range(6))
@pytest.mark.parametrize("n_threads", (1, 4))
def test_oracle_linear_assembly_equals_reference_loop(oracle, k, n_threads):
    """LinearAssembler::assemble: pattern identical (incl. the stored zeros of Mass off its block diagonal), values 1e-13."""
    if not os.path.exists(LINEAR_GOLD_PATH):
        pytest.skip("tests/golden/linear_loops.npz missing")
    G = np.load(LINEAR_GOLD_PATH)
    name, material, mesh = linear_cases()[k]
    assert np.array_equal(G[f"vertices_{name}"], mesh.vertices), "golden inputs are stale: rerun tools/make_golden.py"
    K = oracle.problem_from_mesh(mesh, material, rho=RHO, n_threads=n_threads).assemble()
    assert np.array_equal(K.outer, G[f"outer_{name}"]) and np.array_equal(K.inner, G[f"inner_{name}"])
    close(K.values, G[f"values_{name}"])


def test_live_reference_linear_loop_reproduces_the_golden(oracle):
    if not os.path.exists(LIB_PATH):
        pytest.skip("oracle/_ref/libloopref.so not built (no reference tree)")
    G = np.load(LINEAR_GOLD_PATH)
    for name, material, mesh in linear_cases():
        o, i, v = reference_linear_loop(oracle, material, mesh, 1)
        assert np.array_equal(o, G[f"outer_{name}"]) and np.array_equal(i, G[f"inner_{name}"]) and np.array_equal(v, G[f"values_{name}"])
        o3, i3, v3 = reference_linear_loop(oracle, material, mesh, 3)  # 3 thread storages: same pattern, merge order differs
        assert np.array_equal(o, o3) and np.array_equal(i, i3)
        close(v3, v)
