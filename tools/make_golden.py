#!/usr/bin/env python
"""Regenerate the committed data/golden files from the reference's own sources.

Runs only in the build container (needs /root/reference): `oracle/Makefile` compiles the
reference's `autogen/auto_p_bases.cpp` and `quadrature/TetQuadrature.cpp` unmodified into
`oracle/_ref/libpfref.so`; this script calls it and writes

  polyfem_b200/data/tet_quadrature.json   tet rules, orders 1..8, hex floats (bit exact)
  tests/golden/ref_tables.npz             reference P1..P4 nodes, values and gradients
                                          evaluated by the reference's generated code at
                                          the quadrature points of orders 1,2,4,6 and at 16
                                          fixed interior points

Nothing under tests/ or the product reads /root/reference at run time; they read these files.
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_ref():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpfref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.pfref_tet_quadrature.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int]
    lib.pfref_p_nodes_3d.argtypes = [ctypes.c_int, dp, ctypes.c_int]
    lib.pfref_p_basis_3d.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp]
    return lib


def ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def ref_quadrature(lib, order):
    pts = np.zeros((512, 3))
    w = np.zeros(512)
    n = lib.pfref_tet_quadrature(order, ptr(pts), ptr(w), 512)
    assert n > 0
    return pts[:n].copy(), w[:n].copy()


def ref_nodes(lib, p):
    nodes = np.zeros((64, 3))
    n = lib.pfref_p_nodes_3d(p, ptr(nodes), 64)
    return nodes[:n].copy()


def ref_basis(lib, p, pts):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    n_loc = ref_nodes(lib, p).shape[0]
    val = np.zeros((pts.shape[0], n_loc))
    grad = np.zeros((pts.shape[0], n_loc, 3))
    for li in range(n_loc):
        v = np.zeros(pts.shape[0])
        g = np.zeros((pts.shape[0], 3))
        lib.pfref_p_basis_3d(p, li, pts.shape[0], ptr(pts), ptr(v), ptr(g))
        val[:, li] = v
        grad[:, li, :] = g
    return val, grad


def load_nhref():
    """oracle/_ref/libnhref.so: the reference's own NeoHookean gradient / Hessian function bodies (oracle/refmath)."""
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libnhref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    for f in (lib.ref_nh_energy, lib.ref_nh_gradient, lib.ref_nh_hessian):
        f.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp]
    lib.ref_linear_elasticity_block.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp]
    lib.ref_laplacian_block.argtypes = [ctypes.c_int, dp, dp, dp, dp]
    lib.ref_mass_block.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_double, dp]
    return lib


def linear_reference_blocks(lib, vertices, t, t_mass, lam, mu, rho):
    """All local blocks of one affine element from the reference's own LinearElasticity / Laplacian / Mass
    ::assemble(LinearAssemblerData): (le[n_loc][n_loc][9], lap[n_loc][n_loc], mass[n_loc][n_loc][9]), block entries
    in the reference's index n*size + m. grad_t_m = grad * J^-T and da = det * w are evaluated here with numpy."""
    edges = vertices[1:] - vertices[0]
    jit = np.linalg.inv(edges).T
    det = np.linalg.det(edges)
    gt = np.einsum("qic,cd->qid", t["grad"], jit)
    da = np.ascontiguousarray(det * t["weights"])
    dam = np.ascontiguousarray(det * t_mass["weights"])
    nl = t["grad"].shape[1]
    le, lap, mass = np.zeros((nl, nl, 9)), np.zeros((nl, nl)), np.zeros((nl, nl, 9))
    for i in range(nl):
        gi = np.ascontiguousarray(gt[:, i, :])
        vi = np.ascontiguousarray(t_mass["val"][:, i])
        for j in range(nl):
            gj = np.ascontiguousarray(gt[:, j, :])
            vj = np.ascontiguousarray(t_mass["val"][:, j])
            o1 = np.zeros(1)
            assert lib.ref_linear_elasticity_block(da.size, ptr(gi), ptr(gj), ptr(da), lam, mu, ptr(le[i, j])) == 0
            assert lib.ref_laplacian_block(da.size, ptr(gi), ptr(gj), ptr(da), ptr(o1)) == 0
            lap[i, j] = o1[0]
            assert lib.ref_mass_block(dam.size, ptr(vi), ptr(vj), ptr(dam), rho, ptr(mass[i, j])) == 0
    return le, lap, mass


def nh_reference_local(lib, vertices, u, grads, weights, lam, mu):
    """(energy, gradient[n_loc*3], hessian[N,N]) of one affine element from the reference's own functions; the
    geometry (J^-T, det, ElementAssemblyValues.cpp:81-103) is evaluated here with numpy."""
    edges = vertices[1:] - vertices[0]
    jit = np.linalg.inv(edges).T
    nq, nl = weights.size, grads.shape[1]
    jac_it = np.ascontiguousarray(np.repeat(jit[None], nq, 0).reshape(nq, 9))
    da = np.ascontiguousarray(np.linalg.det(edges) * weights)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
    grads = np.ascontiguousarray(grads)
    g = np.zeros(nl * 3)
    H = np.zeros((nl * 3, nl * 3))
    e = np.zeros(1)
    assert lib.ref_nh_energy(nl, nq, ptr(u), ptr(grads), ptr(jac_it), ptr(da), lam, mu, ptr(e)) == 0
    assert lib.ref_nh_gradient(nl, nq, ptr(u), ptr(grads), ptr(jac_it), ptr(da), lam, mu, ptr(g)) == 0
    assert lib.ref_nh_hessian(nl, nq, ptr(u), ptr(grads), ptr(jac_it), ptr(da), lam, mu, ptr(H)) == 0
    return float(e[0]), g, H


def write_nh_golden():
    """tests/golden/nh_local.npz: single-element NeoHookean cases (P1..P4, jittered tets, random displacement,
    one inverted element) with the gradient and Hessian returned by the reference's own function bodies, and
    all LinearElasticity / Laplacian / Mass local blocks of one element per order."""
    sys.path.insert(0, ROOT)
    from polyfem_b200 import tables
    lib = load_nhref()
    rng = np.random.default_rng(424242)
    lam, mu = 57692.307692307695, 38461.53846153846  # E = 1e5, nu = 0.3 (MatParams.cpp:11-22)
    gold = {"lambda": lam, "mu": mu}
    k = 0
    for p in (1, 2, 3, 4):
        t = tables.reference_tables(p)
        nodes = tables.p_nodes(p)  # reference-element positions of the local nodes
        for rep in range(3):
            verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
            verts = 0.3 * (verts + 0.25 * rng.uniform(-1, 1, (4, 3))) + rng.uniform(-1, 1, 3)
            scale = (0.05 if p < 3 else 0.01) * 0.3
            u = scale * rng.uniform(-1, 1, (nodes.shape[0], 3))
            if p == 2 and rep == 2:  # inverted: log(J <= 0) -> NaN must propagate exactly as in the reference
                u[1] += 3.0 * (verts[0] - verts[1])
            e, g, H = nh_reference_local(lib, verts, u, t["grad"], t["weights"], lam, mu)
            gold[f"energy_{k}"] = e
            gold[f"p_{k}"] = p
            gold[f"vertices_{k}"] = verts
            gold[f"u_{k}"] = u
            gold[f"gradient_{k}"] = g
            gold[f"hessian_{k}"] = H
            k += 1
    gold["n_cases"] = k
    # linear assemblers: every local block of one jittered element per order
    gold["rho"] = 2.5
    for p in (1, 2, 3, 4):
        verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
        verts = 0.4 * (verts + 0.25 * rng.uniform(-1, 1, (4, 3))) + rng.uniform(-1, 1, 3)
        t = tables.reference_tables(p)
        tm = tables.reference_tables(p, tables.quadrature_order(p, is_mass=True))
        le, lap, mass = linear_reference_blocks(lib, verts, t, tm, lam, mu, 2.5)
        gold[f"lin_vertices_p{p}"] = verts
        gold[f"le_blocks_p{p}"] = le
        gold[f"lap_blocks_p{p}"] = lap
        gold[f"mass_blocks_p{p}"] = mass
    # geometry: the reference's own ElementAssemblyValues::finalize3d on the same elements
    geo = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgeomref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    geo.ref_finalize3d.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp]
    for p in (1, 2, 3, 4):
        t = tables.reference_tables(p)
        nq, nl = t["weights"].size, t["grad"].shape[1]
        verts = np.ascontiguousarray(gold[f"lin_vertices_p{p}"])
        det, jit, gt = np.zeros(nq), np.zeros((nq, 3, 3)), np.zeros((nq, nl, 3))
        assert geo.ref_finalize3d(nl, nq, ptr(verts), ptr(np.ascontiguousarray(t["grad"])), ptr(det), ptr(jit), ptr(gt)) == 0
        gold[f"geo_det_p{p}"], gold[f"geo_jac_it_p{p}"], gold[f"geo_grad_t_m_p{p}"] = det, jit, gt
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nh_local.npz"), **gold)
    print(f"wrote tests/golden/nh_local.npz ({k} cases)")


def write_le_energy_golden():
    """tests/golden/le_energy.npz: LinearElasticity::compute_energy (the energy of a linear material inside a nonlinear solve,
    LinearElasticity.cpp:65-68, 103-132, the reference's own function body via libnhref.so) on the non-inverted single-element
    cases of nh_local.npz."""
    lib = load_nhref()
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ref_le_energy.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp]
    sys.path.insert(0, ROOT)
    from polyfem_b200 import tables
    G = np.load(os.path.join(ROOT, "tests", "golden", "nh_local.npz"))
    gold = {}
    for k in range(int(G["n_cases"])):
        t = tables.reference_tables(int(G[f"p_{k}"]))
        verts, u = G[f"vertices_{k}"], np.ascontiguousarray(G[f"u_{k}"]).reshape(-1)
        edges = verts[1:] - verts[0]
        nq, nl = t["weights"].size, t["grad"].shape[1]
        jac_it = np.ascontiguousarray(np.repeat(np.linalg.inv(edges).T[None], nq, 0).reshape(nq, 9))
        da = np.ascontiguousarray(np.linalg.det(edges) * t["weights"])
        e = np.zeros(1)
        assert lib.ref_le_energy(nl, nq, ptr(u), ptr(np.ascontiguousarray(t["grad"])), ptr(jac_it), ptr(da), float(G["lambda"]), float(G["mu"]), ptr(e)) == 0
        gold[f"le_energy_{k}"] = e[0]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "le_energy.npz"), **gold)
    print("wrote tests/golden/le_energy.npz")


def write_cache_golden():
    """tests/golden/cache_sequences.npz: the matrices the reference's own SparseMatrixCache (utils/MatrixCache.cpp compiled
    unmodified, oracle/_ref/libcacheref.so) returns for the call sequences of tests/test_oracle_cache_vs_reference.py."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_oracle_cache_vs_reference as T
    mats = T.golden_sequences(T.RefCache, lambda c: c.get_matrix())
    gold = {"n": len(mats)}
    for k, (o, i, v) in enumerate(mats):
        gold[f"outer_{k}"], gold[f"inner_{k}"], gold[f"values_{k}"] = o, i, v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cache_sequences.npz"), **gold)
    print(f"wrote tests/golden/cache_sequences.npz ({len(mats)} matrices)")


def write_bc_golden():
    """tests/golden/bc_projection.npz: not_constraints_, old_to_new_, projected gradient and projected CSC matrix from the
    reference's own BCLagrangianForm code (oracle/_ref/libbcref.so) for the cases of tests/test_oracle_projection.py."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_oracle_projection as T
    gold = {}
    for k, (csc, constrained, grad) in enumerate(T.golden_cases()):
        nc, o2n, g, (o, i, v) = T.reference_projection(csc, constrained, grad)
        gold[f"nc_{k}"], gold[f"o2n_{k}"], gold[f"g_{k}"] = nc, o2n, g
        gold[f"outer_{k}"], gold[f"inner_{k}"], gold[f"values_{k}"] = o, i, v
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bc_projection.npz"), **gold)
    print("wrote tests/golden/bc_projection.npz")


def write_loop_golden():
    """tests/golden/nl_loops.npz: energy, gradient and CSC Hessian of small multi-element meshes from the reference's own
    global loops NLAssembler::assemble_* (oracle/_ref/libloopref.so) with 1 and 3 per-thread storages."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_oracle_loops_vs_reference as T
    from oracle import pyoracle
    pyoracle.build()
    gold = {}
    for name, mesh, x in T.loop_cases():
        gold[f"x_{name}"] = x
        for t in T.THREADS:
            e, g, o, i, v, epe = T.reference_loops(pyoracle, mesh, x, t)
            gold[f"energy_per_element_{name}"] = epe  # does not depend on the number of storages
            gold[f"energy_{name}_t{t}"], gold[f"gradient_{name}_t{t}"], gold[f"values_{name}_t{t}"] = e, g, v
            if f"outer_{name}" in gold:
                assert np.array_equal(o, gold[f"outer_{name}"]) and np.array_equal(i, gold[f"inner_{name}"])
            gold[f"outer_{name}"], gold[f"inner_{name}"] = o, i
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nl_loops.npz"), **gold)
    print("wrote tests/golden/nl_loops.npz")
    gold = {}
    for name, material, mesh in T.linear_cases():
        gold[f"vertices_{name}"] = mesh.vertices
        gold[f"outer_{name}"], gold[f"inner_{name}"], gold[f"values_{name}"] = T.reference_linear_loop(pyoracle, material, mesh, 1)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "linear_loops.npz"), **gold)
    print("wrote tests/golden/linear_loops.npz")


def write_vd_golden():
    """tests/golden/vd_local.npz: single-element ViscousDamping cases (P1..P3, jittered tets, previous and current displacement,
    two time steps) with the energy, gradient and Hessian returned by the reference's own function bodies
    (oracle/_ref/libvdref.so = ViscousDamping.cpp:5-62, 122-229, 297-342 compiled verbatim against oracle/refmath/mini_eigen.hpp)."""
    sys.path.insert(0, ROOT)
    from polyfem_b200 import tables
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libvdref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ref_vd_local.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, dp, dp]
    rng = np.random.default_rng(535353)
    gold = {"psi": 30.0, "phi": 20.0}
    k = 0
    for p in (1, 2, 3):
        t = tables.reference_tables(p)
        nl, nq = t["grad"].shape[1], t["weights"].size
        grads = np.ascontiguousarray(t["grad"])
        for rep in range(3):
            verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
            verts = 0.3 * (verts + 0.25 * rng.uniform(-1, 1, (4, 3))) + rng.uniform(-1, 1, 3)
            u_prev = 0.03 * rng.uniform(-1, 1, (nl, 3))
            u = u_prev + 0.01 * rng.uniform(-1, 1, (nl, 3))
            dt = (0.05, 0.2, 1e-3)[rep]
            edges = verts[1:] - verts[0]
            jac_it = np.ascontiguousarray(np.repeat(np.linalg.inv(edges).T[None], nq, 0).reshape(nq, 9))
            da = np.ascontiguousarray(np.linalg.det(edges) * t["weights"])
            e, g, H = np.zeros(1), np.zeros(nl * 3), np.zeros((nl * 3, nl * 3))
            assert lib.ref_vd_local(nl, nq, ptr(np.ascontiguousarray(u.reshape(-1))), ptr(np.ascontiguousarray(u_prev.reshape(-1))), ptr(grads), ptr(jac_it),
                                    ptr(da), dt, gold["psi"], gold["phi"], ptr(e), ptr(g), ptr(H)) == 0
            gold[f"p_{k}"], gold[f"dt_{k}"] = p, dt
            gold[f"vertices_{k}"], gold[f"u_{k}"], gold[f"u_prev_{k}"] = verts, u, u_prev
            gold[f"energy_{k}"], gold[f"gradient_{k}"], gold[f"hessian_{k}"] = float(e[0]), g, H
            k += 1
    gold["n_cases"] = k
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vd_local.npz"), **gold)
    print(f"wrote tests/golden/vd_local.npz ({k} cases)")


def write_fc_golden():
    """tests/golden/fc_local.npz: single-element FixedCorotational cases (P1..P3, jittered tets; small, moderate and large
    displacements, one inverted element, one pure rotation) with the energy, gradient and Hessian returned by the reference's own
    function bodies over its own SVD (oracle/_ref/libfcref.so), plus the reference's signed SVD of a few matrices."""
    sys.path.insert(0, ROOT)
    from polyfem_b200 import tables
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libfcref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ref_fc_local.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp]
    lib.ref_svd3.argtypes = [dp, dp, dp, dp]
    rng = np.random.default_rng(646464)
    lam, mu = 57692.307692307695, 38461.53846153846  # E = 1e5, nu = 0.3
    gold = {"lambda": lam, "mu": mu}
    k = 0
    for p in (1, 2, 3):
        t = tables.reference_tables(p)
        nodes = tables.p_nodes(p)
        nl, nq = t["grad"].shape[1], t["weights"].size
        grads = np.ascontiguousarray(t["grad"])
        for rep in range(4):
            verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
            verts = 0.3 * (verts + 0.25 * rng.uniform(-1, 1, (4, 3))) + rng.uniform(-1, 1, 3)
            scale = (1e-4, 0.05, 0.3, 0.05)[rep] * 0.3 * (1.0 if p < 3 else 0.3)
            u = scale * rng.uniform(-1, 1, (nl, 3))
            if rep == 3 and p == 1:  # inverted element (det F < 0): allowed by this material, sigma_2 < 0
                u[1] += 2.5 * (verts[0] - verts[1])
            if rep == 3 and p == 2:  # a rigid rotation of the element: zero energy and stress
                th = 0.8
                R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
                edges0 = verts[1:] - verts[0]
                X = verts[0] + nodes @ edges0
                u = X @ R.T - X
            edges = verts[1:] - verts[0]
            jac_it = np.ascontiguousarray(np.repeat(np.linalg.inv(edges).T[None], nq, 0).reshape(nq, 9))
            da = np.ascontiguousarray(np.linalg.det(edges) * t["weights"])
            e, g, H = np.zeros(1), np.zeros(nl * 3), np.zeros((nl * 3, nl * 3))
            assert lib.ref_fc_local(nl, nq, ptr(np.ascontiguousarray(u.reshape(-1))), ptr(grads), ptr(jac_it), ptr(da), lam, mu, ptr(e), ptr(g), ptr(H)) == 0
            gold[f"p_{k}"] = p
            gold[f"vertices_{k}"], gold[f"u_{k}"] = verts, u
            gold[f"energy_{k}"], gold[f"gradient_{k}"], gold[f"hessian_{k}"] = float(e[0]), g, H
            k += 1
    gold["n_cases"] = k
    mats, Us, Ss, Vs = [], [], [], []
    for trial in range(6):
        A = np.eye(3) + (0.3, 0.3, 1e-3, 1e-6, 0.8, 0.3)[trial] * rng.uniform(-1, 1, (3, 3))
        if trial == 5:
            A[:, 0] *= -1.0
        U, S, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
        lib.ref_svd3(ptr(np.ascontiguousarray(A)), ptr(U), ptr(S), ptr(V))
        mats.append(A), Us.append(U), Ss.append(S), Vs.append(V)
    gold["svd_A"], gold["svd_U"], gold["svd_S"], gold["svd_V"] = np.array(mats), np.array(Us), np.array(Ss), np.array(Vs)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fc_local.npz"), **gold)
    print(f"wrote tests/golden/fc_local.npz ({k} cases)")


def write_mr_golden():
    """tests/golden/mr_local.npz: single-element MooneyRivlin cases (P1..P4, jittered tets) with the energy, gradient and Hessian
    of the reference's own code path: elastic_energy<T> through its own autodiff scalars inside GenericElastic (oracle/_ref/libmrref.so)."""
    sys.path.insert(0, ROOT)
    from polyfem_b200 import tables
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmrref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ref_mr_local.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, dp, dp]
    rng = np.random.default_rng(757575)
    gold = {"c1": 11000.0, "c2": 7000.0, "k": 90000.0}
    k = 0
    for p in (1, 2, 3, 4):
        t = tables.reference_tables(p)
        nl, nq = t["grad"].shape[1], t["weights"].size
        grads = np.ascontiguousarray(t["grad"])
        for rep in range(3):
            verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
            verts = 0.3 * (verts + 0.25 * rng.uniform(-1, 1, (4, 3))) + rng.uniform(-1, 1, 3)
            scale = (1e-3, 0.05, 0.1)[rep] * 0.3 * (1.0, 1.0, 0.4, 0.1)[p - 1]
            u = scale * rng.uniform(-1, 1, (nl, 3))
            edges = verts[1:] - verts[0]
            jac_it = np.ascontiguousarray(np.repeat(np.linalg.inv(edges).T[None], nq, 0).reshape(nq, 9))
            da = np.ascontiguousarray(np.linalg.det(edges) * t["weights"])
            e, g, H = np.zeros(1), np.zeros(nl * 3), np.zeros((nl * 3, nl * 3))
            assert lib.ref_mr_local(nl, nq, ptr(np.ascontiguousarray(u.reshape(-1))), ptr(grads), ptr(jac_it), ptr(da), gold["c1"], gold["c2"], gold["k"],
                                    ptr(e), ptr(g), ptr(H)) == 0
            assert np.isfinite(e[0]) and np.isfinite(H).all()
            gold[f"p_{k}"] = p
            gold[f"vertices_{k}"], gold[f"u_{k}"] = verts, u
            gold[f"energy_{k}"], gold[f"gradient_{k}"], gold[f"hessian_{k}"] = float(e[0]), g, H
            k += 1
    gold["n_cases"] = k
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "mr_local.npz"), **gold)
    print(f"wrote tests/golden/mr_local.npz ({k} cases)")


def write_sv_golden():
    """tests/golden/sv_local.npz: single-element SaintVenant cases (P1..P4, jittered tets; also compressed states, where the
    tangent is indefinite) with the energy, gradient and Hessian of the reference's own code path (oracle/_ref/libsvref.so)."""
    sys.path.insert(0, ROOT)
    from polyfem_b200 import tables
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsvref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ref_sv_local.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp]
    rng = np.random.default_rng(868686)
    lam, mu = 57692.307692307695, 38461.53846153846
    gold = {"lambda": lam, "mu": mu}
    k = 0
    for p in (1, 2, 3, 4):
        t = tables.reference_tables(p)
        nodes = tables.p_nodes(p)
        nl, nq = t["grad"].shape[1], t["weights"].size
        grads = np.ascontiguousarray(t["grad"])
        for rep in range(3):
            verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
            verts = 0.3 * (verts + 0.25 * rng.uniform(-1, 1, (4, 3))) + rng.uniform(-1, 1, 3)
            u = (1e-3, 0.1, 0.05)[rep] * 0.3 * rng.uniform(-1, 1, (nl, 3))
            if rep == 2:  # compress the element to 40 % (SaintVenant allows it; the geometric stiffness turns negative)
                X = verts[0] + nodes @ (verts[1:] - verts[0])
                u = u - 0.6 * (X - X.mean(0))
            edges = verts[1:] - verts[0]
            jac_it = np.ascontiguousarray(np.repeat(np.linalg.inv(edges).T[None], nq, 0).reshape(nq, 9))
            da = np.ascontiguousarray(np.linalg.det(edges) * t["weights"])
            e, g, H = np.zeros(1), np.zeros(nl * 3), np.zeros((nl * 3, nl * 3))
            assert lib.ref_sv_local(nl, nq, ptr(np.ascontiguousarray(u.reshape(-1))), ptr(grads), ptr(jac_it), ptr(da), lam, mu, ptr(e), ptr(g), ptr(H)) == 0
            gold[f"p_{k}"] = p
            gold[f"vertices_{k}"], gold[f"u_{k}"] = verts, u
            gold[f"energy_{k}"], gold[f"gradient_{k}"], gold[f"hessian_{k}"] = float(e[0]), g, H
            k += 1
    gold["n_cases"] = k
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sv_local.npz"), **gold)
    print(f"wrote tests/golden/sv_local.npz ({k} cases)")


def write_le_nl_golden():
    """tests/golden/le_nl_local.npz: LinearElasticity as an NLAssembler (a linear material inside a nonlinear solve): energy and the
    reference's own AUTODIFF gradient / Hessian of it (LinearElasticity.cpp:65-132 through utils/autodiff.h and
    gradient_from_energy / hessian_from_energy; oracle/_ref/libsvref.so::ref_le_nl_local), P1..P4 single elements."""
    sys.path.insert(0, ROOT)
    from polyfem_b200 import tables
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsvref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.ref_le_nl_local.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, ctypes.c_double, ctypes.c_double, dp, dp, dp]
    rng = np.random.default_rng(979797)
    lam, mu = 57692.307692307695, 38461.53846153846
    gold = {"lambda": lam, "mu": mu}
    k = 0
    for p in (1, 2, 3, 4):
        t = tables.reference_tables(p)
        nl, nq = t["grad"].shape[1], t["weights"].size
        grads = np.ascontiguousarray(t["grad"])
        for rep in range(2):
            verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
            verts = 0.3 * (verts + 0.25 * rng.uniform(-1, 1, (4, 3))) + rng.uniform(-1, 1, 3)
            u = (0.05, 1.0)[rep] * 0.3 * rng.uniform(-1, 1, (nl, 3))
            edges = verts[1:] - verts[0]
            jac_it = np.ascontiguousarray(np.repeat(np.linalg.inv(edges).T[None], nq, 0).reshape(nq, 9))
            da = np.ascontiguousarray(np.linalg.det(edges) * t["weights"])
            e, g, H = np.zeros(1), np.zeros(nl * 3), np.zeros((nl * 3, nl * 3))
            assert lib.ref_le_nl_local(nl, nq, ptr(np.ascontiguousarray(u.reshape(-1))), ptr(grads), ptr(jac_it), ptr(da), lam, mu, ptr(e), ptr(g), ptr(H)) == 0
            gold[f"p_{k}"] = p
            gold[f"vertices_{k}"], gold[f"u_{k}"] = verts, u
            gold[f"energy_{k}"], gold[f"gradient_{k}"], gold[f"hessian_{k}"] = float(e[0]), g, H
            k += 1
    gold["n_cases"] = k
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "le_nl_local.npz"), **gold)
    print(f"wrote tests/golden/le_nl_local.npz ({k} cases)")


def write_geom_iso_golden():
    """tests/golden/geom_iso.npz: the reference's own ElementAssemblyValues::finalize3d (oracle/_ref/libgeomref.so) on CURVED
    elements - isoparametric P2 geometry (10 geometric nodes, edge midpoints pushed off their edges) under P2 and P3 bases: det,
    jac_it, grad_t_m per quadrature point."""
    sys.path.insert(0, ROOT)
    from polyfem_b200 import tables
    geo = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgeomref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    geo.ref_finalize3d_iso.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, dp]
    rng = np.random.default_rng(101010)
    gold = {}
    k = 0
    for p in (2, 3):
        t = tables.reference_tables(p)
        nq, nl = t["weights"].size, t["grad"].shape[1]
        # geometric basis: P2 at the quadrature points of the order-p rule
        tg = tables.reference_tables(2, tables.quadrature_order(p))
        assert np.allclose(tg["points"], t["points"])
        nodes2 = tables.p_nodes(2)
        for rep in range(3):
            verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
            verts = 0.3 * (verts + 0.25 * rng.uniform(-1, 1, (4, 3))) + rng.uniform(-1, 1, 3)
            gnodes = verts[0] + nodes2 @ (verts[1:] - verts[0])
            gnodes[4:] += 0.02 * rng.uniform(-1, 1, (6, 3))  # curved edges
            det, jit, gt = np.zeros(nq), np.zeros((nq, 3, 3)), np.zeros((nq, nl, 3))
            assert geo.ref_finalize3d_iso(nl, nq, 10, ptr(np.ascontiguousarray(gnodes)), ptr(np.ascontiguousarray(tg["grad"])),
                                          ptr(np.ascontiguousarray(t["grad"])), ptr(det), ptr(jit), ptr(gt)) == 0
            gold[f"p_{k}"] = p
            gold[f"vertices_{k}"], gold[f"geom_nodes_{k}"] = verts, gnodes
            gold[f"det_{k}"], gold[f"jac_it_{k}"], gold[f"grad_t_m_{k}"] = det, jit, gt
            k += 1
    gold["n_cases"] = k
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "geom_iso.npz"), **gold)
    print(f"wrote tests/golden/geom_iso.npz ({k} cases)")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "geom_iso":
        return write_geom_iso_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "le_nl":
        return write_le_nl_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "saint_venant":
        return write_sv_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "mooney":
        return write_mr_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "corotational":
        return write_fc_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "viscous":
        return write_vd_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "loops":
        return write_loop_golden()
    if len(sys.argv) > 1 and sys.argv[1] == "le_energy":
        return write_le_energy_golden()
    lib = load_ref()
    quad = {"source": "polyfem autogen/auto_tetrahedron.ipp via quadrature/TetQuadrature.cpp (weights /= 6)",
            "orders": {}}
    for order in range(1, 9):
        pts, w = ref_quadrature(lib, order)
        quad["orders"][str(order)] = {
            "points": [[float(c).hex() for c in row] for row in pts],
            "weights": [float(c).hex() for c in w],
        }
    os.makedirs(os.path.join(ROOT, "polyfem_b200", "data"), exist_ok=True)
    with open(os.path.join(ROOT, "polyfem_b200", "data", "tet_quadrature.json"), "w") as f:
        json.dump(quad, f, indent=0)

    rng = np.random.default_rng(20261017)
    bary = rng.dirichlet(np.ones(4), size=16)
    extra = bary[:, 1:4].copy()
    gold = {"extra_points": extra}
    for p in (1, 2, 3, 4):
        gold[f"nodes_p{p}"] = ref_nodes(lib, p)
        for order in (1, 2, 4, 6, 8):
            pts, w = ref_quadrature(lib, order)
            v, g = ref_basis(lib, p, pts)
            gold[f"val_p{p}_q{order}"] = v
            gold[f"grad_p{p}_q{order}"] = g
        v, g = ref_basis(lib, p, extra)
        gold[f"val_p{p}_extra"] = v
        gold[f"grad_p{p}_extra"] = g
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_tables.npz"), **gold)
    print("wrote tet_quadrature.json and tests/golden/ref_tables.npz")
    write_nh_golden()
    write_le_energy_golden()
    write_cache_golden()
    write_bc_golden()
    write_loop_golden()


if __name__ == "__main__":
    sys.exit(main())
