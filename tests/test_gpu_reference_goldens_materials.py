"""CUDA path vs the outputs of the reference's OWN code for the materials added in round 2 (tests/golden/vd_local.npz, fc_local.npz,
mr_local.npz, sv_local.npz: function bodies compiled verbatim from /root/reference - FixedCorotational over the reference's own
SVD, MooneyRivlin and SaintVenant through the reference's own autodiff scalars; tests/test_oracle_*_reference.py). Single-element
meshes: the assembled matrix is the dense local Hessian."""
import os

import numpy as np
import pytest

from polyfem_b200 import tables

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VD = np.load(os.path.join(GOLDEN, "vd_local.npz"))
FC = np.load(os.path.join(GOLDEN, "fc_local.npz"))
MR = np.load(os.path.join(GOLDEN, "mr_local.npz"))
SV = np.load(os.path.join(GOLDEN, "sv_local.npz"))
LE = np.load(os.path.join(GOLDEN, "le_nl_local.npz"))


def dense(h, v, n):
    outer, inner = h.pattern()
    assert h.nnz == n * n
    H = np.zeros((n, n))
    H[inner, np.repeat(np.arange(n), np.diff(outer))] = v
    return H


@pytest.mark.parametrize("k", range(int(VD["n_cases"])))
def test_viscous_damping_equals_the_reference_code(k):
    from polyfem_b200 import capi
    p = int(VD[f"p_{k}"])
    t = tables.reference_tables(p)
    u, u_prev = VD[f"u_{k}"], VD[f"u_prev_{k}"]
    nl = u.shape[0]
    h = capi.Handle("ViscousDamping", np.arange(nl, dtype=np.int32)[None, :], nl, t["weights"], t["grad"], vertices=VD[f"vertices_{k}"][None],
                    lam=float(VD["psi"]), mu=float(VD["phi"]))
    h.set_previous(u_prev.reshape(-1), float(VD[f"dt_{k}"]))
    e, g, v = h.grad_hess(u.reshape(-1))
    e_ref, g_ref, H_ref = float(VD[f"energy_{k}"]), VD[f"gradient_{k}"], VD[f"hessian_{k}"]
    assert abs(e - e_ref) <= 1e-12 * abs(e_ref)
    assert np.abs(g - g_ref).max() <= 1e-12 * np.abs(g_ref).max()
    assert np.abs(dense(h, v, 3 * nl) - H_ref).max() <= 1e-12 * np.abs(H_ref).max()


@pytest.mark.parametrize("k", range(int(FC["n_cases"])))
def test_fixed_corotational_equals_the_reference_code(k):
    from polyfem_b200 import capi
    p = int(FC[f"p_{k}"])
    t = tables.reference_tables(p)
    u = FC[f"u_{k}"]
    nl = u.shape[0]
    verts = FC[f"vertices_{k}"]
    h = capi.Handle("FixedCorotational", np.arange(nl, dtype=np.int32)[None, :], nl, t["weights"], t["grad"], vertices=verts[None],
                    lam=float(FC["lambda"]), mu=float(FC["mu"]))
    e, g, v = h.grad_hess(u.reshape(-1))
    e_ref, g_ref, H_ref = float(FC[f"energy_{k}"]), FC[f"gradient_{k}"], FC[f"hessian_{k}"]
    hs, size = np.abs(H_ref).max(), float(np.linalg.norm(verts[1] - verts[0]))
    # 1e-11 of the tangent scale (energy and stress vanish in the rigid-rotation case); the kernel and the reference take the
    # singular vectors from F^T F by different eigen-solvers
    assert abs(e - e_ref) <= 1e-11 * max(abs(e_ref), hs * size * size * 1e-3)
    assert np.abs(g - g_ref).max() <= 1e-11 * max(np.abs(g_ref).max(), hs * size * 1e-3)
    assert np.abs(dense(h, v, 3 * nl) - H_ref).max() <= 1e-11 * hs


@pytest.mark.parametrize("k", range(int(MR["n_cases"])))
def test_mooney_rivlin_equals_the_reference_code(k):
    from polyfem_b200 import capi
    p = int(MR[f"p_{k}"])
    t = tables.reference_tables(p)
    u, verts = MR[f"u_{k}"], MR[f"vertices_{k}"]
    nl = u.shape[0]
    c1, c2, kk = float(MR["c1"]), float(MR["c2"]), float(MR["k"])
    h = capi.Handle("MooneyRivlin", np.arange(nl, dtype=np.int32)[None, :], nl, t["weights"], t["grad"], vertices=verts[None], lam=c1, mu=c2, param3=kk)
    e, g, v = h.grad_hess(u.reshape(-1))
    e_ref, g_ref, H_ref = float(MR[f"energy_{k}"]), MR[f"gradient_{k}"], MR[f"hessian_{k}"]
    vol = abs(np.linalg.det(verts[1:] - verts[0])) / 6.0
    # (the energy of a tiny strain is a difference of numbers near 3 (c1 + c2) per unit volume on both sides)
    assert abs(e - e_ref) <= 1e-12 * max(abs(e_ref), 3.0 * (c1 + c2) * vol)
    assert np.abs(g - g_ref).max() <= 1e-12 * max(np.abs(g_ref).max(), 1e-6 * np.abs(H_ref).max() * float(np.linalg.norm(verts[1] - verts[0])))
    assert np.abs(dense(h, v, 3 * nl) - H_ref).max() <= 1e-12 * np.abs(H_ref).max()


@pytest.mark.parametrize("k", range(int(SV["n_cases"])))
def test_saint_venant_equals_the_reference_code(k):
    from polyfem_b200 import capi
    p = int(SV[f"p_{k}"])
    t = tables.reference_tables(p)
    u = SV[f"u_{k}"]
    nl = u.shape[0]
    h = capi.Handle("SaintVenant", np.arange(nl, dtype=np.int32)[None, :], nl, t["weights"], t["grad"], vertices=SV[f"vertices_{k}"][None],
                    lam=float(SV["lambda"]), mu=float(SV["mu"]))
    e, g, v = h.grad_hess(u.reshape(-1))
    e_ref, g_ref, H_ref = float(SV[f"energy_{k}"]), SV[f"gradient_{k}"], SV[f"hessian_{k}"]
    assert abs(e - e_ref) <= 1e-12 * abs(e_ref)
    assert np.abs(g - g_ref).max() <= 1e-12 * np.abs(g_ref).max()
    assert np.abs(dense(h, v, 3 * nl) - H_ref).max() <= 1e-12 * np.abs(H_ref).max()


@pytest.mark.parametrize("k", range(int(LE["n_cases"])))
def test_linear_elasticity_nl_path_equals_the_reference_autodiff(k):
    """LinearElasticity inside a nonlinear solve: energy, gradient and Hessian (LinearElasticity.cpp:65-132, autodiff in the reference)"""
    from polyfem_b200 import capi
    p = int(LE[f"p_{k}"])
    t = tables.reference_tables(p)
    u = LE[f"u_{k}"]
    nl = u.shape[0]
    h = capi.Handle("LinearElasticity", np.arange(nl, dtype=np.int32)[None, :], nl, t["weights"], t["grad"], vertices=LE[f"vertices_{k}"][None],
                    lam=float(LE["lambda"]), mu=float(LE["mu"]))
    e, g, v = h.grad_hess(u.reshape(-1))
    e_ref, g_ref, H_ref = float(LE[f"energy_{k}"]), LE[f"gradient_{k}"], LE[f"hessian_{k}"]
    assert abs(e - e_ref) <= 1e-12 * abs(e_ref)
    assert np.abs(g - g_ref).max() <= 1e-12 * np.abs(g_ref).max()
    assert np.abs(dense(h, v, 3 * nl) - H_ref).max() <= 1e-12 * np.abs(H_ref).max()
