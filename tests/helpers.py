"""Shared helpers of the parity tests (tolerances of SURVEY.md §8d, written out here)."""
import numpy as np

from polyfem_b200 import mesh as M
from polyfem_b200 import tables

# north_star: energy, gradient and Hessian values within 1e-12 relative (summation order differs)
REL_TOL = 1e-12


def row_scale(n, inner, values):
    s = np.zeros(n)
    np.maximum.at(s, inner, np.abs(values))
    return s


def assert_values_close(outer, inner, v_gpu, v_ref, tol=REL_TOL, what="hessian"):
    """abs(v_gpu - v_ref) <= tol * max(abs(v_ref), s_row), s_row = max |entry| of that row
    (guards structural zeros); NaN must coincide with NaN."""
    n = outer.size - 1
    nan_ref, nan_gpu = np.isnan(v_ref), np.isnan(v_gpu)
    assert np.array_equal(nan_ref, nan_gpu), f"{what}: NaN pattern differs"
    ok = ~nan_ref
    s = row_scale(n, inner[ok], v_ref[ok])[inner]
    bound = tol * np.maximum(np.abs(v_ref), s)
    err = np.abs(v_gpu - v_ref)
    bad = ok & (err > bound)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} entries off, worst ratio {np.nanmax(err[ok] / np.maximum(bound[ok], 1e-300)):.3g}"


def assert_vector_close(g_gpu, g_ref, tol=REL_TOL, what="gradient"):
    nan_ref, nan_gpu = np.isnan(g_ref), np.isnan(g_gpu)
    assert np.array_equal(nan_ref, nan_gpu), f"{what}: NaN pattern differs"
    ok = ~nan_ref
    if not ok.any():
        return
    s = np.abs(g_ref[ok]).max()
    err = np.abs(g_gpu[ok] - g_ref[ok]).max()
    assert err <= tol * max(s, 1e-300), f"{what}: max err {err:.3g} vs scale {s:.3g}"


def make_case(n, p, jitter=0.0, scale=0.05, seed=42):
    mesh = M.kuhn_cube(n, p, jitter=jitter)
    x = M.random_displacement(mesh, scale=scale, seed=seed)
    return mesh, x, tables.reference_tables(p)


def gpu_handle(mesh, material, t=None, E=1e5, nu=0.3, **kw):
    from polyfem_b200 import capi
    t = t or tables.reference_tables(mesh.p)
    lam, mu = M.lame_from_E_nu(E, nu)
    return capi.Handle(material, mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices,
                       lam=lam, mu=mu, **kw)
