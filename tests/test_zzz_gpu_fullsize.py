"""BASELINE.json configs[2] at FULL size (NeoHookean P2, Kuhn cube n=69, 1 971 054 tets, 8.06 M dofs, 686 M nnz) on the GPU,
checked through size-independent properties (the oracle comparison at configuration size is tests/test_zzz_gpu_config_size.py:
cfg 1 and cfg 2 whole, cfg 3 at n = 30; at n = 69 the oracle mirrors the reference's memory design - a 7 GB slot map and one
5.5 GB value buffer per thread - which is what rules it out here, not its run time of a few minutes):

* the pattern size equals the closed form 9 (230 n^3 + 138 n^2 + 24 n + 1) (tests/test_oracle_properties.py pins it on the oracle);
* energy > 0, every output finite;
* translation invariance: the nodal forces sum to zero and H t = 0 for the three rigid translations t;
* symmetry: a^T H^T b == a^T H b for random a, b (pfa_symv reads a CSC column as a row, i.e. multiplies by H^T);
* the Hessian is the derivative of the gradient: (g(x + eps d) - g(x - eps d)) / (2 eps) == H d;
* at x = 0 the gradient vanishes and the NeoHookean Hessian (owner-computes kernels) equals the LinearElasticity stiffness
  (reference-moment kernel) entry by entry: two independent kernels, same pattern, 1e-12 of the largest entry.

Everything stays device-resident (torch tensors through the raw C-ABI entry points); this file sorts last on purpose."""
import numpy as np
import pytest

from polyfem_b200 import mesh as M, tables

pytestmark = pytest.mark.gpu

N, P = 69, 2


def test_cfg3_full_size_properties():
    import torch
    from polyfem_b200 import capi

    mesh = M.kuhn_cube(N, P)
    assert mesh.n_elements == 6 * N ** 3 == 1971054
    t = tables.reference_tables(P)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    h = capi.Handle("NeoHookean", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu)
    h.set_stream(torch.cuda.current_stream().cuda_stream)  # library launches ordered with the torch ops below (as bench.py does)
    assert h.ndof == 3 * (P * N + 1) ** 3
    assert h.nnz == 9 * (230 * N ** 3 + 138 * N ** 2 + 24 * N + 1)

    dev = "cuda"
    x = M.random_displacement(mesh)[: h.ndof]
    xd = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    e = torch.zeros(1, dtype=torch.float64, device=dev)
    g = torch.zeros(h.ndof, dtype=torch.float64, device=dev)
    v = torch.zeros(h.nnz, dtype=torch.float64, device=dev)
    h.grad_hess_raw(xd, e, g, v)
    h.synchronize()
    assert float(e.item()) > 0.0
    assert bool(torch.isfinite(g).all()) and bool(torch.isfinite(v).all())
    vmax = float(v.abs().max().item())

    # translation invariance of the energy: forces sum to zero, translations are in the null space of H
    assert float(g.view(-1, 3).sum(0).abs().max().item()) <= 1e-10 * float(g.abs().sum().item())
    y = torch.zeros(h.ndof, dtype=torch.float64, device=dev)
    s = torch.zeros(1, dtype=torch.float64, device=dev)
    for c in range(3):
        tc = torch.zeros(h.ndof, dtype=torch.float64, device=dev)
        tc[c::3] = 1.0
        h.inertia_raw(v, tc, None, s, y)  # y = H^T tc
        h.synchronize()
        assert float(y.abs().max().item()) <= 1e-10 * vmax

    # symmetry through two products with H^T
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    a = torch.rand(h.ndof, dtype=torch.float64, device=dev, generator=gen) - 0.5
    b = torch.rand(h.ndof, dtype=torch.float64, device=dev, generator=gen) - 0.5
    ya = torch.zeros_like(a)
    yb = torch.zeros_like(b)
    h.inertia_raw(v, a, None, s, ya)
    h.inertia_raw(v, b, None, s, yb)
    h.synchronize()
    s1, s2 = float(torch.dot(a, yb).item()), float(torch.dot(b, ya).item())  # a^T H^T b, a^T H b
    assert abs(s1 - s2) <= 1e-10 * float((yb.norm() * a.norm()).item())

    # H is the derivative of g: central difference along a random direction (Hd via H^T d, H symmetric as just shown)
    d = a / a.abs().max()
    eps = 1e-5 * mesh.h  # truncation error 4e-9 of |Hd| on the oracle at n = 4 (O(eps^2)), rounding ~ 1e-12
    gp = torch.zeros_like(g)
    gm = torch.zeros_like(g)
    h.grad_hess_raw(xd + eps * d, None, gp, None)
    h.grad_hess_raw(xd - eps * d, None, gm, None)
    h.inertia_raw(v, d, None, s, y)
    h.synchronize()
    fd = (gp - gm) / (2.0 * eps)
    assert float((fd - y).abs().max().item()) <= 1e-6 * float(y.abs().max().item())
    del gp, gm, fd, a, b, ya, yb, d

    # x = 0: zero forces, and the NeoHookean tangent is the LinearElasticity stiffness (another kernel, same pattern)
    xd.zero_()
    h.grad_hess_raw(xd, e, g, v)
    hl = capi.Handle("LinearElasticity", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu)
    hl.set_stream(torch.cuda.current_stream().cuda_stream)
    assert hl.nnz == h.nnz
    k = torch.zeros(hl.nnz, dtype=torch.float64, device=dev)
    hl.linear_stiffness_raw(k)
    h.synchronize()
    hl.synchronize()
    kmax = float(k.abs().max().item())
    assert abs(float(e.item())) <= 1e-12 * kmax * mesh.h ** 2
    assert float(g.abs().max().item()) <= 1e-12 * kmax * mesh.h
    assert float((v - k).abs().max().item()) <= 1e-12 * kmax
    o1, i1 = h.pattern()
    o2, i2 = hl.pattern()
    assert o1.tobytes() == o2.tobytes() and i1.tobytes() == i2.tobytes()
