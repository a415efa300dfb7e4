"""Multi-GPU owner-computes form on the device: every rank's handle (pfa_partition_create + PFA_FLAG_GHOST_GEOMETRY +
pfa_mesh_desc.owned_nodes) writes the FINISHED columns and gradient entries of the nodes it owns and nothing else; the
energies of the ranks add up to the oracle's. The ranks' handles are created one after the other on cuda:0, so the test needs
one GPU only; what it does not cover is the NCCL all-reduce of the energy (tests/test_dist_gloo.py covers the host logic with
gloo, `tools/dist_check.py --owner` the same comparison with one process per GPU)."""
import numpy as np
import pytest

from helpers import REL_TOL, make_case
from polyfem_b200 import dist as pdist, mesh as M

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p,n,world,jitter", [(2, 4, 2, 0.2), (2, 5, 3, 0.0), (1, 6, 4, 0.1), (2, 6, 8, 0.1)])
def test_owned_columns_on_the_device_equal_the_oracle(oracle, p, n, world, jitter):
    import torch
    mesh, x, t = make_case(n, p, jitter=jitter)
    x = x[: mesh.n_bases * 3]
    ref = oracle.problem_from_mesh(mesh, "NeoHookean", n_threads=4)
    H = ref.assemble_hessian(x).to_scipy().tocsc()
    g_ref, e_ref = ref.assemble_gradient(x), ref.assemble_energy(x)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    e_sum, owned_total = 0.0, 0
    for rank in range(world):
        part = pdist.partition_owner_computes(mesh, rank, world)
        h = pdist.owner_handle(part, t, lam, mu)
        h.profile_enable(True)
        xd = torch.from_numpy(np.ascontiguousarray(x.reshape(-1, 3)[part.l2g].reshape(-1))).cuda()
        e = torch.zeros(1, dtype=torch.float64, device="cuda")
        g = torch.full((h.ndof,), 7.0, dtype=torch.float64, device="cuda")      # sentinels: entries of nodes owned elsewhere
        v = torch.full((h.nnz,), 7.0, dtype=torch.float64, device="cuda")       # must stay untouched
        h.grad_hess_raw(xd, e, g, v)
        h.synchronize()
        assert any("column_lane" in k for (k, ms) in h.profile_read())
        e_sum += float(e.item())
        g, v = g.cpu().numpy(), v.cpu().numpy()
        adj_off, adj = h.block_pattern()
        for b in range(part.n_bases):
            sl = slice(9 * adj_off[b], 9 * adj_off[b + 1])
            if not part.owned[b]:
                assert (v[sl] == 7.0).all() and (g[3 * b:3 * b + 3] == 7.0).all()
                continue
            gb = int(part.l2g[b])
            deg = adj_off[b + 1] - adj_off[b]
            rows_g = part.l2g[adj[adj_off[b]:adj_off[b + 1]]]
            for m in range(3):
                assert H[:, 3 * gb + m].nnz == 3 * deg
                mine = v[9 * adj_off[b] + m * 3 * deg: 9 * adj_off[b] + (m + 1) * 3 * deg].reshape(deg, 3)
                want = np.asarray(H[(3 * rows_g[:, None] + np.arange(3)[None, :]).reshape(-1), 3 * gb + m].todense()).reshape(deg, 3)
                assert np.abs(mine - want).max() <= REL_TOL * np.abs(want).max()
            assert np.abs(g[3 * b:3 * b + 3] - g_ref[3 * gb:3 * gb + 3]).max() <= REL_TOL * np.abs(g_ref).max()
            owned_total += 1
        h.close()
    assert owned_total == mesh.n_bases
    assert abs(e_sum - e_ref) <= REL_TOL * abs(e_ref)
