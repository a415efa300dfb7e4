"""TEST HELPER (run under torchrun by tests/test_gpu_multi_rank.py, one rank per GPU): the multi-GPU owner-computes step of
bench.py - pfa_partition_create, ghost elements with geometry, pfa_grad_hess per rank, NCCL all-reduce of the energy -
compared rank by rank with the CPU ORACLE on the whole mesh: owned columns and gradient entries 1e-12, energy 1e-12.
Exits non-zero on any rank's failure."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from oracle import pyoracle  # noqa: E402
from polyfem_b200 import dist as pdist, mesh as M, tables  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", dest="n", type=int, default=6)
    ap.add_argument("--order", dest="p", type=int, default=2)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    mesh = M.kuhn_cube(a.n, a.p, jitter=0.1)
    x = M.random_displacement(mesh)[: mesh.n_bases * 3]
    t = tables.reference_tables(a.p)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    part = pdist.partition_owner_computes(mesh, rank, world)
    h = pdist.owner_handle(part, t, lam, mu, device=local)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    xd = torch.from_numpy(np.ascontiguousarray(x.reshape(-1, 3)[part.l2g].reshape(-1))).to(dev)
    e = torch.zeros(1, dtype=torch.float64, device=dev)
    g = torch.zeros(h.ndof, dtype=torch.float64, device=dev)
    v = torch.zeros(h.nnz, dtype=torch.float64, device=dev)
    for _ in range(2):  # repeated steps give the same answer
        h.grad_hess_raw(xd, e, g, v)
        dist.all_reduce(e)
    torch.cuda.synchronize()

    ref = pyoracle.problem_from_mesh(mesh, "NeoHookean", n_threads=2)
    H = ref.assemble_hessian(x).to_scipy().tocsc()
    g_ref, e_ref = ref.assemble_gradient(x), ref.assemble_energy(x)
    adj_off, adj = h.block_pattern()
    gl, vl = g.cpu().numpy(), v.cpu().numpy()
    ok = abs(float(e.item()) - e_ref) <= 1e-12 * abs(e_ref)
    gerr = verr = 0.0
    scale = np.abs(H.data).max()
    for b in np.flatnonzero(part.owned):
        gb = int(part.l2g[b])
        deg = adj_off[b + 1] - adj_off[b]
        rows_g = part.l2g[adj[adj_off[b]:adj_off[b + 1]]]
        for m in range(3):
            if H[:, 3 * gb + m].nnz != 3 * deg:
                ok = False
            mine = vl[9 * adj_off[b] + m * 3 * deg: 9 * adj_off[b] + (m + 1) * 3 * deg].reshape(deg, 3)
            want = np.asarray(H[(3 * rows_g[:, None] + np.arange(3)[None, :]).reshape(-1), 3 * gb + m].todense()).reshape(deg, 3)
            verr = max(verr, float(np.abs(mine - want).max()) / scale)
        gerr = max(gerr, float(np.abs(gl[3 * b:3 * b + 3] - g_ref[3 * gb:3 * gb + 3]).max()) / np.abs(g_ref).max())
    ok = ok and gerr <= 1e-12 and verr <= 1e-12
    print(f"rank {rank}/{world}: own {part.n_own_elements} + ghost {part.n_ghost_elements} elements, owned nodes {int(part.owned.sum())}, "
          f"energy {float(e.item()):.12e} vs {e_ref:.12e}, grad err {gerr:.2e}, values err {verr:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    rc = 1 if int(flag.item()) else 0
    dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()
