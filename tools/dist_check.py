#!/usr/bin/env python
"""Multi-GPU parity check on real devices (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py [--cells 6 --order 2]

Every rank assembles its element block (interface elements first, exchange under the assembly
of the rest, as bench.py does), then compares the columns / dofs it owns with a single-GPU
assembly of the whole mesh done on its own device. Bars: energy 1e-12 relative, gradient and
values 1e-12 of the largest entry. Prints one line per rank and exits non-zero on failure."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from polyfem_b200 import capi, dist as pdist, mesh as M, tables  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", dest="n", type=int, default=6)
    ap.add_argument("--order", dest="p", type=int, default=2)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--bench-path", action="store_true", help="bench.py's default multi-GPU step: cuts on cell layers, one-tensor exchange after the assembly")
    ap.add_argument("--graph", action="store_true", help="capture the step in a CUDA graph and replay it")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    mesh = M.kuhn_cube(a.n, a.p, jitter=0.1)
    x = M.random_displacement(mesh)
    t = tables.reference_tables(a.p)
    lam, mu = M.lame_from_E_nu(1e5, 0.3)
    part = pdist.partition_elements(mesh, rank, world, align=6 * a.n * a.n if a.bench_path else 1)
    h = capi.Handle("NeoHookean", part.conn, part.n_bases, t["weights"], t["grad"], vertices=part.vertices, lam=lam, mu=mu,
                    device=local, n_ghost_elements=part.n_ghost_elements, n_first_elements=part.n_interface_elements)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    ex = pdist.InterfaceExchange(h, part, rank, world, dev, grad_offset=h.nnz)
    xd = torch.from_numpy(np.ascontiguousarray(x.reshape(-1, 3)[part.l2g].reshape(-1))).to(dev)
    e = torch.zeros(1, dtype=torch.float64, device=dev)
    vg = torch.zeros(h.nnz + h.ndof, dtype=torch.float64, device=dev)
    v, g = vg[:h.nnz], vg[h.nnz:]
    def step():
        if a.bench_path:
            h.grad_hess_raw(xd, e, g, v)
            ex.reduce_combined(e, vg)
        elif a.no_overlap:
            h.grad_hess_raw(xd, e, g, v)
            ex.reduce(e, g, v)
        else:
            h.grad_hess_part_raw(xd, e, g, v, 1)
            ex.start(g, v)
            h.grad_hess_part_raw(xd, e, g, v, 2)
            ex.finish(e)

    if a.graph:  # the way bench.py runs multi-GPU steps: captured once, replayed
        cap = torch.cuda.Stream(device=dev)
        h.set_stream(cap.cuda_stream)
        with torch.cuda.stream(cap):
            for _ in range(2):
                step()
        cap.synchronize()
        dist.barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=cap, capture_error_mode="thread_local"):  # NCCL's watchdog thread also calls CUDA
            step()
        e.fill_(7.0)
        g.fill_(7.0)
        v.fill_(7.0)
        for _ in range(3):
            graph.replay()
    else:
        for _ in range(3):  # repeated steps must give the same answer (buffers are re-cleared)
            step()
    torch.cuda.synchronize()

    hf = capi.Handle("NeoHookean", mesh.conn, mesh.n_bases, t["weights"], t["grad"], vertices=mesh.vertices, lam=lam, mu=mu, device=local)
    e_ref, g_ref, v_ref = hf.grad_hess(x)
    offf, adjf = hf.block_pattern()
    offl, adjl = h.block_pattern()
    gl, vl = g.cpu().numpy().reshape(-1, 3), v.cpu().numpy()
    ok = abs(float(e.item()) - e_ref) <= 1e-12 * abs(e_ref)
    mine = np.flatnonzero(part.owner == rank)
    gerr = np.abs(gl[mine] - g_ref.reshape(-1, 3)[part.l2g[mine]]).max() / np.abs(g_ref).max()
    scale = np.abs(v_ref).max()
    verr = 0.0
    for b in mine:
        gb = part.l2g[b]
        rows_l = part.l2g[adjl[offl[b]:offl[b + 1]]]
        rows_f = adjf[offf[gb]:offf[gb + 1]]
        order = np.argsort(rows_l)
        if not np.array_equal(rows_l[order], rows_f):
            ok = False
            break
        deg = rows_f.size
        bl = vl[9 * offl[b]: 9 * offl[b] + 9 * deg].reshape(3, deg, 3)[:, order, :]
        bf = v_ref[9 * offf[gb]: 9 * offf[gb] + 9 * deg].reshape(3, deg, 3)
        verr = max(verr, float(np.abs(bl - bf).max()) / scale)
    ok = ok and gerr <= 1e-12 and verr <= 1e-12
    print(f"rank {rank}/{world}: own {part.n_own_elements} elements ({part.n_interface_elements} interface), ghost {part.n_ghost_elements}, "
          f"owned nodes {mine.size}, energy {float(e.item()):.12e} vs {e_ref:.12e}, grad err {gerr:.2e}, values err {verr:.2e} -> {'OK' if ok else 'FAIL'}",
          flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    rc = 1 if int(flag.item()) else 0
    if a.graph:  # communicator teardown hangs while a captured graph references its kernels
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(rc)
    dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()
