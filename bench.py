#!/usr/bin/env python
"""bench.py — assembly hot-path benchmark (contract: task brief ④, BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W     CPU reference arm (oracle port)

A "step" is one fused energy + gradient + Hessian assembly (pfa_grad_hess) of NeoHookean
P2 tets on the synthetic Kuhn cube with a random displacement (BASELINE.json configs[2]:
n=69 cells/side, 1 971 054 elements, 8 056 857 dofs, 685 941 705 nnz). Prints ONE JSON line.

  value      elements/s, whole job, x / E / grad / values[] resident in HBM
  e2e        the same metric through the C-ABI call with HOST (pinned) buffers: H2D of x and
             D2H of energy, gradient and the CSC values[] inside the timed region
  roofline   achieved = B_alg(=SURVEY.md §8d bytes/element) * elements / assembly-kernel time
             (CUDA events on the launching stream, live in this run) vs measured HBM copy peak
  cpu_baseline  the CPU oracle (a port of the reference algorithm, NOT PolyFEM itself) timed
             on this box's host cores on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "elements/s (NeoHookean P2 grad+Hessian assembly)"
UNIT = "elements/s"
E_MOD, NU = 1e5, 0.3
RHO, DT = 1000.0, 1e-3  # cfg 5

# BASELINE.json configs[0..4] -> --config 1..5 (3 is the headline and the default; "4le" is the LinearElasticity half of
# configs[3]: 1.43 G nnz, 11.5 GB of values[]; --flags 32 hands the pattern out with int64 indices, POLYSOLVE_LARGE_INDEX).
#   mode: nl     fused energy + gradient + Hessian (pfa_grad_hess)
#         linear LinearAssembler::assemble (pfa_linear_stiffness)
#         euler  one implicit-Euler Newton assembly: dt^2-weighted elastic form + InertiaForm on the mass matrix
# cpu_n: cells per side of the bounded CPU sample of the same material / order (about 1-3 s of host work per step)
CONFIGS = {
    "1": dict(material="LinearElasticity", p=1, n=20, mode="linear", cpu_n=20, what="stiffness assembly (pfa_linear_stiffness)"),
    "2": dict(material="NeoHookean", p=1, n=44, mode="nl", cpu_n=30, what="fused energy+gradient+Hessian (pfa_grad_hess)"),
    "3": dict(material="NeoHookean", p=2, n=69, mode="nl", cpu_n=24, what="fused energy+gradient+Hessian (pfa_grad_hess)"),
    "4": dict(material="Laplacian", p=4, n=32, mode="linear", cpu_n=8, what="stiffness assembly (pfa_linear_stiffness)"),
    "4le": dict(material="LinearElasticity", p=4, n=32, mode="linear", cpu_n=5, what="stiffness assembly (pfa_linear_stiffness)"),
    "5": dict(material="NeoHookean", p=1, n=119, mode="euler", cpu_n=30,
              what="implicit-Euler Newton assembly: dt^2-weighted pfa_grad_hess_weighted + pfa_inertia + H = dt^2 H_el + M"),
}
FP64_PEAK_TFLOPS = 36.9  # DFMA peak measured on this pool's B200 (tools/microbench5.cu, profiles/microbench5_r02.jsonl)


def metric_name(cfg):
    if cfg["material"] == "NeoHookean" and cfg["p"] == 2 and cfg["mode"] == "nl":
        return METRIC
    what = {"nl": "grad+Hessian assembly", "linear": "stiffness assembly", "euler": "implicit-Euler grad+Hessian assembly"}[cfg["mode"]]
    return f"elements/s ({cfg['material']} P{cfg['p']} {what})"


def workload_name(cfg, mesh):
    size = 1 if cfg["material"] == "Laplacian" else 3
    return (f"{cfg['material']} P{cfg['p']} tets, Kuhn cube n={cfg['n']}: {mesh.n_elements} elements, {mesh.n_bases * size} dofs, {cfg['what']}")


def b_alg_bytes_per_element(n_loc, ndof, nnz, n_el, linear=False, laplacian=False):
    """compulsory traffic per element (inputs once, outputs once); the linear assemblers read no x and write
    no gradient, the Laplacian has no Lame parameters."""
    return 4 * n_loc + 80 + (0 if laplacian else 16) + (0 if linear else 2 * 8 * ndof / n_el) + 8 * nnz / n_el


def f_alg_flops_per_element(material, n_loc, n_qp):
    """SURVEY.md §8d algorithmic FLOPs per element (symmetric half of B^T D B for the linear forms)."""
    if material == "Laplacian":
        return n_qp * 3 * n_loc * n_loc
    if material == "LinearElasticity":
        N = 3 * n_loc
        return n_qp * (72 * N + 6 * N * N)
    return n_qp * (2 * (81 * n_loc + 27 * n_loc * (n_loc + 1) / 2) + 300)


def ncu_traffic(kernel_name, n_el, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed ncu --set full capture (profiles/traffic.json, written from tools/ncu_summary.py
    output). Only valid for the single-GPU headline workload it was captured on."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if world != 1 or not os.path.exists(path):
        return None
    with open(path) as f:
        rec = json.load(f)
    if rec.get("elements") != n_el or not kernel_name or rec.get("kernel") not in kernel_name:
        return None
    return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"])


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_first_sample(self, timeout=8.0):
        """nvidia-smi's start-up (NVML initialisation over all GPUs of the box) briefly stalls CUDA work on
        every device; a multi-GPU timed region is only tens of milliseconds long, so wait until the
        sampler is in its steady polling loop before any step is issued."""
        if self.proc is None:
            return
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                return
            time.sleep(0.05)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = [float(r[1]) for r in rows if len(r) >= 9]
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                seen = set()
                for r in rows:
                    for k, nm in enumerate(names):
                        if "Active" in r[5 + k] and "Not" not in r[5 + k]:
                            seen.add(nm)
                out["reasons"] = sorted(seen)
                out["samples"] = len(sm)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


def build_workload(n, p, rank=0, world=1):
    """Mesh + tables of the synthetic Kuhn cube."""
    from polyfem_b200 import mesh as M, tables
    mesh = M.kuhn_cube(n, p)
    x = M.random_displacement(mesh)
    t = tables.reference_tables(p)
    return mesh, x, t


def cpu_baseline_run(cfg, n_sample, steps, warmup, threads, use_cache=False):
    """Times the oracle (reference-algorithm port) on a bounded sample of the workload: the steady-state step of the config
    (pattern build = first call, untimed, like pfa_create). use_cache=False: assembly values recomputed per call, which is
    what the reference does at the headline size (the basis cache is dropped above 900 k bases,
    varforms/ElasticVarForm.cpp:217-232)."""
    from oracle import pyoracle
    from polyfem_b200 import mesh as M
    mesh = M.kuhn_cube(n_sample, cfg["p"])
    x = M.random_displacement(mesh)
    prob = pyoracle.problem_from_mesh(mesh, cfg["material"], E=E_MOD, nu=NU, n_threads=threads, use_cache=use_cache)
    nl = cfg["mode"] != "linear"
    mass = pyoracle.problem_from_mesh(mesh, "Mass", rho=RHO, n_threads=threads) if cfg["mode"] == "euler" else None
    M_mat = mass.assemble() if mass is not None else None
    xt = 0.5 * x if mass is not None else None

    def step():
        if not nl:
            prob.assemble()
            return
        prob.assemble_energy(x)
        prob.assemble_gradient(x)
        prob.assemble_hessian(x)
        if mass is not None:
            pyoracle.inertia(M_mat, x[: M_mat.outer.size - 1], xt[: M_mat.outer.size - 1])

    t0 = time.perf_counter()
    step()  # first call: triplets + pattern + slot map (one-off)
    first = time.perf_counter() - t0
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"elements": mesh.n_elements, "seconds_per_step": dt, "first_call_seconds": first, "value": mesh.n_elements / dt}


def cpu_sample_text(cfg, n_sample, r, use_cache, steps):
    return (f"{cfg['material']} P{cfg['p']} Kuhn cube n={n_sample} ({r['elements']} elements), {steps} steady-state steps ({cfg['what']}), "
            f"basis cache {'on' if use_cache else 'off (as the reference above 900 k bases)'}; first call (pattern build) {r['first_call_seconds']:.2f}s excluded")


def run_reference(args, cfg):
    """--impl reference: the CPU arm. Same metric / unit / config as our arm; every step is a bounded sample of that workload
    (cpu_baseline.sample says which), all host threads, the same K and W."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from polyfem_b200 import mesh as M
    threads = os.cpu_count() or 1
    steps, warmup = max(args.steps, 1), max(args.warmup, 3)
    n_sample = args.cpu_sample_n or cfg["cpu_n"]
    r = cpu_baseline_run(cfg, n_sample, steps, warmup, threads, use_cache=False)
    r_on = cpu_baseline_run(cfg, n_sample, 2, 1, threads, use_cache=True)
    # the config line of our arm (sizes in closed form: the headline mesh itself is not built here)
    n, p = cfg["n"], cfg["p"]
    n_el = 6 * n ** 3
    n_nodes = (p * n + 1) ** 3
    size = 1 if cfg["material"] == "Laplacian" else 3

    class Sizes:
        n_elements, n_bases = n_el, n_nodes
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(cfg, Sizes, None, 1, None),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": cpu_sample_text(cfg, n_sample, r, False, steps),
                         "value_basis_cache_on": r_on["value"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of the reference algorithm (oracle/), not the PolyFEM binary: Eigen/TBB are not available offline; "
                f"ms_per_step is the time of one step on the {r['elements']}-element sample",
    }
    print(json.dumps(line))
    return 0


def config_dict(cfg, mesh, h, world, parallelism):
    """`config` of the JSON line: identical for our arm and the reference arm (same workload, material, inputs)."""
    d = {"workload": workload_name(cfg, mesh), "material": "E=1e5 nu=0.3" if cfg["material"] != "Laplacian" else "-",
         "displacement": "0.05*h*U(-1,1) seed 42" if cfg["mode"] != "linear" else "-",
         "l2": "outputs (values[]) exceed the 126 MB L2 for configs 3-5; configs 1, 2: an L2 flush (256 MB write) runs between timed steps"}
    if cfg["mode"] == "euler":
        d["time_integrator"] = f"implicit Euler dt={DT} rho={RHO}: elastic weight dt^2 (ImplicitEuler.cpp:28-31), InertiaForm on the mass matrix"
    return d


def bind_to_gpu_cpus(index):
    """Multi-rank runs: pin this process to the CPUs NVML reports as closest to its GPU, so that the pinned host buffers of the
    end-to-end leg are first touched on that GPU's NUMA node (the D2H copies of 8 ranks otherwise cross the socket link).
    Best effort: returns the number of CPUs bound to, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hdl = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="3", choices=sorted(CONFIGS), help="BASELINE.json configs[k-1]; 3 = the headline (NeoHookean P2, n=69)")
    ap.add_argument("--n", type=int, default=None, help="cells per side, overrides the config's (69 -> 1 971 054 tets)")
    ap.add_argument("--p", type=int, default=None, help="basis order, overrides the config's")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-n", type=int, default=None, help="cells per side of the bounded CPU sample (default: per config, about 1-3 s of host work per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-align", action="store_true", help="multi-GPU: cut the element range evenly instead of on whole cell layers")
    ap.add_argument("--graph", action="store_true", help="multi-GPU, experimental: capture the step in a CUDA graph and replay it "
                    "(hung on the round-1 stack: NCCL 2.28 point-to-point under capture; never the default)")
    ap.add_argument("--exchange", action="store_true", help="multi-GPU: the round-1 form (row-lane kernels + NCCL point-to-point interface exchange) "
                    "instead of the owner-computes form (no exchange, energy all-reduce only)")
    ap.add_argument("--overlap", action="store_true", help="multi-GPU: interface elements first, exchange on a side stream under the "
                    "assembly of the rest (measured no faster than the plain order in round 1)")
    ap.add_argument("--flags", type=int, default=0, help="pfa_mesh_desc.flags (1 = keep the caller's element order, 2 = in-kernel zero fill, 8 = round-1 row-lane RED kernels)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = dict(CONFIGS[args.config])
    if args.n is not None:
        cfg["n"] = args.n
    if args.p is not None:
        cfg["p"] = args.p
    args.n, args.p = cfg["n"], cfg["p"]
    mode = cfg["mode"]

    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    from polyfem_b200 import capi
    from polyfem_b200 import dist as pdist
    from polyfem_b200.mesh import lame_from_E_nu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the assembly path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None
    if world > 1:
        # the interface rows go point-to-point to the neighbouring slab: let NCCL use more channels
        # than its 1-2 default for send/recv over NVSwitch
        if args.overlap:
            os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")  # NCCL's stream ahead of the assembly kernel's pending CTAs
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "16")
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
        dist.init_process_group("nccl", device_id=dev)

    mesh, x_host, t = build_workload(args.n, args.p)
    lam, mu = lame_from_E_nu(E_MOD, NU)
    exch = None
    if world > 1 and cfg["material"] != "NeoHookean":
        raise SystemExit("bench.py: multi-GPU runs exist for the NeoHookean configs (2, 3, 5); the linear assemblers are single-GPU here")
    owner_mode = world > 1 and not args.exchange
    if owner_mode:
        # owner-computes form: the library's partition, ghost elements with geometry, every rank writes the
        # finished columns / gradient entries of the nodes it owns; the only collective is the all-reduce of the energy
        part = pdist.partition_owner_computes(mesh, rank, world)
        h = pdist.owner_handle(part, t, lam, mu, device=local_rank, flags=args.flags)
    else:
        # round-1 form (--exchange, or one GPU): cuts on whole layers of cells (6 n^2 tets), partial sums of interface
        # columns go point-to-point to their owners
        part = pdist.partition_elements(mesh, rank, world, align=1 if args.no_align else 6 * args.n * args.n)
        h = capi.Handle(cfg["material"], part.conn, part.n_bases, t["weights"], t["grad"], vertices=part.vertices,
                        lam=lam, mu=mu, device=local_rank, n_ghost_elements=part.n_ghost_elements,
                        flags=args.flags | (capi.FLAG_ROW_LANE if world > 1 else 0), n_first_elements=part.n_interface_elements)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    if world > 1 and not owner_mode:
        exch = pdist.InterfaceExchange(h, part, rank, world, dev, grad_offset=h.nnz)

    size = h.size
    x_loc = np.ascontiguousarray(x_host.reshape(-1, 3)[part.l2g].reshape(-1))[: h.ndof] if size == 3 else np.zeros(h.ndof)
    xd = torch.from_numpy(x_loc).to(dev)
    e_d = torch.zeros(1, dtype=torch.float64, device=dev)
    # values[] and the gradient share one allocation so that the interface exchange packs both at once
    vg_d = torch.zeros(h.nnz + h.ndof, dtype=torch.float64, device=dev)
    v_d, g_d = vg_d[:h.nnz], vg_d[h.nnz:]

    # cfg 5: the mass matrix (assembled once, like State::build_mass_matrix) and the InertiaForm buffers. In the
    # owner-computes partition the mass handle takes own AND ghost elements as its elements, so that the columns of the
    # owned nodes are complete without an exchange; the energy of the inertia term is summed over owned dofs only.
    hm = m_d = xt_d = ei_d = gi_d = owned_dofs = None
    if mode == "euler":
        from polyfem_b200 import tables as T
        tm = T.reference_tables(args.p, T.quadrature_order(args.p, is_mass=True))
        hm = capi.Handle("Mass", part.conn, part.n_bases, tm["weights"], None, vertices=part.vertices if owner_mode else part.vertices,
                         device=local_rank, ref_vals=tm["val"], density=RHO) if (owner_mode or world == 1) else None
        if hm is None:
            raise SystemExit("bench.py --config 5 runs on one GPU or in the owner-computes multi-GPU form")
        hm.set_stream(torch.cuda.current_stream().cuda_stream)
        assert hm.nnz == h.nnz  # same connectivity, same pattern: H = dt^2 H_el + M is one axpy
        m_d = torch.zeros(h.nnz, dtype=torch.float64, device=dev)
        hm.linear_stiffness_raw(m_d)
        xt_d = (0.5 * xd).contiguous()  # x_tilde = x_prev + dt v_prev is the caller's (ImplicitEuler.cpp:13-16)
        ei_d = torch.zeros(1, dtype=torch.float64, device=dev)
        gi_d = torch.zeros(h.ndof, dtype=torch.float64, device=dev)
        if owner_mode:
            # weights 1 / 0 per dof (a multiply, not a boolean gather: masked indexing would synchronise the host every step)
            owned_dofs = torch.from_numpy(np.repeat(part.owned.astype(np.float64), 3)).to(dev)

    # configs whose outputs fit the 126 MB L2 (cfg 1, 2): flush L2 between timed steps
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if 8 * h.nnz < (200 << 20) else None

    def step():
        if mode == "linear":
            h.linear_stiffness_raw(v_d)
        elif mode == "euler":
            h.grad_hess_weighted_raw(xd, DT * DT, e_d, g_d, v_d)            # dt^2 (E_el, g_el, H_el)
            hm.inertia_raw(m_d, xd, xt_d, ei_d, gi_d)                        # 1/2 d^T M d, M d
            h.axpy(1.0, gi_d, g_d)
            h.axpy(1.0, m_d, v_d)                                           # H = dt^2 H_el + M
            if owner_mode:
                e_d.add_(0.5 * torch.dot((xd - xt_d) * owned_dofs, gi_d))
                dist.all_reduce(e_d)
            else:
                e_d.add_(ei_d)
        elif owner_mode:
            h.grad_hess_raw(xd, e_d, g_d, v_d)
            dist.all_reduce(e_d)
        elif exch is None:
            h.grad_hess_raw(xd, e_d, g_d, v_d)
        elif not args.overlap:
            h.grad_hess_raw(xd, e_d, g_d, v_d)
            exch.reduce_combined(e_d, vg_d)
        else:
            # interface elements first; their partial sums travel to the owners on a side stream
            # while the remaining elements are assembled
            h.grad_hess_part_raw(xd, e_d, g_d, v_d, 1)
            exch.start(g_d, v_d)
            h.grad_hess_part_raw(xd, e_d, g_d, v_d, 2)
            exch.finish(e_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # experimental (--graph): capture the dozen short launches of a multi-GPU step (memset, kernels,
    # pack, NCCL send/recv, unpack, all-reduce) once in a CUDA graph and replay it
    graph = None
    if world > 1 and args.graph:
        eager_step = step
        try:
            cap = torch.cuda.Stream(device=dev)
            h.set_stream(cap.cuda_stream)
            with torch.cuda.stream(cap):
                for _ in range(2):  # NCCL communicators and staging buffers are set up outside the capture
                    eager_step()
            cap.synchronize()
            dist.barrier()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=cap, capture_error_mode="thread_local"):  # NCCL's watchdog thread also calls CUDA
                eager_step()
            step = graph.replay
        except Exception as ex:  # capture not possible on this stack: measure the eager path and say so
            sys.stderr.write(f"[bench] CUDA graph capture failed on rank {rank} ({type(ex).__name__}: {ex}); eager steps\n")
            graph = None
            step = eager_step
            h.set_stream(torch.cuda.current_stream().cuda_stream)
        ok = torch.tensor([1 if graph is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and graph is not None:  # all ranks or none
            graph, step = None, eager_step
            h.set_stream(torch.cuda.current_stream().cuda_stream)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    if world > 1:
        # set-up, not warm-up: NCCL opens its point-to-point channels and the caching allocator its
        # blocks during the first exchanges
        for _ in range(10):
            step()
    barrier()
    for _ in range(args.warmup):
        step()
    barrier()
    h.profile_enable(True)
    launches0 = h.launch_count() + (exch.launches if exch else 0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if flush is None:
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
        ms_total = ev0.elapsed_time(ev1)
    else:
        # small configs: an L2 flush (256 MB write) before every step, outside the per-step event pairs
        pairs = []
        for _ in range(args.steps):
            flush.zero_()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            step()
            a1.record()
            pairs.append((a0, a1))
        barrier()
        ms_total = float(sum(a0.elapsed_time(a1) for a0, a1 in pairs))
    clocks = sampler.stop() if rank == 0 else None
    recs = h.profile_read()
    h.profile_enable(False)
    launches = h.launch_count() + (exch.launches if exch else 0) - launches0
    prof_steps = args.steps
    if graph is not None:
        # a replayed graph does not pass through the library: per-kernel times and the launch count
        # of one step come from a few eager steps after the timed region
        h.set_stream(torch.cuda.current_stream().cuda_stream)
        prof_steps = 3
        h.profile_enable(True)
        l0 = h.launch_count()
        for _ in range(prof_steps):
            eager_step()
        barrier()
        recs = h.profile_read()
        h.profile_enable(False)
        launches = (h.launch_count() - l0) // prof_steps * args.steps
    per_rank = None
    if world > 1:
        tt = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        # where a multi-GPU step goes, rank by rank (outside the timed region): assembly launches vs the
        # interface exchange, each bracketed by CUDA events on the launching stream, 5 plain steps
        try:
            if graph is None and not args.overlap:
                ka, kb, kc = [], [], []
                for _ in range(5):
                    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    e0.record()
                    h.grad_hess_raw(xd, e_d, g_d, v_d)
                    e1.record()
                    if owner_mode:
                        dist.all_reduce(e_d)
                    else:
                        exch.reduce_combined(e_d, vg_d)
                    e2.record()
                    ka.append((e0, e1))
                    kb.append((e1, e2))
                torch.cuda.synchronize()
                mine = torch.tensor([np.median([a.elapsed_time(b) for a, b in ka]), np.median([a.elapsed_time(b) for a, b in kb]),
                                     float(h.n_elements), float(part.n_ghost_elements if owner_mode else part.n_interface_elements),
                                     float(h.nnz if not owner_mode else 9 * int(np.diff(h.block_pattern()[0])[part.owned == 1].sum()))],
                                    dtype=torch.float64, device=dev)
                allr = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(allr, mine)
                per_rank = {"assembly_ms": [round(float(t[0]), 4) for t in allr], "exchange_ms": [round(float(t[1]), 4) for t in allr],
                            "elements": [int(t[2]) for t in allr],
                            ("ghost_elements" if owner_mode else "interface_elements"): [int(t[3]) for t in allr],
                            ("owned_nnz" if owner_mode else "nnz"): [int(t[4]) for t in allr],
                            "note": ("exchange_ms = the energy all-reduce (the only collective), including waiting for the slowest rank" if owner_mode
                                     else "exchange_ms includes waiting for the neighbours and the energy all-reduce")}
        except Exception as ex:  # diagnostics must never cost the bench line
            sys.stderr.write(f"[bench] per-rank breakdown skipped on rank {rank}: {type(ex).__name__}: {ex}\n")
            per_rank = None
    ms_step = ms_total / args.steps
    n_el_total = mesh.n_elements
    value = n_el_total / (ms_step * 1e-3)

    # dominant kernel: average launch duration from the library's own CUDA events (same stream)
    kern = [ms for (name, ms) in recs if "assemble" in name]
    fill = [ms for (name, ms) in recs if "zero_fill" in name]
    kern_ms = float(np.sum(kern)) / prof_steps if kern else float("nan")  # per step (two launches when the step is split)
    peak, peak_src = measured_peaks()
    linear = mode == "linear"
    if world == 1:
        b_alg = b_alg_bytes_per_element(h.n_loc, h.ndof, h.nnz, h.n_elements, linear, cfg["material"] == "Laplacian")
        achieved = b_alg * h.n_elements / (kern_ms * 1e-3) / 1e9
    else:
        # per GPU: the whole mesh's compulsory bytes per element (closed-form nnz of the Kuhn cube) times this
        # rank's share of the elements, over this rank's kernel time
        nn = args.n
        nnz_g = 9 * ((230 * nn ** 3 + 138 * nn ** 2 + 24 * nn + 1) if args.p == 2 else (15 * nn ** 3 + 21 * nn ** 2 + 9 * nn + 1))
        b_alg = b_alg_bytes_per_element(h.n_loc, 3 * mesh.n_bases, nnz_g, mesh.n_elements)
        achieved = b_alg * (mesh.n_elements / world) / (kern_ms * 1e-3) / 1e9
    kname = sorted({name for (name, ms) in recs if "assemble" in name})[0] if kern else None
    f_alg = f_alg_flops_per_element(cfg["material"], h.n_loc, h.n_qp)
    fp64_tflops = f_alg * (h.n_elements if world == 1 else mesh.n_elements / world) / (kern_ms * 1e-3) / 1e12
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(kname, h.n_elements, world), "kernel": kname,
                "kernel_ms": kern_ms,
                "zero_fill_ms": float(np.sum(fill)) / prof_steps if fill else 0.0,
                "algorithmic_bytes_per_element": b_alg, "peak_source": peak_src + " (of measured)",
                # the FP64 side of the same launch (SURVEY.md §8d F_alg; DFMA peak measured by tools/microbench5.cu): the P4
                # configs are specified FP64-bound, and the P2 headline sits above the FP64 ridge too (6.6 FLOP/B vs 5.6)
                "fp64": {"achieved_tflops": fp64_tflops, "peak_tflops": FP64_PEAK_TFLOPS, "frac": fp64_tflops / FP64_PEAK_TFLOPS,
                         "algorithmic_flops_per_element": f_alg}}

    # e2e through the C ABI with pinned HOST buffers (H2D x, D2H E + grad + values every step)
    e2e = None
    if args.e2e_steps <= 0 or 8 * h.nnz > (8 << 30):
        e2e = None  # (values[] beyond 8 GB: no pinned host copy of it here)
    elif world == 1 and mode != "euler":
        xh = torch.from_numpy(x_loc).pin_memory()
        eh = torch.zeros(1, dtype=torch.float64).pin_memory()
        gh = torch.zeros(h.ndof, dtype=torch.float64).pin_memory()
        vh = torch.zeros(h.nnz, dtype=torch.float64).pin_memory()

        def host_call():
            if linear:
                h.linear_stiffness_raw(vh.numpy())
            else:
                h.grad_hess_raw(xh.numpy(), eh.numpy(), gh.numpy(), vh.numpy())
        host_call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host_call()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        e2e = {"value": n_el_total / dt, "unit": UNIT, "h2d_bytes_per_step": 0 if linear else int(8 * h.ndof),
               "d2h_bytes_per_step": int(8 * h.nnz) if linear else int(8 * (1 + h.ndof + h.nnz)), "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "buffers": "pinned host"}
        if not linear:
            assert abs(float(eh[0]) - float(e_d.item())) <= 1e-9 * abs(float(e_d.item()))
    else:
        # multi-GPU (and the implicit-Euler step): host buffers per rank, device step + copies
        xh = torch.from_numpy(x_loc).pin_memory()
        gh = torch.zeros(h.ndof, dtype=torch.float64).pin_memory()
        vh = torch.zeros(h.nnz, dtype=torch.float64).pin_memory()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            xd.copy_(xh, non_blocking=True)
            step()
            gh.copy_(g_d, non_blocking=True)
            vh.copy_(v_d, non_blocking=True)
            _ = float(e_d.item())
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_el_total / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": int(8 * h.ndof),
               "d2h_bytes_per_step": int(8 * (1 + h.ndof + h.nnz)), "ms_per_step": float(tt.item()) * 1e3,
               "steps": args.e2e_steps, "buffers": "pinned host, per rank", "cpus_bound_per_rank": numa}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_sample = args.cpu_sample_n or cfg["cpu_n"]
        r = cpu_baseline_run(cfg, n_sample, 2, 1, threads, use_cache=False)
        cpu = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample_text(cfg, n_sample, r, False, 2)}

    if rank == 0:
        line = {
            "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, mesh, h, world, None),
            "parallelism": ((f"element partition x{world} (pfa_partition_create), owner-computes columns with ghost elements: no interface "
                             "exchange, NCCL all-reduce of the energy only") if owner_mode else
                            (f"element partition x{world}, interface exchange "
                             + ("under" if args.overlap else "after") + " the assembly, "
                             + ("step replayed from a CUDA graph" if graph is not None else "eager launches"))) if world > 1 else "single GPU",
            "nnz": int(h.nnz) if world == 1 else None,
            "nnz_per_s": (h.nnz / (ms_step * 1e-3)) if world == 1 else None,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "setup_seconds": h.setup_seconds(),
        }
        if per_rank is not None:
            line["per_rank"] = per_rank
        print(json.dumps(line))
    if world > 1:
        if graph is not None:
            # tearing the NCCL communicator down while a captured graph still references its kernels
            # hangs on this stack: release the graph, finish all GPU work and leave without the teardown
            graph = None
            step = None
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
