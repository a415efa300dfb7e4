"""Multi-GPU element partition and interface exchange.

One process per GPU. Elements are split into `world` contiguous blocks (slabs of the Kuhn
cube); a node is owned by the lowest rank that touches it, and each rank ends up with the
complete CSC columns (= rows, the pattern is symmetric) of the nodes it owns:

  * every rank assembles only its own elements (pfa_grad_hess on its local handle);
  * its local pattern is widened by *ghost elements* (elements of other ranks touching a node
    it owns, `pfa_mesh_desc.n_ghost_elements`) so owned columns have their full row set;
  * after the kernel, partial sums that landed in columns owned by another rank are packed and
    sent point-to-point to the owner, which adds them in (Hessian values and gradient entries);
    the scalar energy is all-reduced. No replicated matrix, no full-vector all-reduce.

`torch.distributed` is the plumbing (NCCL on GPUs, gloo in the CPU tests).

Round 2 adds the OWNER-COMPUTES form (`partition_owner_computes`, `owner_handle`), which is what bench.py uses: the
partition comes from the library (`pfa_partition_create`: contiguous element blocks cut so that the (element, node)
incidences of the owned nodes are balanced, owner = rank of the first element touching a node), every rank holds the ghost
elements around its owned nodes WITH their geometry (`PFA_FLAG_GHOST_GEOMETRY`) and the column-lane kernels produce the
finished columns and gradient entries of the owned nodes - there is no interface exchange at all, only the scalar energy
is all-reduced. The exchange form above stays for the row-lane kernels (`PFA_FLAG_ROW_LANE`) and the other materials.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .mesh import TetMesh, first_touch_numbering


@dataclass
class Partition:
    rank: int
    world: int
    conn: np.ndarray         # [n_own + n_ghost, n_loc] local node ids, own elements first
    n_own_elements: int
    n_ghost_elements: int
    vertices: np.ndarray     # [n_own, 4, 3]
    n_bases: int             # local nodes
    l2g: np.ndarray          # [n_bases] global node id of each local node
    owner: np.ndarray        # [n_bases] owning rank of each local node
    own_elements: np.ndarray  # global element ids of the own block (interface elements first)
    n_interface_elements: int = 0  # leading own elements that touch a node owned by another rank


def element_ranges(n_elements: int, world: int, align: int = 1):
    """Contiguous element blocks; with align > 1 the cuts fall on multiples of `align` elements
    (one layer of cells of a structured mesh), which keeps the interface one node plane thick."""
    if align <= 1 or n_elements % align != 0 or n_elements // align < world:
        return [(n_elements * r) // world for r in range(world + 1)]
    layers = n_elements // align
    return [((layers * r) // world) * align for r in range(world + 1)]


def partition_elements(mesh: TetMesh, rank: int, world: int, align: int = 1) -> Partition:
    ne = mesh.n_elements
    if world == 1:
        return Partition(0, 1, mesh.conn, ne, 0, mesh.vertices, mesh.n_bases,
                         np.arange(mesh.n_bases, dtype=np.int64), np.zeros(mesh.n_bases, dtype=np.int32),
                         np.arange(ne, dtype=np.int64))
    bounds = element_ranges(ne, world, align)
    elem_rank = np.searchsorted(np.asarray(bounds[1:]), np.arange(ne), side="right").astype(np.int32)
    node_owner = np.full(mesh.n_bases, world, dtype=np.int32)
    np.minimum.at(node_owner, mesh.conn.reshape(-1), np.repeat(elem_rank, mesh.n_loc))
    own = np.arange(bounds[rank], bounds[rank + 1], dtype=np.int64)
    # interface elements first: they are the only ones that write into columns / dofs of other
    # ranks, so they are assembled first and the exchange runs under the assembly of the rest
    is_iface = (node_owner[mesh.conn[own]] != rank).any(axis=1)
    own = np.concatenate([own[is_iface], own[~is_iface]])
    touches_owned = (node_owner[mesh.conn] == rank).any(axis=1) & (elem_rank != rank)
    ghost = np.nonzero(touches_owned)[0].astype(np.int64)
    elems = np.concatenate([own, ghost])
    conn_local, l2g = first_touch_numbering(mesh.conn[elems].astype(np.int64))
    return Partition(rank, world, np.ascontiguousarray(conn_local), int(own.size), int(ghost.size),
                     np.ascontiguousarray(mesh.vertices[own]), int(l2g.size), l2g.astype(np.int64),
                     node_owner[l2g].astype(np.int32), own, int(is_iface.sum()))


def block_pattern_numpy(conn: np.ndarray, n_bases: int):
    """Node-block pattern (adj_off, adj) from connectivity — same result as the library's
    pfa_block_pattern; used by the CPU (gloo) tests where no handle exists."""
    nl = conn.shape[1]
    b = np.repeat(conn.astype(np.int64), nl, axis=1).reshape(-1)   # column node, each repeated
    a = np.tile(conn.astype(np.int64), (1, nl)).reshape(-1)        # row node
    keys = np.unique(b * n_bases + a)
    cols = keys // n_bases
    adj = (keys % n_bases).astype(np.int32)
    adj_off = np.zeros(n_bases + 1, dtype=np.int32)
    np.add.at(adj_off, cols + 1, 1)
    return np.cumsum(adj_off).astype(np.int32), adj


def pair_value_indices(adj_off: np.ndarray, pair_idx: np.ndarray, pair_col: np.ndarray, size: int) -> np.ndarray:
    """values[] indices of the size x size scalar entries of node pairs `pair_idx` (positions in
    adj) lying in column nodes `pair_col`; order per pair: n (column dof) major, m minor."""
    off = adj_off[pair_col].astype(np.int64)
    deg = (adj_off[pair_col + 1] - adj_off[pair_col]).astype(np.int64)
    k = pair_idx.astype(np.int64) - off
    base = size * size * off + size * k
    n = np.arange(size, dtype=np.int64)
    idx = base[:, None, None] + n[None, :, None] * (size * deg)[:, None, None] + n[None, None, :]
    return idx.reshape(-1)


class InterfaceExchange:
    """Moves partial sums of non-owned columns / dofs to their owners and adds them in."""

    def __init__(self, handle, part: Partition, rank: int, world: int, device, size: int = 3,
                 block_pattern=None, grad_offset=None):
        """grad_offset: when values[] and the gradient live in ONE tensor (gradient starting at this
        element offset), reduce_combined() packs / adds both with a single gather / index_add."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.device, self.size = rank, world, device, size
        self.launches = 0
        self._side, self._done, self._send_bufs = None, None, []
        adj_off, adj = block_pattern if block_pattern is not None else handle.block_pattern()
        nb = part.n_bases
        conn_own = part.conn[:part.n_own_elements].astype(np.int64)
        nl = conn_own.shape[1]

        # pairs (row a, column b) produced by own elements whose column node b is owned elsewhere
        iface = (part.owner[conn_own] != rank).any(axis=1)
        c = conn_own[iface]
        b = np.repeat(c, nl, axis=1).reshape(-1)
        a = np.tile(c, (1, nl)).reshape(-1)
        keep = part.owner[b] != rank
        keys = np.unique(b[keep] * nb + a[keep])
        pb, pa = keys // nb, keys % nb
        # position of each such pair in my adjacency
        my_pair_key = np.repeat(np.arange(nb, dtype=np.int64), np.diff(adj_off)) * nb + adj.astype(np.int64)
        pidx = np.searchsorted(my_pair_key, keys)
        assert np.array_equal(my_pair_key[pidx], keys)
        dest = part.owner[pb]
        # dofs: non-owned nodes touched by own elements
        nodes = np.unique(c.reshape(-1))
        nodes = nodes[part.owner[nodes] != rank]
        ndest = part.owner[nodes]

        send_meta = []
        self.send_vidx, self.send_gidx = {}, {}
        for s in range(world):
            if s == rank:
                send_meta.append(None)
                continue
            sel = dest == s
            nsel = ndest == s
            if not sel.any() and not nsel.any():
                send_meta.append(None)
                continue
            send_meta.append((part.l2g[pb[sel]], part.l2g[pa[sel]], part.l2g[nodes[nsel]]))
            self.send_vidx[s] = torch.from_numpy(pair_value_indices(adj_off, pidx[sel], pb[sel], size)).to(device)
            gi = (nodes[nsel][:, None] * size + np.arange(size)[None, :]).reshape(-1)
            self.send_gidx[s] = torch.from_numpy(gi.astype(np.int64)).to(device)

        # tell every owner what it will receive, in the order it will arrive
        gathered = [None] * world
        dist.all_gather_object(gathered, send_meta)
        order = np.argsort(part.l2g, kind="stable")
        sorted_g = part.l2g[order]

        def to_local(gids):
            pos = np.searchsorted(sorted_g, gids)
            assert np.array_equal(sorted_g[pos], gids), "interface node missing on the owner"
            return order[pos]

        self.recv_vidx, self.recv_gidx = {}, {}
        for s in range(world):
            if s == rank or gathered[s] is None or gathered[s][rank] is None:
                continue
            gb, ga, gn = gathered[s][rank]
            lb, la = to_local(gb), to_local(ga)
            k = np.searchsorted(my_pair_key, lb * nb + la)
            assert np.array_equal(my_pair_key[k], lb * nb + la), "owner pattern lacks a received pair"
            self.recv_vidx[s] = torch.from_numpy(pair_value_indices(adj_off, k, lb, size)).to(device)
            ln = to_local(gn)
            gi = (ln[:, None] * size + np.arange(size)[None, :]).reshape(-1)
            self.recv_gidx[s] = torch.from_numpy(gi.astype(np.int64)).to(device)
        self.peers_send = sorted(self.send_vidx)
        self.peers_recv = sorted(self.recv_vidx)
        self.recv_buf = {s: torch.empty(self.recv_vidx[s].numel() + self.recv_gidx[s].numel(), dtype=torch.float64, device=device)
                         for s in self.peers_recv}
        self.send_idx, self.recv_idx = {}, {}
        if grad_offset is not None:
            for s in self.peers_send:
                self.send_idx[s] = torch.cat([self.send_vidx[s], self.send_gidx[s] + int(grad_offset)])
            for s in self.peers_recv:
                self.recv_idx[s] = torch.cat([self.recv_vidx[s], self.recv_gidx[s] + int(grad_offset)])
        self.interface_bytes = 8 * sum(self.send_vidx[s].numel() + self.send_gidx[s].numel() for s in self.peers_send)

    def _post(self, grad, values):
        torch, dist = self.torch, self.dist
        ops, self._send_bufs = [], []
        for s in self.peers_recv:
            ops.append(dist.P2POp(dist.irecv, self.recv_buf[s], s))
        for s in self.peers_send:
            buf = torch.cat([values[self.send_vidx[s]], grad[self.send_gidx[s]]])
            self._send_bufs.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, s))
        return dist.batch_isend_irecv(ops) if ops else []

    def _add_received(self, reqs, grad, values):
        for r in reqs:
            r.wait()
        for s in self.peers_recv:
            nv = self.recv_vidx[s].numel()
            values.index_add_(0, self.recv_vidx[s], self.recv_buf[s][:nv])
            grad.index_add_(0, self.recv_gidx[s], self.recv_buf[s][nv:])

    def reduce(self, energy, grad, values):
        """energy: all-reduce; grad / values: partial sums of non-owned entries go to the owner."""
        reqs = self._post(grad, values)
        if energy is not None:
            self.dist.all_reduce(energy)
        self._add_received(reqs, grad, values)

    def reduce_combined(self, energy, vg):
        """reduce() for values and gradient stored in one tensor `vg` (see grad_offset): one gather,
        one grouped send/recv, one index_add per peer."""
        dist = self.dist
        ops, self._send_bufs = [], []
        for s in self.peers_recv:
            ops.append(dist.P2POp(dist.irecv, self.recv_buf[s], s))
        for s in self.peers_send:
            buf = vg[self.send_idx[s]]
            self._send_bufs.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, s))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        if energy is not None:
            dist.all_reduce(energy)
        for r in reqs:
            r.wait()
        for s in self.peers_recv:
            vg.index_add_(0, self.recv_idx[s], self.recv_buf[s])

    # ---- overlapped form (CUDA): start() after the interface elements, finish() after the rest ----
    def start(self, grad, values):
        """Call when the interface elements (Partition.n_interface_elements, assembled first) are in
        grad / values: packs and sends their partial sums on a side stream while the caller goes on
        assembling the remaining elements on the current stream. Received sums are added with atomics
        (index_add_), which may run concurrently with the assembly kernel's own reductions."""
        torch = self.torch
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device, priority=-1)  # ahead of the assembly kernel's pending CTAs
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            reqs = self._post(grad, values)
            self._add_received(reqs, grad, values)
            self._done = torch.cuda.Event()
            self._done.record(self._side)

    def finish(self, energy):
        """Joins the side stream and all-reduces the energy (complete only after the last element)."""
        main = self.torch.cuda.current_stream(self.device)
        main.wait_event(self._done)
        if energy is not None:
            self.dist.all_reduce(energy)


# ---------------------------------------------------------------- owner-computes form (no interface exchange)
@dataclass
class OwnerPartition:
    rank: int
    world: int
    conn: np.ndarray          # [n_own + n_ghost, n_loc] local node ids, own elements first
    n_own_elements: int
    n_ghost_elements: int
    vertices: np.ndarray      # [n_own + n_ghost, 4, 3]: the ghost elements carry geometry
    n_bases: int              # local nodes
    l2g: np.ndarray           # [n_bases] global node id of each local node
    owned: np.ndarray         # [n_bases] uint8, 1 = this rank produces the node's columns / gradient entries
    elements: np.ndarray      # [n_own + n_ghost] global element ids


def partition_owner_computes(mesh: TetMesh, rank: int, world: int) -> OwnerPartition:
    """The library's partition (pfa_partition_create, include/pfa.h) applied to a TetMesh."""
    from . import capi
    p = capi.partition(mesh.conn, mesh.n_bases, world, rank)
    el = p["elements"].astype(np.int64)
    return OwnerPartition(rank, world, p["conn"], p["n_own"], p["n_ghost"], np.ascontiguousarray(mesh.vertices[el]),
                          int(p["l2g"].size), p["l2g"].astype(np.int64), p["owned"], el)


def owner_handle(part: OwnerPartition, tables_, lam, mu, device=0, material="NeoHookean", flags=0):
    """One pfa handle per rank: pfa_grad_hess / pfa_hessian write the finished columns of the owned nodes."""
    from . import capi
    lam = np.asarray(lam, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    if lam.ndim > 0:  # per-element arrays of the whole mesh: take this rank's elements
        lam, mu = lam[part.elements], mu[part.elements]
    return capi.Handle(material, part.conn, part.n_bases, tables_["weights"], tables_["grad"], vertices=part.vertices, lam=lam, mu=mu,
                       device=device, n_ghost_elements=part.n_ghost_elements, flags=flags | capi.FLAG_GHOST_GEOMETRY,
                       owned_nodes=part.owned)
