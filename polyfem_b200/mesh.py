"""Synthetic tetrahedral meshes for the benchmarks and tests.

Stands in for PolyFEM's mesh + FE-space build, which is outside the hot path: it produces
exactly what the assembler reads from `std::vector<basis::ElementBases>` — per element the
global basis index of every local basis (`bases[e].bases[j].global()[0].index`) and the P1
geometric nodes of `gbases[e]`.

* Geometry: unit cube, `n` cells per side, each cell split into 6 Kuhn tetrahedra, all
  positively oriented; optional interior-vertex jitter (geometry stays affine).
* Numbering: first touch in element order, local order vertices -> edges -> faces -> cell,
  like `tet_local_to_global` / `MeshNodes::node_id_from_primitive`
  (basis/LagrangeBasis3d.cpp:238-330, mesh/MeshNodes.cpp:159-180).
  Higher-order nodes are identified through their position on the global principal lattice,
  which makes the numbering conforming for every order.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np

from . import tables


@dataclass
class TetMesh:
    p: int                    # basis order
    n_cells: int              # cells per side
    conn: np.ndarray          # [n_el, n_loc] int32 global basis ids
    vertices: np.ndarray      # [n_el, 4, 3] float64 P1 geometric nodes per element
    n_bases: int              # number of global bases (nodes)
    node_xyz: np.ndarray      # [n_bases, 3] node positions (undeformed, for tests/partitioning)

    @property
    def n_elements(self) -> int:
        return self.conn.shape[0]

    @property
    def n_loc(self) -> int:
        return self.conn.shape[1]

    @property
    def h(self) -> float:
        return 1.0 / self.n_cells


def _kuhn_local_tets():
    """6 tets of the unit cell as integer corner offsets [6,4,3], positively oriented."""
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [np.zeros(3, dtype=np.int64)]
        for axis in perm:
            step = np.zeros(3, dtype=np.int64)
            step[axis] = 1
            v.append(v[-1] + step)
        t = np.stack(v)
        d = np.linalg.det((t[1:] - t[0]).astype(np.float64))
        if d < 0:
            t[[2, 3]] = t[[3, 2]]
        tets.append(t)
    return np.stack(tets)


def first_touch_numbering(keys: np.ndarray):
    """Number distinct keys in order of first appearance in the flattened array."""
    flat = keys.reshape(-1)
    uniq, first, inv = np.unique(flat, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv].reshape(keys.shape).astype(np.int32), uniq[order]


def kuhn_cube(n: int, p: int = 1, jitter: float = 0.0, seed: int = 12345) -> TetMesh:
    """Unit cube with n^3 cells x 6 tets and a P_p Lagrange space."""
    if p not in tables.P_NODES_LATTICE:
        raise ValueError(f"unsupported basis order {p}")
    local = _kuhn_local_tets()                                  # [6,4,3]
    ii, jj, kk = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    cells = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1)  # x slowest, z fastest
    corners = (cells[:, None, None, :] + local[None, :, :, :]).reshape(-1, 4, 3)  # [n_el,4,3] ints

    # vertex coordinates (+ optional jitter of interior grid vertices)
    m = n + 1
    grid = np.stack(np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij"), axis=-1).astype(np.float64) / n
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter / n, jitter / n, size=grid.shape)
        d[0, :, :] = d[-1, :, :] = 0.0
        d[:, 0, :] = d[:, -1, :] = 0.0
        d[:, :, 0] = d[:, :, -1] = 0.0
        grid = grid + d
    vertices = grid[corners[..., 0], corners[..., 1], corners[..., 2]]            # [n_el,4,3]

    # lattice position of every local node: p*v0 + i(v1-v0) + j(v2-v0) + k(v3-v0)
    lat = np.array(tables.P_NODES_LATTICE[p], dtype=np.int64)                    # [n_loc,3]
    v0 = corners[:, 0, :]
    edges = corners[:, 1:, :] - v0[:, None, :]                                    # [n_el,3,3]
    pos = p * v0[:, None, :] + np.einsum("lc,ecd->eld", lat, edges)              # [n_el,n_loc,3]
    side = n * p + 1
    keys = (pos[..., 0] * side + pos[..., 1]) * side + pos[..., 2]
    conn, uniq = first_touch_numbering(keys)
    n_bases = int(uniq.size)

    # undeformed node positions through the (possibly jittered) affine element maps
    ref = lat.astype(np.float64) / p                                             # [n_loc,3]
    xyz_el = vertices[:, None, 0, :] + np.einsum("lc,ecd->eld", ref, vertices[:, 1:, :] - vertices[:, None, 0, :])
    node_xyz = np.zeros((n_bases, 3))
    node_xyz[conn.reshape(-1)] = xyz_el.reshape(-1, 3)
    return TetMesh(p=p, n_cells=n, conn=np.ascontiguousarray(conn), vertices=np.ascontiguousarray(vertices),
                   n_bases=n_bases, node_xyz=node_xyz)


def random_displacement(mesh: TetMesh, dim: int = 3, scale: float = 0.05, seed: int = 42) -> np.ndarray:
    """x = scale * h * U(-1,1) i.i.d. per dof, node-major x[node*dim+d].
    (numpy MT19937 stream; the recipe's std::mt19937_64 is not reproduced bit for bit —
    the array itself is what both the oracle and the GPU path consume.)"""
    rng = np.random.Generator(np.random.MT19937(seed))
    return scale * mesh.h * rng.uniform(-1.0, 1.0, size=mesh.n_bases * dim)


def lame_from_E_nu(E: float, nu: float):
    """3D conversion of assembler/MatParams.cpp:11-22."""
    lam = (E * nu) / ((1.0 + nu) * (1.0 - 2.0 * nu))
    mu = E / (2.0 * (1.0 + nu))
    return lam, mu
