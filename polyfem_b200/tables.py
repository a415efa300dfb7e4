"""Reference-element data the hot path consumes: tet quadrature rules and P1..P4
Lagrange bases on the reference tetrahedron.

In PolyFEM these come from `basis::ElementBases` (built once per mesh, outside the hot
path) and reach the assembler as `ElementAssemblyValues`.
The C-ABI (`include/pfa.h`) therefore takes them as *inputs*; this module only exists so
that the synthetic benchmarks / tests can produce the same inputs without PolyFEM.

* Quadrature: `data/tet_quadrature.json` holds the rules of
  `autogen/auto_tetrahedron.ipp` (orders 1..8, weights already divided by 6 as
  `quadrature/TetQuadrature.cpp:54` does), dumped bit-exactly (hex floats) by
  `tools/make_golden.py` through `oracle/_ref`.
* Bases: derived here from first principles (Lagrange polynomials on the principal
  lattice, exact rational coefficients) — NOT transcribed from `autogen/auto_p_bases.cpp`.
  Only the *local node order* is taken from the reference (`auto_p_bases.cpp:1169-1175,
  1437-1449,1969-1991,3018-3055`), stored below as integer lattice coordinates.
  `tests/test_tables.py` pins values and gradients against the reference's own generated
  code (golden file + live `oracle/_ref`).
"""
from __future__ import annotations

import json
import os
from fractions import Fraction
from functools import lru_cache

import numpy as np

_DATA = os.path.join(os.path.dirname(__file__), "data", "tet_quadrature.json")

# Local node order of the reference P_p tetrahedron as lattice coordinates (x, y, z) * p.
# vertices; edges 0-1, 1-2, 2-0, 0-3, 1-3, 2-3; faces 012, 013, 123, 203; cell.
P_NODES_LATTICE = {
    1: [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)],
    2: [(0, 0, 0), (2, 0, 0), (0, 2, 0), (0, 0, 2),
        (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1)],
    3: [(0, 0, 0), (3, 0, 0), (0, 3, 0), (0, 0, 3),
        (1, 0, 0), (2, 0, 0), (2, 1, 0), (1, 2, 0), (0, 2, 0), (0, 1, 0),
        (0, 0, 1), (0, 0, 2), (2, 0, 1), (1, 0, 2), (0, 2, 1), (0, 1, 2),
        (1, 1, 0), (1, 0, 1), (1, 1, 1), (0, 1, 1)],
    4: [(0, 0, 0), (4, 0, 0), (0, 4, 0), (0, 0, 4),
        (1, 0, 0), (2, 0, 0), (3, 0, 0), (3, 1, 0), (2, 2, 0), (1, 3, 0),
        (0, 3, 0), (0, 2, 0), (0, 1, 0), (0, 0, 1), (0, 0, 2), (0, 0, 3),
        (3, 0, 1), (2, 0, 2), (1, 0, 3), (0, 3, 1), (0, 2, 2), (0, 1, 3),
        (1, 1, 0), (1, 2, 0), (2, 1, 0), (1, 0, 1), (1, 0, 2), (2, 0, 1),
        (2, 1, 1), (1, 1, 2), (1, 2, 1), (0, 2, 1), (0, 1, 2), (0, 1, 1),
        (1, 1, 1)],
}

N_LOC = {p: len(v) for p, v in P_NODES_LATTICE.items()}


def quadrature_order(p: int, is_mass: bool = False) -> int:
    """Simplex Lagrange rule of `AssemblerUtils::quadrature_order`
    (assembler/AssemblerUtils.cpp:201-245): mass -> 2p, otherwise max(2(p-1), 1)."""
    if is_mass:
        return max(2 * p, 1)
    return max(2 * (p - 1), 1)


@lru_cache(maxsize=None)
def _quadrature_db():
    with open(_DATA) as f:
        raw = json.load(f)
    out = {}
    for k, v in raw["orders"].items():
        pts = np.array([[float.fromhex(c) for c in row] for row in v["points"]], dtype=np.float64)
        w = np.array([float.fromhex(c) for c in v["weights"]], dtype=np.float64)
        out[int(k)] = (pts, w)
    return out


def tet_quadrature(order: int):
    """(points [n,3], weights [n]) with sum(weights) == 1/6 (TetQuadrature.cpp:43-55)."""
    db = _quadrature_db()
    if order not in db:
        raise ValueError(f"tet quadrature order {order} not tabulated (have {sorted(db)})")
    pts, w = db[order]
    return pts.copy(), w.copy()


def p_nodes(p: int) -> np.ndarray:
    """Reference-element node positions [n_loc, 3] (p_nodes_3d)."""
    return np.array(P_NODES_LATTICE[p], dtype=np.float64) / p


@lru_cache(maxsize=None)
def _lagrange_factors(p: int):
    """For each local node: barycentric multi-index m=(m0..m3) and 1/prod(m_v!)."""
    out = []
    for (i, j, k) in P_NODES_LATTICE[p]:
        m = (p - i - j - k, i, j, k)
        denom = 1
        for mv in m:
            for t in range(1, mv + 1):
                denom *= t
        out.append((m, Fraction(1, denom)))
    return out


def p_basis(p: int, pts: np.ndarray):
    """Values [n_pts, n_loc] and gradients [n_pts, n_loc, 3] of the P_p Lagrange basis.

    phi_m(lambda) = prod_v (1/m_v!) prod_{k<m_v} (p*lambda_v - k) with lambda = (1-x-y-z, x, y, z):
    equals 1 at lattice node m and 0 at every other node of the principal lattice.
    """
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
    lam = np.stack([1.0 - pts[:, 0] - pts[:, 1] - pts[:, 2], pts[:, 0], pts[:, 1], pts[:, 2]], axis=1)
    # d lambda_v / d(x,y,z)
    dlam = np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    n_pts = pts.shape[0]
    fac = _lagrange_factors(p)
    val = np.empty((n_pts, len(fac)))
    grad = np.empty((n_pts, len(fac), 3))
    for n, (m, c) in enumerate(fac):
        # per barycentric direction: product and derivative of the 1-D factor
        f = np.ones((n_pts, 4))
        df = np.zeros((n_pts, 4))
        for v in range(4):
            for k in range(m[v]):
                term = p * lam[:, v] - k
                df[:, v] = df[:, v] * term + f[:, v] * p
                f[:, v] = f[:, v] * term
        val[:, n] = float(c) * f.prod(axis=1)
        g = np.zeros((n_pts, 3))
        for v in range(4):
            others = np.ones(n_pts)
            for u in range(4):
                if u != v:
                    others = others * f[:, u]
            g += (df[:, v] * others)[:, None] * dlam[v][None, :]
        grad[:, n, :] = float(c) * g
    return val, grad


def reference_tables(p: int, order: int | None = None):
    """What a PolyFEM-side caller would read out of `ElementAssemblyValues` for one
    reference element: quadrature points/weights, basis values and reference gradients."""
    if order is None:
        order = quadrature_order(p)
    pts, w = tet_quadrature(order)
    val, grad = p_basis(p, pts)
    return {"p": p, "order": order, "points": pts, "weights": w, "val": val, "grad": grad}
