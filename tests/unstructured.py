"""TEST HELPER: an unstructured tetrahedral mesh (Delaunay triangulation of random points in the unit cube, slivers removed,
elements in random order) as a polyfem_b200.mesh.TetMesh with P1 or P2 nodes in the reference's local order
(vertices, then edge midpoints (0,1) (1,2) (2,0) (0,3) (1,3) (2,3); auto_p_bases.cpp:1437-1449). Vertex valences vary
(5 .. 40 incident tets), unlike the Kuhn cube the other tests use."""
import numpy as np

from polyfem_b200.mesh import TetMesh, first_touch_numbering

EDGES = [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]


def delaunay_mesh(n_points=120, p=1, seed=3):
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = rng.random((n_points, 3))
    tets = Delaunay(pts).simplices.astype(np.int64)
    v = pts[tets]
    vol = np.linalg.det(v[:, 1:] - v[:, :1]) / 6.0
    flip = vol < 0
    tets[flip] = tets[flip][:, [0, 2, 1, 3]]
    vol = np.abs(vol)
    # drop slivers (keeps the Jacobians well conditioned); the rest stays a valid conforming mesh of a sub-domain
    edge_len = np.linalg.norm(v[:, [1, 2, 3, 2, 3, 3]] - v[:, [0, 0, 0, 1, 1, 2]], axis=2).max(axis=1)
    tets = tets[vol > 0.02 * edge_len ** 3]
    tets = tets[rng.permutation(tets.shape[0])]
    if p == 1:
        keys = tets
        xyz_el = pts[tets]
    else:
        nv = n_points
        a = np.stack([tets[:, i] for i, j in EDGES], axis=1)
        b = np.stack([tets[:, j] for i, j in EDGES], axis=1)
        ekey = nv + np.minimum(a, b) * nv + np.maximum(a, b)  # unique id per edge, disjoint from the vertex ids
        keys = np.concatenate([tets, ekey], axis=1)
        xyz_el = np.concatenate([pts[tets], 0.5 * (pts[a] + pts[b])], axis=1)
    conn, _ = first_touch_numbering(keys)
    n_bases = int(conn.max()) + 1
    node_xyz = np.zeros((n_bases, 3))
    node_xyz[conn.reshape(-1)] = xyz_el.reshape(-1, 3)
    return TetMesh(p=p, n_cells=int(round(n_points ** (1 / 3))), conn=np.ascontiguousarray(conn, dtype=np.int32), vertices=np.ascontiguousarray(pts[tets]),
                   n_bases=n_bases, node_xyz=node_xyz)
