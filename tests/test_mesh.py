"""Synthetic mesh generator: sizes of SURVEY.md §8, orientation, conformity, numbering."""
import numpy as np
import pytest

from polyfem_b200 import mesh as M
from polyfem_b200 import tables


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_sizes_and_first_touch(p):
    n = 3
    m = M.kuhn_cube(n, p)
    assert m.n_elements == 6 * n ** 3 and m.n_bases == (p * n + 1) ** 3
    assert m.conn.dtype == np.int32 and m.conn.shape == (m.n_elements, tables.N_LOC[p])
    # first touch: the running maximum of the flattened connectivity grows by at most one
    flat = m.conn.reshape(-1)
    run = np.maximum.accumulate(flat)
    assert flat[0] == 0 and np.all(np.diff(run) <= 1) and run[-1] == m.n_bases - 1
    assert np.array_equal(m.conn[0], np.arange(tables.N_LOC[p]))


@pytest.mark.parametrize("jitter", [0.0, 0.2])
def test_positive_orientation_and_volume(jitter):
    m = M.kuhn_cube(4, 1, jitter=jitter)
    e = m.vertices[:, 1:, :] - m.vertices[:, :1, :]
    det = np.linalg.det(e)
    assert det.min() > 0
    assert abs(det.sum() / 6 - 1.0) < 1e-12


@pytest.mark.parametrize("p", [2, 3, 4])
def test_conforming_nodes(p):
    """A global node has one position no matter which element reports it."""
    m = M.kuhn_cube(2, p, jitter=0.2)
    ref = tables.p_nodes(p)
    xyz = m.vertices[:, None, 0, :] + np.einsum("lc,ecd->eld", ref, m.vertices[:, 1:, :] - m.vertices[:, None, 0, :])
    assert np.abs(xyz - m.node_xyz[m.conn]).max() < 1e-14


def test_config_sizes_of_survey_table():
    # cfg 1: n=20 P1 -> 48 000 tets, 9 261 nodes ; cfg 3 closed forms at n=69
    m = M.kuhn_cube(20, 1)
    assert (m.n_elements, m.n_bases) == (48000, 9261)
    assert 6 * 69 ** 3 == 1971054 and (2 * 69 + 1) ** 3 == 2685619
    assert 9 * (230 * 69 ** 3 + 138 * 69 ** 2 + 24 * 69 + 1) == 685941705
