"""Integer orientation / ordering tables vs the golden vectors of the reference's own tests (bit-exact bar).

Sources: test/test_connection.cpp:14-131 (vertex_inds; extracted to tests/golden/vertex_inds.json by
tests/golden/extract_reference_vectors.py), test/test_Row_index.cpp:6-51, test/test_Hanging_vertex_matcher.cpp:20-96,
include/math.hpp:188-199 (stretched_ind), include/Kernel_connection.hpp:7-37."""
import json
import os

import numpy as np
import pytest

from hexed_b200 import tables as T
from hexed_b200.tables import Connection_direction

HERE = os.path.dirname(os.path.abspath(__file__))


def test_vertex_inds_golden():
    golden = json.load(open(os.path.join(HERE, "golden", "vertex_inds.json")))
    assert len(golden["cases"]) >= 14
    for case in golden["cases"]:
        d = Connection_direction(case["i_dim"], case["face_sign"])
        assert T.vertex_inds(case["n_dim"], d) == case["inds"], case


def test_row_index_golden():
    # test/test_Row_index.cpp:6-51, Row_index(3, 5, i_dim)
    q = lambda d, fq, node: T.row_qpoint(3, 5, d, fq, node)  # noqa: E731
    assert (q(0, 0, 0), q(0, 0, 1), q(0, 1, 0), q(0, 1, 1), q(0, 6, 0), q(0, 6, 1)) == (0, 25, 1, 26, 6, 31)
    assert (q(1, 0, 0), q(1, 0, 1), q(1, 1, 0), q(1, 1, 1), q(1, 6, 0), q(1, 6, 1)) == (0, 5, 1, 6, 26, 31)
    assert (q(2, 0, 0), q(2, 0, 1), q(2, 1, 0), q(2, 1, 1), q(2, 6, 0), q(2, 6, 1)) == (0, 1, 5, 6, 30, 31)
    # every row visits every quadrature point exactly once, for every dimension
    for nd, rs in [(1, 4), (2, 3), (3, 6)]:
        for d in range(nd):
            pts = sorted(T.row_qpoint(nd, rs, d, fq, k) for fq in range(rs**(nd - 1)) for k in range(rs))
            assert pts == list(range(rs**nd))


def test_connection_direction_flags():
    # include/Kernel_connection.hpp:19-36
    d = Connection_direction([0, 0], [1, 0])
    assert not d.flip_normal(0) and not d.flip_normal(1) and not d.flip_tangential() and not d.transpose()
    d = Connection_direction([0, 2], [0, 0])
    assert d.flip_normal(0) and not d.flip_normal(1) and not d.flip_tangential() and d.transpose()
    d = Connection_direction([2, 0], [1, 0])
    assert d.flip_tangential() and d.transpose()
    d = Connection_direction([1, 2], [1, 1])
    assert not d.flip_normal(0) and d.flip_normal(1) and not d.flip_tangential() and not d.transpose()
    assert T.cartesian_direction(1).as_list() == [1, 1, 1, 0]  # include/connection.hpp:30-38


def test_stretched_ind():
    # include/math.hpp:188-199
    assert [T.stretched_ind(3, i, [False, False]) for i in range(4)] == [0, 1, 2, 3]
    assert [T.stretched_ind(3, i, [True, False]) for i in range(4)] == [0, 1, 0, 1]
    assert [T.stretched_ind(3, i, [False, True]) for i in range(4)] == [0, 0, 1, 1]
    assert [T.stretched_ind(3, i, [True, True]) for i in range(4)] == [0, 0, 0, 0]
    assert [T.stretched_ind(2, i, [False, False]) for i in range(2)] == [0, 1]
    assert [T.stretched_ind(2, i, [True, False]) for i in range(2)] == [0, 0]


def test_hanging_vertex_matcher_golden():
    # test/test_Hanging_vertex_matcher.cpp:6-26 (2D)
    v = np.ones((2, 4))
    v[0, 2], v[0, 3], v[1, 2], v[1, 3] = 0.1, 0.4, 0.3, 0.2
    T.hanging_vertex_match(2, v, 0, True)
    assert np.allclose([v[0, 2], v[0, 3], v[1, 2], v[1, 3]], [0.1, 0.15, 0.15, 0.2], rtol=1e-14)
    # :27-52 (3D)
    v = np.ones((4, 8))
    v[0, 0], v[1, 1], v[2, 4], v[3, 5] = 0.2, 0.7, 0.9, 0.5
    T.hanging_vertex_match(3, v, 1, False)
    assert np.allclose([v[0, 0], v[0, 1], v[0, 4], v[0, 5], v[1, 1], v[2, 5], v[3, 0], v[3, 1]],
                       [0.2, 0.45, 0.55, 0.575, 0.7, 0.7, 0.575, 0.6], rtol=1e-14)
    # :53-75 (3D, stretched along dimension 0)
    v = np.ones((2, 8))
    v[0, 0], v[1, 1], v[0, 4], v[1, 5] = 0.2, 0.7, 0.9, 0.5
    T.hanging_vertex_match(3, v, 1, False, (True, False))
    assert np.allclose([v[0, 0], v[0, 1], v[0, 4], v[0, 5], v[1, 1], v[1, 0]], [0.2, 0.45, 0.9, 0.7, 0.7, 0.45], rtol=1e-14)
    # :76-95 (3D, stretched along dimension 1)
    v = np.ones((2, 8))
    v[0, 0], v[0, 1], v[1, 4], v[1, 5] = 0.2, 0.7, 0.9, 0.5
    T.hanging_vertex_match(3, v, 1, False, (False, True))
    assert np.allclose([v[0, 0], v[0, 1], v[0, 4], v[0, 5], v[1, 1], v[1, 5], v[1, 0]], [0.2, 0.7, 0.55, 0.6, 0.6, 0.5, 0.55], rtol=1e-14)


@pytest.mark.parametrize("nd", [2, 3])
def test_face_permutation_matches_vertex_table(nd):
    """the quadrature-point permutation table must be the row_size-node refinement of the (golden-pinned) vertex permutation:
    with row_size 2 and nodes at the vertices both describe the same reordering of face 1 (src/connection.cpp:6-28 vs
    include/Spatial.hpp:85-129)"""
    for d0 in range(nd):
        for d1 in range(nd):
            for s0 in range(2):
                for s1 in range(2):
                    d = Connection_direction([d0, d1], [s0, s1])
                    assert T.face_permutation(nd, 2, d).tolist() == T.face_vertex_inds(nd, d)
                    for rs in (3, 6):
                        p = T.face_permutation(nd, rs, d)
                        assert sorted(p.tolist()) == list(range(rs**(nd - 1)))  # a permutation


def test_refined_fine_order_conforming():
    # include/connection.hpp:249-267: without stretching the i-th fine connection pairs mortar face i with fine element perm[i]
    for d in [Connection_direction([0, 0], [1, 0]), Connection_direction([2, 0], [1, 0]), Connection_direction([1, 2], [0, 0])]:
        for reverse in (False, True):
            elem, mortar, cs = T.refined_fine_order(3, d, reverse, [False, False])
            perm = T.face_vertex_inds(3, d)
            if reverse:
                assert elem == list(range(4)) and mortar == perm
            else:
                assert mortar == list(range(4)) and elem == perm
            assert cs == [False, False]
