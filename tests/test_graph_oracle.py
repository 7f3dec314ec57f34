"""The graph oracle (pure-Python DFS restating differt-core's DiGraph iterators) against the
reference's own graph tests: differt-core/src/geometry/graph.rs:1338-1706 and
differt-core/tests/geometry/test_graph.py."""

from __future__ import annotations

import numpy as np
import pytest

from differt_b200 import scenes
from oracle.graph_oracle import DiGraph, hybrid_path_candidates


@pytest.mark.parametrize("n,depth", [(1, 3), (2, 3), (3, 3), (5, 4), (4, 5), (6, 2)])
def test_complete_graph_and_di_graph_agree(n, depth):  # graph.rs:1566-1580
    g = DiGraph.from_complete_graph(n)
    from_, to = g.insert_from_and_to_nodes()
    paths = list(g.all_paths(from_, to, depth, include_from_and_to=False))
    if depth == 2:
        assert paths == []  # no direct path in the DiGraph unless direct_path=True
        return
    exp = scenes.complete_graph_candidates(n, depth - 2)
    np.testing.assert_array_equal(np.array(paths, np.int32).reshape(len(paths), depth - 2), exp)


def test_insert_from_and_to_nodes_ids():  # test_graph.py:13-25
    g = DiGraph.from_complete_graph(5)
    assert g.insert_from_and_to_nodes() == (5, 6)
    assert g.insert_from_and_to_nodes(direct_path=True) == (7, 8)
    assert g.insert_from_and_to_nodes(direct_path=False) == (9, 10)


def test_direct_path():
    g = DiGraph.from_complete_graph(3)
    from_, to = g.insert_from_and_to_nodes(direct_path=True)
    assert list(g.all_paths(from_, to, 2)) == [[from_, to]]
    assert list(g.all_paths(from_, to, 2, include_from_and_to=False)) == [[]]


@pytest.mark.parametrize("fast_mode", [True, False])
def test_filter_by_mask(fast_mode):  # test_graph.py:51-82
    g = DiGraph.from_complete_graph(8)
    from_, to = g.insert_from_and_to_nodes()
    before = len(list(g.all_paths(from_, to, 3, include_from_and_to=False)))
    g.filter_by_mask(np.array([True, False] * 4), fast_mode=fast_mode)
    after = list(g.all_paths(from_, to, 3, include_from_and_to=False))
    assert 0 < len(after) < before
    assert all(node not in path for path in after for node in (1, 3, 5, 7))
    assert after == [[0], [2], [4], [6]]


@pytest.mark.parametrize("fast_mode", [True, False])
def test_filter_by_mask_all_disconnected(fast_mode):  # test_graph.py:84-95
    g = DiGraph.from_complete_graph(4)
    from_, to = g.insert_from_and_to_nodes(direct_path=False)
    g.filter_by_mask(np.zeros(4, bool), fast_mode=fast_mode)
    assert list(g.all_paths(from_, to, 4, include_from_and_to=False)) == []


def test_filter_by_mask_wrong_size():  # test_graph.py:97-121
    DiGraph.from_complete_graph(5).filter_by_mask(np.array([True, False, True]))
    with pytest.raises(ValueError):
        DiGraph.from_complete_graph(5).filter_by_mask(np.array([True, False, True, False, False, False]))


@pytest.mark.parametrize("fast_mode", [True, False])
def test_masked_graph_equals_complete_graph_of_the_kept_nodes(fast_mode):  # test_graph.py:123-150
    g = DiGraph.from_complete_graph(6)
    from_, to = g.insert_from_and_to_nodes()
    g.filter_by_mask(np.array([True, True, True, False, False, False]), fast_mode=fast_mode)
    for order in range(3):
        paths = list(g.all_paths(from_, to, order + 2, include_from_and_to=False))
        exp = scenes.complete_graph_candidates(3, order)
        np.testing.assert_array_equal(np.array(paths, np.int32).reshape(len(paths), order),
                                      exp if order > 0 else exp[:0])


def test_paths_are_sorted_and_respect_adjacency():  # graph.rs:1539-1552, 1442-1512
    rng = np.random.default_rng(7)
    n = 7
    a, b, m = rng.uniform(size=n) < 0.6, rng.uniform(size=n) < 0.6, rng.uniform(size=n) < 0.8
    for order in (1, 2, 3):
        c = hybrid_path_candidates(n, order, a, b, m)
        assert c.shape[1] == order
        assert [tuple(x) for x in c] == sorted(tuple(x) for x in c)
        assert len({tuple(x) for x in c}) == c.shape[0]
        if c.shape[0]:
            assert a[c[:, 0]].all() and b[c[:, -1]].all() and m[c].all()
            assert (c[:, 1:] != c[:, :-1]).all()
        # brute force: every tuple satisfying the constraints is there
        grid = np.stack(np.meshgrid(*[np.arange(n)] * order, indexing="ij"), -1).reshape(-1, order)
        ok = a[grid[:, 0]] & b[grid[:, -1]] & m[grid].all(-1) & (grid[:, 1:] != grid[:, :-1]).all(-1)
        np.testing.assert_array_equal(c, grid[ok].astype(np.int32))


# reference known-answer table: differt/tests/geometry/test_utils.py:448-490
_ENUMERATION_KATS = [
    (0, 0, np.empty((1, 0), int)),
    (8, 0, np.empty((1, 0), int)),
    (0, 5, np.empty((0, 5), int)),
    (3, 1, np.array([[0], [1], [2]])),
    (3, 2, np.array([[0, 1], [0, 2], [1, 0], [1, 2], [2, 0], [2, 1]])),
    (3, 3, np.array([[0, 1, 0], [0, 1, 2], [0, 2, 0], [0, 2, 1], [1, 0, 1], [1, 0, 2],
                     [1, 2, 0], [1, 2, 1], [2, 0, 1], [2, 0, 2], [2, 1, 0], [2, 1, 2]])),
]


@pytest.mark.parametrize("num_primitives,order,expected", _ENUMERATION_KATS)
def test_complete_graph_enumeration_known_answers(num_primitives, order, expected):
    got = scenes.complete_graph_candidates(num_primitives, order)
    assert got.shape == expected.shape and got.dtype == np.int32
    np.testing.assert_array_equal(got, expected)  # our decode is already in the sorted order
    assert scenes.num_complete_graph_candidates(num_primitives, order) == expected.shape[0]
    # chunks of the linear index reproduce the same rows (reference idiom: chunk_size, graph.rs:64-116)
    parts = [scenes.complete_graph_candidates(num_primitives, order, s, 5) for s in range(0, expected.shape[0], 5)]
    if parts:
        np.testing.assert_array_equal(np.concatenate(parts), expected)


@pytest.mark.parametrize("n,depth,count", [(10, 3, 10), (10, 4, 90), (5, 5, 80), (1, 3, 1), (1, 4, 0), (0, 4, 0)])
def test_complete_graph_counts(n, depth, count):  # n (n-1)^(k-1), graph.rs:356-362 / tests :1338-1378
    assert scenes.num_complete_graph_candidates(n, depth - 2) == count
    assert scenes.complete_graph_candidates(n, depth - 2).shape[0] == count
