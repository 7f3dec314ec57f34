"""CPU oracle for the path-candidate graphs — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Pure-Python restatement (small cases only) of the Rust iterators the reference uses to enumerate path
candidates.  Citations are relative to ``/root/reference/differt-core/src/geometry/graph.rs``.
Only ``tests/`` may import this module.
"""

from __future__ import annotations

from collections import deque

import numpy as np


class DiGraph:
    """``DiGraph`` (graph.rs:596-602): ``edges_list[i]`` = sorted adjacent nodes of node ``i``."""

    def __init__(self, edges_list: list[list[int]]) -> None:
        self.edges_list = edges_list

    @classmethod
    def from_complete_graph(cls, num_nodes: int) -> "DiGraph":
        """graph.rs:1012-1026: every node is adjacent to every other node, ascending."""
        return cls([[j for j in range(num_nodes) if j != i] for i in range(num_nodes)])

    @property
    def num_nodes(self) -> int:
        return len(self.edges_list)

    def insert_from_and_to_nodes(self, direct_path=False, from_adjacency=None, to_adjacency=None):
        """graph.rs:636-691."""
        from_ = self.num_nodes
        to = from_ + 1
        for i, edges in enumerate(self.edges_list):
            if to_adjacency is None or to_adjacency[i]:
                edges.append(to)
        from_edges = [i for i in range(from_) if from_adjacency is None or from_adjacency[i]]
        if direct_path:
            from_edges.append(to)
        self.edges_list.append(from_edges)
        self.edges_list.append([])
        return from_, to

    def filter_by_mask(self, mask, fast_mode: bool = True) -> None:
        """graph.rs:879-915."""
        mask = np.asarray(mask, bool)
        if mask.size > self.num_nodes:
            raise ValueError("'mask' length must be smaller than or equal to the number of nodes")
        for i, keep in enumerate(mask):
            if not keep:
                self.edges_list[i] = []
        if not fast_mode:
            self.edges_list = [
                [n for n in edges if (mask[n] if n < mask.size else True)] for edges in self.edges_list
            ]

    def all_paths(self, from_: int, to: int, depth: int, include_from_and_to: bool = True):
        """``AllPathsFromDiGraphIter`` (graph.rs:1029-1108): DFS, children in list order."""
        stack = [deque(self.edges_list[from_])]
        visited = [from_]
        while stack:
            children = stack[-1]
            if len(visited) + 1 == depth:
                if to in children:  # binary_search on the sorted list
                    yield list(visited) + [to] if include_from_and_to else list(visited[1:])
                stack.pop()
                visited.pop()
            elif children:
                child = children.popleft()
                visited.append(child)
                stack.append(deque(self.edges_list[child]))
            else:
                stack.pop()
                visited.pop()


def hybrid_path_candidates(num_primitives, order, visible_from_tx, visible_from_rx, mask=None) -> np.ndarray:
    """The candidate list of ``HybridPathTracer.generate_path_candidates`` before the quad doubling
    (reference ``differt/src/differt/geometry/_solvers.py:1023-1051``)."""
    graph = DiGraph.from_complete_graph(num_primitives)
    from_, to = graph.insert_from_and_to_nodes(
        from_adjacency=np.asarray(visible_from_tx, bool), to_adjacency=np.asarray(visible_from_rx, bool)
    )
    if mask is not None:
        graph.filter_by_mask(mask, fast_mode=True)
    paths = list(graph.all_paths(from_, to, order + 2, include_from_and_to=False))
    return np.asarray(paths, dtype=np.int32).reshape(len(paths), order)
