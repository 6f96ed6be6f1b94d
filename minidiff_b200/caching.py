"""Graph-structure caching (reference minidiff/caching.py:11-65): inside `with reuse_graph():`
the backward traversal order of structurally identical graphs is computed once and replayed as
index paths into the graph's nested tensor list.  Pure host bookkeeping; no arithmetic.
"""
from __future__ import annotations

from contextvars import ContextVar

_caching_graph = ContextVar("caching_graph", default=False)
_cached_graph_indices = ContextVar("cached_indices", default=None)


class reuse_graph:
    def __enter__(self):
        self._tokens = (_caching_graph.set(True), _cached_graph_indices.set({}))

    def __exit__(self, *exc):
        _caching_graph.reset(self._tokens[0])
        _cached_graph_indices.set({})


def currently_caching() -> bool:
    return _caching_graph.get()


def backward_indices_for_root(root_node):
    """Index paths (into root_node._tensor_graph) of the tensors in backward-traversal order,
    memoised by the structural hash of the graph (caching.py:31-65)."""
    if not _caching_graph.get():
        raise ValueError("Not currently preserving graph")
    table = _cached_graph_indices.get()
    key = root_node.hash
    hit = table.get(key)
    if hit is not None:
        return hit
    ordered = root_node.toposort()
    if not ordered:
        return ()
    where = {id(t): -1 for t in ordered}
    pending = [([i], item) for i, item in enumerate(root_node._tensor_graph)]
    while pending:
        path, item = pending.pop()
        if isinstance(item, list):
            pending.extend((path + [i], sub) for i, sub in enumerate(item))
        elif id(item) in where:
            where[id(item)] = path
    paths = tuple(where[id(t)] for t in ordered)
    table[key] = paths
    return paths
