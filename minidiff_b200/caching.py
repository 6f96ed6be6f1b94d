"""Graph-structure caching: `with reuse_graph():` promises that the same graph is built again and
again, so the backward order of a graph is worked out once per STRUCTURE and replayed afterwards
(same contract as the reference's minidiff/caching.py:15-65: `reuse_graph`, `currently_caching`,
`backward_indices_for_root`).  Host bookkeeping only.  The device-side pay-off of the same promise is
`md.capture_graph` (graphs.py): the whole step as one CUDA-graph replay.

How it works here: every OpNode built while caching carries (a) a structural key -- nested tuples of
the forward functions' ids -- and (b) `_tensor_graph`, a nested list holding the tensors of its
sub-graph in construction order.  For a new key the tensors of `toposort()` are located inside that
nested list once and remembered as index paths; a later graph with an equal key is traversed by
following the remembered paths into ITS nested list, which yields its tensors in the same order
without sorting.
"""
from __future__ import annotations

from contextvars import ContextVar

_active = ContextVar("mdb_reuse_graph_active", default=False)
_paths_by_structure = ContextVar("mdb_reuse_graph_paths", default=None)


class reuse_graph:
    """Context manager; nesting keeps caching on until the outermost block exits."""

    def __enter__(self):
        self._restore = _active.set(True)
        _paths_by_structure.set({})
        return self

    def __exit__(self, *exc):
        _active.reset(self._restore)
        _paths_by_structure.set({})


def currently_caching() -> bool:
    return _active.get()


def _locate(nested, wanted):
    """Index path of every tensor whose id is in `wanted`, searching the nested list breadth-wise
    with an explicit work list (graphs can be deeper than the recursion limit).  A tensor that occurs
    more than once keeps the first path found; any of its paths leads to the same object."""
    found = {}
    work = [((), nested)]
    while work and len(found) < len(wanted):
        prefix, items = work.pop()
        for i, item in enumerate(items):
            if type(item) is list:
                work.append((prefix + (i,), item))
            else:
                key = id(item)
                if key in wanted and key not in found:
                    found[key] = prefix + (i,)
    return found


def backward_indices_for_root(root_node):
    """Paths (into `root_node._tensor_graph`) of the tensors below `root_node`, in the order
    `toposort()` gives them; computed once per graph structure."""
    if not _active.get():
        raise ValueError("Not currently preserving graph")
    memo = _paths_by_structure.get()
    structure = root_node.hash
    try:
        return memo[structure]
    except KeyError:
        pass
    tensors = root_node.toposort()
    if not tensors:
        return ()
    where = _locate(root_node._tensor_graph, {id(t) for t in tensors})
    memo[structure] = paths = tuple(where.get(id(t), -1) for t in tensors)
    return paths
