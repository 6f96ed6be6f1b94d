"""Import shim: the reference imports `graphviz` at module top (minidiff/utils.py:6) although
only its drawing helper uses it; graphviz is not installed in this image.  Test infrastructure."""


class Graph:
    def __init__(self, **kw):
        pass

    def node(self, *a, **k):
        pass

    def edge(self, *a, **k):
        pass


class Digraph(Graph):
    pass
