"""Random graph states for `simulate()`: a sampler that partitions the qubits into library sub-graphs and joins them
with random connector edges -- the reference's GraphState (graph_state.py:40-215) over the reference's graph library
(graph_data.py, here `qcdenoise_b200.graph_data`).  `sample()` makes the same `random` / `np.random` calls in the same
order over the same library and the same list of combinations, so under the same seeds it returns the reference's
graphs edge for edge (tests/golden/graphs.json was written by the reference's own code).  `sample_batch()` draws whole
batches natively (C ABI qcd_sample_graph_states, Philox streams): the same distribution, checked against statistics
of the reference's sampler (tests/golden/graph_sampling_stats.json)."""
import ctypes
import random

import networkx as nx
import numpy as np
from networkx.algorithms.approximation import vertex_cover

from . import _lib
from .graph_data import GraphData, GraphDB, HeinGraphData, atlas_graph_data, partition_GraphData  # noqa: F401

__all__ = ["GraphData", "GraphDB", "GraphState", "HeinGraphData", "partition_GraphData", "atlas_graph_data"]


def _partitions(n, start=2):
    """Non-decreasing tuples of integers >= start that add up to n (the reference's `partitions`,
    graph_state.py:23-37): (n,) itself, then every (i,) + partition of n - i into parts >= i."""
    out = [(n,)]
    for i in range(start, n // 2 + 1):
        out += [(i,) + p for p in _partitions(n - i, i)]
    return out


class GraphState:
    """Random graph states: a random combination of library sub-graphs whose orders add up to n_qubits, joined block
    after block, screened on the size of an approximate minimum vertex cover (graph_state.py:40-215)."""

    def __init__(self, graph_db, n_qubits=5, min_num_edges=2, max_num_edges=7, directed=False):
        if n_qubits < 2:
            raise AssertionError(f"n_qubits={n_qubits} < 2")
        self.graph_db = graph_db
        self.n_qubits = n_qubits
        self.all_graphs = graph_db.get_edge_sorted_graphs()         # {order: [graphs]}
        avail = [k for k, lst in self.all_graphs.items() if lst]
        self.largest_subgraph = max(avail) if max_num_edges is None else min(max_num_edges, max(avail))
        if min_num_edges > n_qubits:
            min_num_edges = int((n_qubits // 2) - 1)
        self.smallest_subgraph = max(2, min_num_edges)
        self.directed = directed
        # the reference's list, in the reference's order: list(set(partitions)) filtered on the part sizes
        # (graph_state.py:145-172; the set's iteration order decides which combination a draw selects).  Combinations
        # naming an order the library has no graph of are dropped (the reference would fail on drawing one).
        self.graph_combs = [c for c in list(set(_partitions(n_qubits, self.smallest_subgraph)))
                            if all(self.smallest_subgraph <= k <= self.largest_subgraph and self.all_graphs.get(k)
                                   for k in c)]
        if not self.graph_combs:
            raise ValueError("no combination of library sub-graphs adds up to n_qubits")
        self._tables = None

    def pick_subgraphs(self):
        """A random combination, then a random library graph per order (graph_state.py:200-215)."""
        comb = self.graph_combs[random.randint(0, len(self.graph_combs) - 1)]
        subs = []
        for k in comb:
            pool = self.all_graphs[k]
            subs.append(pool[random.randint(0, len(pool) - 1)])
        return subs

    def combine_subgraphs(self, sub_graphs):
        """Disjoint union, block after block.  Every block after the first is tied to ONE random vertex of the graph
        so far by order(block) draws among the block's first order(block) - 1 vertices; repeated draws collapse
        (graph_state.py:176-198, including its np.random.randint upper bound)."""
        g = nx.DiGraph() if self.directed else nx.Graph()
        for sg in sub_graphs:
            have = g.order()
            if have > 1:
                first = random.randint(0, have - 1)
                targets = np.random.randint(low=have, high=have + sg.order() - 1, size=sg.order())
                g = nx.disjoint_union(g, sg)
                g.add_weighted_edges_from((first, int(t), 1.0) for t in targets)
            else:
                g = nx.disjoint_union(g, sg)
        return g

    def sample(self, min_vertex_cover=1, max_itr=10, graph_plot=False):
        g = None
        for _ in range(max(1, max_itr)):
            g = self.combine_subgraphs(self.pick_subgraphs())
            cover = vertex_cover.min_weighted_vertex_cover(g, weight="weight")
            if len(cover) // 2 >= min_vertex_cover:
                break
        return g

    # ---- whole batches, natively (qcd_sample_graph_states)
    def _library_tables(self):
        """The library and the combinations as the flat arrays the C ABI takes; vertices are numbered in node-iteration
        order because that is how networkx's disjoint_union relabels a block."""
        if self._tables is None:
            lib_edges, lib_off, lib_orders, order_off, order_graphs = [], [0], [], [0], []
            for k in range(max(self.all_graphs) + 1):
                for sg in self.all_graphs.get(k, []):
                    idx = {v: i for i, v in enumerate(sg.nodes())}
                    lib_edges += [(idx[u], idx[v]) for u, v in sg.edges()]
                    lib_off.append(len(lib_edges))
                    lib_orders.append(sg.order())
                    order_graphs.append(len(lib_orders) - 1)
                order_off.append(len(order_graphs))
            comb_sizes = [k for c in self.graph_combs for k in c]
            comb_off = np.cumsum([0] + [len(c) for c in self.graph_combs])
            self._tables = (np.asarray(lib_edges, dtype=np.uint8).reshape(-1, 2), np.asarray(lib_off, dtype=np.uint32),
                            np.asarray(lib_orders, dtype=np.uint8), np.asarray(order_off, dtype=np.uint32),
                            np.asarray(order_graphs, dtype=np.uint32), np.asarray(comb_sizes, dtype=np.uint8),
                            np.asarray(comb_off, dtype=np.uint32))
        return self._tables

    def sample_batch(self, n_graphs, min_vertex_cover=1, max_itr=10, seed=None, first_graph=0):
        """n_graphs graph states at once -> (edges uint8 [n_graphs, n (n - 1) / 2, 2], n_edges uint32 [n_graphs]);
        graph i's edges are edges[i, :n_edges[i]], in the order `graph.edges()` would list them.  Same algorithm as
        sample(); the draws come from Philox4x32-10 keyed by `seed` (default: 63 bits from `random`), counter = graph
        index, so any split of a batch over ranks (first_graph) gives the same graphs."""
        if self.directed:
            raise NotImplementedError("sample_batch builds undirected graphs; use sample() for directed ones")
        lib = _lib.load()
        le, lo, lord, oo, og, cs, co = self._library_tables()
        if seed is None:
            seed = random.getrandbits(63)
        n = self.n_qubits
        max_edges = max(1, n * (n - 1) // 2)
        edges = np.zeros((n_graphs, max_edges, 2), dtype=np.uint8)
        n_edges = np.zeros(n_graphs, dtype=np.uint32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)       # noqa: E731
        rc = lib.qcd_sample_graph_states(p(le), p(lo), p(lord), len(lord), p(oo), p(og), len(oo) - 2, p(cs), p(co),
                                         len(self.graph_combs), n, n_graphs, min_vertex_cover, max_itr, seed,
                                         first_graph, p(edges), max_edges, p(n_edges))
        if rc:
            raise _lib.QcdError(rc, "qcd_sample_graph_states: invalid graph library or combination table")
        return edges, n_edges
