"""Random graph states for `simulate()`: a library of small connected graphs and a sampler that partitions the
qubits into library sub-graphs and joins them with random connector edges.  Same role and call surface as the
reference's GraphDB / GraphState (graph_data.py:274-390, graph_state.py:40-215).  NOT a copy of its data: the
reference ships the 45 local-unitary-inequivalent graphs of Hein et al.'s table; this library is every connected
graph on 2..7 vertices from the networkx graph atlas (a superset up to isomorphism), thinned per order.  Parity
fixtures that need the reference's exact graphs use tests/golden/graphs.json instead."""
import ctypes
import random

import networkx as nx
import numpy as np
from networkx.algorithms.approximation import vertex_cover

from . import _lib

__all__ = ["GraphData", "GraphDB", "GraphState"]


class GraphData:
    """Edge-list container: {name: {"G": edge tuples}}."""

    def __init__(self, data=None):
        self.data = data if data is not None else self._atlas()

    _atlas_cache = {}

    @staticmethod
    def _atlas(max_per_order=12):
        if max_per_order not in GraphData._atlas_cache:
            GraphData._atlas_cache[max_per_order] = GraphData._scan_atlas(max_per_order)
        return {k: {"G": list(v["G"])} for k, v in GraphData._atlas_cache[max_per_order].items()}

    @staticmethod
    def _scan_atlas(max_per_order):
        out, per = {}, {}
        for g in nx.graph_atlas_g():
            k = g.number_of_nodes()
            if k < 2 or k > 7 or not nx.is_connected(g) or per.get(k, 0) >= max_per_order:
                continue
            # spread the picks over the edge-count range instead of taking the sparsest ones only
            per[k] = per.get(k, 0) + 1
            out[f"{k}_{per[k]}"] = {"G": sorted(tuple(sorted(e)) for e in g.edges())}
        return out


class GraphDB(dict):
    """name -> {"G": nx.Graph | nx.DiGraph, "LUclass": None, "2Color": None}."""

    def __init__(self, graph_data=None, directed=False):
        super().__init__()
        gd = graph_data if graph_data is not None else GraphData()
        self.directed = directed
        for name, ent in gd.data.items():
            g = nx.DiGraph() if directed else nx.Graph()
            g.add_edges_from(ent["G"], weight=1)
            self[name] = {"G": nx.convert_node_labels_to_integers(g, ordering="sorted"), "LUclass": None,
                          "2Color": None}

    def get_edge_sorted_graphs(self):
        """Lists of graphs indexed by node count (index k -> graphs with k nodes), up to the largest order."""
        top = max(ent["G"].order() for ent in self.values())
        out = [[] for _ in range(top + 3)]
        for ent in self.values():
            out[ent["G"].order()].append(ent["G"])
        return out


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
        self.all_graphs = graph_db.get_edge_sorted_graphs()
        avail = [k for k, lst in enumerate(self.all_graphs) if lst]
        self.largest_subgraph = max(avail) if max_num_edges is None else min(max_num_edges, max(avail))
        if min_num_edges > n_qubits:
            min_num_edges = int((n_qubits // 2) - 1)
        self.smallest_subgraph = max(2, min_num_edges)
        self.directed = directed
        self.graph_combs = sorted(c for c in set(_partitions(n_qubits, self.smallest_subgraph))
                                  if all(self.smallest_subgraph <= k <= self.largest_subgraph and self.all_graphs[k]
                                         for k in c))
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
            for k, pool in enumerate(self.all_graphs):
                for sg in pool:
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
