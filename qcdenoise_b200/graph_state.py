"""Random graph states for `simulate()`: a library of small connected graphs and a sampler that partitions the
qubits into library sub-graphs and joins them with random connector edges.  Same role and call surface as the
reference's GraphDB / GraphState (graph_data.py:274-390, graph_state.py:40-215).  NOT a copy of its data: the
reference ships the 45 local-unitary-inequivalent graphs of Hein et al.'s table; this library is every connected
graph on 2..7 vertices from the networkx graph atlas (a superset up to isomorphism), thinned per order.  Parity
fixtures that need the reference's exact graphs use tests/golden/graphs.json instead."""
import random

import networkx as nx
from networkx.algorithms.approximation import vertex_cover

__all__ = ["GraphData", "GraphDB", "GraphState"]


class GraphData:
    """Edge-list container: {name: {"G": edge tuples}}."""

    def __init__(self, data=None):
        self.data = data if data is not None else self._atlas()

    @staticmethod
    def _atlas(max_per_order=12):
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


def _partitions(n, lo, hi):
    """All multisets of integers in [lo, hi] summing to n."""
    def rec(rest, cap):
        if rest == 0:
            yield []
            return
        for k in range(min(cap, rest), lo - 1, -1):
            if rest - k == 0 or rest - k >= lo:
                for tail in rec(rest - k, k):
                    yield [k] + tail
    return list(rec(n, hi))


class GraphState:
    def __init__(self, graph_db, n_qubits=5, min_num_edges=2, max_num_edges=7, directed=False):
        if n_qubits < 2:
            raise AssertionError(f"n_qubits={n_qubits} < 2")
        self.graph_db = graph_db
        self.n_qubits = n_qubits
        self.all_graphs = graph_db.get_edge_sorted_graphs()
        avail = [k for k, lst in enumerate(self.all_graphs) if lst]
        self.largest_subgraph = min(max_num_edges, max(avail))
        if min_num_edges > n_qubits:
            min_num_edges = int((n_qubits // 2) - 1)
        self.smallest_subgraph = max(2, min_num_edges)
        self.directed = directed
        self.graph_combs = [p for p in _partitions(n_qubits, self.smallest_subgraph, self.largest_subgraph)
                            if all(self.all_graphs[k] for k in p)]
        if not self.graph_combs:
            raise ValueError("no combination of library sub-graphs adds up to n_qubits")

    def pick_subgraphs(self):
        comb = random.choice(self.graph_combs)
        return [random.choice(self.all_graphs[k]) for k in comb]

    def combine_subgraphs(self, sub_graphs):
        g = nx.DiGraph() if self.directed else nx.Graph()
        offset, blocks = 0, []
        for sg in sub_graphs:
            g.add_edges_from(((u + offset, v + offset) for u, v in sg.edges()), weight=1)
            blocks.append(range(offset, offset + sg.order()))
            offset += sg.order()
        for a, b in zip(blocks[:-1], blocks[1:]):       # one random connector edge between consecutive blocks
            g.add_edge(random.choice(a), random.choice(b), weight=1)
        return g

    def sample(self, min_vertex_cover=1, max_itr=10, graph_plot=False):
        g = None
        for _ in range(max(1, max_itr)):
            g = self.combine_subgraphs(self.pick_subgraphs())
            cover = vertex_cover.min_weighted_vertex_cover(g.to_undirected(), weight="weight")
            if len(cover) // 2 >= min_vertex_cover:
                break
        return g
