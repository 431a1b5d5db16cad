"""Graph library for random graph states: the data format and the database of the reference (graph_data.py:39-390).

`GraphData.data` maps a name to a list of (u, v, weight) triples with 1-BASED vertex labels, exactly what the reference
stores, so a GraphData built for the reference works here unchanged and vice versa.  The default library,
`HeinGraphData`, is Table V of Hein et al. (quant-ph/0602096) as the reference lists it (graph_data.py:74-178); it is
kept as data in `data/hein_graphs.json` (written by tools/make_hein_table.py from the reference checkout, irregular
entries included: entry 9 repeats edge (1, 4) and skips vertex 5, so it loads as a 5-vertex star).

Vertex numbering matters downstream (gate order, histogram bit order), so `GraphDB` builds each graph the way the
reference does: `add_weighted_edges_from` on the 0-based triples, i.e. vertices in first-appearance order -- the order
`networkx.disjoint_union` relabels by when GraphState joins blocks."""
import json
import math
import os
import random
from dataclasses import dataclass
from typing import Dict, List, Tuple

import networkx as nx

__all__ = ["GraphData", "HeinGraphData", "partition_GraphData", "GraphDB", "atlas_graph_data"]


def offset(edge_list):
    """1-based (u, v, ..., w) -> 0-based (u, v, w)  (graph_data.py:25-36)."""
    return [(e[0] - 1, e[1] - 1, e[-1]) for e in edge_list]


@dataclass
class GraphData:
    """name -> [(u, v, weight), ...], 1-based vertices (graph_data.py:39-62)."""
    data: Dict[str, List[Tuple]] = None

    def __post_init__(self):
        if self.data is None:          # GraphData() = the default library, a private copy
            self.data = {k: list(v) for k, v in HeinGraphData.data.items()}

    def groupby_edges(self):
        """Regroup IN PLACE by the number of edges of an entry: {"<edges>": [entry, ...]} (graph_data.py:45-56)."""
        grouped = {}
        for itm in self.data.values():
            grouped.setdefault(str(len(itm)), []).append(itm)
        self.data = grouped

    def keys(self):
        return self.data.keys()

    def items(self):
        return self.data.items()


def _load_hein():
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "hein_graphs.json")) as f:
        raw = json.load(f)["graphs"]
    return GraphData(data={k: [(int(u), int(v), float(w)) for u, v, w in lst] for k, lst in raw.items()})


HeinGraphData = _load_hein()


def atlas_graph_data(max_per_order=12, max_order=7):
    """An alternative, larger library in the same format: the first `max_per_order` connected graphs of every order
    2..max_order of the networkx graph atlas (a superset of the Hein table up to isomorphism)."""
    data, per = {}, {}
    for g in nx.graph_atlas_g():
        k = g.number_of_nodes()
        if k < 2 or k > max_order or not nx.is_connected(g) or per.get(k, 0) >= max_per_order:
            continue
        per[k] = per.get(k, 0) + 1
        data[str(len(data) + 1)] = [(u + 1, v + 1, 1.0) for u, v in sorted(tuple(sorted(e)) for e in g.edges())]
    return GraphData(data=data)


def partition_GraphData(graph_data, ratio=0.2, not_empty=True):
    """Train / test split balanced over the number of edges per graph (graph_data.py:181-271); returns (train, test).
    Groups with more than two graphs are shuffled (`random.shuffle`) and the first ceil(len * ratio) go to the test
    side; a group whose train side would hold a single graph gives its test part to BOTH sides; groups of one or two
    graphs go to both sides when `not_empty`, else to one side by a fair coin.  `graph_data` is regrouped in place, as
    in the reference.  Entries of the results are renumbered from 0."""
    graph_data.groupby_edges()
    test_groups, train_groups = [], []
    for itm in graph_data.data.values():
        if len(itm) > 2:
            random.shuffle(itm)
            part = math.ceil(len(itm) * ratio)
            test, train = itm[:part], itm[part:]
            test_groups.append(test)
            train_groups.append(train if len(train) > 1 else test)
        elif not_empty:
            test_groups.append(itm)
            train_groups.append(itm)
        else:
            # the reference compares the LIST random.choices returns with a string, so this coin always lands on the
            # train side (graph_data.py:217-222); kept, draw included
            side = random.choices(["test", "train"], weights=[0.5, 0.5])
            (test_groups if side == "test" else train_groups).append(itm)

    def collapse(groups):
        # a group contributes all its graphs when it holds more than one, else its only graph (graph_data.py:230-246)
        out = []
        for items in groups:
            if len(items) > 1 and isinstance(items[0], list):
                out.extend(items)
            else:
                out.append(items[0])
        return GraphData(data=dict(enumerate(out)))

    return collapse(train_groups), collapse(test_groups)


class GraphDB(dict):
    """'1'..'N' -> {"G": nx.Graph | nx.DiGraph, "V": order, "LUclass": None, "2Color": None} (graph_data.py:274-307).
    Entry i holds the i-th graph of `graph_data` whatever its name, as the reference's zip over both dicts does."""

    def __init__(self, graph_data=None, directed=False):
        super().__init__()
        self.directed = directed
        self.graph_data = HeinGraphData if graph_data is None else graph_data
        for i, (_, triples) in enumerate(self.graph_data.items(), start=1):
            ent = {"G": None, "V": None, "LUclass": None, "2Color": None}
            if triples:
                g = nx.DiGraph() if directed else nx.Graph()
                g.add_weighted_edges_from(offset(triples))
                ent["G"], ent["V"] = g, g.order()
            self[str(i)] = ent

    @property
    def db(self):
        return self

    def get_edge_sorted_graphs(self):
        """{order: [graphs with that many vertices]} for every order from 1 up (graph_data.py:367-381; despite the
        name the key is the number of VERTICES).  The reference sizes the dict by the number of database entries and
        fails on a graph larger than that; here every order present gets a key."""
        graphs = [ent["G"] for ent in self.values() if ent["G"] is not None]
        top = max([len(self) - 1] + [g.order() for g in graphs])
        out = {k: [] for k in range(1, top + 1)}
        for g in graphs:
            out[g.order()].append(g)
        return out
