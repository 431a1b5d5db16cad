#!/usr/bin/env python
"""Writes qcdenoise_b200/data/hein_graphs.json: the graph library the reference ships as a constant
(HeinGraphData, qcdenoise/graph_data.py:74-178 -- Table V of Hein et al., quant-ph/0602096), read from the reference
checkout by file path and stored verbatim: 1-based vertex labels, (u, v, weight) triples, INCLUDING the table's own
irregularities (entry 9 lists edge (1, 4) twice and skips vertex 5, so it loads as a 5-vertex star).

Also writes tests/golden/graph_sampling_stats.json: statistics of >= 10^4 graphs sampled with the reference's OWN
GraphDB() / GraphState.sample() per configuration -- the fixture the native batch sampler is tested against on machines
without the reference checkout.

    python tools/make_hein_table.py [/root/reference]
"""
import collections
import importlib.util
import json
import os
import random
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIGS = [  # (n_qubits, min_num_edges, max_num_edges, min_vertex_cover, max_itr, samples)
    (4, 2, 7, 1, 1, 20000), (8, 2, 7, 1, 1, 20000), (8, 2, 7, 3, 10, 10000), (9, 3, 5, 1, 1, 10000),
    (14, 2, 7, 1, 1, 10000), (20, 2, 7, 5, 10, 10000)]


def load_reference(ref):
    pkg = "_qcdref"
    root = types.ModuleType(pkg)
    root.__path__ = [os.path.join(ref, "qcdenoise")]
    sys.modules[pkg] = root
    mods = {}
    for name in ("config", "graph_data", "graph_state"):
        spec = importlib.util.spec_from_file_location(f"{pkg}.{name}", os.path.join(ref, "qcdenoise", name + ".py"))
        mods[name] = importlib.util.module_from_spec(spec)
        sys.modules[f"{pkg}.{name}"] = mods[name]
        spec.loader.exec_module(mods[name])
    return mods["graph_data"], mods["graph_state"]


def graph_stats(graphs, n):
    """Edge-count histogram, degree-sequence histogram of vertex 0 and the largest vertex, and how often each pair (u, v)
    is an edge -- enough to catch a wrong join rule, a wrong library or a wrong vertex numbering."""
    edges = collections.Counter(g.number_of_edges() for g in graphs)
    pair = np.zeros((n, n))
    for g in graphs:
        for u, v in g.edges():
            pair[min(u, v), max(u, v)] += 1
    return {"samples": len(graphs), "edge_count_hist": {str(k): v for k, v in sorted(edges.items())},
            "pair_frequency": (pair / len(graphs)).round(6).tolist()}


def main(ref):
    gd, gs = load_reference(ref)
    table = {k: [[int(u), int(v), float(w)] for u, v, w in lst] for k, lst in gd.HeinGraphData.data.items()}
    os.makedirs(os.path.join(ROOT, "qcdenoise_b200", "data"), exist_ok=True)
    with open(os.path.join(ROOT, "qcdenoise_b200", "data", "hein_graphs.json"), "w") as f:
        json.dump({"source": "Hein et al., quant-ph/0602096, Table V, as listed in qcdenoise/graph_data.py:74-178 "
                             "(keys = graph numbers; 1-based vertices; [u, v, weight])", "graphs": table}, f,
                  separators=(",", ":"))
    out = {"how": "tools/make_hein_table.py: reference GraphDB() + GraphState(db, n, min_e, max_e).sample(cover, itr); "
                  "random.seed(77 + n); np.random.seed(77 + n)", "configs": []}
    for n, lo, hi, cover, itr, count in CONFIGS:
        random.seed(77 + n)
        np.random.seed(77 + n)
        state = gs.GraphState(gd.GraphDB(), n_qubits=n, min_num_edges=lo, max_num_edges=hi)
        graphs = [state.sample(min_vertex_cover=cover, max_itr=itr) for _ in range(count)]
        st = graph_stats(graphs, n)
        st.update({"n_qubits": n, "min_num_edges": lo, "max_num_edges": hi, "min_vertex_cover": cover, "max_itr": itr,
                   "graph_combs": [list(c) for c in state.graph_combs]})
        out["configs"].append(st)
    with open(os.path.join(ROOT, "tests", "golden", "graph_sampling_stats.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote hein_graphs.json, graph_sampling_stats.json")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
