"""The benchmark's C4 workload (BASELINE.json configs[3]) restated on the ORACLE side, from the reference's own
construction rules, without importing the product package: the CPU reference arm (bench.py --impl reference) builds its
gate lists here.  TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).
tests/test_oracle_workload_cpu.py checks gate for gate that these are the programs bench.py uploads to the GPU.

Construction (all draws from the global np.random stream, in the reference's order):
  graph         tests/golden/graphs.json (sampled with the reference's GraphDB / GraphState code)
  edge order    networkx Graph.edges() of add_nodes_from(range(n)); add_edges_from(edges)
  gate list     CXGateCircuit.build, stochastic=True: one np.random.choice(2) coin per edge   graph_circuit.py:81-99
  noise         UnitaryNoiseSampler.build_noise_model: per labelled site two U(0, max_prob) draws, amplitude damping
                (quirk Q1), channel of draw 1 on the first qubit of the label, draw 0 on the second   samplers.py:388-451
  settings      TothStabilizer: per vertex the generator X_v prod_{w ~ v} Z_w, then the identity; sub-circuit X -> h,
                I / Z -> id appended to the graph circuit                                   stabilizers.py:104-168, :194-199
"""
import json
import os

import numpy as np

from . import noise, ref_glue

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def nx_edge_order(n, edges):
    """Order in which networkx yields Graph.edges() for nodes 0..n-1 added first, then `edges` in list order."""
    adj = {q: {} for q in range(n)}
    for a, b in edges:
        adj[a][b] = None
        adj[b][a] = None
    seen, out = set(), []
    for u in range(n):
        for v in adj[u]:
            if v not in seen:
                out.append((u, v))
        seen.add(u)
    return out


def c4_workload(graph_index, n=20, seed=1234, max_prob=0.1):
    """-> dict(edges, n_sites, circuits = [graph circuit, n + 1 settings] as gate lists, names, masks)."""
    with open(os.path.join(ROOT, "tests", "golden", "graphs.json")) as f:
        graphs = json.load(f)["graphs"][str(n)]
    edges = nx_edge_order(n, [tuple(e) for e in graphs[graph_index % len(graphs)]])
    np.random.seed(seed + graph_index)
    coins = [int(np.random.choice(2)) for _ in edges]
    ops, labels = ref_glue.cx_gate_circuit(n, edges, coins)
    draws = [np.random.uniform(high=max_prob, size=2) for _ in labels]
    noisy, site = [], 0
    for op in ops:
        noisy.append(op)
        if op[0] == "unitary" and len(op) > 3:
            on_node, on_ngbr = noise.reference_unitary_noise_site(draws[site])
            site += 1
            noisy.append(("kraus", (op[1][0],), on_node))
            noisy.append(("kraus", (op[1][1],), on_ngbr))
    nbrs = {v: [] for v in range(n)}
    for a, b in edges:
        nbrs[a].append(b)
        nbrs[b].append(a)
    names = ["".join("X" if q == v else ("Z" if q in nbrs[v] else "I") for q in range(n))[::-1] for v in range(n)]
    names.append("I" * n)
    circuits = [noisy]
    for name in names:
        sub = [("h", (q,)) if ch == "X" else ("id", (q,)) for q, ch in enumerate(name[::-1])]
        circuits.append(noisy + sub)
    masks = [0] + [sum(1 << q for q, ch in enumerate(nm[::-1]) if ch != "I") for nm in names]
    return {"edges": edges, "n_sites": len(labels), "circuits": circuits, "names": names, "masks": masks, "n": n}
