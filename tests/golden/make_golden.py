#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ from the reference checkout.

Run ONCE in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

It produces
  * witness_goldens.json  -- G1..G5 of SURVEY.md section 4: recorded cell outputs of
      examples/graph_witness_refractored.ipynb  (#c18,#c29,#c34-36,#c41,#c45-47,#c53,#c55)
      examples/graph_witness_example.ipynb      (#c23,#c25)
      examples/graph_circuits.ipynb             (#c28,#c30)
      examples/circuit_samplers.ipynb           (#c4: printed (phases, amps) draws -> quirk Q1)
    The notebook outputs are python reprs; they are parsed with ast.literal_eval after replacing the
    `array(...)` / `<qiskit ... at 0x..>` wrappers, nothing is recomputed here.
  * graphs.json -- edge lists sampled with the reference's OWN graph code
      qcdenoise/graph_data.py (GraphDB) + qcdenoise/graph_state.py (GraphState.sample), loaded by file
      path (they only need networkx), seeds random.seed(1234); np.random.seed(1234), for
      n = 4, 5, 7, 8, 9, 10, 14, 20, 28 (SURVEY.md section 8d "Concrete synthetic inputs").
      These are the benchmark / parity-test inputs on the GPU box, where /root/reference is absent.
"""
import ast
import importlib.util
import json
import os
import random
import re
import sys
import types

import numpy as np

REF = os.environ.get("QCD_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _cell_out(nb, idx):
    cell = nb["cells"][idx]
    chunks = []
    for o in cell.get("outputs", []):
        t = o.get("text") or o.get("data", {}).get("text/plain")
        if t:
            chunks.append("".join(t))
    return chunks


def _parse_repr(text):
    text = re.sub(r"<qiskit[^>]*>", "None", text)
    text = re.sub(r"array\(", "(", text)
    text = text.replace("OrderedDict(", "(")
    text = re.sub(r"Witness\(W_ij=", "dict(W_ij=", text)
    # dict(...) calls are not literals: rewrite into {...}
    def _dictcall(m):
        body = m.group(1)
        parts = re.findall(r"(\w+)=((?:\([^()]*\)|[^,()]+)+)", body)
        return "{" + ", ".join(f"'{k}': {v}" for k, v in parts) + "}"
    text = re.sub(r"dict\(((?:[^()]|\([^()]*\))*)\)", _dictcall, text)
    return ast.literal_eval(text.strip())


def witness_goldens():
    out = {}
    nb = json.load(open(os.path.join(REF, "examples", "graph_witness_refractored.ipynb")))
    # G4: ideal run on the c9 graph; key order of #c18
    out["G4"] = {
        "source": "examples/graph_witness_refractored.ipynb#c18,#c34,#c35,#c36",
        "stabilizer_order_c18": [k for k, _ in _parse_repr(_cell_out(nb, 18)[0])],
        "stabilizer_order_c29": [k for k, _ in _parse_repr(_cell_out(nb, 29)[0])],
        "ideal_counts": _parse_repr(_cell_out(nb, 34)[0]),
        "stabilizer_measurements": [[k, list(v)] for k, v in _parse_repr(_cell_out(nb, 35)[0])],
        "genuine": _parse_repr(_cell_out(nb, 36)[0]),
        "n_shots": 1024,
    }
    # G3
    g3 = _parse_repr(_cell_out(nb, 41)[0])
    out["G3"] = {
        "source": "examples/graph_witness_refractored.ipynb#c41",
        "bisep": [{"W_ij": list(v["W_ij"]), "value": v["value"], "variance": float(v["variance"])}
                  for _, v in sorted(g3.items())],
    }
    # G1
    out["G1"] = {
        "source": "examples/graph_witness_refractored.ipynb#c45,#c46,#c47",
        "stabilizer_order": [k for k, _ in _parse_repr(_cell_out(nb, 47)[0])],
        "counts": _parse_repr(_cell_out(nb, 45)[0]),
        "diagonals": [[k, list(v)] for k, v in _parse_repr(_cell_out(nb, 46)[0])],
        "stabilizer_measurements": [[k, list(v)] for k, v in _parse_repr(_cell_out(nb, 47)[0])],
        "n_shots": 1024,
        "n_qubits": 4,
        # graph of #c29/#c41: generators IZZX, IIXZ, ZXIZ, XZII  <=> edges below
        "edges": [[0, 1], [0, 2], [2, 3]],
    }
    # G2
    g2 = _parse_repr(_cell_out(nb, 55)[0])
    out["G2"] = {
        "source": "examples/graph_witness_refractored.ipynb#c53,#c55",
        "genuine_value_c53": float(_cell_out(nb, 53)[0]),
        "genuine": {"value": g2["value"], "variance": float(g2["variance"])},
    }
    # G5
    nb2 = json.load(open(os.path.join(REF, "examples", "graph_witness_example.ipynb")))
    nb3 = json.load(open(os.path.join(REF, "examples", "graph_circuits.ipynb")))
    out["G5"] = [
        {"source": "examples/graph_witness_example.ipynb#c23,#c25",
         "generators": _parse_repr(_cell_out(nb2, 23)[0]),
         "stabilizers": _parse_repr(_cell_out(nb2, 25)[0])},
        {"source": "examples/graph_circuits.ipynb#c28,#c30",
         "generators": _parse_repr(_cell_out(nb3, 28)[0]),
         "stabilizers": _parse_repr(_cell_out(nb3, 30)[0])},
    ]
    # Q1 evidence / G6 layout
    nb4 = json.load(open(os.path.join(REF, "examples", "circuit_samplers.ipynb")))
    printed = _cell_out(nb4, 4)[0].strip().splitlines()
    out["Q1"] = {
        "source": "examples/circuit_samplers.ipynb#c4",
        "printed_params": [[list(p[0]), list(p[1])] for p in (_parse_repr(l) for l in printed)],
        "filtered_probs": _parse_repr(_cell_out(nb4, 4)[1]),
    }
    return out


def _load_ref_module(name, pkg):
    path = os.path.join(REF, "qcdenoise", name + ".py")
    spec = importlib.util.spec_from_file_location(f"{pkg}.{name}", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[f"{pkg}.{name}"] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_graphs():
    """Sample graphs with the reference's own GraphDB / GraphState (file-path import, no qiskit)."""
    pkg = "_qcdref"
    root = types.ModuleType(pkg)
    root.__path__ = [os.path.join(REF, "qcdenoise")]
    sys.modules[pkg] = root
    _load_ref_module("config", pkg)
    gd = _load_ref_module("graph_data", pkg)
    gs = _load_ref_module("graph_state", pkg)
    out = {"how": "reference GraphDB() + GraphState(db, n, 2, 7).sample(min_vertex_cover=1, max_itr=1); "
                  "random.seed(1234); np.random.seed(1234) before each n",
           "graphs": {}}
    for n, count in ((4, 16), (5, 16), (7, 16), (8, 64), (9, 16), (10, 16), (14, 16), (20, 8), (28, 4)):
        random.seed(1234)
        np.random.seed(1234)
        db = gd.GraphDB()
        state = gs.GraphState(db, n_qubits=n, min_num_edges=2, max_num_edges=7)
        lst = []
        for _ in range(count):
            g = state.sample(min_vertex_cover=1, max_itr=1)
            assert sorted(g.nodes()) == list(range(n)), (n, sorted(g.nodes()))
            lst.append([[int(a), int(b)] for a, b in g.edges()])
        out["graphs"][str(n)] = lst
    # the fixed C1 graph: Hein #4 = path P4 (graph_data.py:78-79), 0-based
    out["P4"] = [[0, 1], [1, 2], [2, 3]]
    return out


def main():
    with open(os.path.join(HERE, "witness_goldens.json"), "w") as f:
        json.dump(witness_goldens(), f, indent=1, sort_keys=True)
    with open(os.path.join(HERE, "graphs.json"), "w") as f:
        json.dump(reference_graphs(), f, separators=(",", ":"))
    print("wrote witness_goldens.json, graphs.json")


if __name__ == "__main__":
    main()
