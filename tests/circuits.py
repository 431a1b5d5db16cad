"""Shared test inputs: reference-shaped circuits with explicit channel parameters, as gate lists.
Uses the oracle's restatement of the reference builders (oracle/ref_glue.py) and channel constructors
(oracle/noise.py) -- test-side only."""
import json
import os

import numpy as np

from oracle import noise, ref_glue

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def ref_graph(n, i=0):
    with open(os.path.join(GOLDEN_DIR, "graphs.json")) as f:
        g = json.load(f)
    lst = g["graphs"][str(n)]
    return [tuple(e) for e in lst[i % len(lst)]]


def graph_circuit_amp_damp(n, edges, rng, max_prob=0.1, builder="cx", all_sites=False):
    """Reference UnitaryNoiseSampler semantics (samplers.py:388-451, quirk Q1) on a CX/CZ/CPhase graph circuit:
    coin per edge, two U(0, max_prob) draws per labelled site, amplitude damping on (node, ngbr)."""
    build = {"cx": ref_glue.cx_gate_circuit, "cz": ref_glue.cz_gate_circuit, "cphase": ref_glue.cphase_gate_circuit}[builder]
    ncoin = len(edges) * (2 if builder == "cphase" else 1)
    coins = [1] * ncoin if all_sites else list(rng.integers(0, 2, ncoin))
    ops, labels = build(n, edges, coins)
    out = []
    for op in ops:
        out.append(op)
        if op[0] == "unitary" and len(op) > 3:
            draws = rng.uniform(0, max_prob, 2)
            kn, kg = noise.reference_unitary_noise_site(draws)
            out.append(("kraus", (op[1][0],), kn))
            out.append(("kraus", (op[1][1],), kg))
    return out


def depolarizing_circuit(n, edges, p1=0.001, p2=0.01):
    """C1: 1q depolarizing after every 1q gate, 2q depolarizing after every cx (circuit_constructors.py:118-131)."""
    ops, _ = ref_glue.cx_gate_circuit(n, edges)
    out = []
    for op in ops:
        out.append(op)
        if len(op[1]) == 1:
            out.append(("pauli", op[1], noise.depolarizing_pauli_probs(p1, 1)))
        else:
            out.append(("pauli", op[1], noise.depolarizing_pauli_probs(p2, 2)))
    return out


def device_noise_circuit(n, edges, rng, gate_error=0.05, t1_us=(50, 150), gate_len_ns=(50, 500), excited=False,
                         builder="cx"):
    """Device-model semantics (samplers.py:577-727, backend=None branch): per (gate, qubits) a depolarizing Pauli
    mixture then thermal relaxation on each gate qubit; parameters drawn uniformly."""
    ops, _ = {"cx": ref_glue.cx_gate_circuit, "cz": ref_glue.cz_gate_circuit,
              "cphase": ref_glue.cphase_gate_circuit}[builder](n, edges)
    t1 = rng.uniform(*t1_us, n)
    t2 = rng.uniform(0.3, 1.9, n) * t1
    table = {}
    out = []
    for op in ops:
        out.append(op)
        key = (op[0], op[1])
        if key not in table:
            k = len(op[1])
            table[key] = noise.device_gate_channels(
                k, rng.uniform(0, gate_error), rng.uniform(*gate_len_ns), [t1[q] for q in op[1]], [t2[q] for q in op[1]])
        probs, relax = table[key]
        if probs is not None:
            out.append(("pauli", op[1], probs))
        if relax is not None:
            for q, kr in zip(op[1], relax):
                out.append(("kraus", (q,), kr))
    conf = np.array([noise.readout_confusion(rng.uniform(0, 0.1), rng.uniform(0, 0.1)) for _ in range(n)])
    return out, conf


def random_circuit(n, depth, rng, noise_sites=True):
    """Every opcode of the IR, random placement."""
    ops = []
    one = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "id"]
    for _ in range(depth):
        r = rng.integers(0, 12)
        q = int(rng.integers(0, n))
        q2 = int((q + 1 + rng.integers(0, n - 1)) % n)
        if r < 3:
            ops.append((one[rng.integers(0, len(one))], (q,)))
        elif r == 3:
            ops.append((["rx", "ry", "rz"][rng.integers(0, 3)], (q,), float(rng.uniform(-3, 3))))
        elif r == 4:
            a = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
            u, _ = np.linalg.qr(a)
            ops.append(("unitary", (q,), u))
        elif r < 8:
            ops.append((["cx", "cz", "swap"][rng.integers(0, 3)], (q, q2)))
        elif r == 8:
            a = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
            u, _ = np.linalg.qr(a)
            ops.append(("unitary", (q, q2), u))
        elif not noise_sites:
            ops.append(("h", (q,)))
        elif r == 9:
            ops.append(("kraus", (q,), noise.amplitude_damping_kraus(float(rng.uniform(0, 0.3)))))
        elif r == 10:
            if rng.integers(0, 2):
                ops.append(("pauli", (q,), noise.depolarizing_pauli_probs(float(rng.uniform(0, 0.3)), 1)))
            else:
                ops.append(("pauli", (q, q2), noise.depolarizing_pauli_probs(float(rng.uniform(0, 0.3)), 2)))
        else:
            if rng.integers(0, 2):
                ops.append(("reset", (q,)))
            else:
                t1 = float(rng.uniform(20, 100))
                ops.append(("kraus", (q,), noise.thermal_relaxation_kraus(t1, float(rng.uniform(0.2, 2.0)) * t1,
                                                                          float(rng.uniform(0.1, 30)))))
    return ops
