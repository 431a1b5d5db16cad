"""SURVEY 8 f3 on the CPU: the dense-matrix oracle of the full stabilizer group and the noise-robust witnesses
(oracle/ref_glue.py) against the identities that define them, and the product's host enumeration (bit arithmetic)
against the oracle's string arithmetic (sigma_prod, stabilizers.py:39-79) and golden G5."""
import json
import os

import networkx as nx
import numpy as np
import pytest

import qcdenoise_b200 as qcd
from oracle import ref_glue, sim
from tests import circuits

HERE = os.path.dirname(os.path.abspath(__file__))


def _gens(n, edges):
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_edges_from(edges)
    return g, ref_glue.get_generators(n, list(g.nodes()), {v: list(g.neighbors(v)) for v in g.nodes()})


def _graph_state(n, edges):
    ops, _ = ref_glue.cx_gate_circuit(n, edges)
    psi = sim.ideal_statevector(ops, n)
    return np.outer(psi, psi.conj())


@pytest.mark.parametrize("n,idx", [(4, 0), (5, 1), (6, 2)])
def test_signed_group_sums_to_the_graph_state_projector(n, idx):
    edges = circuits.ref_graph(n if n != 6 else 5, idx) + ([(4, 5)] if n == 6 else [])
    g, gens = _gens(n, edges)
    group = ref_glue.full_stabilizer_group(gens, n)
    assert len(group) == 2 ** n and len({s[1:] for s in group}) == 2 ** n
    proj = ref_glue.graph_state_projector(group, n)
    rho = _graph_state(n, edges)
    np.testing.assert_allclose(proj, rho, atol=1e-12)                       # |G><G| = 2^-n sum_S S
    ev = ref_glue.group_expectations_dense(rho, group)
    np.testing.assert_allclose(list(ev.values()), 1.0, atol=1e-12)           # every signed element stabilizes |G>
    assert abs(ref_glue.projector_witness_dense(rho, group, n) + 0.5) < 1e-12
    mixed = np.eye(2 ** n) / 2 ** n
    assert abs(ref_glue.projector_witness_dense(mixed, group, n) - (0.5 - 2.0 ** -n)) < 1e-12
    # the product's bit arithmetic gives the same signed labels in the same order
    assert qcd.JungStabilizer(g, n).find_stabilizers() == group


def test_golden_G5_group_from_masks():
    with open(os.path.join(HERE, "golden", "witness_goldens.json")) as f:
        g5 = json.load(f)["G5"]
    for case in g5:
        gens = case["generators"]
        n = len(gens[0])
        js = qcd.JungStabilizer(nx.empty_graph(n), n)
        js.generators = gens
        assert js.find_stabilizers() == case["stabilizers"]
        masks = js.generator_masks()
        assert all(x & z == 0 for x, z in masks) and len(masks) == n


@pytest.mark.parametrize("n,edges", [(4, [(0, 1), (1, 2), (2, 3)]), (5, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0)][:4]),
                                     (4, [(0, 1), (0, 2), (0, 3)])])
def test_two_colour_witness_bounds_the_projector_witness(n, edges):
    """W_1 - 2 W_2 = (1 - P_A)(1 - P_B) >= 0 as an operator, so <W_1> >= 2 <W_2> on EVERY state; both are -1 / -1/2 on
    the graph state itself."""
    g, gens = _gens(n, edges)
    group = ref_glue.full_stabilizer_group(gens, n)
    colour = nx.bipartite.color(g)
    colours = [colour[v] for v in g.nodes()]
    rho = _graph_state(n, edges)
    assert abs(ref_glue.two_colour_witness_dense(rho, gens, colours, n) + 1.0) < 1e-12
    rng = np.random.default_rng(n)
    for _ in range(6):
        a = rng.normal(size=(2 ** n, 2 ** n)) + 1j * rng.normal(size=(2 ** n, 2 ** n))
        r = a @ a.conj().T
        r /= np.trace(r).real
        mix = rng.uniform(0, 1)
        r = mix * r + (1 - mix) * rho
        w1 = ref_glue.two_colour_witness_dense(r, gens, colours, n)
        w2 = ref_glue.projector_witness_dense(r, group, n)
        assert w1 >= 2 * w2 - 1e-10
    # white noise rho_p = (1 - p) |G><G| + p 1/2^n: the projector witness detects up to p < 1 / (2 (1 - 2^-n)), the
    # generator witness (level 0) only up to p < 1 / n
    p = 0.45
    noisy = (1 - p) * rho + p * np.eye(2 ** n) / 2 ** n
    assert ref_glue.projector_witness_dense(noisy, group, n) < 0
    w0 = (n - 1) - sum(np.real(np.trace(noisy @ ref_glue.pauli_matrix(gk))) for gk in gens)
    assert w0 > 0


def test_adjacency_masks():
    g = nx.DiGraph()
    g.add_nodes_from(range(5))
    g.add_edges_from([(0, 3), (3, 1), (4, 0)])
    assert qcd.graph_adjacency_masks(g, 5).tolist() == [0b11000, 0b01000, 0, 0b00011, 0b00001]
