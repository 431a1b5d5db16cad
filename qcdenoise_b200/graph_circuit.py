"""Graph -> gate list.  Same classes / `build(graph) -> {"circuit", "ops"}` contract as the reference's
graph_circuit.py:68-173; "circuit" is an `ir.Circuit` (ordered op list) instead of a qiskit QuantumCircuit."""
import networkx as nx
import numpy as np

from .ir import Circuit

__all__ = ["GraphCircuit", "CXGateCircuit", "CZGateCircuit", "CPhaseGateCircuit"]

_I4 = np.identity(4)


class GraphCircuit:
    """Base builder.  `stochastic`: each entangling gate is followed, with probability 1/2 (np.random.choice, like the
    reference), by a labelled identity `unitary` that a sampler later turns into a noise site."""

    def __init__(self, n_qubits=3, stochastic=True, gate_type="none"):
        self.n_qubits = n_qubits
        self.stochastic = stochastic
        self.gate_type = gate_type
        self.circuit = None
        self.graph_iter = []

    def _start(self, graph):
        if not isinstance(graph, (nx.Graph, nx.DiGraph)):
            raise AssertionError("input graph must be of type networkx.Graph or networkx.DiGraph")
        if self.n_qubits != len(graph.nodes):
            raise AssertionError(f"n_qubits={self.n_qubits} != # of graph nodes={len(graph.nodes)}")
        if isinstance(graph, nx.DiGraph):
            self.graph_iter = [(u, v) for u, nbrs in graph.adjacency() for v in nbrs]
        else:
            self.graph_iter = list(graph.edges())
        self.circuit = Circuit(self.n_qubits)

    def _maybe_site(self, node, ngbr, labels):
        if bool(np.random.choice(2)) and self.stochastic:
            label = f"unitary_{node}_{ngbr}"
            labels.append(label)
            self.circuit.unitary(_I4, [node, ngbr], label=label)

    def build(self, graph):
        raise NotImplementedError


class CXGateCircuit(GraphCircuit):
    """H on every qubit, then per edge (a, b):  H(b) CX(a, b) [site] H(b)."""

    def build(self, graph):
        self._start(graph)
        self.gate_type = "CX"
        labels = []
        for q in range(self.n_qubits):
            self.circuit.h(q)
        for node, ngbr in self.graph_iter:
            self.circuit.h(ngbr)
            self.circuit.cx(node, ngbr)
            self._maybe_site(node, ngbr, labels)
            self.circuit.h(ngbr)
        return {"circuit": self.circuit.copy(), "ops": labels}


class CZGateCircuit(GraphCircuit):
    """H on every qubit, then per edge (a, b):  H(a) CZ(a, b) [site] H(a)."""

    def build(self, graph):
        self._start(graph)
        self.gate_type = "CZ"
        labels = []
        for q in range(self.n_qubits):
            self.circuit.h(q)
        for node, ngbr in self.graph_iter:
            self.circuit.h(node)
            self.circuit.cz(node, ngbr)
            self._maybe_site(node, ngbr, labels)
            self.circuit.h(node)
        return {"circuit": self.circuit.copy(), "ops": labels}


class CPhaseGateCircuit(GraphCircuit):
    """Per edge (a, b):  Ry(L/2, b) CX(a, b) [site] Ry(-T/2, b) CX(a, b) [site] Ry(L/2, b)."""

    def build(self, graph, Lambda=np.pi, Theta=np.pi / 2):
        self._start(graph)
        self.gate_type = "CPHASE"
        labels = []
        for node, ngbr in self.graph_iter:
            self.circuit.ry(Lambda / 2, ngbr)
            self.circuit.cx(node, ngbr)
            self._maybe_site(node, ngbr, labels)
            self.circuit.ry(-Theta / 2, ngbr)
            self.circuit.cx(node, ngbr)
            self._maybe_site(node, ngbr, labels)
            self.circuit.ry(Lambda / 2, ngbr)
        return {"circuit": self.circuit.copy(), "ops": labels}
