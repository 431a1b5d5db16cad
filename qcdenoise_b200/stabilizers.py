"""Graph-state stabilizers: generator strings, per-setting basis-change sub-circuits and the sampler that runs all
settings (reference qcdenoise/stabilizers.py:33-223).  Pauli strings read like qiskit keys: rightmost character =
qubit 0.  Unlike the reference there is no 10-qubit cap (quirk Q8) and the setting order is deterministic
(generators in node order, identity last) instead of `list(set(...))` hash order (quirk Q9)."""
from collections import OrderedDict
from copy import deepcopy

import networkx as nx

from .ir import Circuit
from .samplers import CircuitSampler, NoiseModel

__all__ = ["TothStabilizer", "JungStabilizer", "StabilizerSampler", "StabilizerCircuit", "get_unique_operators",
           "sigma_prod", "pauli_mask"]

# single-qubit Pauli products: (a, b) -> (power of i, result) with a.b = i^power * result
_IDX = {"I": 0, "X": 1, "Y": 2, "Z": 3}
_LBL = "IXYZ"


def _mul(a, b):
    ia, ib = _IDX[a], _IDX[b]
    if ia == 0 or ib == 0:
        return 0, _LBL[ia or ib]
    if ia == ib:
        return 0, "I"
    c = 6 - ia - ib                      # the third Pauli
    return (1 if (ib - ia) % 3 == 1 else 3), _LBL[c]


def sigma_prod(op_str):
    """Product of the single-qubit Paulis in `op_str`, evaluated right to left as (... (op[-2] . op[-1])); returns
    (complex coefficient, label) like the reference's sympy-based helper (stabilizers.py:39-79)."""
    ops = list(op_str)
    power = 0
    while len(ops) > 1:
        last = ops.pop()
        prev = ops.pop()
        p, r = _mul(prev, last)
        power = (power + p) % 4
        ops.append(r)
    return (1, 1j, -1, -1j)[power] + 0j, ops[0]


def get_unique_operators(stabilizers):
    """Strip the leading sign; unique, in first-seen order."""
    seen = OrderedDict()
    for s in stabilizers:
        seen.setdefault(s[1:], None)
    return list(seen)


def pauli_mask(name):
    """Bit q set iff the operator acts non-trivially on qubit q (character n-1-q != 'I'); a leading sign is
    ignored.  D[x] = (-1)^popcount(x & mask) is the diagonal the reference builds with np.kron."""
    s = name[1:] if name[0] in "+-" else name
    return sum(1 << i for i, ch in enumerate(s[::-1]) if ch != "I")


class StabilizerCircuit:
    def __init__(self, graph, n_qubits=3):
        if not isinstance(graph, (nx.Graph, nx.DiGraph)):
            raise AssertionError("graph should be an instance of networkx.Graph or networkx.DiGraph")
        self.graph = graph
        self.n_qubits = n_qubits
        self.generators = self._get_generators()
        self.stabilizers = None
        self.circuits = []

    def _get_generators(self):
        gens = []
        for v in self.graph.nodes():
            s = ["I"] * self.n_qubits
            s[v] = "X"
            for w in self.graph.neighbors(v):
                s[w] = "Z"
            gens.append("".join(s)[::-1])
        return gens

    def find_stabilizers(self, **kwargs):
        raise NotImplementedError

    def build(self, drop_coef=False, **kwargs):
        self.find_stabilizers(**kwargs)
        out = OrderedDict()
        for name in get_unique_operators(self.stabilizers):
            circ = Circuit(self.n_qubits, name=name)
            paulis = name[1:] if drop_coef else name
            for q, ch in enumerate(paulis[::-1]):
                if ch == "X":
                    circ.h(q)
                elif ch == "Y":
                    circ.sdg(q)
                    circ.h(q)
                elif ch in "IZ":
                    circ.id(q)
            self.circuits.append(circ)
            out[name] = circ
        return out


class TothStabilizer(StabilizerCircuit):
    """The n generators + identity (Toth-Guehne witness settings)."""

    def find_stabilizers(self):
        self.stabilizers = ["+" + g for g in self.generators] + ["+" + "I" * self.n_qubits]
        return deepcopy(self.stabilizers)

    def build(self):
        return super().build()


class JungStabilizer(StabilizerCircuit):
    """Full signed stabilizer group (2^n elements), the enumeration of the reference's legacy code
    (circuit_constructors.py:554-590) as bit arithmetic on (x, z) masks: products of generators by symplectic
    composition with the phase tracked mod 4."""

    def find_stabilizers(self, max_elements=None):
        n = self.n_qubits
        gens = []
        for g in self.generators:
            x = sum(1 << q for q, ch in enumerate(g[::-1]) if ch in "XY")
            z = sum(1 << q for q, ch in enumerate(g[::-1]) if ch in "ZY")
            gens.append((x, z))
        out = []
        total = 2 ** len(gens)
        for sel in range(total if max_elements is None else min(total, max_elements)):
            # the reference selects generator j by character j of np.binary_repr(sel, n) (MSB first)
            x = z = 0
            power = 0          # element = i^power * X^x Z^z
            for j, (gx, gz) in enumerate(gens):
                if not (sel >> (len(gens) - 1 - j)) & 1:
                    continue
                # (X^x Z^z)(X^gx Z^gz) = (-1)^{popc(z & gx)} X^(x^gx) Z^(z^gz)
                power = (power + 2 * bin(z & gx).count("1")) % 4
                x ^= gx
                z ^= gz
            # X^x Z^z with both bits set is XZ = -iY  ->  label Y carries i^{-popc(x & z)}
            power = (power - bin(x & z).count("1")) % 4
            label = "".join("IXZY"[((x >> q) & 1) | (((z >> q) & 1) << 1)] for q in range(n))[::-1]
            if power not in (0, 2):
                raise ArithmeticError("non-Hermitian product in a stabilizer group")
            out.append(("+" if power == 0 else "-") + label)
        self.stabilizers = out
        return deepcopy(out)


class StabilizerSampler(CircuitSampler):
    """Appends each setting's basis change to the graph circuit and simulates all settings in ONE batched engine
    call (the reference simulates them one after another: stabilizers.py:183-215).  Ideal unless `noise_model` is
    given (reference quirk Q7)."""

    def __init__(self, backend=None, n_shots=1024, **kw):
        super().__init__(backend=backend, n_shots=n_shots, **kw)

    def sample(self, stabilizer_circuits, graph_circuit, execute=False, noise_model=None):
        if not isinstance(graph_circuit, Circuit):
            raise AssertionError("graph_circuit should be an instance of qcdenoise_b200.Circuit")
        circuits = []
        for name, stab in stabilizer_circuits.items():
            circ = graph_circuit.copy(name=name)
            circ.append(stab.to_gate(label=name), qargs=range(stab.num_qubits))
            circ.measure_all()
            circuits.append(circ)
        self.transpile_circuit(circuits, {})
        self.build_noise_model(noise_model)
        return self.simulate_circuit(circuits)

    def build_noise_model(self, noise_model):
        if noise_model is not None and not noise_model.is_ideal():
            self.noise_model = noise_model
            self.noisy = True
        else:
            self.noise_model = NoiseModel()
            self.noisy = False
