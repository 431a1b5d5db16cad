"""Graph-state stabilizers: generator strings, per-setting basis-change sub-circuits and the sampler that runs all
settings (reference qcdenoise/stabilizers.py:33-223).  Pauli strings read like qiskit keys: rightmost character =
qubit 0.  Unlike the reference there is no 10-qubit cap (quirk Q8) and the setting order is deterministic
(generators in node order, identity last) instead of `list(set(...))` hash order (quirk Q9)."""
from collections import OrderedDict
from copy import deepcopy

import networkx as nx
import numpy as np

from .ir import Circuit
from .samplers import CircuitSampler, NoiseModel

__all__ = ["TothStabilizer", "JungStabilizer", "StabilizerSampler", "StabilizerCircuit", "get_unique_operators",
           "sigma_prod", "pauli_mask", "single_qubit_settings", "run_settings_batch", "graph_adjacency_masks"]

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

    def generator_masks(self):
        """(x, z) bit masks of the generators: bit q of x / z set <=> X / Z on qubit q (Y: both)."""
        return [(sum(1 << q for q, ch in enumerate(g[::-1]) if ch in "XY"),
                 sum(1 << q for q, ch in enumerate(g[::-1]) if ch in "ZY")) for g in self.generators]

    def find_stabilizers(self, max_elements=None, engine=None):
        """All 2^n signed elements, in the reference's order (element `sel` multiplies the generators picked by the
        characters of np.binary_repr(sel, n)).  With `engine` the enumeration runs on the GPU (qcd_stabilizer_group) and
        only the label strings are formed here."""
        n = self.n_qubits
        gens = self.generator_masks()
        total = 2 ** len(gens)
        count = total if max_elements is None else min(total, max_elements)
        if engine is not None:
            xs, zs, sg = engine.stabilizer_group([g[0] for g in gens], [g[1] for g in gens], 0, count)
            xs = xs.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            zs = zs.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            q = np.arange(n - 1, -1, -1)
            code = ((xs[:, None] >> q) & 1) | (((zs[:, None] >> q) & 1) << 1)          # [count, n], qubit n-1 first
            letters = np.array(list("IXZY"))[code]
            out = [("-" if s else "+") + "".join(row) for s, row in zip(sg.cpu().numpy(), letters)]
            self.stabilizers = out
            return deepcopy(out)
        out = []
        for sel in range(count):
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


def graph_adjacency_masks(graph, n_qubits):
    """uint32 [n]: bit j of row k set <=> edge (k, j) (either direction for a DiGraph) -- the form
    qcd_run_density_matrix_group takes."""
    adj = np.zeros(n_qubits, dtype=np.uint32)
    for u, v in graph.edges():
        if u != v:
            adj[u] |= np.uint32(1 << v)
            adj[v] |= np.uint32(1 << u)
    return adj


_SQ2 = 2 ** -0.5
_GATE_1Q = {"id": np.eye(2, dtype=np.complex128), "h": np.array([[_SQ2, _SQ2], [_SQ2, -_SQ2]], dtype=np.complex128),
            "s": np.diag([1, 1j]).astype(np.complex128), "sdg": np.diag([1, -1j]).astype(np.complex128),
            "x": np.array([[0, 1], [1, 0]], dtype=np.complex128), "y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
            "z": np.diag([1, -1]).astype(np.complex128)}
_PAULIS = [_GATE_1Q["id"], _GATE_1Q["x"], _GATE_1Q["y"], _GATE_1Q["z"]]


def _errors_after(model, name, qs):
    """Channels `model` attaches to an unlabelled gate `name` on qubits `qs` (the rule of NoiseModel.instantiate),
    without materialising a calibration-defined model."""
    from . import channels
    from .samplers import DeviceNoiseModel
    if model is None:
        return []
    if isinstance(model, DeviceNoiseModel) and model.native:
        ent = model.gate_table.get((name, tuple(qs)))
        if ent is None:
            return []
        props = model.props["qubits"]
        t1 = [max(props[q].get("T1", float("inf")), 1e-9) for q in qs]
        t2 = [max(props[q].get("T2", float("inf")), 1e-9) for q in qs]
        probs, relax = channels.device_gate_error(len(qs), ent[0], ent[1], t1, t2)
        out = []
        if probs is not None:
            out.append(("pauli", tuple(range(len(qs))), probs))
        if relax is not None:
            out += [("kraus", (i,), k) for i, k in enumerate(relax)]
        return out
    errs = list(model.label_errors.get(name, [])) + list(model.gate_errors.get((name, tuple(qs)), [])) \
        + list(model.all_qubit_errors.get(name, []))
    return [ch for err in errs for ch in err]


def single_qubit_settings(stabilizer_circuits, model, n_qubits):
    """If every measurement setting differs from the bare circuit by single-qubit steps on at most ONE qubit (the Toth
    settings: H on the generator's vertex, identities elsewhere), return (qubits uint8 [S], superops complex [S, 4, 4])
    for Engine.run_density_matrix_settings; else None.  The superoperator is the ordered product of the setting's
    gates on that qubit and of the channels `model` attaches to them, S = sum_j conj(K_j) (x) K_j."""
    qubits, sops = [], []
    for stab in stabilizer_circuits.values():
        if stab.n_qubits != n_qubits:
            return None
        target, S = None, np.eye(4, dtype=np.complex128)
        for op in stab.ops:
            name, qs = op[0], op[1]
            if name in ("barrier", "measure_all"):
                continue
            if name not in _GATE_1Q or len(qs) != 1:
                return None
            errs = _errors_after(model, name, qs)
            if name == "id" and not errs:
                continue
            if target is None:
                target = qs[0]
            elif target != qs[0]:
                return None
            S = np.kron(_GATE_1Q[name].conj(), _GATE_1Q[name]) @ S
            for ch in errs:
                if ch[0] == "kraus" and tuple(ch[1]) == (0,):
                    T = sum(np.kron(np.asarray(k).conj(), np.asarray(k)) for k in ch[2])
                elif ch[0] == "pauli" and tuple(ch[1]) == (0,):
                    T = sum(p * np.kron(P.conj(), P) for p, P in zip(ch[2], _PAULIS))
                else:
                    return None
                S = T @ S
        qubits.append(0xFF if target is None else target)
        sops.append(S)
    return np.array(qubits, dtype=np.uint8), np.array(sops, dtype=np.complex128)


def run_settings_batch(sampler, circuits, models, settings):
    """Density-matrix route for measurement settings: ONE simulation per circuit, every setting's distribution read
    off its density matrix (qcd_run_density_matrix_settings), then sampled.  settings[i] = single_qubit_settings(...)
    of circuit i (same number of settings for all).  -> list (per circuit) of lists (per setting) of Counts."""
    from .samplers import Counts, get_engine, upload_batch
    n = circuits[0].n_qubits
    eng = get_engine(sampler.device)
    readout, _ = upload_batch(eng, circuits, models)
    q = np.stack([s[0] for s in settings])
    sop = np.stack([s[1] for s in settings])
    B, S = q.shape
    probs = eng.run_density_matrix_settings(q, sop, sampler.precision).reshape(B * S, 2 ** n)
    sampler.last_probs = probs
    ro = np.repeat(readout, S, axis=0) if readout is not None else None
    hist = eng.sample_from_probs(probs, sampler.n_shots, seed=sampler._next_seed(), readout=ro)
    return [[Counts(None, n, device_hist=hist[i * S + j:i * S + j + 1]) for j in range(S)] for i in range(B)]


class StabilizerSampler(CircuitSampler):
    """Appends each setting's basis change to the graph circuit and simulates all settings in ONE batched engine
    call (the reference simulates them one after another: stabilizers.py:183-215).  Ideal unless `noise_model` is
    given (reference quirk Q7)."""

    settings_from_one_state = True     # False: always re-simulate the circuit once per setting (A/B, tests)

    def __init__(self, backend=None, n_shots=1024, **kw):
        super().__init__(backend=backend, n_shots=n_shots, **kw)

    def sample(self, stabilizer_circuits, graph_circuit, execute=False, noise_model=None):
        if not isinstance(graph_circuit, Circuit):
            raise AssertionError("graph_circuit should be an instance of qcdenoise_b200.Circuit")
        self.build_noise_model(noise_model)
        if self.noisy and self.settings_from_one_state and \
                self._pick_method(graph_circuit.n_qubits, True) == "density_matrix":
            # density-matrix route: simulate the graph circuit ONCE and read every setting off its density matrix
            settings = single_qubit_settings(stabilizer_circuits, self.noise_model, graph_circuit.n_qubits)
            if settings is not None:
                self.transpile_circuit(graph_circuit, {})
                return run_settings_batch(self, [graph_circuit], [self.noise_model], [settings])[0]
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
