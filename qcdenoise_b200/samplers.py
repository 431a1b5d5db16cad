"""Sampler classes with the reference's names and call signatures (qcdenoise/samplers.py:205-552), running on the
CUDA engine instead of `qk.execute(circuit, AerSimulator, noise_model, shots)` (samplers.py:308-326).

What differs from the reference, deliberately:
  * "circuit" is an `ir.Circuit` gate list; there is no transpiler, so device noise is attached by
    (gate name, qubits) to the gate list as written -- the reference's `backend=None` semantics
    (samplers.py:597-611, :703-727).  `backend` may be None or a plain properties dict (see DeviceNoiseSampler).
  * a NoiseModel here is explicit data: channels keyed by instruction label or (gate, qubits) + readout matrices.
  * `sample()` does NOT mutate the caller's circuit (reference quirk Q6) unless `faithful=True`.
  * the reference's `stateprep_errors` (quirk Q3: an `initialize` appended AFTER the gates, which erases the graph
    state) is only reproduced with `faithful=True`.
"""
import random
from collections.abc import Mapping
from copy import deepcopy
from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch

from . import channels
from .engine import Engine
from .ir import Circuit

__all__ = ["NoiseSpec", "unitary_noise_spec", "device_noise_spec", "NoiseModel", "DeviceNoiseModel", "Counts", "CircuitSampler",
           "UnitaryNoiseSampler", "DeviceNoiseSampler", "convert_to_prob_vector", "generate_binary_strings",
           "get_backendProp", "get_engine"]

_ENGINES = {}


def get_engine(device=None, slot=0):
    """Process-wide engine per CUDA device (one context per (process, device)).  slot > 0: an additional context on the
    same device -- simulate() alternates between two so that one lowers the next chunk of a large batch on the host
    while the other's kernels run."""
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    key = device if slot == 0 else (device, slot)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(device)
    return _ENGINES[key]


def _to_host(t):
    """Device tensor -> numpy through a pinned staging buffer (pageable cudaMemcpy is ~3x slower for the 4 MiB-per-
    circuit histograms of a 20-qubit run)."""
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()


@dataclass(init=True)
class NoiseSpec:
    """Noise-model specification (same container as the reference's, samplers.py:205-209)."""
    specs: Dict = None


unitary_noise_spec = NoiseSpec(specs={"type": "phase_amplitude_damping_error", "max_prob": 0.1})
_GATE = {"gate_error": 0.5, "gate_length": 0.5}
device_noise_spec = NoiseSpec(specs={
    "qubit": {"T1": 0.5, "T2": 0.5, "frequency": 0.5, "readout_error": 0.5, "prob_meas0_prep1": 0.5,
              "prob_meas1_prep0": 0.5},
    "cx": dict(_GATE), "id": dict(_GATE), "u1": dict(_GATE), "u2": dict(_GATE), "u3": dict(_GATE)})


def get_backendProp(backend_dict):
    """The reference wraps its random properties dict in a qiskit `BackendProperties` (samplers.py:94-112); here the
    plain dict IS the properties object the samplers consume."""
    if not isinstance(backend_dict, dict) or "qubits" not in backend_dict or "gates" not in backend_dict:
        raise ValueError("backend properties must be a dict with 'qubits' and 'gates'")
    return backend_dict


def generate_binary_strings(n_qubits):
    """All n-bit strings in numeric order (samplers.py:37-68), built without recursion."""
    return [format(x, f"0{n_qubits}b") for x in range(2 ** n_qubits)]


class Counts(Mapping):
    """qiskit-`Counts`-like read-only mapping bitstring -> int (key MSB = qubit n-1), backed by the dense histogram
    the engine produced; the Python dict is only materialised if somebody iterates it."""

    def __init__(self, hist, n_qubits, device_hist=None):
        self._hist = None if hist is None else np.asarray(hist)
        self.n_qubits = n_qubits
        self.device_hist = device_hist      # [1, 2^n] int32 view of the engine's output (stays on the GPU)
        self._d = None

    @property
    def hist(self):
        """Dense host histogram; copied from the device the first time somebody asks for it."""
        if self._hist is None:
            self._hist = _to_host(self.device_hist)[0]
        return self._hist

    @classmethod
    def from_dict(cls, d, n_qubits=None):
        n = n_qubits if n_qubits is not None else len(next(iter(d)))
        h = np.zeros(2 ** n, dtype=np.int64)
        for k, v in d.items():
            h[int(k.replace(" ", ""), 2)] = v
        return cls(h, n)

    def _dict(self):
        if self._d is None:
            nz = np.nonzero(self.hist)[0]
            self._d = {format(int(x), f"0{self.n_qubits}b"): int(self.hist[x]) for x in nz}
        return self._d

    def __getitem__(self, key):
        return self._dict()[key]

    def __iter__(self):
        return iter(self._dict())

    def __len__(self):
        return int(np.count_nonzero(self.hist))

    def shots(self):
        return int(self.hist.sum())

    def __repr__(self):
        return repr(self._dict()) if len(self) <= 64 else f"Counts(n_qubits={self.n_qubits}, outcomes={len(self)})"


def convert_to_prob_vector(prob_dict, n_qubits, n_shots):
    """counts -> dense probability vector, p[int(key, 2)] = count / n_shots (samplers.py:195-202)."""
    if isinstance(prob_dict, Counts):
        return prob_dict.hist.astype(np.float64) / n_shots
    p = np.zeros(2 ** n_qubits, dtype=np.float64)
    for key, itm in prob_dict.items():
        p[int(key.replace(" ", ""), 2)] = itm / n_shots
    return p


class NoiseModel:
    """Explicit noise description.  An *error* is a list of channel ops in instruction-local qubit numbering:
    ("kraus", (0,), [K...]) / ("pauli", (0, 1), probs) / ("reset", (0,)); local index i = i-th qubit of the
    instruction the error follows."""

    def __init__(self):
        self.label_errors = {}       # instruction label -> [error, ...]   (composed in insertion order)
        self.gate_errors = {}        # (gate name, qubits tuple) -> [error, ...]
        self.all_qubit_errors = {}   # gate name -> [error, ...]
        self.readout = {}            # qubit -> 2x2 assignment matrix (row = prepared, col = measured)
        self.basis_gates = []

    def add_all_qubit_quantum_error(self, error, label):
        """Attach `error` to every instruction whose label (or, failing that, gate name) equals `label`; adding a
        second error for the same label composes them, like Aer."""
        self.label_errors.setdefault(label, []).append(error)

    def add_quantum_error(self, error, gate, qubits):
        self.gate_errors.setdefault((gate, tuple(qubits)), []).append(error)

    def add_readout_error(self, matrix, qubit):
        self.readout[int(qubit)] = np.asarray(matrix, dtype=np.float64)

    def add_basis_gates(self, gates):
        self.basis_gates += list(gates)

    def is_ideal(self):
        return not (self.label_errors or self.gate_errors or self.all_qubit_errors or self.readout)

    def instantiate(self, circuit):
        """-> (Circuit with the channels written out after the instructions they follow, readout [n,2,2] | None)."""
        out = Circuit(circuit.n_qubits, name=circuit.name)
        for op in circuit.ops:
            out.ops.append(op)
            if op[0] in ("barrier", "measure_all"):
                continue
            qs = op[1]
            errs = []
            label = op[3] if op[0] == "unitary" and len(op) > 3 else None
            if label is not None and label in self.label_errors:
                errs += self.label_errors[label]
            elif label is None and op[0] in self.label_errors:
                errs += self.label_errors[op[0]]
            errs += self.gate_errors.get((op[0], tuple(qs)), [])
            errs += self.all_qubit_errors.get(op[0], [])
            for err in errs:
                for ch in err:
                    tgt = tuple(qs[i] for i in ch[1])
                    out.ops.append((ch[0], tgt) + tuple(ch[2:]))
        ro = None
        if self.readout:
            ro = np.tile(np.eye(2), (circuit.n_qubits, 1, 1))
            for q, m in self.readout.items():
                if q < circuit.n_qubits:
                    ro[q] = m
        return out, ro


class DeviceNoiseModel(NoiseModel):
    """Noise model given by calibration numbers (`props`: {"qubits": [...], "gates": [...]}, see DeviceNoiseSampler).
    The channels are built natively, for the whole batch, when the circuits are uploaded
    (qcd_upload_device_noise_programs).  Reading `gate_errors` / `readout` (or adding errors) materialises the same
    channels explicitly in Python instead; both routes give the same program, operator for operator."""

    def __init__(self, props):
        self.props = props
        self._explicit = False
        self._gate_errors, self._readout = {}, {}
        super().__init__()
        nq = len(props["qubits"])
        # calibration entry per (gate, qubits); entries that name a qubit without properties are ignored
        self.gate_table = {(g["gate"], tuple(g["qubits"])): (g.get("gate_error"), g.get("gate_length"))
                           for g in props["gates"] if all(q < nq for q in g["qubits"])}

    @property
    def native(self):
        """True while nothing but the calibration numbers defines this model."""
        return not (self._explicit or self.label_errors or self.all_qubit_errors)

    def _materialize(self):
        if self._explicit:
            return
        self._explicit = True
        props = self.props
        for (gate, qs), (g_err, g_len) in self.gate_table.items():
            t1 = [max(props["qubits"][q].get("T1", float("inf")), 1e-9) for q in qs]
            t2 = [max(props["qubits"][q].get("T2", float("inf")), 1e-9) for q in qs]
            probs, relax = channels.device_gate_error(len(qs), g_err, g_len, t1, t2)
            err = []
            if probs is not None:
                err.append(("pauli", tuple(range(len(qs))), probs))
            if relax is not None:
                err += [("kraus", (i,), k) for i, k in enumerate(relax)]
            if err:
                self._gate_errors.setdefault((gate, qs), []).append(err)
        for q, m in enumerate(self.readout_rows()):
            if m is not None:
                self._readout[q] = np.array(m, dtype=np.float64)

    def readout_rows(self):
        """Per qubit: [[1 - p10, p10], [p01, 1 - p01]] or None (no readout entry)."""
        out = []
        for qp in self.props["qubits"]:
            if "prob_meas1_prep0" in qp and "prob_meas0_prep1" in qp:
                a, b = qp["prob_meas1_prep0"], qp["prob_meas0_prep1"]
            elif "readout_error" in qp:
                a = b = qp["readout_error"]
            else:
                out.append(None)
                continue
            out.append([[1.0 - a, a], [b, 1.0 - b]])
        return out

    def is_ideal(self):
        if self.native:
            return not (self.gate_table or any(m is not None for m in self.readout_rows()))
        return super().is_ideal()

    gate_errors = property(lambda self: (self._materialize(), self._gate_errors)[1],
                           lambda self, v: setattr(self, "_gate_errors", v))
    readout = property(lambda self: (self._materialize(), self._readout)[1],
                       lambda self, v: setattr(self, "_readout", v))


_NAN2 = (float("nan"), float("nan"))
_NOT_GATES = ("barrier", "measure_all")


def device_noise_arrays(circuits, models):
    """Calibration numbers of a batch as the arrays qcd_upload_device_noise_programs takes:
    (qubit_t1_t2 [B, n, 2], op_gate_props [total ops, 2], readout [B, n, 2, 2] | None)."""
    n = circuits[0].n_qubits
    inf = float("inf")
    t1t2, gp, ro, any_ro = [], [], [], False
    eye = [[1.0, 0.0], [0.0, 1.0]]
    for circ, model in zip(circuits, models):
        qp = model.props["qubits"]
        row = [[max(q.get("T1", inf), 1e-9), max(q.get("T2", inf), 1e-9)] for q in qp[:n]]
        row += [[inf, inf]] * (n - len(row))
        t1t2.append(row)
        rr = model.readout_rows()[:n]
        any_ro = any_ro or any(m is not None for m in rr)
        rr = [m if m is not None else eye for m in rr]
        rr += [eye] * (n - len(rr))
        ro.append(rr)
        table = model.gate_table
        for op in circ.ops:
            name = op[0]
            if name in _NOT_GATES:
                continue
            ent = table.get((name, tuple(op[1])), _NAN2) if name not in ("kraus", "pauli", "reset") else _NAN2
            gp.append((float("nan") if ent[0] is None else ent[0], float("nan") if ent[1] is None else ent[1]))
    return (np.array(t1t2, dtype=np.float64).reshape(len(circuits), n, 2),
            np.array(gp, dtype=np.float64).reshape(-1, 2),
            np.array(ro, dtype=np.float64) if any_ro else None)


def upload_batch(eng, circuits, models):
    """Upload circuits with THEIR noise models (None = ideal) -> (readout [B, n, 2, 2] | None, noisy)."""
    from .ir import encode_programs
    n = circuits[0].n_qubits
    if models and all(isinstance(m, DeviceNoiseModel) and m.native for m in models):
        t1t2, gp, readout = device_noise_arrays(circuits, models)
        eng.upload_device_noise(encode_programs(circuits), t1t2, gp)
        with np.errstate(invalid="ignore"):
            noisy = bool((gp > 0).any()) or any(op[0] in ("kraus", "pauli", "reset") for c in circuits for op in c.ops)
        return readout, noisy
    inst = [(m if m is not None else NoiseModel()).instantiate(c) for c, m in zip(circuits, models)]
    progs = [c for c, _ in inst]
    readout = None
    if any(r is not None for _, r in inst):
        readout = np.stack([r if r is not None else np.tile(np.eye(2), (n, 1, 1)) for _, r in inst])
    noisy = any(op[0] in ("kraus", "pauli", "reset") for c in progs for op in c.ops)
    eng.upload(progs)
    return readout, noisy


def run_batch(sampler, circuits, models, fetch=False):
    """Simulate a batch in ONE engine call, every circuit under its own noise model -> list of Counts."""
    n = circuits[0].n_qubits
    eng = get_engine(sampler.device)
    readout, noisy = upload_batch(eng, circuits, models)
    seed = sampler._next_seed()
    if sampler._pick_method(n, noisy) == "density_matrix":
        probs = eng.run_density_matrix(sampler.precision)
        sampler.last_probs = probs
        hist = eng.sample_from_probs(probs, sampler.n_shots, seed=seed, readout=readout)
    else:
        sampler.last_probs = None
        hist, _ = eng.run_trajectories(sampler.n_shots, seed=seed, precision=sampler.precision, readout=readout)
    if fetch:       # one batched device->host copy (the caller reads every histogram)
        h = _to_host(hist)
        return [Counts(h[i], n, device_hist=hist[i:i + 1]) for i in range(len(circuits))]
    return [Counts(None, n, device_hist=hist[i:i + 1]) for i in range(len(circuits))]


class CircuitSampler:
    """Base class (samplers.py:234-352).  `method`: "automatic" (Aer's rule: noisy and n_shots > 2^n and the
    density matrix fits -> exact density matrix + multinomial sampling, else statevector trajectories),
    "statevector" or "density_matrix".  `precision`: 32 (complex64) or 64 (complex128)."""

    dm_max_qubits = 12

    def __init__(self, backend=None, n_shots=1024, noise_specs=None, *, precision=32, method="automatic",
                 seed=None, device=None, faithful=False):
        self.n_shots = n_shots
        if backend is not None and not isinstance(backend, dict):
            raise AssertionError("backend must be None or a backend-properties dict "
                                 "({'qubits': [...], 'gates': [...]}); qiskit backends are not used by this engine")
        self.backend = backend
        self.transpiled_circuit = None
        self.circuit_dag = None
        self.noise_model = None
        self.noisy = False
        self.noise_specs = noise_specs
        if noise_specs:
            if not isinstance(noise_specs, NoiseSpec):
                raise AssertionError("noise models specs is not an instance of NoiseSpec")
            self.noisy = bool(noise_specs.specs)
        self.precision = precision
        self.method = method
        self.faithful = faithful
        self.device = device
        self._seed = seed if seed is not None else int(np.random.randint(0, 2 ** 31 - 1))
        self._calls = 0
        self.last_probs = None     # exact distribution(s) of the last density-matrix run (device tensor)

    # ---- reference API
    def sample(self, *args, **kwargs):
        raise NotImplementedError

    def build_noise_model(self, *args, **kwargs):
        raise NotImplementedError

    def execute_circuit(self, circuit, execute=False, job_name="placeholder"):
        raise NotImplementedError("execution on IBMQ hardware is outside this engine (reference samplers.py:288-306)")

    def transpile_circuit(self, circuit, transpile_kwargs=None):
        """No transpiler: the gate list is executed as written; kept so that reference call sites run."""
        self.transpiled_circuit = circuit
        self.circuit_dag = circuit
        return circuit

    def _next_seed(self):
        self._calls += 1
        return (self._seed + 0x9E3779B97F4A7C15 * self._calls) & 0xFFFFFFFFFFFFFFFF

    def _pick_method(self, n, noisy):
        if self.method != "automatic":
            return self.method
        if noisy and self.n_shots > 2 ** n and n <= self.dm_max_qubits:
            return "density_matrix"
        return "statevector"

    def simulate_circuit(self, circuit):
        """circuit | list of circuits -> Counts | list of Counts (samplers.py:308-326)."""
        single = isinstance(circuit, Circuit)
        circs = [circuit] if single else list(circuit)
        model = self.noise_model if (self.noise_model is not None and self.noisy) else None
        out = run_batch(self, circs, [model] * len(circs))
        return out[0] if single else out


class UnitaryNoiseSampler(CircuitSampler):
    """Labelled identity `unitary` gates become noise sites: per label two U(0, max_prob) draws from np.random ->
    one damping channel on each of the site's two qubits (samplers.py:388-451).  Every `sample()` draws a new
    noise model."""

    def __init__(self, backend=None, n_shots=1024, noise_specs=unitary_noise_spec, **kw):
        super().__init__(backend=backend, n_shots=n_shots, noise_specs=noise_specs, **kw)

    @staticmethod
    def get_phase_amp_damp_error(func=None, max_prob=0.1):
        """The reference's 'phase-amplitude damping' site (quirk Q1): draws `phases`, forces `amps = min(0, 0.5 -
        phase)` and calls func(phases[i], amps[i]) -- i.e. amplitude damping of strength phases[i] and a phase
        parameter that is 0 for max_prob <= 0.5.  `q_error_1.tensor(q_error_2)` puts error 2 on the FIRST qubit."""
        phases = np.random.uniform(high=max_prob, size=2)
        amps = [min(0, 0.5 - p) for p in phases]
        err = [("kraus", (0,), channels.damping_kraus(phases[1], max(amps[1], 0.0))),
               ("kraus", (1,), channels.damping_kraus(phases[0], max(amps[0], 0.0)))]
        return err, (phases, amps)

    @staticmethod
    def get_amp_damp_error(func=None, max_prob=0.1):
        amps = np.random.uniform(high=max_prob, size=2)
        err = [("kraus", (0,), channels.damping_kraus(amps[1])), ("kraus", (1,), channels.damping_kraus(amps[0]))]
        return err, amps

    def build_noise_model(self, ops_labels):
        self.noise_model = NoiseModel()
        specs = deepcopy(self.noise_specs.specs) if self.noise_specs and self.noise_specs.specs else {}
        kind = specs.pop("type", "phase_amplitude_damping_error")
        funcs = {"phase_amplitude_damping_error": UnitaryNoiseSampler.get_phase_amp_damp_error,
                 "amplitude_damping_error": UnitaryNoiseSampler.get_amp_damp_error}
        if self.noisy:
            if kind not in funcs:
                raise KeyError(f"unsupported error type {kind!r}")
            for op in ops_labels:
                err, _ = funcs[kind](None, **specs)
                self.noise_model.add_all_qubit_quantum_error(err, op)
            self.noise_model.add_basis_gates(["unitary"])

    def sample(self, circuit, ops_labels):
        if not isinstance(circuit, Circuit):
            raise AssertionError("passed circuit is not an instance of qcdenoise_b200.Circuit")
        if self.faithful:
            circuit.measure_all()      # the reference mutates the caller's circuit (Q6)
        self.build_noise_model(ops_labels)
        counts = self.simulate_circuit(circuit)
        return convert_to_prob_vector(counts, circuit.num_qubits, self.n_shots)


class DeviceNoiseSampler(CircuitSampler):
    """Calibration-style noise: per qubit T1/T2/readout, per (gate, qubits) error + length; every gate is followed by
    depolarizing o thermal-relaxation, measurement by a classical assignment error (samplers.py:483-727).

    backend=None: random properties within `noise_specs` (the reference's fake-backend branch, :597-611).
    backend=<dict>: {"qubits": [{"T1", "T2", "readout_error" | "prob_meas0_prep1"/"prob_meas1_prep0"}, ...],
                     "gates": [{"gate", "qubits", "gate_error", "gate_length"}, ...]}   (T in us, length in ns)."""

    def __init__(self, backend=None, n_shots=1024, noise_specs=device_noise_spec, **kw):
        super().__init__(backend=backend, n_shots=n_shots, noise_specs=noise_specs, **kw)
        self.backend_props = None

    def _get_random_value(self, name, op):
        try:
            return random.uniform(0, self.noise_specs.specs[op][name])
        except KeyError:
            return 1e-6            # reference quirk Q4: unknown (gate, property) -> 1e-6

    def _get_gate_specs(self):
        specs = {}
        for op in self.circuit_dag.ops:
            if op[0] in ("barrier", "measure_all", "kraus", "pauli", "reset"):
                continue
            key = "_".join([op[0]] + [str(q) for q in op[1]])
            specs.setdefault(key, {"gate": op[0], "qubits": list(op[1])})
        return specs

    def _generate_qbit_props(self):
        """One qubit's properties, each ~ U(0, spec) (samplers.py:613-637)."""
        return {nm: self._get_random_value(nm, "qubit") for nm in self.noise_specs.specs["qubit"].keys()}

    def _generate_gate_props(self, name, specs):
        """One (gate, qubits) entry; the property names always come from the 'cx' spec (reference quirk Q5,
        samplers.py:639-659)."""
        g = {"name": name, "gate": specs["gate"], "qubits": specs["qubits"]}
        for nm in self.noise_specs.specs["cx"].keys():
            g[nm] = self._get_random_value(nm, specs["gate"])
        return g

    def _generate_backend_dict(self, n_qubits):
        return {"qubits": [self._generate_qbit_props() for _ in range(n_qubits)],
                "gates": [self._generate_gate_props(key, spec) for key, spec in self._get_gate_specs().items()],
                "general": [], "backend_name": "qcdenoise-fake-backend", "backend_version": "0.0.0"}

    def build_noise_model(self, n_qubits, user_backend=True):
        self.noise_model = NoiseModel()
        if not self.noisy:
            return
        self.backend_props = self.backend if self.backend else self._generate_backend_dict(n_qubits)
        self.noise_model = DeviceNoiseModel(self.backend_props)

    def stateprep_errors(self, circuit, zero_init=True):
        """Reference quirk Q3, reproduced literally (only called when faithful=True): appends reset + |0>/|1>
        preparation AFTER the gates, with the spec's max values as weights."""
        if self.noisy:
            q_props = self.noise_specs.specs["qubit"]
            for q in range(circuit.n_qubits):
                pe = q_props["prob_meas1_prep0"]
                zero_w = [1 - pe, pe]
                pe = q_props["prob_meas0_prep1"]
                one_w = [pe, 1 - pe]
                prep = random.choices([[0, 1], [1, 0]], weights=zero_w if zero_init else one_w)[0]
                circuit.reset(q)
                if prep == [0, 1]:
                    circuit.x(q)

    def sample(self, circuit, user_backend=False):
        if not isinstance(circuit, Circuit):
            raise AssertionError("passed circuit is not an instance of qcdenoise_b200.Circuit")
        self.transpile_circuit(circuit, {})
        self.build_noise_model(circuit.num_qubits, user_backend=user_backend)
        if self.faithful:
            self.stateprep_errors(circuit)
            circuit.measure_all()
        counts = self.simulate_circuit(circuit)
        return convert_to_prob_vector(counts, circuit.num_qubits, self.n_shots)
