"""Whole-batch host stages of `simulate()` as arrays: gate lists straight from edge arrays, device-noise calibration
draws for every circuit at once, adjacency tensors -- no per-circuit Python objects.  Once the GPU needs ~4 us per
8-qubit circuit (BASELINE config C2: 10 000 circuits), building a `Circuit` + `NoiseModel` per circuit in Python
(~3 ms) is the job's bottleneck by three orders of magnitude; these functions feed the same C ABI from numpy.

The object route (graph_circuit.py builders, DeviceNoiseSampler.build_noise_model) stays the definition: the tests
check that every function here produces exactly what the object route produces for the same graphs / numbers."""
import ctypes

import numpy as np

from . import _lib
from ._lib import OPCODES, OP_DTYPE
from .graph_circuit import CPhaseGateCircuit, CXGateCircuit, CZGateCircuit

__all__ = ["build_gate_lists", "draw_device_props", "device_props_dict", "adjacency_tensors_batch", "edge_lists"]

_GATE_NAMES = {OPCODES[nm]: nm for nm in ("id", "h", "x", "y", "z", "s", "sdg", "t", "tdg", "rx", "ry", "rz", "cx", "cz",
                                          "swap")}
_GATE_NAMES[OPCODES["u1q"]] = _GATE_NAMES[OPCODES["u2q"]] = "unitary"


def edge_lists(edges, n_edges):
    """(edges [B, M, 2], n_edges [B]) -> list of B lists of (u, v) tuples (the `graph_data` entry of the results)."""
    n_edges = np.asarray(n_edges, dtype=np.int64)
    keep = np.arange(edges.shape[1])[None, :] < n_edges[:, None]
    pairs = list(zip(edges[..., 0][keep].tolist(), edges[..., 1][keep].tolist()))
    ends = np.cumsum(n_edges).tolist()
    return [pairs[e - k:e] for e, k in zip(ends, n_edges.tolist())]


def build_gate_lists(builder, edges, n_edges, **build_options):
    """Gate lists of a batch of graphs, as the encoded arrays `Engine.upload` takes: (ops, offsets, params, n).
    builder: a CXGateCircuit / CZGateCircuit / CPhaseGateCircuit instance with stochastic=False (noise sites are drawn
    per circuit by the object route).  Same gates in the same order as builder.build(graph) for a graph whose
    `edges()` are edges[i, :n_edges[i]] (graph_circuit.py:81-99, :114-130, :151-173)."""
    if builder.stochastic:
        raise ValueError("build_gate_lists covers stochastic=False builders; labelled noise sites need the object route")
    n = builder.n_qubits
    edges = np.asarray(edges, dtype=np.uint8)
    n_edges = np.asarray(n_edges, dtype=np.int64)
    B = len(n_edges)
    if edges.ndim != 3 or edges.shape[0] != B or edges.shape[2] != 2:
        raise ValueError("edges must have shape (n_graphs, max_edges, 2)")
    keep = np.arange(edges.shape[1])[None, :] < n_edges[:, None]
    a = edges[..., 0][keep].astype(np.uint8)       # flat, circuit after circuit
    b = edges[..., 1][keep].astype(np.uint8)
    if len(a) and max(int(a.max()), int(b.max())) >= n:
        raise AssertionError(f"n_qubits={n} but an edge names vertex {max(int(a.max()), int(b.max()))}")
    e_start = np.concatenate([[0], np.cumsum(n_edges)])
    e_local = np.arange(len(a)) - np.repeat(e_start[:-1], n_edges)
    circ = np.repeat(np.arange(B), n_edges)
    params = np.zeros(0, dtype=np.float64)
    H, CX, CZ, RY = OPCODES["h"], OPCODES["cx"], OPCODES["cz"], OPCODES["ry"]
    if isinstance(builder, (CXGateCircuit, CZGateCircuit)):
        per_edge, prefix = 3, n
    elif isinstance(builder, CPhaseGateCircuit):
        per_edge, prefix = 5, 0
    else:
        raise TypeError(f"no batched gate-list builder for {type(builder).__name__}")
    offsets = np.concatenate([[0], np.cumsum(prefix + per_edge * n_edges)])
    if offsets[-1] >= 2 ** 32:
        raise ValueError("batch too large for one upload (2^32 ops)")
    ops = np.zeros(int(offsets[-1]), dtype=OP_DTYPE)
    ops["q1"] = 0xFF
    base = offsets[:-1][circ] + prefix + per_edge * e_local
    if prefix:
        pos = (offsets[:-1][:, None] + np.arange(n)[None, :]).ravel()
        ops["opcode"][pos] = H
        ops["q0"][pos] = np.tile(np.arange(n, dtype=np.uint8), B)
    if isinstance(builder, CXGateCircuit):          # H(b) CX(a, b) H(b)
        for k in (0, 2):
            ops["opcode"][base + k] = H
            ops["q0"][base + k] = b
        ops["opcode"][base + 1] = CX
        ops["q0"][base + 1] = a
        ops["q1"][base + 1] = b
    elif isinstance(builder, CZGateCircuit):        # H(a) CZ(a, b) H(a)
        for k in (0, 2):
            ops["opcode"][base + k] = H
            ops["q0"][base + k] = a
        ops["opcode"][base + 1] = CZ
        ops["q0"][base + 1] = a
        ops["q1"][base + 1] = b
    else:                                           # Ry(L/2, b) CX Ry(-T/2, b) CX Ry(L/2, b)
        lam = build_options.get("Lambda", np.pi)
        theta = build_options.get("Theta", np.pi / 2)
        params = np.array([lam / 2, -theta / 2], dtype=np.float64)
        for k, off in ((0, 0), (2, 1), (4, 0)):
            ops["opcode"][base + k] = RY
            ops["q0"][base + k] = b
            ops["param_off"][base + k] = off
        for k in (1, 3):
            ops["opcode"][base + k] = CX
            ops["q0"][base + k] = a
            ops["q1"][base + k] = b
    return ops, offsets.astype(np.uint32), params, n


def draw_device_props(noise_specs, enc, rng):
    """The random calibration numbers DeviceNoiseSampler draws per circuit when backend is None (samplers.py:613-727),
    for the whole batch: every qubit property ~ U(0, spec["qubit"][name]); every distinct (gate, qubits) of a circuit
    gets, for each property NAME of spec["cx"] (reference quirk Q5), U(0, spec[gate][name]) or 1e-6 when the spec has
    no such entry (quirk Q4).  rng: np.random.Generator.
    -> dict: "qubit": {name: [B, n]}, "gate_key": uint32 [U] (circuit-major, first-appearance order is NOT kept: sorted
       by (circuit, opcode, q0, q1)), "gate_circuit": [U], "gate": {name: [U]}, "op_entry": [n_ops] index into U."""
    ops, offsets, _, n = enc
    specs = noise_specs.specs
    B = len(offsets) - 1
    qubit = {nm: rng.uniform(0.0, float(top), size=(B, n)) for nm, top in specs["qubit"].items()}
    circ = np.repeat(np.arange(B, dtype=np.int64), np.diff(offsets.astype(np.int64)))
    key = (ops["opcode"].astype(np.int64) << 16) | (ops["q0"].astype(np.int64) << 8) | ops["q1"].astype(np.int64)
    full, op_entry = np.unique((circ << 24) | key, return_inverse=True)
    opcode = (full >> 16) & 0xFF
    gate = {}
    for nm in specs["cx"].keys():
        top = np.full(len(full), np.nan)
        for code in np.unique(opcode):
            gname = _GATE_NAMES.get(int(code))
            try:
                top[opcode == code] = float(specs[gname][nm])
            except KeyError:
                pass
        draw = rng.uniform(0.0, 1.0, size=len(full)) * np.where(np.isnan(top), 0.0, top)
        gate[nm] = np.where(np.isnan(top), 1e-6, draw)
    return {"qubit": qubit, "gate_key": (full & 0xFFFFFF).astype(np.uint32), "gate_circuit": (full >> 24).astype(np.int64),
            "gate": gate, "op_entry": op_entry.astype(np.int64), "n_circuits": B, "n_qubits": n}


def device_noise_inputs(props):
    """draw_device_props output -> (qubit_t1_t2 [B, n, 2], op_gate_props [n_ops, 2], readout [B, n, 2, 2] | None):
    the arrays qcd_upload_device_noise_programs and the run calls take."""
    B, n = props["n_circuits"], props["n_qubits"]
    q = props["qubit"]
    t1t2 = np.full((B, n, 2), np.inf)
    if "T1" in q:
        t1t2[..., 0] = np.maximum(q["T1"], 1e-9)
    if "T2" in q:
        t1t2[..., 1] = np.maximum(q["T2"], 1e-9)
    g = props["gate"]
    nan = np.full(len(props["gate_key"]), np.nan)
    ent = np.stack([g.get("gate_error", nan), g.get("gate_length", nan)], axis=1)
    gp = ent[props["op_entry"]]
    readout = None
    if "prob_meas1_prep0" in q and "prob_meas0_prep1" in q:
        p10, p01 = q["prob_meas1_prep0"], q["prob_meas0_prep1"]
    elif "readout_error" in q:
        p10 = p01 = q["readout_error"]
    else:
        p10 = None
    if p10 is not None:
        readout = np.empty((B, n, 2, 2))
        readout[..., 0, 0] = 1.0 - p10
        readout[..., 0, 1] = p10
        readout[..., 1, 0] = p01
        readout[..., 1, 1] = 1.0 - p01
    return t1t2, np.ascontiguousarray(gp), readout


def device_props_dict(props, i):
    """Circuit i's calibration numbers as the backend-properties dict DeviceNoiseSampler(backend=<dict>) takes."""
    qubits = [{nm: float(v[i, q]) for nm, v in props["qubit"].items()} for q in range(props["n_qubits"])]
    gates = []
    for j in np.nonzero(props["gate_circuit"] == i)[0]:
        key = int(props["gate_key"][j])
        code, q0, q1 = key >> 16, (key >> 8) & 0xFF, key & 0xFF
        qs = [q0] if q1 == 0xFF else [q0, q1]
        ent = {"gate": _GATE_NAMES[code], "qubits": qs, "name": "_".join([_GATE_NAMES[code]] + [str(x) for x in qs])}
        ent.update({nm: float(v[j]) for nm, v in props["gate"].items()})
        gates.append(ent)
    return {"qubits": qubits, "gates": gates, "general": [], "backend_name": "qcdenoise-fake-backend",
            "backend_version": "0.0.0"}


def adjacency_tensors_batch(enc, encoding, tensor_shape=(32, 32, 32), directed=False):
    """Adjacency tensors of every circuit of an encoded batch, natively (qcd_adjacency_tensors):
    float64 [B, planes, rows, cols].  `encoding`: gate name -> code, as for dataset.generate_adjacency_tensor."""
    ops, offsets, _, n = enc
    planes, rows, cols = tensor_shape
    if n > rows or n > cols:
        raise IndexError(f"tensor_shape {tensor_shape} is too small for {n} qubits")
    codes = np.zeros(len(OPCODES), dtype=np.int32)
    present = np.unique(ops["opcode"])
    for code in present:
        if int(code) >= OPCODES["kraus1q"]:
            continue
        codes[int(code)] = encoding[_GATE_NAMES[int(code)]]       # KeyError for gates missing from the encoding
    B = len(offsets) - 1
    out = np.empty((B, planes, rows, cols), dtype=np.float64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)       # noqa: E731
    rc = _lib.load().qcd_adjacency_tensors(p(ops) if len(ops) else None, p(offsets), B, p(codes), planes, rows, cols,
                                           int(bool(directed)), p(out))
    if rc:
        raise _lib.QcdError(rc, "qcd_adjacency_tensors: invalid arguments")
    return out
