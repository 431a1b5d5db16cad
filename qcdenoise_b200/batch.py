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

__all__ = ["build_gate_lists", "sites_per_edge", "slice_programs", "device_gate_superops_1q", "toth_masks",
           "toth_witnesses", "draw_device_props", "device_props_dict", "adjacency_tensors_batch", "edge_lists"]

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


def sites_per_edge(builder):
    """Potential labelled noise sites per edge: 1 for the CX / CZ builders, 2 for CPhase (one after each CX)."""
    return 2 if isinstance(builder, CPhaseGateCircuit) else 1


def build_gate_lists(builder, edges, n_edges, coins=None, gammas=None, **build_options):
    """Gate lists of a batch of graphs, as the encoded arrays `Engine.upload` takes: (ops, offsets, params, n).
    builder: a CXGateCircuit / CZGateCircuit / CPhaseGateCircuit instance.  Same gates in the same order as
    builder.build(graph) for a graph whose `edges()` are edges[i, :n_edges[i]] (graph_circuit.py:81-99, :114-130,
    :151-173).

    Stochastic builders: coins uint8 [total edges, sites_per_edge(builder)] says which potential sites exist (the
    object route's np.random.choice(2) per site, graph_circuit.py:91); gammas float64 [coins.sum(), 2] are the two
    U(0, max_prob) draws of each site in site order (samplers.py:423-451).  A site is the labelled identity `unitary`
    followed by the amplitude-damping channels UnitaryNoiseSampler attaches to its label: draw 1 on the first qubit,
    draw 0 on the second (reference quirk Q1).  Two sites of one CPhase edge share the label `unitary_a_b`, so each
    instance carries both sites' channels, like the reference's composed label errors."""
    n = builder.n_qubits
    edges = np.asarray(edges, dtype=np.uint8)
    n_edges = np.asarray(n_edges, dtype=np.int64)
    B = len(n_edges)
    if edges.ndim != 3 or edges.shape[0] != B or edges.shape[2] != 2:
        raise ValueError("edges must have shape (n_graphs, max_edges, 2)")
    keep = np.arange(edges.shape[1])[None, :] < n_edges[:, None]
    a = edges[..., 0][keep].astype(np.uint8)       # flat, circuit after circuit
    b = edges[..., 1][keep].astype(np.uint8)
    E = len(a)
    if E and max(int(a.max()), int(b.max())) >= n:
        raise AssertionError(f"n_qubits={n} but an edge names vertex {max(int(a.max()), int(b.max()))}")
    S = sites_per_edge(builder)
    if coins is None:
        if builder.stochastic:
            raise ValueError("a stochastic builder needs `coins` (and `gammas`)")
        coins = np.zeros((E, S), dtype=np.int64)
    coins = np.asarray(coins).astype(np.int64).reshape(E, S)
    n_sites = int(coins.sum())
    if n_sites:
        gammas = np.asarray(gammas, dtype=np.float64)
        if gammas.shape != (n_sites, 2):
            raise ValueError(f"gammas must have shape ({n_sites}, 2)")
        if (gammas < 0).any() or (gammas > 1).any():
            raise ValueError("damping parameters out of range")
    circ = np.repeat(np.arange(B), n_edges)
    H, CX, CZ, RY = OPCODES["h"], OPCODES["cx"], OPCODES["cz"], OPCODES["ry"]
    U2Q, K1Q = OPCODES["u2q"], OPCODES["kraus1q"]
    if isinstance(builder, (CXGateCircuit, CZGateCircuit)):
        n_gates, prefix = 3, n
    elif isinstance(builder, CPhaseGateCircuit):
        n_gates, prefix = 5, 0
    else:
        raise TypeError(f"no batched gate-list builder for {type(builder).__name__}")
    k_e = coins.sum(axis=1)                         # chosen sites of the edge = channels pairs behind each instance
    inst_len = 1 + 2 * k_e                          # marker + 2 channels per chosen site of the edge
    block = coins * inst_len[:, None]               # [E, S] ops of each site instance (0 if not chosen)
    edge_len = n_gates + block.sum(axis=1)
    per_circ = np.bincount(circ, weights=edge_len, minlength=B).astype(np.int64) if E else np.zeros(B, dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(prefix + per_circ)])
    if offsets[-1] >= 2 ** 32:
        raise ValueError("batch too large for one upload (2^32 ops)")
    ops = np.zeros(int(offsets[-1]), dtype=OP_DTYPE)
    ops["q1"] = 0xFF
    base = (circ + 1) * prefix + np.concatenate([[0], np.cumsum(edge_len)[:-1]]) if E else np.zeros(0, dtype=np.int64)

    def put(pos, code, q0, q1=None, off=None, sel=None):
        if sel is not None:
            pos, q0 = pos[sel], q0[sel]
            q1 = q1[sel] if q1 is not None else None
            off = off[sel] if isinstance(off, np.ndarray) else off
        ops["opcode"][pos] = code
        ops["q0"][pos] = q0
        if q1 is not None:
            ops["q1"][pos] = q1
        if off is not None:
            ops["param_off"][pos] = off

    if prefix:
        pos = (offsets[:-1][:, None] + np.arange(n)[None, :]).ravel()
        put(pos, H, np.tile(np.arange(n, dtype=np.uint8), B))
    # parameters: [builder angles] [I4 of the site markers] [17 doubles per (site, draw): 2, K0, K1]
    lam = build_options.get("Lambda", np.pi)
    theta = build_options.get("Theta", np.pi / 2)
    head = [lam / 2, -theta / 2] if isinstance(builder, CPhaseGateCircuit) else []
    i4_off = len(head)
    k_base = i4_off + (32 if n_sites else 0)
    params = np.zeros(k_base + 34 * n_sites, dtype=np.float64)
    params[:len(head)] = head
    if n_sites:
        params[i4_off + np.arange(4) * 10] = 1.0             # identity 4x4, (re, im) pairs, row-major
        kb = params[k_base:].reshape(n_sites, 2, 17)
        kb[..., 0] = 2.0
        kb[..., 1] = 1.0                                      # K0 = diag(1, sqrt(1 - g))
        kb[..., 7] = np.sqrt(np.maximum(0.0, 1.0 - gammas))
        kb[..., 11] = np.sqrt(gammas)                         # K1 = sqrt(g) |0><1|
        if (np.sqrt(gammas) <= 1e-12).any():
            raise ValueError("a damping parameter of (numerically) zero: use the object route")
    site_idx = (np.cumsum(coins.ravel()) - coins.ravel()).reshape(E, S)      # global index of each chosen site

    def site_block(pos, j):
        """Instance of site j of every edge that has it, starting at pos: marker, then per chosen site jj of the edge
        K(a, draw 1), K(b, draw 0)."""
        sel = coins[:, j] == 1
        put(pos, U2Q, a, b, i4_off, sel)
        slot = np.ones(E, dtype=np.int64)
        for jj in range(S):
            both = sel & (coins[:, jj] == 1)
            put(pos + slot, K1Q, a, None, k_base + (2 * site_idx[:, jj] + 1) * 17, both)
            put(pos + slot + 1, K1Q, b, None, k_base + (2 * site_idx[:, jj]) * 17, both)
            slot = slot + 2 * coins[:, jj]

    if isinstance(builder, (CXGateCircuit, CZGateCircuit)):
        hq = b if isinstance(builder, CXGateCircuit) else a       # H(b) CX(a, b) [site] H(b)  /  H(a) CZ(a, b) [site] H(a)
        put(base, H, hq)
        put(base + 1, CX if isinstance(builder, CXGateCircuit) else CZ, a, b)
        site_block(base + 2, 0)
        put(base + 2 + block[:, 0], H, hq)
    else:                                           # Ry(L/2, b) CX [site] Ry(-T/2, b) CX [site] Ry(L/2, b)
        put(base, RY, b, None, 0)
        put(base + 1, CX, a, b)
        site_block(base + 2, 0)
        p2 = base + 2 + block[:, 0]
        put(p2, RY, b, None, 1)
        put(p2 + 1, CX, a, b)
        site_block(p2 + 2, 1)
        put(p2 + 2 + block[:, 1], RY, b, None, 0)
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
    # distinct (circuit, gate, qubits), sorted as numpy.unique((circuit << 24) | opcode << 16 | q0 << 8 | q1) would list them,
    # natively and per circuit (qcd_gate_entries): the sort over all ops of the batch was 14 of this stage's 20 ms
    lib = _lib.load()
    n_ops = len(ops)
    ekey = np.empty(max(n_ops, 1), dtype=np.uint32)
    ecirc = np.empty(max(n_ops, 1), dtype=np.uint32)
    op_entry = np.empty(max(n_ops, 1), dtype=np.uint32)
    n_ent = ctypes.c_size_t()
    off32 = np.ascontiguousarray(offsets, dtype=np.uint32)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)       # noqa: E731
    rc = lib.qcd_gate_entries(ptr(ops) if n_ops else None, ptr(off32), B, ptr(ekey), ptr(ecirc), ptr(op_entry),
                              ctypes.byref(n_ent))
    if rc:
        raise _lib.QcdError(rc, "qcd_gate_entries: malformed gate lists")
    U = n_ent.value
    ekey, ecirc = ekey[:U], ecirc[:U]
    op_entry = op_entry[:n_ops]
    opcode = ((ekey >> np.uint32(16)) & np.uint32(0xFF)).astype(np.intp)
    present = np.flatnonzero(np.bincount(opcode, minlength=256))
    gate = {}
    for nm in specs["cx"].keys():
        table = np.full(256, np.nan)          # spec value by opcode (NaN: the gate's spec has no such entry, quirk Q4)
        for code in present:
            try:
                table[code] = float(specs[_GATE_NAMES.get(int(code))][nm])
            except KeyError:
                pass
        top = table[opcode]
        draw = rng.uniform(0.0, 1.0, size=U)
        if np.isnan(table[present]).any():
            missing = np.isnan(top)
            gate[nm] = np.where(missing, 1e-6, draw * np.where(missing, 0.0, top))
        else:
            gate[nm] = draw * top
    return {"qubit": qubit, "gate_key": ekey.copy(), "gate_circuit": ecirc.astype(np.int64),
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


def slice_programs(enc, lo, hi):
    """Circuits [lo, hi) of an encoded batch as an encoded batch of their own (the parameter array is shared, so
    param_off stays valid) -> ((ops, offsets, params, n), (op_lo, op_hi)): one rank's share of a sharded batch."""
    ops, offsets, params, n = enc
    o_lo, o_hi = int(offsets[lo]), int(offsets[hi])
    return (ops[o_lo:o_hi], (offsets[lo:hi + 1].astype(np.int64) - o_lo).astype(np.uint32), params, n), (o_lo, o_hi)


_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) * 2 ** -0.5


def device_gate_superops_1q(unitary, gate_error, gate_length, t1_us, t2_us):
    """4x4 superoperators (index = ket + 2 * bra) of a noisy single-qubit gate for a whole batch, natively
    (qcd_device_gate_superops_1q): `unitary` followed by the calibration-derived error of each entry (NaN, NaN = no
    error).  gate_error / gate_length / t1_us / t2_us: float64 arrays of one shape -> complex128 [shape + (4, 4)]."""
    ge = np.ascontiguousarray(gate_error, dtype=np.float64)
    gl, t1, t2 = (np.ascontiguousarray(np.broadcast_to(a, ge.shape), dtype=np.float64) for a in (gate_length, t1_us, t2_us))
    u = np.ascontiguousarray(unitary, dtype=np.complex128).reshape(4)
    out = np.empty(ge.shape + (4, 4), dtype=np.complex128)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)       # noqa: E731
    rc = _lib.load().qcd_device_gate_superops_1q(ge.size, p(u.view(np.float64)), p(ge), p(gl), p(t1), p(t2),
                                                 p(out.view(np.float64)))
    if rc:
        raise _lib.QcdError(rc, "qcd_device_gate_superops_1q: relaxation parameters out of range")
    return out


def toth_masks(edges, n_edges, n):
    """Parity masks of the n + 1 Toth settings of every graph: generator k = X on vertex k, Z on its neighbours
    (stabilizers.py:104-122) -> bit k and the neighbours' bits; identity last (mask 0).  uint32 [B, n + 1]."""
    B = len(n_edges)
    keep = np.arange(edges.shape[1])[None, :] < np.asarray(n_edges)[:, None]
    circ = np.repeat(np.arange(B), n_edges)
    a = edges[..., 0][keep].astype(np.int64)
    b = edges[..., 1][keep].astype(np.int64)
    masks = np.zeros((B, n + 1), dtype=np.uint32)
    masks[:, :n] = (1 << np.arange(n, dtype=np.uint32))[None, :]
    np.bitwise_or.at(masks, (circ, a), (1 << b).astype(np.uint32))
    np.bitwise_or.at(masks, (circ, b), (1 << a).astype(np.uint32))
    return masks


def toth_witnesses(kind, signed, totals, edges, n_edges, n):
    """Witness values from the settings' parity sums (witnesses.py:59-67, :74-144), for a whole batch.
    signed, totals: int64 [B, n + 1] (generator settings, identity last).  kind "genuine": W = (n - 1) <I> - sum_k <g_k>,
    sigma = sqrt(sum std^2) -> (value [B], sigma [B]); "biseparable": per edge (i, j) W_ij = 1 - <g_i> - <g_j>,
    sigma = sqrt(std_i^2 + std_j^2) -> (list of B value lists, list of B sigma lists), edge order = graph.edges()."""
    mean = signed / totals
    std = np.sqrt(np.maximum(0.0, (1.0 - mean * mean) / totals))
    if kind == "genuine":
        value = (n - 1) * mean[:, n] - mean[:, :n].sum(axis=1)
        return value, np.sqrt((std ** 2).sum(axis=1))      # std-devs enter unscaled (witnesses.py:136-144)
    keep = np.arange(edges.shape[1])[None, :] < np.asarray(n_edges)[:, None]
    circ = np.repeat(np.arange(len(n_edges)), n_edges)
    a = edges[..., 0][keep].astype(np.int64)
    b = edges[..., 1][keep].astype(np.int64)
    val = 1.0 - mean[circ, a] - mean[circ, b]
    sig = np.sqrt(std[circ, a] ** 2 + std[circ, b] ** 2)
    ends = np.cumsum(n_edges).tolist()
    val, sig = val.tolist(), sig.tolist()
    return ([val[e - k:e] for e, k in zip(ends, np.asarray(n_edges).tolist())],
            [sig[e - k:e] for e, k in zip(ends, np.asarray(n_edges).tolist())])
