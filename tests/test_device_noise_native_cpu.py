"""Native (C++) construction of the device-noise programs (qcd_expand_device_noise, the host half of
qcd_upload_device_noise_programs) against the explicit Python construction (DeviceNoiseSampler + channels.py, which
tests/test_oracle_channels.py pins to the oracle).  Host-only: no GPU."""
import math
import random

import networkx as nx
import numpy as np
import pytest

import qcdenoise_b200 as qcd
from qcdenoise_b200 import channels
from qcdenoise_b200._lib import OPCODES
from qcdenoise_b200.engine import expand_device_noise
from qcdenoise_b200.ir import encode_programs
from qcdenoise_b200.samplers import DeviceNoiseModel, device_noise_arrays

_PLEN = {OPCODES["rx"]: 1, OPCODES["ry"]: 1, OPCODES["rz"]: 1, OPCODES["u1q"]: 8, OPCODES["u2q"]: 32,
         OPCODES["pauli1q"]: 4, OPCODES["pauli2q"]: 16}


def _payload(op, params):
    code = int(op["opcode"])
    if code in (OPCODES["kraus1q"], OPCODES["kraus2q"]):
        m = int(params[op["param_off"]])
        return params[op["param_off"]: op["param_off"] + 1 + m * (8 if code == OPCODES["kraus1q"] else 32)]
    k = _PLEN.get(code, 0)
    return params[op["param_off"]: op["param_off"] + k]


def _same_programs(a, b, tol=1e-14):
    (ops_a, off_a, par_a, n_a), (ops_b, off_b, par_b, n_b) = a, b
    assert n_a == n_b
    assert np.array_equal(off_a, off_b)
    assert len(ops_a) == len(ops_b)
    for x, y in zip(ops_a, ops_b):
        assert (x["opcode"], x["q0"], x["q1"]) == (y["opcode"], y["q0"], y["q1"])
        px, py = _payload(x, par_a), _payload(y, par_b)
        assert len(px) == len(py)
        assert np.allclose(px, py, rtol=0, atol=tol)


def _random_graph(n, rng):
    g = nx.gnm_random_graph(n, rng.randint(n - 1, 2 * n), seed=rng.randint(0, 10 ** 6))
    g.add_nodes_from(range(n))
    return g


@pytest.mark.parametrize("builder", [qcd.CXGateCircuit, qcd.CZGateCircuit, qcd.CPhaseGateCircuit])
def test_native_expansion_equals_python_instantiate(builder):
    rng = random.Random(7)
    random.seed(11)
    np.random.seed(11)
    n = 6
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 150.0, "readout_error": 0.2, "prob_meas0_prep1": 0.1,
                                          "prob_meas1_prep0": 0.1},
                                "cx": {"gate_error": 0.25, "gate_length": 500.0}, "cz": {"gate_error": 0.1, "gate_length": 300.0},
                                "h": {"gate_error": 0.05, "gate_length": 50.0}, "ry": {"gate_error": 0.02, "gate_length": 80.0}})
    s = qcd.DeviceNoiseSampler(None, n_shots=64, noise_specs=spec)
    circuits, models = [], []
    for _ in range(70):          # more than one host thread's block (blocks of >= 64 circuits)
        c = builder(n_qubits=n, stochastic=False).build(_random_graph(n, rng))["circuit"]
        s.transpile_circuit(c)
        s.build_noise_model(n)
        assert isinstance(s.noise_model, DeviceNoiseModel) and s.noise_model.native
        circuits.append(c)
        models.append(s.noise_model)
    t1t2, gp, ro = device_noise_arrays(circuits, models)
    native = expand_device_noise(encode_programs(circuits), t1t2, gp)
    assert all(m.native for m in models)          # building the arrays does not materialise the models
    inst = [m.instantiate(c) for m, c in zip(models, circuits)]
    explicit = encode_programs([p for p, _ in inst])
    _same_programs(native, explicit)
    assert np.allclose(ro, np.stack([r for _, r in inst]), rtol=0, atol=0)
    # T1 < T2 <= 2 T1 (Choi branch) and T2 <= T1 (mixture branch) both occurred
    t1, t2 = t1t2[..., 0], t1t2[..., 1]
    assert (t2 > t1).any() and (t2 <= t1).any()


def _one_gate(k, g_err, g_len, t1, t2):
    c = qcd.Circuit(2)
    (c.h(0) if k == 1 else c.cx(0, 1))
    gp = np.array([[g_err, g_len]], dtype=np.float64)
    tq = np.array([[[t1[0], t2[0]], [t1[-1], t2[-1]]]], dtype=np.float64)
    ops, off, par, _ = expand_device_noise(encode_programs([c]), tq, gp)
    return ops, par


@pytest.mark.parametrize("k", [1, 2])
def test_native_gate_error_edge_cases(k):
    nan, inf = float("nan"), float("inf")
    rng = np.random.default_rng(5)
    cases = [(0.1, 100.0, [50.0] * k, [70.0] * k), (0.1, nan, [50.0] * k, [70.0] * k), (nan, 100.0, [50.0] * k, [20.0] * k),
             (0.0, 0.0, [50.0] * k, [50.0] * k), (0.9, 1e5, [1.0] * k, [2.0] * k), (1e-9, 10.0, [inf] * k, [inf] * k),
             (0.3, 500.0, [1e-12] * k, [1e-12] * k), (0.2, 400.0, [30.0] * k, [1000.0] * k)]
    cases += [(float(rng.uniform(0, 0.5)), float(rng.uniform(0, 1000)), list(rng.uniform(0.01, 200, k)),
               list(rng.uniform(0.01, 200, k))) for _ in range(200)]
    for g_err, g_len, t1, t2 in cases:
        ops, par = _one_gate(k, g_err, g_len, t1, t2)
        t1c, t2c = [max(t, 1e-9) for t in t1], [max(t, 1e-9) for t in t2]
        probs, relax = channels.device_gate_error(k, None if math.isnan(g_err) else g_err,
                                                  None if math.isnan(g_len) else g_len, t1c, t2c)
        want = []
        if probs is not None:
            want.append(("pauli", probs))
        if relax is not None:
            want += [("kraus", ks) for ks in relax]
        got = ops[1:]
        assert len(got) == len(want), (g_err, g_len, t1, t2)
        for op, (kind, val) in zip(got, want):
            if kind == "pauli":
                assert int(op["opcode"]) == OPCODES["pauli1q" if k == 1 else "pauli2q"]
                assert np.allclose(_payload(op, par), val, rtol=0, atol=1e-15)
            else:
                assert int(op["opcode"]) == OPCODES["kraus1q"]
                flat = _payload(op, par)
                assert int(flat[0]) == len(val)
                mats = flat[1:].reshape(len(val), 2, 2, 2)
                for m, kr in zip(mats, val):
                    assert np.allclose(m[..., 0] + 1j * m[..., 1], kr, rtol=0, atol=1e-15)
                # trace preserving
                tot = sum((m[..., 0] + 1j * m[..., 1]).conj().T @ (m[..., 0] + 1j * m[..., 1]) for m in mats)
                assert np.allclose(tot, np.eye(2), atol=1e-12)


def test_native_expansion_rejects_bad_input():
    c = qcd.Circuit(2)
    c.h(0)
    enc = encode_programs([c])
    with pytest.raises(ValueError):
        expand_device_noise(enc, np.zeros((1, 3, 2)), np.zeros((1, 2)))
    ops, off, par, n = enc
    bad = ops.copy()
    bad["q0"] = 5                                    # qubit outside the register
    with pytest.raises(qcd._lib.QcdError, match="qubit index out of range"):
        expand_device_noise((bad, off, par, n), np.ones((1, 2, 2)), np.array([[0.1, 100.0]]))
    with pytest.raises(qcd._lib.QcdError, match="relaxation parameters"):
        expand_device_noise(enc, np.full((1, 2, 2), np.nan), np.array([[0.1, 100.0]]))
