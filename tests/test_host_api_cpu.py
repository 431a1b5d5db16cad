"""CPU tests of the host-side mirror of the reference API (no simulation calls): builders, stabilizer strings,
noise-model instantiation, graph sampling, library exports."""
import ctypes
import os
import re

import networkx as nx
import numpy as np
import pytest

import qcdenoise_b200 as qcd
from oracle import noise, ref_glue
from qcdenoise_b200 import _lib, channels


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(os.path.dirname(_lib.HERE), "include", "qcdsim.h")).read()
    declared = set(re.findall(r"QCD_API\s+[\w\s\*]+?\b(qcd_\w+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS) and len(declared) == 26
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.qcd_abi_version() == _lib.QCD_ABI_VERSION == 4


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.qcd_create(0, 0, ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.qcd_last_error(None)
    with pytest.raises(RuntimeError):
        qcd.Engine(0)


@pytest.mark.parametrize("builder,ref", [("CXGateCircuit", ref_glue.cx_gate_circuit),
                                         ("CZGateCircuit", ref_glue.cz_gate_circuit),
                                         ("CPhaseGateCircuit", ref_glue.cphase_gate_circuit)])
def test_builders_match_reference_gate_lists(builder, ref, ref_graphs):
    edges = [tuple(e) for e in ref_graphs["graphs"]["7"][0]]
    g = nx.Graph()
    g.add_nodes_from(range(7))
    g.add_edges_from(edges)
    np.random.seed(3)
    out = getattr(qcd, builder)(n_qubits=7, stochastic=True).build(g)
    np.random.seed(3)
    ncoin = len(edges) * (2 if builder == "CPhaseGateCircuit" else 1)
    coins = [int(np.random.choice(2)) for _ in range(ncoin)]
    want_ops, want_labels = ref(7, list(g.edges()), coins)
    assert out["ops"] == want_labels
    got = out["circuit"].ops
    assert len(got) == len(want_ops)
    for a, b in zip(got, want_ops):
        assert a[0] == b[0] and tuple(a[1]) == tuple(b[1])
        if a[0] == "ry":
            assert a[2] == pytest.approx(b[2])
        if a[0] == "unitary":
            assert a[3] == b[3] and np.array_equal(a[2], np.identity(4))
    assert getattr(qcd, builder)(n_qubits=7, stochastic=False).build(g)["ops"] == []


def test_toth_stabilizers_P4_golden(goldens):
    g = nx.path_graph(4)
    stab = qcd.TothStabilizer(g, n_qubits=4)
    circs = stab.build()
    assert len(circs) == 5                                       # reference tests/test_stabilizers.py:37-44
    assert sorted(circs) == sorted(goldens["G4"]["stabilizer_order_c18"])
    assert list(circs)[-1] == "IIII"
    for name, c in circs.items():
        assert [(o[0], o[1]) for o in c.ops] == [(o[0], o[1]) for o in ref_glue.stabilizer_basis_change(name)]
        assert qcd.pauli_mask(name) == int("".join("0" if ch == "I" else "1" for ch in name), 2)


def test_no_ten_qubit_cap():
    g = nx.path_graph(14)
    assert len(qcd.TothStabilizer(g, n_qubits=14).build()) == 15


def test_full_stabilizer_group_golden_G5(goldens):
    for entry in goldens["G5"]:
        gens = entry["generators"]
        n = len(gens[0])
        js = qcd.JungStabilizer.__new__(qcd.JungStabilizer)
        js.n_qubits, js.generators = n, gens
        assert js.find_stabilizers() == entry["stabilizers"]
        assert js.find_stabilizers() == ref_glue.full_stabilizer_group(gens, n)


def test_sigma_prod_matches_reference_restatement():
    import itertools
    for s in itertools.product("IXYZ", repeat=3):
        assert qcd.sigma_prod("".join(s)) == ref_glue.sigma_prod("".join(s))


def test_channels_match_oracle_constructors():
    rng = np.random.default_rng(0)
    sup = lambda ks: noise.kraus_to_superop(ks)
    for _ in range(10):
        ga, gp = rng.uniform(0, 0.5), rng.uniform(0, 0.4)
        np.testing.assert_allclose(sup(channels.damping_kraus(ga, gp)), sup(noise.phase_amplitude_damping_kraus(ga, gp)),
                                   atol=1e-14)
        t1 = rng.uniform(10, 100)
        for f in (rng.uniform(0.1, 1.0), rng.uniform(1.0, 2.0)):
            t = rng.uniform(0.01, 50)
            np.testing.assert_allclose(sup(channels.thermal_relaxation_kraus(t1, f * t1, t)),
                                       sup(noise.thermal_relaxation_kraus(t1, f * t1, t)), atol=1e-12)
        p = rng.uniform(0, 1)
        for k in (1, 2):
            np.testing.assert_allclose(channels.depolarizing_probs(p, k), noise.depolarizing_pauli_probs(p, k))
        for k in (1, 2):
            args = (k, rng.uniform(0, 0.2), rng.uniform(10, 500), list(rng.uniform(20, 100, k)), list(rng.uniform(20, 100, k)))
            pa, ra = channels.device_gate_error(*args)
            pb, rb = noise.device_gate_channels(*args)
            assert (pa is None) == (pb is None)
            if pa is not None:
                np.testing.assert_allclose(pa, pb, atol=1e-13)
            for x, y in zip(ra, rb):
                np.testing.assert_allclose(sup(x), sup(y), atol=1e-12)
    for ks in (channels.damping_kraus(0.3), channels.thermal_relaxation_kraus(50, 80, 10)):
        np.testing.assert_allclose(sum(k.conj().T @ k for k in ks), np.eye(2), atol=1e-13)       # trace preserving


def test_unitary_noise_model_follows_reference_quirk_Q1(goldens):
    """Two U(0, max_prob) draws per label; pure amplitude damping; error 2 lands on the site's first qubit."""
    c = qcd.Circuit(3).h(0).cx(0, 1).unitary(np.identity(4), [0, 1], label="unitary_0_1").h(1)
    s = qcd.UnitaryNoiseSampler(None, n_shots=10)
    np.random.seed(11)
    s.build_noise_model(["unitary_0_1"])
    np.random.seed(11)
    draws = np.random.uniform(high=0.1, size=2)
    inst, ro = s.noise_model.instantiate(c)
    assert ro is None
    ks = [op for op in inst.ops if op[0] == "kraus"]
    assert [k[1] for k in ks] == [(0,), (1,)]
    kn, kg = noise.reference_unitary_noise_site(draws)
    for got, want in zip(ks, (kn, kg)):
        np.testing.assert_allclose(noise.kraus_to_superop(got[2]), noise.kraus_to_superop(want), atol=1e-14)
    # the same label twice composes two errors on every matching instruction
    s.build_noise_model(["unitary_0_1", "unitary_0_1"])
    inst, _ = s.noise_model.instantiate(c)
    assert sum(op[0] == "kraus" for op in inst.ops) == 4


def test_device_noise_model_attaches_by_gate_and_qubits():
    import random
    random.seed(2)
    c = qcd.Circuit(3).h(0).h(1).cx(0, 1).h(1).cx(0, 1)
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0, "readout_error": 0.1, "prob_meas0_prep1": 0.1,
                                          "prob_meas1_prep0": 0.1},
                                "cx": {"gate_error": 0.1, "gate_length": 300.0},
                                "h": {"gate_error": 0.01, "gate_length": 50.0}})
    s = qcd.DeviceNoiseSampler(None, n_shots=10, noise_specs=spec)
    s.transpile_circuit(c)
    s.build_noise_model(3)
    assert {g["name"] for g in s.backend_props["gates"]} == {"h_0", "h_1", "cx_0_1"}
    inst, ro = s.noise_model.instantiate(c)
    assert ro.shape == (3, 2, 2) and np.allclose(ro.sum(axis=2), 1.0)
    kinds = [op[0] for op in inst.ops]
    assert kinds.count("cx") == 2 and kinds.count("h") == 3
    # identical (gate, qubits) pairs carry identical channels
    cx_pos = [i for i, k in enumerate(kinds) if k == "cx"]
    a, b = inst.ops[cx_pos[0] + 1], inst.ops[cx_pos[1] + 1]
    assert a[0] == b[0] == "pauli" and np.array_equal(a[2], b[2])


def test_graph_state_sampler():
    import random
    random.seed(5)
    db = qcd.GraphDB()
    for n in (4, 7, 9, 20):
        g = qcd.GraphState(db, n_qubits=n, min_num_edges=2, max_num_edges=7).sample(min_vertex_cover=1, max_itr=3)
        assert g.number_of_nodes() == n and sorted(g.nodes()) == list(range(n)) and nx.is_connected(g)


def test_counts_mapping_and_prob_vector():
    h = np.array([3, 0, 0, 5, 0, 0, 0, 2])
    c = qcd.Counts(h, 3)
    assert dict(c) == {"000": 3, "011": 5, "111": 2} and len(c) == 3 and c.shots() == 10
    np.testing.assert_array_equal(qcd.convert_to_prob_vector(c, 3, 10), h / 10)
    np.testing.assert_array_equal(qcd.convert_to_prob_vector(dict(c), 3, 10), ref_glue.convert_to_prob_vector(dict(c), 3, 10))
    assert qcd.generate_binary_strings(3) == ref_glue.generate_binary_strings(3)


def test_adjacency_tensor_gate_list_walk():
    """Same integer scatter as the reference's DAG walk (samplers.py:119-192), on the gate list as written."""
    c = qcd.Circuit(3).h(0).h(1).h(0).cx(0, 1).cx(0, 1).cx(1, 2)
    enc = {"h": 1, "cx": 2}
    a = qcd.generate_adjacency_tensor(c, enc, tensor_shape=(4, 3, 3))
    assert a.shape == (4, 3, 3)
    assert a[0, 0, 0] == 1 and a[1, 0, 0] == 1 and a[2, 0, 0] == 0          # two h on qubit 0 stack in planes
    assert a[0, 1, 1] == 1 and a[1, 1, 1] == 0
    assert a[0, 0, 1] == 2 and a[0, 1, 0] == 2 and a[1, 0, 1] == 2 and a[1, 1, 0] == 2
    assert a[0, 1, 2] == 2 and a[0, 2, 1] == 2
    d = qcd.generate_adjacency_tensor(c, enc, tensor_shape=(4, 3, 3), directed=True)
    assert d[0, 0, 1] == 2 and d[0, 1, 0] == 0
    t = qcd.generate_adjacency_tensor(c, enc, tensor_shape=(4, 3, 3), fixed_size=False)
    assert t.shape == (2, 3, 3)
    assert qcd.encode_basis_gates(["id", "u1", "cx"]) == {"id": 0, "u1": 1, "cx": 2}
    with pytest.raises(KeyError):
        qcd.generate_adjacency_tensor(c, {"h": 1}, tensor_shape=(4, 3, 3))


def test_stateprep_errors_quirk_Q3():
    """DeviceNoiseSampler.stateprep_errors (reference samplers.py:554-575, quirk Q3): per qubit a reset and, with the
    spec's assignment-error weights, a flip to |1> -- appended AFTER the gates, exactly like the reference."""
    import random
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 1.0, "T2": 1.0, "prob_meas0_prep1": 0.3, "prob_meas1_prep0": 0.4},
                                "cx": {"gate_error": 0.1, "gate_length": 1.0}})
    s = qcd.DeviceNoiseSampler(None, n_shots=8, noise_specs=spec)
    c = qcd.Circuit(6)
    c.h(0)
    c.cx(0, 1)
    random.seed(1)
    want = [random.choices([[0, 1], [1, 0]], weights=[0.6, 0.4])[0] == [0, 1] for _ in range(6)]
    random.seed(1)
    s.stateprep_errors(c)
    tail = [op for op in c.ops[2:]]
    assert [op[0] for op in c.ops[:2]] == ["h", "cx"]
    i = 0
    for q in range(6):
        assert tail[i][0] == "reset" and tuple(tail[i][1]) == (q,)
        i += 1
        if want[q]:
            assert tail[i][0] == "x" and tuple(tail[i][1]) == (q,)
            i += 1
    assert i == len(tail) and any(want) and not all(want)
    quiet = qcd.DeviceNoiseSampler(None, n_shots=8, noise_specs=qcd.NoiseSpec(specs=None))
    c2 = qcd.Circuit(2)
    quiet.stateprep_errors(c2)
    assert len(c2.ops) == 0                        # not noisy: nothing appended
