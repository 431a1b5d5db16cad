"""GPU tests of the reference-facing Python API: the reference's own tests (shape / length / "runs" assertions,
tests/test_samplers.py, test_stabilizers.py, test_simulate.py) plus the value-level goldens G1-G4."""
import random

import networkx as nx
import numpy as np
import pytest

from oracle import sim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qcd():
    import qcdenoise_b200 as q
    return q


def _graph(n, seed, qcd):
    random.seed(seed)
    np.random.seed(seed)
    return qcd.GraphState(qcd.GraphDB(), n_qubits=n, min_num_edges=2, max_num_edges=7).sample()


def test_unitary_noise_sampler_shape_and_distribution(qcd):
    n = 5
    g = _graph(n, 1, qcd)
    cd = qcd.CXGateCircuit(n_qubits=n, stochastic=True).build(g)
    sampler = qcd.UnitaryNoiseSampler(None, n_shots=4096, noise_specs=qcd.NoiseSpec(
        specs={"type": "phase_amplitude_damping_error", "max_prob": 0.3}), seed=5)
    n_ops = len(cd["circuit"])
    p = sampler.sample(cd["circuit"], cd["ops"])
    assert p.shape == (2 ** n,) and p.dtype == np.float64 and abs(p.sum() - 1) < 1e-12     # reference test: shape
    assert len(cd["circuit"]) == n_ops                                                     # caller's circuit untouched
    # 4096 > 2^5 and noisy -> density-matrix method: exact probs equal the oracle on the instantiated circuit
    inst, _ = sampler.noise_model.instantiate(cd["circuit"])
    exact = sampler.last_probs.cpu().numpy()[0]
    np.testing.assert_allclose(exact, sim.density_matrix_probs(inst.ops, n), atol=1e-5)
    assert 0.5 * np.abs(p - exact).sum() < 0.08
    # forced trajectories give a statistically equal histogram
    s2 = qcd.UnitaryNoiseSampler(None, n_shots=4096, method="statevector", seed=6)
    s2.noise_model, s2.noisy = sampler.noise_model, True
    c2 = s2.simulate_circuit(cd["circuit"])
    assert c2.shots() == 4096
    assert 0.5 * np.abs(c2.hist / 4096 - exact).sum() < 0.08


def test_device_noise_sampler_shape(qcd):
    n = 5
    g = _graph(n, 2, qcd)
    cd = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(g)
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0, "readout_error": 0.05, "prob_meas0_prep1": 0.05,
                                          "prob_meas1_prep0": 0.05},
                                "cx": {"gate_error": 0.1, "gate_length": 400.0},
                                "h": {"gate_error": 0.01, "gate_length": 50.0}})
    sampler = qcd.DeviceNoiseSampler(None, n_shots=2048, noise_specs=spec, precision=64, seed=1)
    p = sampler.sample(cd["circuit"], user_backend=False)
    assert p.shape == (2 ** n,) and abs(p.sum() - 1) < 1e-12
    inst, ro = sampler.noise_model.instantiate(cd["circuit"])
    want = sim.apply_readout_to_probs(sim.density_matrix_probs(inst.ops, n), n, ro)
    assert 0.5 * np.abs(p - want).sum() < 0.12
    np.testing.assert_allclose(sampler.last_probs.cpu().numpy()[0], sim.density_matrix_probs(inst.ops, n), atol=1e-10)


def test_stabilizer_sampler_ideal_G4(qcd, goldens):
    """Ideal P4 graph state: n+1 Counts, the recorded 8-outcome supports, <S> = 1, genuine witness -1."""
    g = nx.path_graph(4)
    cd = qcd.CXGateCircuit(n_qubits=4, stochastic=False).build(g)
    stab = qcd.TothStabilizer(g, n_qubits=4).build()
    counts = qcd.StabilizerSampler(None, n_shots=1024, seed=3).sample(stab, cd["circuit"])
    assert len(counts) == 5                                            # reference tests/test_stabilizers.py:46-53
    g4 = goldens["G4"]
    order = g4["stabilizer_order_c18"]
    for name, c in zip(stab, counts):
        assert c.shots() == 1024
        assert set(c.keys()) <= set(g4["ideal_counts"][order.index(name)].keys())
    w = qcd.GenuineWitness(4, stab, counts)
    assert all(m == (1.0, 0.0) for m in w.stabilizer_measurements.values())
    res = w.estimate(g)
    assert res.value == -1.0 and res.variance == 0.0
    for r in qcd.BiSeparableWitness(4, stab, counts).estimate(g).values():
        assert r.value == -1.0


def test_witnesses_from_recorded_counts_G1_G2_G3(qcd, goldens):
    from collections import OrderedDict
    g1 = goldens["G1"]
    names = g1["stabilizer_order"]
    circs = OrderedDict((nm, None) for nm in names)
    wit = qcd.GenuineWitness(4, circs, g1["counts"])
    for name, (mean, std) in g1["stabilizer_measurements"]:
        assert wit.stabilizer_measurements[name][0] == mean
        assert wit.stabilizer_measurements[name][1] == pytest.approx(std, rel=1e-13, abs=1e-18)
    for name, diag in g1["diagonals"]:
        np.testing.assert_array_equal(wit.diagonals[name], np.array(diag))
    g = nx.Graph()
    g.add_edges_from([tuple(e) for e in g1["edges"]])
    res = wit.estimate(g)
    assert res.value == goldens["G2"]["genuine"]["value"]
    assert res.variance == pytest.approx(goldens["G2"]["genuine"]["variance"], abs=5e-9)
    # identity NOT last (the reference's dict is hash-ordered, quirk Q9): sigma = sqrt(sum std_k^2) all the same --
    # uncertainties' propagation of uarray(coef * means, stds).sum() (witnesses.py:136-144)
    i = names.index("IIII")
    order = [i] + [j for j in range(len(names)) if j != i]
    wit2 = qcd.GenuineWitness(4, OrderedDict((names[j], None) for j in order), [g1["counts"][j] for j in order])
    res2 = wit2.estimate(g)
    expected = sum(ms[1] ** 2 for _, ms in g1["stabilizer_measurements"]) ** 0.5
    assert res2.value == goldens["G2"]["genuine"]["value"] and res2.variance == pytest.approx(expected, rel=1e-13)
    bis = qcd.BiSeparableWitness(4, OrderedDict((nm, None) for nm in names), g1["counts"]).estimate(g)
    for idx, exp in enumerate(goldens["G3"]["bisep"]):
        assert list(bis[idx].W_ij) == exp["W_ij"] and bis[idx].value == exp["value"]
        assert bis[idx].variance == pytest.approx(exp["variance"], abs=5e-9)


@pytest.mark.parametrize("with_witness", [False, True])
def test_simulate(qcd, with_witness):
    """Reference tests/test_simulate.py: 4 qubits, 10 circuits, runs and returns the documented dict."""
    random.seed(0)
    np.random.seed(0)
    sim_specs = qcd.SimulationSpecs(n_qubits=4, n_circuits=10, data=qcd.GraphData(),
                                    circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=True))
    sampler_specs = qcd.SamplingSpecs(sampler=qcd.UnitaryNoiseSampler, noise_specs=qcd.unitary_noise_spec, n_shots=1024)
    ent = qcd.EntanglementSpecs(witness=qcd.GenuineWitness, stabilizer=qcd.TothStabilizer, n_shots=1024) \
        if with_witness else None
    res = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=4), entangle_specs=ent)
    assert sorted(res) == list(range(10))
    for r in res.values():
        assert r["prob_vec"].shape == (16,) and abs(r["prob_vec"].sum() - 1) < 1e-12 and r["graph_data"]
        if with_witness:
            assert r["entanglement"]["value"] == -1.0        # stabilizer settings are ideal (reference quirk Q7)


def test_dataset_rows_noisy_ideal(qcd):
    """(N, 2^n, 2) rows: column 1 (ideal) of a CX graph circuit is the uniform |amplitude|^2 = 2^-n support of the
    graph state; column 0 differs by the drawn damping noise; both are normalised histograms."""
    import torch
    n = 5
    np.random.seed(4)
    builder = qcd.CXGateCircuit(n_qubits=n, stochastic=True)
    cds = [builder.build(_graph(n, s, qcd)) for s in range(6)]
    sampler = qcd.UnitaryNoiseSampler(None, n_shots=8192, noise_specs=qcd.NoiseSpec(
        specs={"type": "amplitude_damping_error", "max_prob": 0.3}), seed=2)
    rows = qcd.sample_prob_rows(cds, sampler)
    assert rows.shape == (6, 2 ** n, 2) and rows.dtype == torch.float32 and rows.is_cuda
    r = rows.cpu().numpy()
    np.testing.assert_allclose(r.sum(axis=1), 1.0, atol=1e-6)
    for i, cd in enumerate(cds):
        p_ideal = np.abs(sim.ideal_statevector(cd["circuit"].ops, n)) ** 2
        assert 0.5 * np.abs(r[i, :, 1] - p_ideal).sum() < 0.06
        if cd["ops"]:
            assert 0.5 * np.abs(r[i, :, 0] - p_ideal).sum() > 0.01
    adj = qcd.adjacency_tensors(cds, {"h": 1, "cx": 2, "unitary": 3}, tensor_shape=(16, n, n))
    assert adj.shape == (6, 16, n, n)


def test_simulate_with_adjacency_tensor(qcd):
    random.seed(0)
    np.random.seed(0)
    sim_specs = qcd.SimulationSpecs(n_qubits=4, n_circuits=3, data=qcd.GraphData(),
                                    circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=True))
    sampler_specs = qcd.SamplingSpecs(sampler=qcd.UnitaryNoiseSampler, noise_specs=qcd.unitary_noise_spec, n_shots=256)
    res = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=4),
                       adj_tensor_specs=qcd.AdjTensorSpecs(tensor_shape=(16, 32, 32)))
    for r in res.values():
        assert r["adj_tensor"].shape == (16, 32, 32) and r["adj_tensor"].any()       # reference test: (16, 32, 32)


def test_simulate_device_noise_sampler(qcd):
    """Reference tests/test_simulate.py configuration: DeviceNoiseSampler, non-stochastic CX circuits, 4 qubits."""
    random.seed(1)
    np.random.seed(1)
    sim_specs = qcd.SimulationSpecs(n_qubits=4, n_circuits=10, data=qcd.GraphData(),
                                    circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=False))
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0, "readout_error": 0.05, "prob_meas0_prep1": 0.05,
                                          "prob_meas1_prep0": 0.05},
                                "cx": {"gate_error": 0.1, "gate_length": 400.0},
                                "h": {"gate_error": 0.01, "gate_length": 50.0}})
    sampler_specs = qcd.SamplingSpecs(sampler=qcd.DeviceNoiseSampler, noise_specs=spec, n_shots=1024)
    ent = qcd.EntanglementSpecs(witness=qcd.BiSeparableWitness, stabilizer=qcd.TothStabilizer, n_shots=512, noisy=True)
    res = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=4), entangle_specs=ent)
    assert sorted(res) == list(range(10))
    for r in res.values():
        assert r["prob_vec"].shape == (16,) and abs(r["prob_vec"].sum() - 1) < 1e-12
        assert len(r["entanglement"]["value"]) == len(r["graph_data"])
        assert all(-1.0 <= v <= 1.0 for v in r["entanglement"]["value"])
        assert any(v > -1.0 for v in r["entanglement"]["value"])       # noisy settings: no longer the ideal -1
    with pytest.raises(AssertionError):
        qcd.simulate(qcd.SimulationSpecs(n_qubits=4, n_circuits=1, data=qcd.GraphData(),
                                         circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=True)),
                     sampler_specs, qcd.GraphSpecs(max_num_edges=4))


_DEV_SPEC = {"qubit": {"T1": 100.0, "T2": 150.0, "readout_error": 0.05, "prob_meas0_prep1": 0.05, "prob_meas1_prep0": 0.03},
             "cx": {"gate_error": 0.1, "gate_length": 400.0}, "h": {"gate_error": 0.01, "gate_length": 50.0}}


def test_device_noise_native_route_equals_explicit_route(qcd):
    """Channels written natively at upload (qcd_upload_device_noise_programs) vs. written out in Python and uploaded as
    an explicit program: same exact distributions, and the same complex128 trajectories shot for shot."""
    import torch
    from qcdenoise_b200.samplers import get_engine, upload_batch
    n = 5
    random.seed(4)
    np.random.seed(4)
    s = qcd.DeviceNoiseSampler(None, n_shots=256, noise_specs=qcd.NoiseSpec(specs=_DEV_SPEC))
    circuits, models = [], []
    for i in range(9):
        c = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(_graph(n, 10 + i, qcd))["circuit"]
        s.transpile_circuit(c)
        s.build_noise_model(n)
        circuits.append(c)
        models.append(s.noise_model)
    eng = get_engine()
    assert all(m.native for m in models)
    ro, noisy = upload_batch(eng, circuits, models)
    assert noisy and all(m.native for m in models)
    p_nat = eng.run_density_matrix(64).clone()
    h_nat, o_nat = eng.run_trajectories(256, seed=77, precision=64, readout=ro, want_outcomes=True)
    inst = [m.instantiate(c) for m, c in zip(models, circuits)]          # materialises the Python channels
    assert not any(m.native for m in models)
    ro2, noisy2 = upload_batch(eng, circuits, models)                     # explicit route now
    np.testing.assert_array_equal(ro, ro2)
    p_exp = eng.run_density_matrix(64)
    h_exp, o_exp = eng.run_trajectories(256, seed=77, precision=64, readout=ro2, want_outcomes=True)
    assert float((p_nat - p_exp).abs().max()) < 1e-14
    assert torch.equal(o_nat, o_exp) and torch.equal(h_nat, h_exp)
    # and both equal the oracle's density matrix on the explicit program
    for i in (0, 4, 8):
        np.testing.assert_allclose(p_nat[i].cpu().numpy(), sim.density_matrix_probs(inst[i][0].ops, n), atol=1e-10)


def test_simulate_array_route_matches_oracle(qcd):
    """simulate()'s whole-batch route (native graphs, numpy gate lists and calibration draws, native channels): the
    exact distributions behind its histograms equal the oracle on the object route's explicit program built from
    the same numbers; adjacency tensors equal the Python walk."""
    from qcdenoise_b200 import batch
    from qcdenoise_b200.samplers import get_engine
    n, B = 5, 12
    spec = qcd.NoiseSpec(specs=_DEV_SPEC)
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n, max_num_edges=4)
    edges, n_edges = gs.sample_batch(B, seed=21)
    builder = qcd.CXGateCircuit(n_qubits=n, stochastic=False)
    enc = batch.build_gate_lists(builder, edges, n_edges)
    props = batch.draw_device_props(spec, enc, np.random.default_rng(3))
    t1t2, gp, ro = batch.device_noise_inputs(props)
    eng = get_engine()
    eng.upload_device_noise(enc, t1t2, gp)
    probs = eng.apply_readout(eng.run_density_matrix(64).clone(), ro).cpu().numpy()
    for i in (0, 5, 11):
        g = nx.Graph()
        g.add_nodes_from(range(n))
        g.add_edges_from(batch.edge_lists(edges, n_edges)[i])
        c = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(g)["circuit"]
        s = qcd.DeviceNoiseSampler(batch.device_props_dict(props, i), n_shots=16, noise_specs=spec)
        s.transpile_circuit(c)
        s.build_noise_model(n)
        inst, ro_i = s.noise_model.instantiate(c)
        want = sim.apply_readout_to_probs(sim.density_matrix_probs(inst.ops, n), n, ro_i)
        np.testing.assert_allclose(probs[i], want, atol=1e-10)
    # the public call
    random.seed(5)
    sim_specs = qcd.SimulationSpecs(n_qubits=n, n_circuits=40, data=qcd.GraphData(),
                                    circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=False))
    sampler_specs = qcd.SamplingSpecs(sampler=qcd.DeviceNoiseSampler, noise_specs=spec, n_shots=2048)
    adj_specs = qcd.AdjTensorSpecs(tensor_shape=(16, 8, 8))
    res = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=4), adj_tensor_specs=adj_specs)
    random.seed(5)
    res2 = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=4), adj_tensor_specs=adj_specs)
    assert sorted(res) == list(range(40))
    for i, r in res.items():
        assert r["prob_vec"].shape == (2 ** n,) and abs(r["prob_vec"].sum() - 1) < 1e-12
        assert r["adj_tensor"].shape == (16, 8, 8)
        g = nx.Graph()
        g.add_nodes_from(range(n))
        g.add_edges_from(r["graph_data"])
        assert nx.is_connected(g)
        c = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(g)["circuit"]
        np.testing.assert_array_equal(r["adj_tensor"], qcd.generate_adjacency_tensor(c, {"cx": 1, "h": 2}, (16, 8, 8)))
        assert r["graph_data"] == res2[i]["graph_data"]                  # random.seed() fixes the graphs ...
        assert abs(r["prob_vec"] - res2[i]["prob_vec"]).sum() < 0.35    # ... and the noise (shots: sampler seed differs)
    # noise visibly present: the ideal graph state is uniform on its support, these are not
    assert np.std([r["prob_vec"].max() for r in res.values()]) > 0
    # forcing the object route still works and has the same shape of results
    res3 = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=4), adj_tensor_specs=adj_specs,
                        route="objects")
    assert sorted(res3) == list(range(40)) and res3[0]["adj_tensor"].shape == (16, 8, 8)


def test_simulate_array_route_unitary_noise_matches_oracle(qcd, monkeypatch):
    """Array route for stochastic builders + UnitaryNoiseSampler: exact distributions of the batched programs equal
    the oracle's density matrix on the object route's explicit programs built from the same coins and damping draws."""
    from qcdenoise_b200 import batch
    from qcdenoise_b200.samplers import get_engine
    n, B = 5, 10
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n, max_num_edges=4)
    edges, n_edges = gs.sample_batch(B, seed=31)
    rng = np.random.default_rng(8)
    eng = get_engine()
    for cls in (qcd.CXGateCircuit, qcd.CPhaseGateCircuit):
        builder = cls(n_qubits=n, stochastic=True)
        coins = rng.integers(0, 2, size=(int(n_edges.sum()), batch.sites_per_edge(builder)))
        gammas = rng.uniform(0, 0.3, size=(int(coins.sum()), 2))
        eng.upload(batch.build_gate_lists(builder, edges, n_edges, coins=coins, gammas=gammas))
        probs = eng.run_density_matrix(64).cpu().numpy()
        hist, _ = eng.run_trajectories(4096, seed=3, precision=32)
        coin_iter, gamma_iter = iter(coins.ravel().tolist()), iter(gammas)
        with monkeypatch.context() as mp:
            mp.setattr(np.random, "choice", lambda k: next(coin_iter))
            mp.setattr(np.random, "uniform", lambda high, size: next(gamma_iter).copy())
            s = qcd.UnitaryNoiseSampler(None, n_shots=16, seed=1,
                                        noise_specs=qcd.NoiseSpec(specs={"type": "amplitude_damping_error", "max_prob": 0.3}))
            for i, ed in enumerate(batch.edge_lists(edges, n_edges)):
                g = nx.Graph()
                g.add_nodes_from(range(n))
                g.add_edges_from(ed)
                cd = cls(n_qubits=n, stochastic=True).build(g)
                s.build_noise_model(cd["ops"])
                inst, _ = s.noise_model.instantiate(cd["circuit"])
                want = sim.density_matrix_probs(inst.ops, n)
                np.testing.assert_allclose(probs[i], want, atol=1e-10)
                assert 0.5 * np.abs(hist[i].cpu().numpy() / 4096 - want).sum() < 0.08
    # the public call (reference examples/simulate_example.py: UnitaryNoiseSampler, stochastic CX, adjacency tensors)
    random.seed(2)
    sim_specs = qcd.SimulationSpecs(n_qubits=6, n_circuits=30, data=qcd.GraphData(),
                                    circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=True))
    sampler_specs = qcd.SamplingSpecs(sampler=qcd.UnitaryNoiseSampler, noise_specs=qcd.unitary_noise_spec, n_shots=1024)
    res = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=5),
                       adj_tensor_specs=qcd.AdjTensorSpecs(tensor_shape=(16, 32, 32)))
    assert sorted(res) == list(range(30))
    with_sites = 0
    for r in res.values():
        assert r["prob_vec"].shape == (64,) and abs(r["prob_vec"].sum() - 1) < 1e-12
        assert r["adj_tensor"].shape == (16, 32, 32)
        codes = set(np.unique(r["adj_tensor"]).tolist())
        assert {0.0, 1.0, 2.0} <= codes <= {0.0, 1.0, 2.0, 3.0}           # cx, h and (where a coin fell) unitary
        with_sites += 3.0 in codes
    assert 0 < with_sites <= 30


def test_c2_full_size_properties(qcd):
    """BASELINE config C2 at full size (8 qubits, 10 000 circuits, 1024 shots, device noise) through the array route,
    checked with size-independent properties: every histogram holds every shot; with unital gate noise (depolarizing
    only: T1 = T2 -> infinity) the pre-readout distribution of a graph state is exactly uniform, so the marginal of
    every measured bit is 1/2 + (p10 - p01) / 2 of ITS circuit's readout draw -- compared over 10.24 M shots; two runs
    over different circuit partitions give identical histograms."""
    import torch
    from qcdenoise_b200 import batch
    from qcdenoise_b200.samplers import get_engine
    n, B, shots = 8, 10000, 1024
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 1e15, "T2": 1e15, "prob_meas0_prep1": 0.3, "prob_meas1_prep0": 0.1},
                                "cx": {"gate_error": 0.25, "gate_length": 500.0}, "h": {"gate_error": 0.05, "gate_length": 50.0}})
    gs = qcd.GraphState(qcd.GraphDB(), n_qubits=n, max_num_edges=7)
    edges, n_edges = gs.sample_batch(B, seed=77)
    enc = batch.build_gate_lists(qcd.CXGateCircuit(n_qubits=n, stochastic=False), edges, n_edges)
    props = batch.draw_device_props(spec, enc, np.random.default_rng(5))
    props["qubit"]["T1"][:] = 1e15        # U(0, 1e15) draws can be small: pin both to "no relaxation"
    props["qubit"]["T2"][:] = 1e15
    t1t2, gp, ro = batch.device_noise_inputs(props)
    eng = get_engine()
    eng.upload_device_noise(enc, t1t2, gp)
    probs = eng.run_density_matrix(32)
    assert float((probs - 1.0 / 2 ** n).abs().max()) < 2e-6                     # unital noise keeps the uniform distribution
    hist = eng.sample_from_probs(probs, shots, seed=11, readout=ro)
    h = hist.cpu().numpy().astype(np.int64)
    assert (h.sum(axis=1) == shots).all()
    x = np.arange(2 ** n)
    p10, p01 = props["qubit"]["prob_meas1_prep0"], props["qubit"]["prob_meas0_prep1"]
    for q in range(n):
        ones = int(h[:, ((x >> q) & 1) == 1].sum())
        pq = 0.5 + 0.5 * (p10[:, q] - p01[:, q])                                # per circuit
        mean, var = shots * pq.sum(), shots * (pq * (1 - pq)).sum()
        assert abs(ones - mean) < 5 * np.sqrt(var), (q, ones, mean)
    # partition invariance: circuits [0, 4000) and [4000, 10000) uploaded separately give the same rows
    parts = []
    for lo, hi in ((0, 4000), (4000, B)):
        mine, (o_lo, o_hi) = batch.slice_programs(enc, lo, hi)
        eng.upload_device_noise(mine, t1t2[lo:hi], gp[o_lo:o_hi])
        pr = eng.run_density_matrix(32)
        parts.append(eng.sample_from_probs(pr, shots, seed=11, first_circuit=lo, readout=ro[lo:hi]))
    assert torch.equal(torch.cat(parts, dim=0), hist)


@pytest.mark.parametrize("noise", ["unitary", "device"])
def test_settings_read_off_one_density_matrix_equal_resimulation(qcd, noise):
    """StabilizerSampler in density-matrix mode simulates the graph circuit once and reads the n + 1 Toth settings off
    its density matrix (qcd_run_density_matrix_settings); the reference (and the fallback route) re-simulate the whole
    circuit per setting.  Same exact distributions, to 1e-12 in complex128, and they equal the oracle's."""
    n = 5
    g = _graph(n, 3, qcd)
    np.random.seed(2)
    random.seed(2)
    if noise == "unitary":
        cd = qcd.CXGateCircuit(n_qubits=n, stochastic=True).build(g)
        s = qcd.UnitaryNoiseSampler(None, n_shots=4096, noise_specs=qcd.NoiseSpec(
            specs={"type": "amplitude_damping_error", "max_prob": 0.3}))
        s.build_noise_model(cd["ops"])
    else:
        cd = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(g)
        s = qcd.DeviceNoiseSampler(None, n_shots=4096, noise_specs=qcd.NoiseSpec(specs=_DEV_SPEC))
        s.transpile_circuit(cd["circuit"])
        s.build_noise_model(n)
    stab = qcd.TothStabilizer(g, n_qubits=n).build()
    one = qcd.StabilizerSampler(None, n_shots=4096, precision=64, seed=9)
    c_one = one.sample(stab, cd["circuit"], noise_model=s.noise_model)
    p_one = one.last_probs.cpu().numpy()
    assert one._pick_method(n, True) == "density_matrix" and p_one.shape == (n + 1, 2 ** n)
    many = qcd.StabilizerSampler(None, n_shots=4096, precision=64, seed=9)
    many.settings_from_one_state = False
    c_many = many.sample(stab, cd["circuit"], noise_model=s.noise_model)
    p_many = many.last_probs.cpu().numpy()
    np.testing.assert_allclose(p_one, p_many, atol=1e-12)
    # oracle: explicit program of the graph circuit + each setting's sub-circuit
    for i, (name, sc) in enumerate(stab.items()):
        inst, _ = s.noise_model.instantiate(cd["circuit"].copy().compose(sc))
        np.testing.assert_allclose(p_one[i], sim.density_matrix_probs(inst.ops, n), atol=1e-10)
    # same seeds and (to rounding) the same distributions: the sampled histograms agree almost everywhere
    diff = sum(int(np.abs(a.hist.astype(np.int64) - b.hist.astype(np.int64)).sum()) for a, b in zip(c_one, c_many))
    assert diff <= 8
    w1 = qcd.GenuineWitness(n, stab, c_one).estimate(g)
    w2 = qcd.GenuineWitness(n, stab, c_many).estimate(g)
    assert abs(w1.value - w2.value) < 0.02


@pytest.mark.parametrize("kind", ["device-genuine", "unitary-biseparable", "device-ideal-settings"])
def test_simulate_array_route_with_entanglement_matches_oracle(qcd, kind, monkeypatch):
    """simulate() with an entanglement stage on the array route (Toth settings read off one density matrix per
    circuit, parity sums on the device, witnesses vectorised): witness values agree with the exact expectation values
    the oracle computes from the object route's explicit programs built from the same drawn numbers."""
    from qcdenoise_b200 import batch
    n, B, shots = 5, 8, 65536
    random.seed(13)
    device = kind.startswith("device")
    if device:
        sampler_specs = qcd.SamplingSpecs(sampler=qcd.DeviceNoiseSampler, noise_specs=qcd.NoiseSpec(specs=_DEV_SPEC), n_shots=1024)
    else:
        sampler_specs = qcd.SamplingSpecs(sampler=qcd.UnitaryNoiseSampler, n_shots=1024, noise_specs=qcd.NoiseSpec(
            specs={"type": "amplitude_damping_error", "max_prob": 0.3}))
    sim_specs = qcd.SimulationSpecs(n_qubits=n, n_circuits=B, data=qcd.GraphData(),
                                    circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=not device))
    noisy = kind != "device-ideal-settings"
    wit = qcd.BiSeparableWitness if kind == "unitary-biseparable" else qcd.GenuineWitness
    ent = qcd.EntanglementSpecs(witness=wit, stabilizer=qcd.TothStabilizer, n_shots=shots, noisy=noisy)
    diag = {}
    res = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=4), entangle_specs=ent, diagnostics=diag)
    lb = diag["batch"]
    assert "entanglement" in diag["timing"]          # the array route ran
    coin_iter = iter(lb["coins"].ravel().tolist()) if not device else None
    gamma_iter = iter(lb["gammas"]) if not device else None
    for i, ed in enumerate(batch.edge_lists(lb["edges"], lb["n_edges"])):
        assert res[i]["graph_data"] == ed
        g = nx.Graph()
        g.add_nodes_from(range(n))
        g.add_edges_from(ed)
        if device:
            cd = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(g)
            s = qcd.DeviceNoiseSampler(batch.device_props_dict(lb["device_props"], i), n_shots=16,
                                       noise_specs=qcd.NoiseSpec(specs=_DEV_SPEC))
            s.transpile_circuit(cd["circuit"])
            s.build_noise_model(n)
        else:
            with monkeypatch.context() as mp:
                mp.setattr(np.random, "choice", lambda k: next(coin_iter))
                mp.setattr(np.random, "uniform", lambda high, size: next(gamma_iter).copy())
                cd = qcd.CXGateCircuit(n_qubits=n, stochastic=True).build(g)
                s = qcd.UnitaryNoiseSampler(None, n_shots=16, noise_specs=sampler_specs.noise_specs)
                s.build_noise_model(cd["ops"])
        model = s.noise_model if noisy else qcd.NoiseModel()
        stab = qcd.TothStabilizer(g, n_qubits=n).build()
        x = np.arange(2 ** n)
        means = []
        for name, sc in stab.items():
            inst, ro = model.instantiate(cd["circuit"].copy().compose(sc))
            p = sim.density_matrix_probs(inst.ops, n)
            if ro is not None:
                p = sim.apply_readout_to_probs(p, n, ro)
            sign = 1 - 2 * (np.array([bin(v & qcd.pauli_mask(name)).count("1") for v in x]) & 1)
            means.append(float((p * sign).sum()))
        means = np.array(means)
        tol = 6 * np.sqrt(n / shots) + 1e-9
        if wit is qcd.GenuineWitness:
            want = (n - 1) * means[n] - means[:n].sum()
            assert abs(res[i]["entanglement"]["value"] - want) < tol, (i, res[i]["entanglement"], want)
            assert 0 <= res[i]["entanglement"]["variance"] < 0.05
            if not noisy:
                assert abs(res[i]["entanglement"]["value"] + 1.0) < 1e-12       # ideal graph state: every <g_k> = 1
        else:
            vals = res[i]["entanglement"]["value"]
            assert len(vals) == len(ed) == len(res[i]["entanglement"]["variance"])
            for (a, b), v in zip(ed, vals):
                assert abs(v - (1 - means[a] - means[b])) < tol


def test_dataset_rows_device_noise_native_upload(qcd):
    """sample_prob_rows with a DeviceNoiseSampler: one calibration-defined model for the whole batch, channels written
    natively at upload; the noisy column follows the oracle's distribution of the explicit program."""
    n = 5
    random.seed(8)
    cds = [qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(_graph(n, 40 + s, qcd)) for s in range(4)]
    sampler = qcd.DeviceNoiseSampler(None, n_shots=16384, noise_specs=qcd.NoiseSpec(specs=_DEV_SPEC), seed=4)
    # one backend for all circuits: every (gate, qubits) of the batch gets an entry
    merged = qcd.Circuit(n, [op for cd in cds for op in cd["circuit"].ops])
    sampler.transpile_circuit(merged)
    sampler.build_noise_model(n)
    assert sampler.noise_model.native
    rows = qcd.sample_prob_rows(cds, sampler).cpu().numpy()
    assert sampler.noise_model.native and rows.shape == (4, 2 ** n, 2)
    for i, cd in enumerate(cds):
        inst, ro = sampler.noise_model.instantiate(cd["circuit"])
        want = sim.apply_readout_to_probs(sim.density_matrix_probs(inst.ops, n), n, ro)
        assert 0.5 * np.abs(rows[i, :, 0] - want).sum() < 0.05
        assert abs(rows[i, :, 1].sum() - 1) < 1e-6


def test_simulate_pipelined_chunks_equal_single_upload(qcd, monkeypatch):
    """Large density-matrix batches run as 4 chunks on two alternating contexts (host lowering of chunk k + 1 overlaps
    the kernels of chunk k).  Same seeds -> bit-identical histograms with and without the chunking."""
    import sys
    simmod = sys.modules["qcdenoise_b200.simulate"]      # the package attribute `simulate` is the function
    n, B = 5, 600
    spec = qcd.NoiseSpec(specs=_DEV_SPEC)
    sim_specs = qcd.SimulationSpecs(n_qubits=n, n_circuits=B, data=qcd.GraphData(),
                                    circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=False))
    sampler_specs = qcd.SamplingSpecs(sampler=qcd.DeviceNoiseSampler, noise_specs=spec, n_shots=512)
    out = {}
    for name, threshold in (("chunked", 64), ("single", 10 ** 9)):
        monkeypatch.setattr(simmod, "PIPELINE_MIN", threshold)
        random.seed(99)
        np.random.seed(99)        # the sampler's shot seed comes from np.random
        out[name] = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=4))
    for i in range(B):
        assert out["chunked"][i]["graph_data"] == out["single"][i]["graph_data"]
        np.testing.assert_array_equal(out["chunked"][i]["prob_vec"], out["single"][i]["prob_vec"])
