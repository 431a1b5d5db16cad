#!/usr/bin/env python
"""Run the BASELINE.json configurations other than the bench.py headline (C4) once each on cuda:0 and print one JSON
line per configuration: throughput, sweep-kernel algorithmic GB/s (CUDA events), launches.  These are the numbers
quoted in DESIGN.md; bench.py stays the contract benchmark.

    python tools/run_configs.py [C1] [C2] [C3] [C5] [--circuits N]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import networkx as nx  # noqa: E402
import torch  # noqa: E402

import qcdenoise_b200 as qcd  # noqa: E402


def graphs(n):
    with open(os.path.join(ROOT, "tests", "golden", "graphs.json")) as f:
        return [[tuple(e) for e in g] for g in json.load(f)["graphs"][str(n)]]


def nxg(n, edges):
    g = nx.Graph()
    g.add_nodes_from(range(n))
    g.add_edges_from(edges)
    return g


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def report(name, eng, ms, units, unit, extra=None):
    st = eng.stats()
    gbs = st["sweep_bytes"] / (st["sweep_ms"] * 1e-3) / 1e9 if st["sweep_ms"] > 0 else None
    line = {"config": name, "ms": ms, "value": units / (ms * 1e-3), "unit": unit, "sweep_GBps_algorithmic": gbs,
            "sweep_launches": st["sweep_launches"], "states_swept": st["node_sweeps"], "sweep_ms": st["sweep_ms"]}
    line.update(extra or {})
    print(json.dumps(line), flush=True)


def c1(eng):
    """4-qubit linear graph, depolarizing noise, 1024 shots (the reference's own CPU-runnable case)."""
    n, edges = 4, [(0, 1), (1, 2), (2, 3)]
    cd = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(nxg(n, edges))
    nm = qcd.NoiseModel()
    for op in cd["circuit"].ops:
        k = len(op[1])
        nm.gate_errors.setdefault((op[0], tuple(op[1])), [[("pauli", tuple(range(k)),
                                                            qcd.channels.depolarizing_probs(0.001 if k == 1 else 0.01, k))]])
    prog, _ = nm.instantiate(cd["circuit"])
    eng.upload([prog])
    ms, probs = timed(lambda: eng.run_density_matrix(64))
    report("C1 density matrix c128", eng, ms, 1, "circuits/s")
    ms, _ = timed(lambda: eng.run_trajectories(1024, seed=1, precision=32))
    report("C1 trajectories c64, 1024 shots", eng, ms, 1024, "shots/s")


def c2(eng, n_circ):
    """8-qubit random graphs, batch of circuits, device noise model, count histograms (density matrix + sampling)."""
    import random
    n = 8
    gs = graphs(n)
    random.seed(1234)
    np.random.seed(1234)
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0, "readout_error": 0.25, "prob_meas0_prep1": 0.1,
                                          "prob_meas1_prep0": 0.1},
                                "cx": {"gate_error": 0.25, "gate_length": 500.0},
                                "h": {"gate_error": 0.05, "gate_length": 50.0}})
    sampler = qcd.DeviceNoiseSampler(None, n_shots=1024, noise_specs=spec)
    progs, ros = [], []
    t0 = time.perf_counter()
    for i in range(n_circ):
        cd = qcd.CXGateCircuit(n_qubits=n, stochastic=False).build(nxg(n, gs[i % len(gs)]))
        sampler.transpile_circuit(cd["circuit"])
        sampler.build_noise_model(n)
        p, ro = sampler.noise_model.instantiate(cd["circuit"])
        progs.append(p)
        ros.append(ro)
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng.upload(progs)
    t_up = time.perf_counter() - t0
    ro = np.stack(ros)

    def run():
        probs = eng.run_density_matrix(32)
        st = eng.stats()
        hist = eng.sample_from_probs(probs, 1024, seed=1234, readout=ro)
        return st, hist
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, hist = run()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    assert int(hist.sum()) == 1024 * n_circ
    gbs = st["sweep_bytes"] / (st["sweep_ms"] * 1e-3) / 1e9
    print(json.dumps({"config": f"C2 n=8 device noise, {n_circ} circuits, density matrix c64 + 1024-shot sampling",
                      "ms": ms, "value": n_circ / (ms * 1e-3), "unit": "circuits/s", "shots_per_s": 1024 * n_circ / (ms * 1e-3),
                      "sweep_GBps_algorithmic": gbs, "sweep_launches": st["sweep_launches"],
                      "states_swept": st["node_sweeps"], "host_build_s": t_build, "upload_lower_s": t_up}), flush=True)


def c2_simulate(n_circ):
    """C2 end to end through the public call: qcd.simulate() with DeviceNoiseSampler on 8-qubit random graph states,
    host stages included (graph sampling, gate lists, calibration draws, channel construction, lowering, upload,
    simulation, histograms back to the host, adjacency tensors)."""
    import random
    random.seed(1234)
    np.random.seed(1234)
    n = 8
    spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0, "readout_error": 0.25, "prob_meas0_prep1": 0.1,
                                          "prob_meas1_prep0": 0.1},
                                "cx": {"gate_error": 0.25, "gate_length": 500.0},
                                "h": {"gate_error": 0.05, "gate_length": 50.0}})
    sim_specs = qcd.SimulationSpecs(n_qubits=n, n_circuits=n_circ, data=qcd.GraphData(),
                                    circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=False))
    sampler_specs = qcd.SamplingSpecs(sampler=qcd.DeviceNoiseSampler, noise_specs=spec, n_shots=1024)
    adj = qcd.AdjTensorSpecs(tensor_shape=(16, 8, 8))
    ent = qcd.EntanglementSpecs(witness=qcd.GenuineWitness, stabilizer=qcd.TothStabilizer, n_shots=1024, noisy=True)
    for route, count, with_ent in (("auto", n_circ, False), ("auto", n_circ, False), ("objects", min(n_circ, 1000), False),
                                   ("auto", n_circ, True), ("auto", n_circ, True), ("objects", min(n_circ, 500), True)):
        sim_specs.n_circuits = count
        t0 = time.perf_counter()
        diag = {}
        res = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=7), adj_tensor_specs=adj, route=route,
                           entangle_specs=ent if with_ent else None, diagnostics=diag)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert len(res) == count and abs(res[count - 1]["prob_vec"].sum() - 1) < 1e-9
        if with_ent:
            assert all(-8.0 <= r["entanglement"]["value"] <= 8.0 for r in res.values())
        print(json.dumps({"config": f"C2 end to end: simulate() n=8 device noise, {count} circuits x 1024 shots + adjacency "
                                    f"tensors{' + 9 Toth settings x 1024 shots + genuine witness' if with_ent else ''}, "
                                    f"route={route}", "s": dt, "value": count / dt, "unit": "circuits/s",
                          "shots_per_s": 1024 * count / dt,
                          "stages_s": {k: round(v, 4) for k, v in diag["timing"].items()}
                          if route == "auto" else None}), flush=True)


def c3(eng, n_circ):
    """14-qubit density matrix with custom-unitary-noise Kraus channels + witness estimation."""
    n = 14
    gs = graphs(n)
    np.random.seed(1234)
    sampler = qcd.UnitaryNoiseSampler(None, n_shots=1024, method="density_matrix")
    progs, masks = [], None
    for i in range(n_circ):
        g = nxg(n, gs[i % len(gs)])
        cd = qcd.CXGateCircuit(n_qubits=n, stochastic=True).build(g)
        sampler.build_noise_model(cd["ops"])
        progs.append(sampler.noise_model.instantiate(cd["circuit"])[0])
    eng.upload(progs)

    def run():
        probs = eng.run_density_matrix(32)
        st = eng.stats()
        hist = eng.sample_from_probs(probs, 1024, seed=1234)
        return st, probs, hist
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st, probs, hist = run()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    assert abs(float(probs.sum()) - n_circ) < 1e-2 * n_circ
    gbs = st["sweep_bytes"] / (st["sweep_ms"] * 1e-3) / 1e9
    print(json.dumps({"config": f"C3 n=14 density matrix c64, {n_circ} circuits, amplitude-damping sites", "ms": ms,
                      "value": n_circ / (ms * 1e-3), "unit": "circuits/s", "sweep_GBps_algorithmic": gbs,
                      "sweep_launches": st["sweep_launches"], "states_swept": st["node_sweeps"]}), flush=True)


def c3_witness(eng, n_circ):
    """C3 with its witness estimation: per circuit the 15 Toth settings -> counts -> genuine witness.  Route A reads all
    settings off ONE density-matrix simulation per circuit (qcd_run_density_matrix_settings); route B re-simulates the
    circuit once per setting, which is what the reference does (stabilizers.py:183-215)."""
    from qcdenoise_b200.samplers import run_batch
    from qcdenoise_b200.stabilizers import run_settings_batch, single_qubit_settings
    n = 14
    gs = graphs(n)
    np.random.seed(1234)
    sampler = qcd.UnitaryNoiseSampler(None, n_shots=1024, method="density_matrix")
    gl, circuits, models, stabs = [], [], [], []
    for i in range(n_circ):
        g = nxg(n, gs[i % len(gs)])
        cd = qcd.CXGateCircuit(n_qubits=n, stochastic=True).build(g)
        sampler.build_noise_model(cd["ops"])
        gl.append(g)
        circuits.append(cd["circuit"])
        models.append(sampler.noise_model)
        stabs.append(qcd.TothStabilizer(g, n).build())
    out = {}
    for route in ("one simulation per circuit", "one simulation per setting"):
        ssam = qcd.StabilizerSampler(None, n_shots=1024, method="density_matrix", seed=5)
        ssam.noisy = True
        for it in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if route.endswith("circuit"):
                settings = [single_qubit_settings(sd, m, n) for sd, m in zip(stabs, models)]
                per = run_settings_batch(ssam, circuits, models, settings)
            else:
                allc = [c.copy(name=nm).compose(sc) for c, sd in zip(circuits, stabs) for nm, sc in sd.items()]
                allm = [m for m, sd in zip(models, stabs) for _ in sd]
                flat = run_batch(ssam, allc, allm)
                per = [flat[i * (n + 1):(i + 1) * (n + 1)] for i in range(n_circ)]
            w = [qcd.GenuineWitness(n, sd, cs).estimate(g).value for sd, cs, g in zip(stabs, per, gl)]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out[route] = (dt, w)
        print(json.dumps({"config": f"C3 n=14 density matrix + Toth witnesses, {n_circ} circuits x 15 settings x 1024 "
                                    f"shots: {route}", "s": dt, "value": n_circ / dt, "unit": "circuits (with witness)/s",
                          "witness_mean": float(np.mean(w))}), flush=True)
    a, b = out["one simulation per circuit"][1], out["one simulation per setting"][1]
    assert max(abs(x - y) for x, y in zip(a, b)) < 0.05, (a, b)


def c5(eng, n_circ, shots):
    """28-qubit statevector noisy trajectories (one GPU's share of the 8-GPU configuration)."""
    n = 28
    gs = graphs(n)
    np.random.seed(1234)
    sampler = qcd.UnitaryNoiseSampler(None, n_shots=shots, method="statevector")
    progs = []
    for i in range(n_circ):
        cd = qcd.CXGateCircuit(n_qubits=n, stochastic=True).build(nxg(n, gs[i % len(gs)]))
        sampler.build_noise_model(cd["ops"])
        progs.append(sampler.noise_model.instantiate(cd["circuit"])[0])
    eng.upload(progs)
    t0 = time.perf_counter()
    _, out = eng.run_trajectories(shots, seed=1234, precision=32, want_outcomes=True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    st = eng.stats()
    # 29 Toth generator parities per circuit from the sparse outcome lists (computational-basis settings only here)
    masks = np.array([[qcd.pauli_mask(nm.replace("X", "Z")) for nm in qcd.TothStabilizer(nxg(n, gs[i % len(gs)]), n).build()]
                      for i in range(n_circ)], dtype=np.uint32)
    ss = eng.parity_outcomes(out.reshape(n_circ, shots), masks)
    assert ss.shape == (n_circ, n + 1) and int(ss[:, -1].min()) == shots
    gbs = st["sweep_bytes"] / (st["sweep_ms"] * 1e-3) / 1e9
    print(json.dumps({"config": f"C5 n=28 trajectories c64, {n_circ} circuits x {shots} shots (single GPU share)", "ms": ms,
                      "value": n_circ * shots / (ms * 1e-3), "unit": "shots/s", "sweep_GBps_algorithmic": gbs,
                      "sweep_launches": st["sweep_launches"], "states_swept": st["node_sweeps"],
                      "leaves": st["leaves"], "max_live_states": st["max_live_states"]}), flush=True)


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["C1", "C2", "C3", "C3W", "C5"]
    eng = qcd.get_engine(0)
    eng.set_option("time_sweeps", 1)
    if "C1" in which:
        c1(eng)
    if "C2" in which:
        c2(eng, 10000)
    if "C2" in which or "C2S" in which:
        c2_simulate(10000)
    if "C3" in which:
        c3(eng, 16)
    if "C3W" in which:
        c3_witness(eng, 8)
    if "C5" in which:
        c5(eng, 8, 8)
