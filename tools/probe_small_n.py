#!/usr/bin/env python
"""qcd.simulate() on the array route for the reference's example shape (UnitaryNoiseSampler, stochastic CX circuits,
1024 shots: examples/simulate_example.py, examples/unitary_graphstates/) at n = 9 ... 13: which engine path each size
takes (density matrix up to 2^n < shots, tile kernel below 13 qubits, register-only kernel from 13) and what it costs."""
import sys, time, random, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch
import qcdenoise_b200 as qcd
for n, B in ((9, 4000), (11, 4000), (12, 2000), (13, 1000)):
    random.seed(1)
    sim_specs = qcd.SimulationSpecs(n_qubits=n, n_circuits=B, data=qcd.GraphData(),
                                    circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=True))
    sampler_specs = qcd.SamplingSpecs(sampler=qcd.UnitaryNoiseSampler, noise_specs=qcd.unitary_noise_spec, n_shots=1024)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        diag = {}; res = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=7), diagnostics=diag)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = qcd.get_engine().stats()
    print(json.dumps({"n": n, "circuits": B, "s": round(dt, 4), "circuits_per_s": round(B / dt), "stages": {k: round(v, 3) for k, v in diag["timing"].items()},
                      "sweep_launches": st["sweep_launches"], "node_sweeps": st["node_sweeps"], "leaves": st["leaves"]}), flush=True)
