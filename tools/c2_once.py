#!/usr/bin/env python
"""One C2-shaped qcd.simulate() call (8 qubits, device noise, 1024 shots; default 4096 circuits) -- short enough to
run under `ncu --metrics gpu__time_duration.sum` for a launch list of the density-matrix array route."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import qcdenoise_b200 as qcd  # noqa: E402

n_circ = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
random.seed(1234)
spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0, "readout_error": 0.25, "prob_meas0_prep1": 0.1,
                                      "prob_meas1_prep0": 0.1},
                            "cx": {"gate_error": 0.25, "gate_length": 500.0}, "h": {"gate_error": 0.05, "gate_length": 50.0}})
sim_specs = qcd.SimulationSpecs(n_qubits=8, n_circuits=n_circ, data=qcd.GraphData(),
                                circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=False))
sampler_specs = qcd.SamplingSpecs(sampler=qcd.DeviceNoiseSampler, noise_specs=spec, n_shots=1024)
res = qcd.simulate(sim_specs, sampler_specs, qcd.GraphSpecs(max_num_edges=7))
torch.cuda.synchronize()
print(len(res), float(res[0]["prob_vec"].sum()))
