#!/usr/bin/env python
"""qcd.simulate() end to end in the density-matrix regime (1024 shots, n <= 9) for both noise models and both routes
(QCD_DM_ON_CHIP=1 on-chip interpreter / 0 lowered sweeps): which route should be the default at which size."""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import qcdenoise_b200 as qcd  # noqa: E402

dev_spec = qcd.NoiseSpec(specs={"qubit": {"T1": 100.0, "T2": 100.0, "readout_error": 0.25, "prob_meas0_prep1": 0.1,
                                          "prob_meas1_prep0": 0.1},
                                "cx": {"gate_error": 0.25, "gate_length": 500.0}, "h": {"gate_error": 0.05, "gate_length": 50.0}})
for n in [int(a) for a in sys.argv[1:]] or [7, 8, 9]:
    for kind in ("unitary", "device"):
        B = 4000
        random.seed(1)
        if kind == "unitary":
            sim_specs = qcd.SimulationSpecs(n_qubits=n, n_circuits=B, data=qcd.GraphData(),
                                            circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=True))
            sam = qcd.SamplingSpecs(sampler=qcd.UnitaryNoiseSampler, noise_specs=qcd.unitary_noise_spec, n_shots=1024)
        else:
            sim_specs = qcd.SimulationSpecs(n_qubits=n, n_circuits=B, data=qcd.GraphData(),
                                            circuit=qcd.CircuitSpecs(builder=qcd.CXGateCircuit, stochastic=False))
            sam = qcd.SamplingSpecs(sampler=qcd.DeviceNoiseSampler, noise_specs=dev_spec, n_shots=1024)
        best = None
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            qcd.simulate(sim_specs, sam, qcd.GraphSpecs(max_num_edges=7))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if it and (best is None or dt < best):
                best = dt
        print(json.dumps({"n": n, "noise": kind, "on_chip": os.environ.get("QCD_DM_ON_CHIP", "1"), "circuits": B,
                          "ms": round(1e3 * best, 1), "circuits_per_s": round(B / best)}), flush=True)
