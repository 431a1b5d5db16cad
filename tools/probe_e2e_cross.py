#!/usr/bin/env python
"""Sibling groups through the reference-facing API: one heavy graph of the bench job, sampler + stabilizer sampler calls,
engine statistics and wall time with the option on and off."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import qcdenoise_b200 as qcd  # noqa: E402
from qcdenoise_b200.samplers import get_engine  # noqa: E402

shots = 1_000_000
w = bench.build_workload(2, shots)
g, cd, stab = w["graph"], w["circuit_dict"], w["stab"]
sampler = qcd.UnitaryNoiseSampler(None, n_shots=shots, method="statevector", precision=32, seed=bench.SEED)
ssam = qcd.StabilizerSampler(None, n_shots=shots, method="statevector", precision=32, seed=bench.SEED + 1)
eng = get_engine(None)
for on in (1, 0, 1, 0):
    eng.set_option("leaf_cross", on)
    np.random.set_state(w["rng_state"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p = sampler.sample(cd["circuit"], cd["ops"])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    s1 = eng.stats()
    counts = ssam.sample(stab, cd["circuit"], noise_model=sampler.noise_model)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    s2 = eng.stats()
    wit = qcd.GenuineWitness(20, stab, counts).estimate(g)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"leaf_cross={on}: sample {1e3 * (t1 - t0):.1f} ms (leaves {s1['leaves']}, crossed {s1['cross_leaves']}), "
          f"stabilizer settings {1e3 * (t2 - t1):.1f} ms (leaves {s2['leaves']}, crossed {s2['cross_leaves']}, passes "
          f"{s2['cross_passes']}, node sweeps {s2['node_sweeps']}), witness {1e3 * (t3 - t2):.1f} ms  W = {wit.value:.5f}")
