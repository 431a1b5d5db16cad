#!/usr/bin/env python
"""Where the end-to-end C4 step of bench.py spends its host time: cProfile of UnitaryNoiseSampler.sample +
StabilizerSampler.sample + GenuineWitness.estimate (the `e2e` leg), plus the engine's own trace (QCD_TRACE=1)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("QCD_TRACE", "1")

import torch  # noqa: E402

import bench  # noqa: E402
import qcdenoise_b200 as qcd  # noqa: E402

shots = 1_000_000
work = bench.build_workload(0, shots)
g, cd, stab = work["graph"], work["circuit_dict"], work["stab"]
sampler = qcd.UnitaryNoiseSampler(None, n_shots=shots, method="statevector", precision=32, seed=1)
ssam = qcd.StabilizerSampler(None, n_shots=shots, method="statevector", precision=32, seed=2)


def step():
    np.random.seed(1234)
    t0 = time.perf_counter()
    p = sampler.sample(cd["circuit"], cd["ops"])
    t1 = time.perf_counter()
    counts = ssam.sample(stab, cd["circuit"], noise_model=sampler.noise_model)
    t2 = time.perf_counter()
    w = qcd.GenuineWitness(bench.N_QUBITS, stab, counts).estimate(g)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2


step()
step()
print("sample / stabilizer sample / witness (s):", step(), flush=True)
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
